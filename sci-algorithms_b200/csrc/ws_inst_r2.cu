// Instantiates the warp-specialised fused GAP-TV kernels with R = 2 dual updates (tv_iter_max = 3).
#include "gap_tv_ws.cuh"
namespace scipnp { namespace wsk {
SCIPNP_INSTANTIATE_WS_R(2)
} }
