// Instantiates the warp-specialised fused GAP-TV kernels with R = 3 dual updates (tv_iter_max = 4).
#include "gap_tv_ws.cuh"
namespace scipnp { namespace wsk {
SCIPNP_INSTANTIATE_WS_R(3)
} }
