// Instantiates the fused GAP-TV kernels with R = 4 dual updates (tv_iter_max = 5).
#include "gap_tv_stream.cuh"
namespace scipnp { namespace fusedk {
SCIPNP_INSTANTIATE_FUSED_R(4)
} }
