// Instantiates the fused GAP-TV kernels with R = 3 dual updates (tv_iter_max = 4).
#include "gap_tv_stream.cuh"
namespace scipnp { namespace fusedk {
SCIPNP_INSTANTIATE_FUSED_R(3)
} }
