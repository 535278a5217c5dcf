// Instantiates the CASSI variants of the fused GAP-TV kernel (coded aperture read at per-band index
// offsets instead of a shifted mask stack), R = 4 dual updates (tv_iter_max = 5).
#include "gap_tv_stream.cuh"
namespace scipnp { namespace fusedk {
int launch_stream_cassi_r4(int mode, int K, const FusedParams& fp, const FusedMaps& maps, const CassiParams& cp,
                           dim3 grid, cudaStream_t st) {
    if (mode == MODE_GAP_ACC) return launch_stream_mode<4, MODE_GAP_ACC, true>(K, fp, maps, grid, st, cp);
    return launch_stream_mode<4, MODE_GAP_PLAIN, true>(K, fp, maps, grid, st, cp);
}
} }
