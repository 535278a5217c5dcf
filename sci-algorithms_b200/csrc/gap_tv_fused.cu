// placeholder until the one-pass kernel lands
#include "internal.cuh"
namespace scipnp {
bool fused_supported(int, int, int, int, int, int) { return false; }
size_t fused_workspace_bytes(int, int, int, int, int) { return 0; }
int launch_fused(const FusedArgs&, cudaStream_t) { set_error("fused path not built"); return SCIPNP_ESTATE; }
}
extern "C" {
size_t scipnp_gap_tv_workspace_bytes(int B, int H, int W, int C, int T) { return scipnp::fused_workspace_bytes(B, H, W, C, T); }
int scipnp_gap_tv_fused(const float*, float*, const float*, float*, const float*, const float*, const float*,
                        float, int, double, double, int, int, int, int, int, int, void*, size_t, int*, void*) {
    scipnp::set_error("fused path not built"); return SCIPNP_ESTATE;
}
}
