// Cross-file internal declarations of libscipnp.
#pragma once
#include "common.cuh"

namespace scipnp {

enum { MODE_GAP_ACC = 0, MODE_GAP_PLAIN = 1, MODE_ADMM = 2, MODE_TV = 3 };   // MODE_TV: warp-specialised kernel only

// ops.cu
int launch_project(int mode, const float* a_in, const float* b_in, float* x_out, float* f_out,
                   const float* y1_in, float* y1_out, const float* y, const float* Phi,
                   const float* Phi_sum, float lambda, float gamma, int B, int H, int W, int C,
                   int phi_batched, cudaStream_t st);

int launch_clip01(float* x, size_t n, cudaStream_t st);
// Phi_sum and x0 = At(y) in one pass over the mask stack; SCIPNP_EINVAL (nothing launched) if the shape is not covered
int launch_init_x0_phisum(const float* y, const float* Phi, float* x, float* phisum, int B, int H, int W, int C,
                          int phi_batched, cudaStream_t st);
int launch_sq_err(const float* a, const float* b, size_t n_per_batch, int B, double* sums,
                  cudaStream_t st);

// tv_exact.cu
size_t tv_workspace_bytes(int B, int H, int W, int C);
// Row-tiled scenes: energies over rows [e_lo, e_hi) only, `reduce` (if set) sums the 2*B*C partial energies of
// a dual iteration over the ranks before the stopping rule is applied (scipnp_energy_reduce_fn of scipnp.h).
struct TvTiling {
    int e_lo = 0, e_hi = 0;                       // e_hi == 0: every row
    long long total_rows = 0;                     // rows of the whole scene (0: H)
    int (*reduce)(double* dev, int n, void* stream, void* user) = nullptr;
    void* user = nullptr;
};
int tv_chambolle_exact(const float* in, float* out, double weight, double eps, int T, int B, int H,
                       int W, int C, void* workspace, size_t ws_bytes, int* n_exec_dev,
                       double* energy_dev, int energy_cap, cudaStream_t st, const TvTiling* tiling = nullptr);

// gap_tv_fused.cu
struct FusedArgs {
    const float* x_in;  float* x_out;
    const float* y1_in; float* y1_out;
    const float* y; const float* Phi; const float* Phi_sum;
    // ADMM (mode == MODE_ADMM): x_in = theta, b_in/b_out multiplier, xproj_out = x (may be null)
    const float* b_in; float* b_out; float* xproj_out;
    float lambda, gamma;
    double tv_weight, tv_eps;
    int tv_iter_max;
    int mode;
    int B, H, W, C, phi_batched;
    void* workspace; size_t workspace_bytes;
    int* flag;                // set nonzero if the energy criterion would have fired
    bool workspace_clean;     // the energy accumulators are known to be zero (see launch_fused)
    // CASSI: Phi == nullptr and the coded aperture mask2d [H][mask_w] is read at offset step*c
    const float* mask2d; int cassi_step; int mask_w;
    int clip01;               // clip the TV output to [0,1]
    // Warp-specialised kernel only: produce rows [out_lo, out_hi) of the H local rows (out_hi == 0: all of them), keep
    // this launch's [B][C][R] energies in energy_log, push the rows next to a tile seam to the neighbours (TilePush).
    int out_lo = 0, out_hi = 0;
    double* energy_log = nullptr;
    const struct TilePush* push = nullptr;
};
// Halo push of the row-tiled mode (gap_tv_ws.cuh): the neighbours' OUTPUT buffers of this iteration (IPC-mapped),
// their local row counts and the global row of their local row 0, the flags.
struct TilePush {
    float* x_up = nullptr; float* x_dn = nullptr;         // neighbour's x_out (null: no neighbour on that side)
    float* y1_up = nullptr; float* y1_dn = nullptr;       // neighbour's y1_out (accelerated GAP)
    float* b_up = nullptr; float* b_dn = nullptr;         // neighbour's b_out (ADMM)
    int up_rows = 0, dn_rows = 0;                         // local rows of the neighbours' buffers
    int up_shift = 0, dn_shift = 0;                       // my local row + shift = the neighbour's local row
    const int* wait_up = nullptr; const int* wait_dn = nullptr;
    int* sig_up = nullptr; int* sig_dn = nullptr;
    int wait_epoch = 0, sig_epoch = 0;
    int* timeout_flag = nullptr;
};
bool fused_supported(int mode, int B, int H, int W, int C, int tv_iter_max);
bool fused_cassi_supported(int mode, int B, int H, int W, int C, int tv_iter_max);
size_t fused_workspace_bytes(int B, int H, int W, int C, int tv_iter_max);
int launch_fused(const FusedArgs& a, cudaStream_t st);
// gap_tv_ws.cu: warp-specialised second-generation kernel (GAP modes, C <= 24, W % 4 == 0)
bool fused_ws_supported(const FusedArgs& a);
int launch_fused_ws(const FusedArgs& a, cudaStream_t st);
int fused_variant();

}  // namespace scipnp
