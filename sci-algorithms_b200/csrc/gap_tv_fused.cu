// One-pass fused GAP-TV iteration for sm_100a: Euclidean projection + all dual
// iterations of the Chambolle TV denoiser in a single sweep over HBM
// (replaces pnp_sci_algo.py:640-650 with denoiser='tv', tvm='tv_chambolle').
//
// Algorithmic traffic per outer iteration: read x, Phi (2NC) + y, y1, Phi_sum (3N),
// write x (NC) + y1 (N)  =  4N(3C+4) bytes; intermediates (f, the dual field p,
// the T-1 partial results) never touch HBM.
//
// Structure (row streaming with a one-row lag per dual iteration)
//   * a warp owns 32 consecutive pixels (lane = pixel) of one 4-channel chunk and
//     walks down the rows of its segment; the K = C/4 chunk-warps of a pixel group
//     sit in the same CTA because the projection couples the channels of a pixel;
//   * rows are staged global -> shared by the TMA unit in blocks of 4 rows (3-D tensor
//     maps over [rows][W][C]: a box is 4 rows x 32 pixels x whole pixels, i.e. contiguous
//     rows of global memory; pixels outside the image are zero-filled), four-slot ring,
//     one mbarrier per slot;
//   * phase A, once per staged block: one thread per (row, pixel) forms the dot product
//     over all C channels, updates y1 and leaves lambda*s in a shared-memory plane; a
//     split arrive/wait on a fifth mbarrier publishes it, so warps may drift by a block;
//   * step rho: stage 0 turns row rho into f = x + (lambda*s)*Phi; stage i (1..R,
//     R = tv_iter_max-1) advances dual iteration i on row rho-i using the row it
//     kept from the previous step, so out_R(rho-R) leaves the pipeline every step.
//     Horizontal neighbours come from warp shuffles (out from lane+1, p1 from
//     lane-1), vertical neighbours from the previous step's registers;
//   * R halo pixels per side of a pixel group and R warm-up rows above a row
//     segment are recomputed (projection is pointwise, so y1 stays consistent);
//     Chambolle's boundary rules apply only at true image edges;
//   * x and y1 ping-pong between two buffers (a neighbour's halo must see the old x).
//
// skimage's energy early stop needs a global reduction per dual iteration, which a
// single pass cannot wait for.  The kernel accumulates E_i per (b, c) slice on the
// side; a tiny check kernel replays the criterion and raises *flag if any slice
// would have stopped before tv_iter_max, and the solver then redoes the run on the
// exact path (tv_exact.cu).  At the reference's parameters it never fires.
#include "gap_tv_stream.cuh"

namespace scipnp {

using namespace fusedk;

namespace {

// replay skimage's stopping rule on the accumulated energies
__global__ void energy_check_kernel(double* __restrict__ energy, int nslice, int R, double eps,
                                    int* __restrict__ flag) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslice) return;
    double* e = energy + (size_t)s * R;
    double e_init = e[0], e_prev = e[0];
    bool fired = false;
    for (int i = 1; i < R; ++i) {             // a stop at i = R (the last iteration) changes nothing
        if (fabs(e_prev - e[i]) < eps * e_init) fired = true;
        e_prev = e[i];
    }
    for (int i = 0; i < R; ++i) e[i] = 0.0;   // leave the accumulators clean for the next launch
    if (fired) atomicOr(flag, 1);
}

// ---- TMA descriptors -------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

}  // namespace

// generic tiled float32 tensor map (shared with gap_tv_ws.cu)
int make_tensor_map_f32(CUtensorMap* tm, const float* base, int rank, const unsigned long long* dims,
                        const unsigned long long* strides_bytes, const unsigned* box, int l2_promotion_128, int swizzle128) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return SCIPNP_ECUDA; }
    cuuint64_t d[5], s[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), d, s, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     l2_promotion_128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(rank %d) failed with CUresult %d", rank, (int)r); return SCIPNP_ECUDA; }
    return SCIPNP_OK;
}

namespace {

// frame stack [rows][W][C] -> boxes of RB rows x 32 pixels x 4*SUBK channels (fused_subk)
int make_frame_map(CUtensorMap* tm, const float* base, long long rows, int W, int C) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return SCIPNP_ECUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)rows};
    cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4};
    cuuint32_t box[3] = {(cuuint32_t)(4 * fused_subk(C / 4)), 32, (cuuint32_t)RB};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(frame) failed with CUresult %d", (int)r); return SCIPNP_ECUDA; }
    return SCIPNP_OK;
}

// measurement plane [rows][W] -> boxes of RB rows x 32 pixels
int make_plane_map(CUtensorMap* tm, const float* base, long long rows, int W) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return SCIPNP_ECUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)W * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)RB};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(plane) failed with CUresult %d", (int)r); return SCIPNP_ECUDA; }
    return SCIPNP_OK;
}

}  // namespace

bool fused_supported(int mode, int B, int H, int W, int C, int tv_iter_max) {
    if (mode != MODE_GAP_ACC && mode != MODE_GAP_PLAIN && mode != MODE_ADMM) return false;
    if (C % 4 != 0 || C / 4 > kMaxWarps) return false;
    if (tv_iter_max < 3 || tv_iter_max > 5) return false;      // R = 2..4 are instantiated
    if (B < 1 || B > 65535 || H < 1 || W < 1) return false;
    if ((long long)H * W * C >= (1LL << 31)) return false;      // the kernel keeps 32-bit element offsets per frame
    return true;
}

bool fused_cassi_supported(int mode, int B, int H, int W, int C, int tv_iter_max) {
    return mode != MODE_ADMM && fused_supported(mode, B, H, W, C, tv_iter_max) && tv_iter_max == 5;   // R = 4 only
}

size_t fused_workspace_bytes(int B, int H, int W, int C, int tv_iter_max) {
    (void)H; (void)W;
    int R = tv_iter_max > 1 ? tv_iter_max - 1 : 1;
    // energy accumulators + 16 bytes for the last-CTA ticket of the warp-specialised kernel
    return ((size_t)B * C * R * sizeof(double) + 16 + 255) & ~(size_t)255;
}

int launch_fused(const FusedArgs& a, cudaStream_t st) {
    if (!fused_supported(a.mode, a.B, a.H, a.W, a.C, a.tv_iter_max)) {
        set_error("fused GAP-TV kernel does not cover this configuration (mode=%d C=%d tv_iter_max=%d)",
                  a.mode, a.C, a.tv_iter_max);
        return SCIPNP_EINVAL;
    }
    const bool cassi = a.mask2d != nullptr;
    if (cassi && !fused_cassi_supported(a.mode, a.B, a.H, a.W, a.C, a.tv_iter_max)) {
        set_error("fused CASSI kernel is built for tv_iter_max = 5 only");
        return SCIPNP_EINVAL;
    }
    if (!aligned16(a.x_in) || !aligned16(a.x_out) || (!cassi && !aligned16(a.Phi))) {
        set_error("fused GAP-TV kernel needs 16-byte aligned frame pointers");
        return SCIPNP_EINVAL;
    }
    if (a.workspace_bytes < fused_workspace_bytes(a.B, a.H, a.W, a.C, a.tv_iter_max) || !a.workspace) {
        set_error("fused workspace too small");
        return SCIPNP_EINVAL;
    }
    if (a.x_in == a.x_out || (a.mode == MODE_GAP_ACC && (a.y1_in == a.y1_out || !a.y1_in || !a.y1_out))) {
        set_error("fused GAP-TV kernel ping-pongs: in and out buffers must differ");
        return SCIPNP_EINVAL;
    }
    if (fused_ws_supported(a)) return launch_fused_ws(a, st);
    const int R = a.tv_iter_max - 1;
    FusedParams fp{};
    fp.x_in = a.x_in; fp.x_out = a.x_out; fp.y1_in = a.y1_in; fp.y1_out = a.y1_out;
    fp.y = a.y; fp.Phi = a.Phi; fp.Phi_sum = a.Phi_sum;
    fp.energy = reinterpret_cast<double*>(a.workspace);
    fp.lambda = a.lambda;
    fp.tv_c = (float)(0.25 / a.tv_weight);
    fp.tv_w = (float)a.tv_weight;
    fp.H = a.H; fp.W = a.W; fp.C = a.C;
    fp.K = a.C / 4;
    fp.NG = fused_groups(fp.K);
    const int own = 32 - 2 * R;
    fp.ngroups = (a.W + own - 1) / own;
    const int gx = (fp.ngroups + fp.NG - 1) / fp.NG;
    fp.B = a.B;
    fp.phi_bstride = a.phi_batched ? (long long)a.H * a.W * a.C : 0;
    fp.ps_bstride = a.phi_batched ? (long long)a.H * a.W : 0;
    fp.phi_batched = a.phi_batched ? 1 : 0;
    dim3 grid(gx, 1, a.B);

    // TMA descriptors of this launch (x ping-pongs, so they are rebuilt per call: ~1 us each on the host)
    alignas(64) FusedMaps maps;
    const long long rows = (long long)a.B * a.H, phi_rows = a.phi_batched ? rows : a.H;
    if (int e = make_frame_map(&maps.x, a.x_in, rows, a.W, a.C)) return e;
    if (cassi) maps.phi = maps.x;          // unused
    else if (int e = make_frame_map(&maps.phi, a.Phi, phi_rows, a.W, a.C)) return e;
    const CassiParams cp{a.mask2d, a.cassi_step, a.mask_w, a.b_in, a.b_out, a.xproj_out, a.gamma, a.clip01};
    if (a.mode == MODE_ADMM && (!a.b_in || !a.b_out || !a.xproj_out || a.b_in == a.b_out || !aligned16(a.b_in))) {
        set_error("fused ADMM-TV needs distinct, aligned multiplier buffers and an x output");
        return SCIPNP_EINVAL;
    }
    // plane boxes start at pixel grp*(32-2R) - R: the TMA unit needs that start 16-byte aligned,
    // which holds for R = 4 (tv_iter_max = 5, the reference's setting); otherwise cp.async
    fp.small_tma = (a.W % 4 == 0) && (R % 4 == 0) && aligned16(a.y) && aligned16(a.Phi_sum) &&
                   (a.mode != MODE_GAP_ACC || aligned16(a.y1_in));
    if (fp.small_tma) {
        if (int e = make_plane_map(&maps.y, a.y, rows, a.W)) return e;
        if (int e = make_plane_map(&maps.ps, a.Phi_sum, phi_rows, a.W)) return e;
        if (int e = make_plane_map(&maps.y1, a.mode == MODE_GAP_ACC ? a.y1_in : a.y, rows, a.W)) return e;
    } else {
        maps.y = maps.x; maps.ps = maps.x; maps.y1 = maps.x;    // unused
    }

    // the check kernel leaves the accumulators zeroed; a caller that owns the workspace and always
    // passes a flag (the solver) therefore clears it once, everybody else per launch
    if (!(a.workspace_clean && a.flag && R > 1))
        SCIPNP_CUDA(cudaMemsetAsync(fp.energy, 0, fused_workspace_bytes(a.B, a.H, a.W, a.C, a.tv_iter_max), st));
    int rc = SCIPNP_OK;
    if (cassi) rc = launch_stream_cassi_r4(a.mode, fp.K, fp, maps, cp, grid, st);
    else switch (R) {
        case 2: rc = launch_stream_r<2>(a.mode, fp.K, fp, maps, cp, grid, st); break;
        case 3: rc = launch_stream_r<3>(a.mode, fp.K, fp, maps, cp, grid, st); break;
        case 4: rc = launch_stream_r<4>(a.mode, fp.K, fp, maps, cp, grid, st); break;
        default: set_error("unsupported tv_iter_max"); return SCIPNP_EINVAL;
    }
    if (rc) return rc;
    count_launch();
    if (int e = check_launch("gap_tv_stream_kernel")) return e;
    if (a.flag && R > 1) {
        const int nslice = a.B * a.C;
        energy_check_kernel<<<(nslice + 127) / 128, 128, 0, st>>>(fp.energy, nslice, R, a.tv_eps, a.flag);
        count_launch();
        if (int e = check_launch("energy_check_kernel")) return e;
    }
    return SCIPNP_OK;
}

}  // namespace scipnp

using namespace scipnp;

extern "C" {

size_t scipnp_gap_tv_workspace_bytes(int B, int H, int W, int C, int tv_iter_max) {
    return fused_workspace_bytes(B, H, W, C, tv_iter_max);
}

int scipnp_gap_tv_fused(const float* x_in, float* x_out, const float* y1_in, float* y1_out,
                        const float* y, const float* Phi, const float* Phi_sum, float lambda,
                        int accelerate, double tv_weight, double tv_eps, int tv_iter_max, int B, int H,
                        int W, int C, int phi_batched, void* workspace, size_t workspace_bytes,
                        int* flags_dev, void* stream) {
    SCIPNP_REQUIRE(x_in && x_out && y && Phi && Phi_sum, "null pointer");
    SCIPNP_REQUIRE(tv_weight > 0.0, "tv_weight must be positive");
    FusedArgs a{};
    a.x_in = x_in; a.x_out = x_out; a.y1_in = y1_in; a.y1_out = y1_out;
    a.y = y; a.Phi = Phi; a.Phi_sum = Phi_sum;
    a.lambda = lambda; a.tv_weight = tv_weight; a.tv_eps = tv_eps; a.tv_iter_max = tv_iter_max;
    a.mode = accelerate ? MODE_GAP_ACC : MODE_GAP_PLAIN;
    a.B = B; a.H = H; a.W = W; a.C = C; a.phi_batched = phi_batched;
    a.workspace = workspace; a.workspace_bytes = workspace_bytes; a.flag = flags_dev;
    return launch_fused(a, (cudaStream_t)stream);
}

}  // extern "C"
