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
//   * rows are staged global -> shared with cp.async (16 B per request, coalesced
//     on the global side, transposed to [chunk][pixel] on the shared side so that
//     lane-per-pixel LDS.128 is conflict-free), double buffered, RB rows per block;
//   * step rho: stage 0 turns row rho into f = x + lambda*s*Phi; stage i (1..R,
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
#include "internal.cuh"

namespace scipnp {

namespace {

constexpr int RB = 4;             // rows per staged block
constexpr int PADL = 33;          // padded lane stride of the transposed tiles (float4 units)
constexpr int kMaxWarps = 8;

struct FusedParams {
    const float* x_in; float* x_out;
    const float* y1_in; float* y1_out;
    const float* y; const float* Phi; const float* Phi_sum;
    double* energy;               // [B][C][R] partial sums of d^2 + w*|g|
    float lambda, tv_c, tv_w;     // tv_c = tau / weight
    int H, W, C, K, NG, ngroups;  // K = C/4 chunk-warps per pixel group, NG groups per CTA
    int seg_rows;
    long long phi_bstride, ps_bstride;   // batch strides (0 when shared)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ float fast_sqrt(float v) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float fast_rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }

struct F4 { float v[4]; };
__device__ __forceinline__ F4 lds4(const float4* p) { float4 t = *p; F4 r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r; }

// shared-memory carve-up (per CTA), all offsets in bytes
struct Smem {
    int tile_f4_per_row;     // float4 slots of one of {x, Phi} for one row: NG*K*PADL
    int row_bytes;           // one staged row: 2 tiles + y, y1, Phi_sum lanes
    int buf_bytes;           // RB rows
    int part_off;            // partial dot products [RB][NG][32][KP]
    int KP;
    int total;
};
__host__ __device__ inline Smem smem_layout(int K, int NG) {
    Smem s;
    s.tile_f4_per_row = NG * K * PADL;
    s.row_bytes = 2 * s.tile_f4_per_row * 16 + 3 * NG * 32 * 4;
    s.buf_bytes = RB * s.row_bytes;
    s.KP = (K + 3) & ~3;
    s.part_off = 2 * s.buf_bytes;
    s.total = s.part_off + RB * NG * 32 * s.KP * 4;
    return s;
}

template <int R, int MODE, bool CHECK>
__global__ void __launch_bounds__(kMaxWarps * 32, 2)
gap_tv_stream_kernel(const FusedParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K, NG = p.NG, W = p.W, H = p.H, C = p.C;
    const Smem L = smem_layout(K, NG);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int gi = warp / K, k = warp - gi * K;
    const int b = blockIdx.z;
    constexpr int OWN = 32 - 2 * R;              // owned pixels per group
    const int group0 = blockIdx.x * NG;          // first pixel group of this CTA
    const int grp = group0 + gi;
    const bool grp_live = grp < p.ngroups;
    const int px = grp * OWN - R + lane;         // this lane's pixel column
    const bool px_in = grp_live && px >= 0 && px < W;
    const bool has_left = px > 0, has_right = px < W - 1;
    const bool own_px = px_in && lane >= R && lane < 32 - R;

    const int r0 = blockIdx.y * p.seg_rows;
    const int r1 = min(H, r0 + p.seg_rows);
    const int rs = max(0, r0 - R), rend = r1 + R;       // steps rho in [rs, rend)
    const int load_end = min(H, rend);
    const int nblk = (rend - rs + RB - 1) / RB;

    const size_t frame_b = (size_t)b * H * W * C;        // batch offsets
    const size_t meas_b = (size_t)b * H * W;
    const float* gx = p.x_in + frame_b;
    const float* gphi = p.Phi + (size_t)b * p.phi_bstride;
    const float* gy = p.y + meas_b;
    const float* gy1 = (MODE == MODE_GAP_ACC) ? p.y1_in + meas_b : nullptr;
    const float* gps = p.Phi_sum + (size_t)b * p.ps_bstride;

    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);

    // ---- producer: stage RB rows of block `blk` into buffer blk&1 -----------------------
    auto issue = [&](int blk) {
        const uint32_t buf = smem_base + (blk & 1) * L.buf_bytes;
        const int chunks = NG * 32 * K;                  // 16-byte chunks per tile row
#pragma unroll 1
        for (int j = 0; j < RB; ++j) {
            const int row = rs + blk * RB + j;
            if (row >= load_end) break;
            const uint32_t rowb = buf + j * L.row_bytes;
            for (int c = tid; c < chunks; c += nthreads) {
                const int g2 = c / (32 * K);
                const int rem = c - g2 * 32 * K;
                const int ln = rem / K, kk = rem - ln * K;
                const int gpx = (group0 + g2) * OWN - R + ln;
                const bool ok = gpx >= 0 && gpx < W && (group0 + g2) < p.ngroups;
                const size_t off = ((size_t)row * W + (ok ? gpx : 0)) * C + 4 * kk;
                const uint32_t d = rowb + ((g2 * K + kk) * PADL + ln) * 16;
                cp_async16(d, gx + off, ok ? 16 : 0);
                cp_async16(d + L.tile_f4_per_row * 16, gphi + off, ok ? 16 : 0);
            }
            const uint32_t small = rowb + 2 * L.tile_f4_per_row * 16;
            for (int c = tid; c < NG * 32; c += nthreads) {
                const int g2 = c >> 5, ln = c & 31;
                const int gpx = (group0 + g2) * OWN - R + ln;
                const bool ok = gpx >= 0 && gpx < W && (group0 + g2) < p.ngroups;
                const size_t off = (size_t)row * W + (ok ? gpx : 0);
                cp_async4(small + c * 4, gy + off, ok ? 4 : 0);
                if (MODE == MODE_GAP_ACC) cp_async4(small + (NG * 32 + c) * 4, gy1 + off, ok ? 4 : 0);
                cp_async4(small + (2 * NG * 32 + c) * 4, gps + off, ok ? 4 : 0);
            }
        }
        cp_async_commit();
    };

    // ---- pipeline state (registers) ---------------------------------------------------------
    float o_prev[R][4], g1_prev[R][4];
    float P0[R + 1][4], P1[R + 1][4];        // P[i] = p^i(rho-i-1); P[0] stays 0
    float fd[R][4];                          // fd[j] = f(rho-1-j)
    float en[R][4];                          // energy partials of iterations 0..R-1
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) { o_prev[i][c] = 0.f; g1_prev[i][c] = 0.f; fd[i][c] = 0.f; en[i][c] = 0.f; }
#pragma unroll
    for (int i = 0; i <= R; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) { P0[i][c] = 0.f; P1[i][c] = 0.f; }

    const float tau = 0.25f, tvc = p.tv_c, tvw = p.tv_w, lam = p.lambda;
    float* part = reinterpret_cast<float*>(smem_raw + L.part_off);
    float* xo = p.x_out + frame_b;
    float* y1o = (MODE == MODE_GAP_ACC) ? p.y1_out + meas_b : nullptr;

    issue(0);
    for (int blk = 0; blk < nblk; ++blk) {
        if (blk + 1 < nblk) { issue(blk + 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const unsigned char* buf = smem_raw + (blk & 1) * L.buf_bytes;

        // ---- phase A: partial dot products of this warp's chunk ------------------------------
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const int row = rs + blk * RB + j;
            if (row < load_end) {
                const float4* tx = reinterpret_cast<const float4*>(buf + j * L.row_bytes) + (gi * K + k) * PADL + lane;
                F4 xv = lds4(tx), pv = lds4(tx + L.tile_f4_per_row);
                float d = xv.v[0] * pv.v[0];
                d = fmaf(xv.v[1], pv.v[1], d);
                d = fmaf(xv.v[2], pv.v[2], d);
                d = fmaf(xv.v[3], pv.v[3], d);
                part[((j * NG + gi) * 32 + lane) * L.KP + k] = d;
                if (k == 0)
                    for (int kk = K; kk < L.KP; ++kk) part[((j * NG + gi) * 32 + lane) * L.KP + kk] = 0.f;
            }
        }
        __syncthreads();

        // ---- phase B: RB pipeline steps ----------------------------------------------------------
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const int rho = rs + blk * RB + j;
            if (rho < rend) {
                float o_new[4];
                float f_new[4] = {0.f, 0.f, 0.f, 0.f};
                if (rho < H) {
                    // stage 0: Euclidean projection of row rho
                    const unsigned char* rowp = buf + j * L.row_bytes;
                    const float4* tx = reinterpret_cast<const float4*>(rowp) + (gi * K + k) * PADL + lane;
                    F4 xv = lds4(tx), pv = lds4(tx + L.tile_f4_per_row);
                    const float4* pp = reinterpret_cast<const float4*>(part + ((j * NG + gi) * 32 + lane) * L.KP);
                    float yb = 0.f;
                    for (int q = 0; q < L.KP / 4; ++q) { float4 t = pp[q]; yb += (t.x + t.y) + (t.z + t.w); }
                    const float* sm = reinterpret_cast<const float*>(rowp + 2 * L.tile_f4_per_row * 16);
                    const float yv = sm[gi * 32 + lane];
                    const float psv = sm[2 * NG * 32 + gi * 32 + lane];
                    float s;
                    if (MODE == MODE_GAP_ACC) {
                        const float y1n = sm[NG * 32 + gi * 32 + lane] + (yv - yb);
                        if (k == 0 && own_px && rho >= r0 && rho < r1) y1o[(size_t)rho * W + px] = y1n;
                        s = __fdividef(y1n - yb, psv);
                    } else {
                        s = __fdividef(yv - yb, psv);
                    }
                    s = px_in ? s * lam : 0.f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) f_new[c] = fmaf(s, pv.v[c], xv.v[c]);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) o_new[c] = f_new[c];

                float pend0[4], pend1[4];
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const int row_new = rho - i;          // row of o_new = out_i(row_new)
                    const int u = row_new - 1;            // row whose dual variable advances
                    const bool valid_u = (u >= rs) && (u < H);
                    const bool down_ok = row_new < H;     // g0 = 0 on the last image row
                    const bool own_u = CHECK && own_px && u >= r0 && u < r1;
                    float pi0[4], pi1[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) { pi0[c] = P0[i][c]; pi1[c] = P1[i][c]; }
                    if (i > 0) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) { P0[i][c] = pend0[c]; P1[i][c] = pend1[c]; }
                    }
                    float o_next[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float o_right = __shfl_down_sync(0xffffffffu, o_new[c], 1);
                        const float g0 = down_ok ? o_new[c] - o_prev[i][c] : 0.f;
                        const float g1 = g1_prev[i][c];
                        const float nrm = fast_sqrt(fmaf(g0, g0, g1 * g1));
                        const float r = fast_rcp(fmaf(tvc, nrm, 1.f));
                        float pn0 = fmaf(-tau, g0, pi0[c]) * r;
                        float pn1 = fmaf(-tau, g1, pi1[c]) * r;
                        pn0 = valid_u ? pn0 : 0.f;
                        pn1 = valid_u ? pn1 : 0.f;
                        float p1l = __shfl_up_sync(0xffffffffu, pn1, 1);
                        p1l = has_left ? p1l : 0.f;
                        const float d = (P0[i + 1][c] - pn0) + (p1l - pn1);     // D(p^{i+1})(u)
                        o_next[c] = fd[i][c] + d;
                        if (CHECK) {
                            if (own_u) {
                                en[i][c] = fmaf(tvw, nrm, en[i][c]);                 // w*|grad out_i|(u)
                                if (i + 1 < R) en[i + 1][c] = fmaf(d, d, en[i + 1][c]);   // D(p^{i+1})(u)^2
                            }
                        }
                        g1_prev[i][c] = has_right ? o_right - o_new[c] : 0.f;
                        o_prev[i][c] = o_new[c];
                        pend0[c] = pn0;
                        pend1[c] = pn1;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) o_new[c] = o_next[c];
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) { P0[R][c] = pend0[c]; P1[R][c] = pend1[c]; }
                // f delay line
#pragma unroll
                for (int i = R - 1; i > 0; --i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) fd[i][c] = fd[i - 1][c];
#pragma unroll
                for (int c = 0; c < 4; ++c) fd[0][c] = f_new[c];
                // out_R(rho-R) leaves the pipeline
                const int orow = rho - R;
                if (own_px && orow >= r0 && orow < r1)
                    *reinterpret_cast<float4*>(xo + ((size_t)orow * W + px) * C + 4 * k) =
                        make_float4(o_new[0], o_new[1], o_new[2], o_new[3]);
            }
        }
        __syncthreads();
    }

    if (CHECK) {
        // reduce the energy partials over the pixels of the warp, one atomic per (chunk channel, i)
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v = en[i][c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && grp_live) atomicAdd(p.energy + ((size_t)b * C + 4 * k + c) * R + i, (double)v);
            }
    }
}

// replay skimage's stopping rule on the accumulated energies
__global__ void energy_check_kernel(const double* __restrict__ energy, int nslice, int R, double eps,
                                    int* __restrict__ flag) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslice) return;
    const double* e = energy + (size_t)s * R;
    double e_init = e[0], e_prev = e[0];
    for (int i = 1; i < R; ++i) {             // a stop at i = R (the last iteration) changes nothing
        if (fabs(e_prev - e[i]) < eps * e_init) { atomicOr(flag, 1); return; }
        e_prev = e[i];
    }
}

template <int R, int MODE>
int launch_stream(const FusedParams& fp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    auto kfn = gap_tv_stream_kernel<R, MODE, true>;
    SCIPNP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kfn<<<grid, threads, smem, st>>>(fp);
    return SCIPNP_OK;
}

template <int R>
int launch_stream_mode(int mode, const FusedParams& fp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    if (mode == MODE_GAP_ACC) return launch_stream<R, MODE_GAP_ACC>(fp, grid, threads, smem, st);
    return launch_stream<R, MODE_GAP_PLAIN>(fp, grid, threads, smem, st);
}

}  // namespace

bool fused_supported(int mode, int B, int H, int W, int C, int tv_iter_max) {
    if (mode != MODE_GAP_ACC && mode != MODE_GAP_PLAIN) return false;   // ADMM: exact path
    if (C % 4 != 0 || C / 4 > kMaxWarps) return false;
    if (tv_iter_max < 2 || tv_iter_max > 6) return false;
    if (B < 1 || B > 65535 || H < 1 || W < 1) return false;
    return true;
}

size_t fused_workspace_bytes(int B, int H, int W, int C, int tv_iter_max) {
    (void)H; (void)W;
    int R = tv_iter_max > 1 ? tv_iter_max - 1 : 1;
    return ((size_t)B * C * R * sizeof(double) + 255) & ~(size_t)255;
}

int launch_fused(const FusedArgs& a, cudaStream_t st) {
    if (!fused_supported(a.mode, a.B, a.H, a.W, a.C, a.tv_iter_max)) {
        set_error("fused GAP-TV kernel does not cover this configuration (mode=%d C=%d tv_iter_max=%d)",
                  a.mode, a.C, a.tv_iter_max);
        return SCIPNP_EINVAL;
    }
    if (!aligned16(a.x_in) || !aligned16(a.x_out) || !aligned16(a.Phi)) {
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
    fp.NG = kMaxWarps / fp.K;
    const int own = 32 - 2 * R;
    fp.ngroups = (a.W + own - 1) / own;
    fp.NG = fp.NG < 1 ? 1 : (fp.NG > fp.ngroups ? fp.ngroups : fp.NG);
    const int gx = (fp.ngroups + fp.NG - 1) / fp.NG;
    // row segments: enough CTAs for ~4 waves of 2 CTAs/SM, but segments of >= 32 rows
    const long long target = 8LL * num_sms();
    long long nseg = (target + (long long)gx * a.B - 1) / ((long long)gx * a.B);
    long long max_seg = (a.H + 31) / 32;
    if (nseg > max_seg) nseg = max_seg;
    if (nseg < 1) nseg = 1;
    fp.seg_rows = (int)((a.H + nseg - 1) / nseg);
    nseg = (a.H + fp.seg_rows - 1) / fp.seg_rows;
    fp.phi_bstride = a.phi_batched ? (long long)a.H * a.W * a.C : 0;
    fp.ps_bstride = a.phi_batched ? (long long)a.H * a.W : 0;
    if (nseg > 65535) { set_error("too many row segments"); return SCIPNP_EINVAL; }
    dim3 grid(gx, (unsigned)nseg, a.B);
    const int threads = fp.NG * fp.K * 32;
    const Smem L = smem_layout(fp.K, fp.NG);

    SCIPNP_CUDA(cudaMemsetAsync(fp.energy, 0, (size_t)a.B * a.C * R * sizeof(double), st));
    int rc = SCIPNP_OK;
    switch (R) {
        case 1: rc = launch_stream_mode<1>(a.mode, fp, grid, threads, L.total, st); break;
        case 2: rc = launch_stream_mode<2>(a.mode, fp, grid, threads, L.total, st); break;
        case 3: rc = launch_stream_mode<3>(a.mode, fp, grid, threads, L.total, st); break;
        case 4: rc = launch_stream_mode<4>(a.mode, fp, grid, threads, L.total, st); break;
        case 5: rc = launch_stream_mode<5>(a.mode, fp, grid, threads, L.total, st); break;
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
