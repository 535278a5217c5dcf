// Exact-path Chambolle TV denoiser (R6): one dual iteration per pair of launches,
// dual field in HBM, energy-based early stop evaluated on the device per (b, c)
// slice exactly as skimage's _denoise_tv_chambolle_nd does per channel.
//
// IEEE single precision in the statement order of the published NumPy source
// (no FMA contraction; correctly rounded sqrt and divide), so `out` is
// reproducible bit for bit against oracle/tv_chambolle.py.  The energies are
// accumulated in double (NumPy: float32 pairwise), which can only matter when
// |E_prev - E| sits within rounding of eps*E_init.
//
// The bandwidth-optimised path (all dual iterations of one outer iteration in a
// single pass, fused with the projection) is gap_tv_fused.cu; this file is its
// fallback when the early stop fires and the general-shape implementation
// (any C, any n_iter_max).
#include "internal.cuh"

namespace scipnp {

struct TvSlice {          // one per (b, c)
    double acc_d;         // sum D(p)^2 of the current iteration
    double acc_n;         // sum |g|     of the current iteration
    double e_init, e_prev;
    int done;             // early stop fired
    int n_exec;           // iterations executed so far
};

constexpr int kTvThreads = 128;
constexpr int kTvRows = 8;

// layout of a frame stack: [B][H][WC] floats, WC = W*C, channel of flat col j is j % C
template <int V>
__device__ __forceinline__ void load_vec(const float* p, float* o) {
    if (V == 4) { float4 v = *reinterpret_cast<const float4*>(p); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
    else o[0] = *p;
}
template <int V>
__device__ __forceinline__ void store_vec(float* p, const float* o) {
    if (V == 4) *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    else *p = o[0];
}

// out_i = f + D(p^i)   (i >= 1); accumulates sum D^2 per slice
template <int V>
__global__ void __launch_bounds__(kTvThreads)
tv_out_kernel(const float* __restrict__ f, const float* __restrict__ p0, const float* __restrict__ p1,
              float* __restrict__ out, TvSlice* __restrict__ sl, int H, int W, int C, int e_lo, int e_hi) {
    extern __shared__ double sacc[];   // [C]
    const int WC = W * C, nvec = WC / V;
    const int b = blockIdx.z;
    for (int c = threadIdx.x; c < C; c += kTvThreads) sacc[c] = 0.0;
    __syncthreads();
    const int jv = blockIdx.x * kTvThreads + threadIdx.x;
    if (jv < nvec) {
        const int j = jv * V;
        const int c0 = j % C;
        bool active[V];
        bool any = false;
#pragma unroll
        for (int i = 0; i < V; ++i) { active[i] = !sl[b * C + c0 + i].done; any |= active[i]; }
        if (any) {
            double acc[V];
#pragma unroll
            for (int i = 0; i < V; ++i) acc[i] = 0.0;
            const int r0 = blockIdx.y * kTvRows, r1 = min(H, r0 + kTvRows);
            const bool has_left = j >= C;
            for (int r = r0; r < r1; ++r) {
                size_t o = ((size_t)b * H + r) * WC + j;
                float a0[V], a1[V], up[V], lf[V], fv[V], ov[V];
                load_vec<V>(p0 + o, a0);
                load_vec<V>(p1 + o, a1);
                load_vec<V>(f + o, fv);
                if (r > 0) load_vec<V>(p0 + o - WC, up);
                if (has_left) load_vec<V>(p1 + o - C, lf);
                if (V == 4 && !(active[0] && active[1] && active[2] && active[3])) load_vec<V>(out + o, ov);
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    float d = -__fadd_rn(a0[i], a1[i]);
                    if (r > 0) d = __fadd_rn(d, up[i]);
                    if (has_left) d = __fadd_rn(d, lf[i]);
                    if (active[i]) {
                        ov[i] = __fadd_rn(fv[i], d);
                        if (r >= e_lo && r < e_hi) acc[i] += (double)__fmul_rn(d, d);
                    }
                }
                store_vec<V>(out + o, ov);
            }
#pragma unroll
            for (int i = 0; i < V; ++i) if (active[i]) atomicAdd(&sacc[c0 + i], acc[i]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kTvThreads)
        if (sacc[c] != 0.0) atomicAdd(&sl[b * C + c].acc_d, sacc[c]);
}

// p^{i+1} = (p^i - tau*g) / (1 + (tau/w)|g|),  g = forward differences of out_i
template <int V>
__global__ void __launch_bounds__(kTvThreads)
tv_dual_kernel(const float* __restrict__ outi, float* __restrict__ p0, float* __restrict__ p1,
               TvSlice* __restrict__ sl, int H, int W, int C, float tau, float tau_over_w, int e_lo, int e_hi) {
    extern __shared__ double sacc[];
    const int WC = W * C, nvec = WC / V;
    const int b = blockIdx.z;
    for (int c = threadIdx.x; c < C; c += kTvThreads) sacc[c] = 0.0;
    __syncthreads();
    const int jv = blockIdx.x * kTvThreads + threadIdx.x;
    if (jv < nvec) {
        const int j = jv * V;
        const int c0 = j % C;
        bool active[V];
        bool any = false;
#pragma unroll
        for (int i = 0; i < V; ++i) { active[i] = !sl[b * C + c0 + i].done; any |= active[i]; }
        if (any) {
            double acc[V];
#pragma unroll
            for (int i = 0; i < V; ++i) acc[i] = 0.0;
            const int r0 = blockIdx.y * kTvRows, r1 = min(H, r0 + kTvRows);
            const bool has_right = j + C < WC;
            float cur[V], nxt[V];
            load_vec<V>(outi + ((size_t)b * H + r0) * WC + j, cur);
            for (int r = r0; r < r1; ++r) {
                size_t o = ((size_t)b * H + r) * WC + j;
                float rt[V], a0[V], a1[V];
                const bool has_down = r + 1 < H;
                if (has_down) load_vec<V>(outi + o + WC, nxt);
                if (has_right) load_vec<V>(outi + o + C, rt);
                load_vec<V>(p0 + o, a0);
                load_vec<V>(p1 + o, a1);
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    float g0 = has_down ? __fsub_rn(nxt[i], cur[i]) : 0.f;
                    float g1 = has_right ? __fsub_rn(rt[i], cur[i]) : 0.f;
                    float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(g0, g0), __fmul_rn(g1, g1)));
                    if (active[i]) {
                        if (r >= e_lo && r < e_hi) acc[i] += (double)nrm;
                        float den = __fadd_rn(__fmul_rn(nrm, tau_over_w), 1.f);
                        a0[i] = __fdiv_rn(__fsub_rn(a0[i], __fmul_rn(tau, g0)), den);
                        a1[i] = __fdiv_rn(__fsub_rn(a1[i], __fmul_rn(tau, g1)), den);
                    }
                    cur[i] = nxt[i];
                }
                store_vec<V>(p0 + o, a0);
                store_vec<V>(p1 + o, a1);
            }
#pragma unroll
            for (int i = 0; i < V; ++i) if (active[i]) atomicAdd(&sacc[c0 + i], acc[i]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kTvThreads)
        if (sacc[c] != 0.0) atomicAdd(&sl[b * C + c].acc_n, sacc[c]);
}

// per-slice energy bookkeeping after iteration i
__global__ void tv_check_kernel(TvSlice* __restrict__ sl, int nslice, int i, double weight,
                                double eps, double size, double* __restrict__ energy, int energy_cap) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslice) return;
    TvSlice t = sl[s];
    if (t.done) return;
    double E = (t.acc_d + weight * t.acc_n) / size;
    if (energy && i < energy_cap) energy[(size_t)s * energy_cap + i] = E;
    t.n_exec = i + 1;
    if (i == 0) { t.e_init = E; t.e_prev = E; }
    else if (fabs(t.e_prev - E) < eps * t.e_init) t.done = 1;
    else t.e_prev = E;
    t.acc_d = 0.0;
    t.acc_n = 0.0;
    sl[s] = t;
}

// gather / scatter of the per-slice partial energies around the cross-rank sum (row-tiled scenes)
__global__ void tv_acc_gather_kernel(const TvSlice* __restrict__ sl, int nslice, double* __restrict__ buf) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nslice) { buf[2 * s] = sl[s].acc_d; buf[2 * s + 1] = sl[s].acc_n; }
}
__global__ void tv_acc_scatter_kernel(TvSlice* __restrict__ sl, int nslice, const double* __restrict__ buf) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nslice) { sl[s].acc_d = buf[2 * s]; sl[s].acc_n = buf[2 * s + 1]; }
}

__global__ void tv_finish_kernel(const TvSlice* __restrict__ sl, int nslice, int* __restrict__ n_exec,
                                 int* __restrict__ flag, int n_iter_max) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslice) return;
    if (n_exec) n_exec[s] = sl[s].n_exec;
    if (flag && sl[s].n_exec < n_iter_max) atomicOr(flag, 1);
}

__global__ void fill_nan_kernel(double* p, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = __longlong_as_double(0x7ff8000000000000LL);
}

size_t tv_workspace_bytes(int B, int H, int W, int C) {
    size_t field = (size_t)B * H * W * C * sizeof(float);
    field = (field + 255) & ~(size_t)255;
    size_t slices = ((size_t)B * C * sizeof(TvSlice) + 255) & ~(size_t)255;
    size_t accbuf = ((size_t)B * C * 2 * sizeof(double) + 255) & ~(size_t)255;    // staging of the cross-rank energy sum
    return 2 * field + slices + accbuf;
}

// Runs the whole denoiser on `st`.  full_energy: also evaluate the (dead) dual
// update of the last iteration so that E and n_exec of that iteration exist.
int tv_chambolle_exact(const float* in, float* out, double weight, double eps, int T, int B, int H,
                       int W, int C, void* workspace, size_t ws_bytes, int* n_exec_dev,
                       double* energy_dev, int energy_cap, cudaStream_t st, const TvTiling* tiling) {
    if (ws_bytes < tv_workspace_bytes(B, H, W, C)) {
        set_error("tv workspace too small: %zu < %zu", ws_bytes, tv_workspace_bytes(B, H, W, C));
        return SCIPNP_EINVAL;
    }
    const size_t nelem = (size_t)B * H * W * C;
    size_t field = (nelem * sizeof(float) + 255) & ~(size_t)255;
    float* p0 = reinterpret_cast<float*>(workspace);
    float* p1 = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + field);
    TvSlice* sl = reinterpret_cast<TvSlice*>(reinterpret_cast<char*>(workspace) + 2 * field);
    const int nslice = B * C;
    double* accbuf = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 2 * field +
                                               (((size_t)nslice * sizeof(TvSlice) + 255) & ~(size_t)255));
    const int e_lo = (tiling && tiling->e_hi > 0) ? tiling->e_lo : 0;
    const int e_hi = (tiling && tiling->e_hi > 0) ? tiling->e_hi : H;
    const double size = (double)((tiling && tiling->total_rows > 0) ? tiling->total_rows : H) * W;
    SCIPNP_CUDA(cudaMemsetAsync(workspace, 0, 2 * field + (size_t)nslice * sizeof(TvSlice), st));
    if (energy_dev && energy_cap > 0) {
        size_t n = (size_t)nslice * energy_cap;
        fill_nan_kernel<<<(unsigned)ceil_div_ll((long long)n, 256), 256, 0, st>>>(energy_dev, n);
        count_launch();
    }
    const bool want_stats = (n_exec_dev != nullptr) || (energy_dev != nullptr);
    const bool v4 = (C % 4 == 0) && aligned16(in) && aligned16(out) && ((size_t)W * C % 4 == 0);
    const int V = v4 ? 4 : 1;
    const int WC = W * C;
    dim3 grid((unsigned)ceil_div_ll(WC / V, kTvThreads), (unsigned)ceil_div_ll(H, kTvRows), B);
    const size_t smem = (size_t)C * sizeof(double);
    const float tau = 0.25f;
    const float tow = (float)(0.25 / weight);
    const unsigned cgrid = (unsigned)ceil_div_ll(nslice, 128);
    // sum the partial energies of this dual iteration over the ranks (row-tiled scenes)
    auto reduce_acc = [&]() -> int {
        if (!tiling || !tiling->reduce) return SCIPNP_OK;
        tv_acc_gather_kernel<<<cgrid, 128, 0, st>>>(sl, nslice, accbuf);
        if (int rc = tiling->reduce(accbuf, 2 * nslice, (void*)st, tiling->user)) {
            set_error("energy reduction callback failed with %d", rc);
            return SCIPNP_ESTATE;
        }
        tv_acc_scatter_kernel<<<cgrid, 128, 0, st>>>(sl, nslice, accbuf);
        count_launch(2);
        return SCIPNP_OK;
    };
    if (T <= 1) {
        SCIPNP_CUDA(cudaMemcpyAsync(out, in, nelem * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (T == 1 && want_stats) {   // E_0 needs |grad f|
            if (V == 4) tv_dual_kernel<4><<<grid, kTvThreads, smem, st>>>(in, p0, p1, sl, H, W, C, tau, tow, e_lo, e_hi);
            else tv_dual_kernel<1><<<grid, kTvThreads, smem, st>>>(in, p0, p1, sl, H, W, C, tau, tow, e_lo, e_hi);
            if (int e = reduce_acc()) return e;
            tv_check_kernel<<<cgrid, 128, 0, st>>>(sl, nslice, 0, weight, eps, size, energy_dev, energy_cap);
            count_launch(2);
        }
    }
    for (int i = 0; i < T && T > 1; ++i) {
        const float* src = in;
        if (i > 0) {
            if (V == 4) tv_out_kernel<4><<<grid, kTvThreads, smem, st>>>(in, p0, p1, out, sl, H, W, C, e_lo, e_hi);
            else tv_out_kernel<1><<<grid, kTvThreads, smem, st>>>(in, p0, p1, out, sl, H, W, C, e_lo, e_hi);
            count_launch();
            src = out;
        }
        if (i == T - 1 && !want_stats) break;      // last dual update is dead code
        if (V == 4) tv_dual_kernel<4><<<grid, kTvThreads, smem, st>>>(src, p0, p1, sl, H, W, C, tau, tow, e_lo, e_hi);
        else tv_dual_kernel<1><<<grid, kTvThreads, smem, st>>>(src, p0, p1, sl, H, W, C, tau, tow, e_lo, e_hi);
        if (int e = reduce_acc()) return e;
        tv_check_kernel<<<cgrid, 128, 0, st>>>(sl, nslice, i, weight, eps, size, energy_dev, energy_cap);
        count_launch(2);
    }
    if (n_exec_dev) {
        tv_finish_kernel<<<cgrid, 128, 0, st>>>(sl, nslice, n_exec_dev, nullptr, T);
        count_launch();
    }
    return check_launch("tv_chambolle_exact");
}

}  // namespace scipnp

using namespace scipnp;

extern "C" {

size_t scipnp_tv_workspace_bytes(int B, int H, int W, int C) {
    if (B < 1 || H < 1 || W < 1 || C < 1) return 0;
    return tv_workspace_bytes(B, H, W, C);
}

int scipnp_tv_chambolle(const float* in, float* out, double weight, double eps, int n_iter_max, int B,
                        int H, int W, int C, void* workspace, size_t workspace_bytes, int* n_exec_dev,
                        double* energy_dev, int energy_cap, void* stream) {
    SCIPNP_REQUIRE(B >= 1 && H >= 1 && W >= 1 && C >= 1 && B <= 65535, "bad dimensions");
    SCIPNP_REQUIRE(in && out && workspace, "null pointer");
    SCIPNP_REQUIRE(in != out, "in and out must not alias");
    SCIPNP_REQUIRE(weight > 0.0, "weight must be positive");
    SCIPNP_REQUIRE(n_iter_max >= 0, "negative n_iter_max");
    return tv_chambolle_exact(in, out, weight, eps, n_iter_max, B, H, W, C, workspace, workspace_bytes,
                              n_exec_dev, energy_dev, energy_cap, (cudaStream_t)stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// MATLAB twin's default TV: anisotropic TV by iterative clipping, per 2-D frame,
// PnP_SCI/matlab/algorithms/tvdenoisers/TV_denoising.m:1-44 (the 'ATV_ClipA' branch of
// gapdenoise.m:93-94).  alpha = 5; per iteration
//     x0 = ((y0 - dht(zh)) + (y0 - dvt(zv))) / 2                      (:26-28)
//     zh = clip(zh + (1/alpha) dh(x0), lambda/2),  zv likewise         (:29-30)
// with dh/dv forward differences and dht/dvt their transposes (:51-64); the x0 of the last
// iteration is returned, so the last z update is dead.  IEEE single precision in the statement
// order of the .m file (a float32 NumPy restatement is reproduced bit for bit).
// ---------------------------------------------------------------------------------------------
namespace scipnp {
namespace {

// zh, zv: [B][H][W][C]; column W-1 of zh and row H-1 of zv are never read
__global__ void atv_x_kernel(const float* __restrict__ y0, const float* __restrict__ zh, const float* __restrict__ zv,
                             float* __restrict__ x0, int H, int W, int C, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t pix = i / C;
    const int w = (int)(pix % W), h = (int)((pix / W) % H);
    const size_t sw = (size_t)C, sh = (size_t)W * C;
    float dht, dvt;                                    // TV_denoising.m:57-58, :53-54
    if (w == 0) dht = -zh[i];
    else if (w == W - 1) dht = zh[i - sw];
    else dht = -__fsub_rn(zh[i], zh[i - sw]);
    if (h == 0) dvt = -zv[i];
    else if (h == H - 1) dvt = zv[i - sh];
    else dvt = -__fsub_rn(zv[i], zv[i - sh]);
    const float v = y0[i];
    x0[i] = __fmul_rn(__fadd_rn(__fsub_rn(v, dht), __fsub_rn(v, dvt)), 0.5f);
}

__device__ __forceinline__ float atv_clip(float v, float t) {      // :67-68  sign(x).*min(abs(x), t)
    const float m = fminf(fabsf(v), t);
    return v > 0.f ? m : (v < 0.f ? -m : 0.f * m);
}

__global__ void atv_z_kernel(const float* __restrict__ x0, float* __restrict__ zh, float* __restrict__ zv,
                             float inv_alpha, float half_lambda, int H, int W, int C, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t pix = i / C;
    const int w = (int)(pix % W), h = (int)((pix / W) % H);
    const size_t sw = (size_t)C, sh = (size_t)W * C;
    const float v = x0[i];
    if (w < W - 1) zh[i] = atv_clip(__fadd_rn(zh[i], __fmul_rn(inv_alpha, __fsub_rn(x0[i + sw], v))), half_lambda);
    if (h < H - 1) zv[i] = atv_clip(__fadd_rn(zv[i], __fmul_rn(inv_alpha, __fsub_rn(x0[i + sh], v))), half_lambda);
}

}  // namespace
}  // namespace scipnp

extern "C" {

size_t scipnp_tv_atv_clip_workspace_bytes(int B, int H, int W, int C) {
    if (B < 1 || H < 1 || W < 1 || C < 1) return 0;
    return 2 * (size_t)B * H * W * C * sizeof(float);
}

int scipnp_tv_atv_clip(const float* in, float* out, float lambda, int iters, int B, int H, int W, int C,
                       void* workspace, size_t workspace_bytes, void* stream) {
    using namespace scipnp;
    SCIPNP_REQUIRE(B >= 1 && H >= 2 && W >= 2 && C >= 1, "bad dimensions (frames must be at least 2x2)");
    SCIPNP_REQUIRE(in && out && workspace, "null pointer");
    SCIPNP_REQUIRE(in != out, "in and out must not alias");
    SCIPNP_REQUIRE(iters >= 1, "iters must be >= 1");
    const size_t n = (size_t)B * H * W * C;
    if (workspace_bytes < 2 * n * sizeof(float)) { set_error("ATV workspace too small"); return SCIPNP_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    float* zh = reinterpret_cast<float*>(workspace);
    float* zv = zh + n;
    SCIPNP_CUDA(cudaMemsetAsync(zh, 0, 2 * n * sizeof(float), st));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    const float inv_alpha = (float)(1.0 / 5.0), half = (float)((double)lambda / 2.0);
    for (int it = 0; it < iters; ++it) {
        atv_x_kernel<<<blocks, 256, 0, st>>>(in, zh, zv, out, H, W, C, n);
        count_launch();
        if (it + 1 < iters) {
            atv_z_kernel<<<blocks, 256, 0, st>>>(out, zh, zv, inv_alpha, half, H, W, C, n);
            count_launch();
        }
    }
    return check_launch("atv kernels");
}

}  // extern "C"
