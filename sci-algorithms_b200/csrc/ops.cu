// Exact-path kernels of libscipnp: sensing operators, Euclidean projections,
// PSNR reduction, Bayer / CASSI re-indexing.
//
// "Exact" means IEEE round-to-nearest single precision with the statement order
// of the reference's NumPy code (no FMA contraction, correctly rounded divide),
// including NumPy's pairwise summation order for the Cr-axis sum of utils.A_
// (utils.py:15), so results are reproducible bit for bit against the reference.
// The bandwidth-optimised one-pass kernel lives in gap_tv_fused.cu.
#include "internal.cuh"

namespace scipnp {

// ---------------------------------------------------------------------------
// NumPy's float32 add-reduce order over one contiguous run of n values
// (numpy/core/src/umath/loops_utils.h pairwise sum: 8 running sums combined as a
// balanced tree for 8 <= n <= 128, plain left-to-right below 8, halving above).
// `a` is indexed through stride `st` so callers can keep a transposed tile.
// ---------------------------------------------------------------------------
__device__ float numpy_sum(const float* a, int n, int st) {
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i * st]);
        return r;
    }
    if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = a[j * st];
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[(i + j) * st]);
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i * st]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(numpy_sum(a, n2, st), numpy_sum(a + (size_t)n2 * st, n - n2, st));
}

// ---------------------------------------------------------------------------
// R1 / R3: one thread per pixel (operators outside the solver loop)
// ---------------------------------------------------------------------------
constexpr int kMaxLocalC = 128;

__global__ void A_kernel(const float* __restrict__ x, const float* __restrict__ Phi,
                         float* __restrict__ y, long long npix, int C, long long phi_stride) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int b = blockIdx.y;
    if (p >= npix) return;
    const float* xp = x + ((size_t)b * npix + p) * C;
    const float* pp = Phi + (size_t)b * phi_stride + (size_t)p * C;
    float prod[kMaxLocalC];
    float acc = 0.f;
    for (int c0 = 0; c0 < C; c0 += kMaxLocalC) {      // C <= 128: single exact pass
        int n = min(kMaxLocalC, C - c0);
        for (int c = 0; c < n; ++c) prod[c] = __fmul_rn(xp[c0 + c], pp[c0 + c]);
        float s = numpy_sum(prod, n, 1);
        acc = (c0 == 0) ? s : __fadd_rn(acc, s);
    }
    y[(size_t)b * npix + p] = acc;
}

__global__ void phi_sum_kernel(const float* __restrict__ Phi, float* __restrict__ out,
                               long long npix, int C) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int b = blockIdx.y;
    if (p >= npix) return;
    const float* pp = Phi + ((size_t)b * npix + p) * C;
    const bool vec = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(Phi) & 15u) == 0;
    float v[kMaxLocalC];
    float acc = 0.f;
    for (int c0 = 0; c0 < C; c0 += kMaxLocalC) {
        int n = min(kMaxLocalC, C - c0);
        if (vec) {
            const float4* p4 = reinterpret_cast<const float4*>(pp + c0);
            for (int c = 0; c < n / 4; ++c) {
                const float4 t = p4[c];
                v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
            }
        } else {
            for (int c = 0; c < n; ++c) v[c] = pp[c0 + c];
        }
        float s = numpy_sum(v, n, 1);
        acc = (c0 == 0) ? s : __fadd_rn(acc, s);
    }
    out[(size_t)b * npix + p] = (acc == 0.f) ? 1.f : acc;
}

__global__ void At_kernel(const float* __restrict__ y, const float* __restrict__ Phi,
                          float* __restrict__ x, long long nelem, int C, long long phi_stride) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int b = blockIdx.y;
    if (e >= nelem) return;
    long long p = e / C;
    x[(size_t)b * nelem + e] = __fmul_rn(y[(size_t)b * (nelem / C) + p], Phi[(size_t)b * phi_stride + e]);
}

// C % 4 == 0 and 16-byte aligned stacks: four channels of one pixel per thread
__global__ void At_kernel_v4(const float* __restrict__ y, const float4* __restrict__ Phi,
                             float4* __restrict__ x, long long nelem4, int C4, long long phi_stride4) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int b = blockIdx.y;
    if (e >= nelem4) return;
    const float yv = y[(size_t)b * (nelem4 / C4) + e / C4];
    const float4 m = Phi[(size_t)b * phi_stride4 + e];
    x[(size_t)b * nelem4 + e] = make_float4(__fmul_rn(yv, m.x), __fmul_rn(yv, m.y), __fmul_rn(yv, m.z), __fmul_rn(yv, m.w));
}

// Start of a reconstruction in one pass over the mask stack: Phi_sum (mask_sum, pnp_sci_algo.py:491-492, NumPy's
// summation order) and the default initial guess x0 = At(y) (pnp_sci_algo.py:621-622).  A CTA owns kInitP consecutive
// pixels: every thread first issues all of its coalesced 128-bit loads of Phi (up to kInitL in flight), writes x0
// from the registers and parks the masks in shared memory (row stride C+1: conflict-free for the reader), then one
// thread per pixel sums its C values.
constexpr int kInitP = 256, kInitThreads = 256, kInitL = 8;
__global__ void __launch_bounds__(kInitThreads)
init_x0_phisum_kernel(const float* __restrict__ y, const float4* __restrict__ Phi, float4* __restrict__ x,
                      float* __restrict__ phisum, long long npix, int C4, long long phi_stride4, int phisum_batched) {
    extern __shared__ float sm_init[];
    const int b = blockIdx.y, C = 4 * C4, ld = C + 1;
    const long long p0 = (long long)blockIdx.x * kInitP;
    const int n = (int)((npix - p0) < kInitP ? (npix - p0) : kInitP);
    const float4* src = Phi + (size_t)b * phi_stride4 + (size_t)p0 * C4;
    float4* dst = x + ((size_t)b * npix + p0) * C4;
    const float* yb = y + (size_t)b * npix + p0;
    const int total = n * C4;
    for (int base = 0; base < total; base += kInitThreads * kInitL) {
        float4 m[kInitL];
#pragma unroll
        for (int l = 0; l < kInitL; ++l) {
            const int i = base + l * kInitThreads + threadIdx.x;
            if (i < total) m[l] = src[i];
        }
#pragma unroll
        for (int l = 0; l < kInitL; ++l) {
            const int i = base + l * kInitThreads + threadIdx.x;
            if (i < total) {
                const int px = i / C4, k = i - px * C4;
                const float yv = yb[px];
                dst[i] = make_float4(__fmul_rn(yv, m[l].x), __fmul_rn(yv, m[l].y), __fmul_rn(yv, m[l].z), __fmul_rn(yv, m[l].w));
                float* r = sm_init + px * ld + 4 * k;
                r[0] = m[l].x; r[1] = m[l].y; r[2] = m[l].z; r[3] = m[l].w;
            }
        }
    }
    if (!phisum_batched && b != 0) return;            // one Phi_sum for all measurements: the first slice writes it
    __syncthreads();
    if ((int)threadIdx.x < n) {
        const float acc = numpy_sum(sm_init + threadIdx.x * ld, C, 1);
        phisum[(size_t)(phisum_batched ? b : 0) * npix + p0 + threadIdx.x] = (acc == 0.f) ? 1.f : acc;
    }
}

// ---------------------------------------------------------------------------
// R4 / R5 projection, coalesced: a CTA owns P consecutive pixels; products go
// through shared memory (transposed, [C][P]) so that one thread per pixel can
// sum them in NumPy's order, then every thread applies the per-pixel step.
// ---------------------------------------------------------------------------

constexpr int kProjThreads = 256;
constexpr int kProjK = 4;   // vector chunks per thread

template <int V> struct Vec;
template <> struct Vec<4> { using T = float4; };
template <> struct Vec<1> { using T = float; };

template <int V> __device__ __forceinline__ void unpack(const typename Vec<V>::T& v, float* o);
template <> __device__ __forceinline__ void unpack<4>(const float4& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
template <> __device__ __forceinline__ void unpack<1>(const float& v, float* o) { o[0] = v; }
template <int V> __device__ __forceinline__ typename Vec<V>::T pack(const float* o);
template <> __device__ __forceinline__ float4 pack<4>(const float* o) { return make_float4(o[0], o[1], o[2], o[3]); }
template <> __device__ __forceinline__ float pack<1>(const float* o) { return o[0]; }

template <int V, int MODE>
__global__ void __launch_bounds__(kProjThreads)
project_kernel(const float* __restrict__ a_in,    // GAP: x_in          ADMM: theta
               const float* __restrict__ b_in,    // GAP: unused        ADMM: b
               float* __restrict__ x_out,         // GAP: x_out         ADMM: x
               float* __restrict__ f_out,         // GAP: unused        ADMM: f = x-b
               const float* __restrict__ y1_in, float* __restrict__ y1_out,
               const float* __restrict__ y, const float* __restrict__ Phi,
               const float* __restrict__ Phi_sum, float lambda, float gamma,
               long long npix, int C, int P, long long phi_stride, long long phisum_stride) {
    using VT = typename Vec<V>::T;
    extern __shared__ float smem[];
    const int Ppad = P | 1;
    float* prod = smem;                // [C][Ppad]
    float* sval = smem + (size_t)C * Ppad;   // [P]

    const int b = blockIdx.y;
    const long long pix0 = (long long)blockIdx.x * P;
    const int np = (int)min((long long)P, npix - pix0);
    const int cpp = C / V;
    const int nchunk = np * cpp;
    const size_t base = ((size_t)b * npix + pix0) * C;
    const size_t pbase = (size_t)b * phi_stride + (size_t)pix0 * C;

    float u[kProjK][V], ph[kProjK][V], bb[kProjK][V];
#pragma unroll
    for (int k = 0; k < kProjK; ++k) {
        int j = threadIdx.x + k * kProjThreads;
        if (j < nchunk) {
            VT av = *reinterpret_cast<const VT*>(a_in + base + (size_t)j * V);
            VT pv = *reinterpret_cast<const VT*>(Phi + pbase + (size_t)j * V);
            unpack<V>(av, u[k]);
            unpack<V>(pv, ph[k]);
            if (MODE == MODE_ADMM) {
                VT bv = *reinterpret_cast<const VT*>(b_in + base + (size_t)j * V);
                unpack<V>(bv, bb[k]);
#pragma unroll
                for (int i = 0; i < V; ++i) u[k][i] = __fadd_rn(u[k][i], bb[k][i]);   // theta + b
            }
            int p = j / cpp, c0 = (j - p * cpp) * V;
#pragma unroll
            for (int i = 0; i < V; ++i) prod[(size_t)(c0 + i) * Ppad + p] = __fmul_rn(u[k][i], ph[k][i]);
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < np; p += kProjThreads) {
        float yb = numpy_sum(prod + p, C, Ppad);
        size_t gp = (size_t)b * npix + pix0 + p;
        float ps = Phi_sum[(size_t)b * phisum_stride + pix0 + p];
        float s;
        if (MODE == MODE_GAP_ACC) {
            float y1n = __fadd_rn(y1_in[gp], __fsub_rn(y[gp], yb));
            y1_out[gp] = y1n;
            s = __fdiv_rn(__fsub_rn(y1n, yb), ps);
        } else if (MODE == MODE_GAP_PLAIN) {
            s = __fdiv_rn(__fsub_rn(y[gp], yb), ps);
        } else {
            s = __fdiv_rn(__fsub_rn(y[gp], yb), __fadd_rn(ps, gamma));
        }
        sval[p] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kProjK; ++k) {
        int j = threadIdx.x + k * kProjThreads;
        if (j < nchunk) {
            int p = j / cpp;
            float s = sval[p];
            float xo[V], fo[V];
#pragma unroll
            for (int i = 0; i < V; ++i) {
                float t = __fmul_rn(lambda, __fmul_rn(s, ph[k][i]));
                xo[i] = __fadd_rn(u[k][i], t);
                if (MODE == MODE_ADMM) fo[i] = __fsub_rn(xo[i], bb[k][i]);
            }
            *reinterpret_cast<VT*>(x_out + base + (size_t)j * V) = pack<V>(xo);
            if (MODE == MODE_ADMM) *reinterpret_cast<VT*>(f_out + base + (size_t)j * V) = pack<V>(fo);
        }
    }
}

__global__ void admm_dual_kernel(float* __restrict__ b, const float* __restrict__ x,
                                 const float* __restrict__ theta, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) b[i] = __fsub_rn(b[i], __fsub_rn(x[i], theta[i]));
}

// joint_pnp_sci_algo.py:633  theta = np.clip(theta, 0, 1)
__global__ void clip01_kernel(float* __restrict__ x, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) x[i] = fminf(fmaxf(x[i], 0.f), 1.f);
}

// R10: sum of squared differences (double accumulation)
__global__ void sq_err_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n,
                              double* __restrict__ out) {
    // blockIdx.y = batch element: n values each, one sum each
    a += (size_t)blockIdx.y * n;
    b += (size_t)blockIdx.y * n;
    out += blockIdx.y;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (; i < n; i += stride) {
        float d = __fsub_rn(a[i], b[i]);
        acc += (double)__fmul_rn(d, d);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double ws[32];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) ws[w] = acc;
    __syncthreads();
    if (w == 0) {
        acc = (l < (blockDim.x >> 5)) ? ws[l] : 0.0;
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (l == 0) atomicAdd(out, acc);
    }
}

// R8: Bayer sub-lattices.  quad[q][h2][w2][c] = full[2*h2+r(q)][2*w2+s(q)][c]
__global__ void bayer_kernel(const float* __restrict__ src, float* __restrict__ dst,
                             int H, int W, int C, int merge) {
    size_t n = (size_t)H * W * C;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = (int)(i % C);
    size_t pw = i / C;
    int w = (int)(pw % W), h = (int)(pw / W);
    int q = ((h & 1) << 1) | (w & 1);
    size_t qi = (((size_t)q * (H / 2) + (h >> 1)) * (W / 2) + (w >> 1)) * C + c;
    if (merge) dst[i] = src[qi];
    else dst[qi] = src[i];
}

// R9: CASSI shifted mask stack
__global__ void cassi_mask_kernel(const float* __restrict__ m, float* __restrict__ Phi,
                                  int H, int W, int C, int step) {
    int Wc = W + (C - 1) * step;
    size_t n = (size_t)H * Wc * C;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = (int)(i % C);
    size_t pw = i / C;
    int wc = (int)(pw % Wc), h = (int)(pw / Wc);
    int w = wc - step * k;
    Phi[i] = (w >= 0 && w < W) ? m[(size_t)h * W + w] : 0.f;
}

// ---------------------------------------------------------------------------
// host wrappers
// ---------------------------------------------------------------------------
static int check_dims(int B, int H, int W, int C) {
    if (B < 1 || H < 1 || W < 1 || C < 1) {
        set_error("bad dimensions B=%d H=%d W=%d C=%d", B, H, W, C);
        return SCIPNP_EINVAL;
    }
    if (B > 65535) { set_error("B=%d exceeds 65535", B); return SCIPNP_EINVAL; }
    return SCIPNP_OK;
}

int launch_project(int mode, const float* a_in, const float* b_in, float* x_out, float* f_out,
                   const float* y1_in, float* y1_out, const float* y, const float* Phi,
                   const float* Phi_sum, float lambda, float gamma, int B, int H, int W, int C,
                   int phi_batched, cudaStream_t st) {
    long long npix = (long long)H * W;
    bool v4 = (C % 4 == 0) && aligned16(a_in) && aligned16(Phi) && aligned16(x_out) &&
              (b_in == nullptr || aligned16(b_in)) && (f_out == nullptr || aligned16(f_out));
    int V = v4 ? 4 : 1;
    int cpp = C / V;
    int Pmax = (kProjThreads * kProjK) / cpp;
    if (Pmax < 1) { set_error("C=%d too large for the projection kernel (max %d)", C, kProjThreads * kProjK * 4); return SCIPNP_EINVAL; }
    // enough CTAs to fill the machine on small scenes
    long long want = ceil_div_ll(npix * B, 4LL * num_sms());
    int P = (int)min((long long)Pmax, max(1LL, want));
    P = min(P, 1024);
    size_t smem = ((size_t)C * (P | 1) + P) * sizeof(float);
    while (smem > 200 * 1024 && P > 1) { P /= 2; smem = ((size_t)C * (P | 1) + P) * sizeof(float); }
    dim3 grid((unsigned)ceil_div_ll(npix, P), B);
    long long phi_stride = phi_batched ? npix * C : 0, ps_stride = phi_batched ? npix : 0;
#define LAUNCH(VV, MM)                                                                             \
    do {                                                                                           \
        auto kfn = project_kernel<VV, MM>;                                                         \
        if (smem > 48 * 1024) SCIPNP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kfn<<<grid, kProjThreads, smem, st>>>(a_in, b_in, x_out, f_out, y1_in, y1_out, y, Phi,    \
                                              Phi_sum, lambda, gamma, npix, C, P, phi_stride, ps_stride); \
    } while (0)
    if (V == 4) {
        if (mode == MODE_GAP_ACC) LAUNCH(4, MODE_GAP_ACC);
        else if (mode == MODE_GAP_PLAIN) LAUNCH(4, MODE_GAP_PLAIN);
        else LAUNCH(4, MODE_ADMM);
    } else {
        if (mode == MODE_GAP_ACC) LAUNCH(1, MODE_GAP_ACC);
        else if (mode == MODE_GAP_PLAIN) LAUNCH(1, MODE_GAP_PLAIN);
        else LAUNCH(1, MODE_ADMM);
    }
#undef LAUNCH
    count_launch();
    return check_launch("project_kernel");
}

int launch_clip01(float* x, size_t n, cudaStream_t st) {
    if (n == 0) return SCIPNP_OK;
    unsigned grid = (unsigned)min((long long)ceil_div_ll((long long)n, 256 * 4), 8LL * num_sms());
    clip01_kernel<<<grid, 256, 0, st>>>(x, n);
    count_launch();
    return check_launch("clip01_kernel");
}

int launch_sq_err(const float* a, const float* b, size_t n_per_batch, int B, double* sums,
                  cudaStream_t st) {
    if (n_per_batch == 0 || B < 1) return SCIPNP_OK;
    long long per = ceil_div_ll((long long)n_per_batch, 256 * 8);
    long long cap = max(1LL, (8LL * num_sms()) / B);
    dim3 grid((unsigned)min(per, cap), B);
    sq_err_kernel<<<grid, 256, 0, st>>>(a, b, n_per_batch, sums);
    count_launch();
    return check_launch("sq_err_kernel");
}

}  // namespace scipnp

using namespace scipnp;

// x0 = At(y) and Phi_sum in one pass (internal; see init_x0_phisum_kernel).  Returns SCIPNP_EINVAL without launching
// when the shape is not covered (the caller then uses scipnp_phi_sum + scipnp_At).
namespace scipnp {
int launch_init_x0_phisum(const float* y, const float* Phi, float* x, float* phisum, int B, int H, int W, int C,
                          int phi_batched, cudaStream_t st) {
    const size_t smem = (size_t)kInitP * (C + 1) * sizeof(float);
    if ((C & 3) != 0 || C > kMaxLocalC || smem > 48 * 1024 || !aligned16(Phi) || !aligned16(x)) return SCIPNP_EINVAL;
    const long long npix = (long long)H * W;
    dim3 grid((unsigned)ceil_div_ll(npix, kInitP), B);
    init_x0_phisum_kernel<<<grid, kInitThreads, smem, st>>>(y, reinterpret_cast<const float4*>(Phi), reinterpret_cast<float4*>(x),
                                                           phisum, npix, C / 4, phi_batched ? npix * (C / 4) : 0, phi_batched ? 1 : 0);
    count_launch();
    return check_launch("init_x0_phisum_kernel");
}
}  // namespace scipnp

extern "C" {

int scipnp_A(const float* x, const float* Phi, float* y, int B, int H, int W, int C,
             int phi_batched, void* stream) {
    if (int e = check_dims(B, H, W, C)) return e;
    SCIPNP_REQUIRE(x && Phi && y, "null pointer");
    long long npix = (long long)H * W;
    dim3 grid((unsigned)ceil_div_ll(npix, 128), B);
    A_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, Phi, y, npix, C, phi_batched ? npix * C : 0);
    count_launch();
    return check_launch("A_kernel");
}

int scipnp_At(const float* y, const float* Phi, float* x, int B, int H, int W, int C,
              int phi_batched, void* stream) {
    if (int e = check_dims(B, H, W, C)) return e;
    SCIPNP_REQUIRE(x && Phi && y, "null pointer");
    long long nelem = (long long)H * W * C;
    if ((C & 3) == 0 && aligned16(Phi) && aligned16(x)) {
        const long long n4 = nelem / 4;
        dim3 grid4((unsigned)ceil_div_ll(n4, 256), B);
        At_kernel_v4<<<grid4, 256, 0, (cudaStream_t)stream>>>(y, reinterpret_cast<const float4*>(Phi),
                                                               reinterpret_cast<float4*>(x), n4, C / 4,
                                                               phi_batched ? n4 : 0);
        count_launch();
        return check_launch("At_kernel_v4");
    }
    dim3 grid((unsigned)ceil_div_ll(nelem, 256), B);
    At_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, Phi, x, nelem, C, phi_batched ? nelem : 0);
    count_launch();
    return check_launch("At_kernel");
}

int scipnp_phi_sum(const float* Phi, float* out, int B, int H, int W, int C, void* stream) {
    if (int e = check_dims(B, H, W, C)) return e;
    SCIPNP_REQUIRE(Phi && out, "null pointer");
    long long npix = (long long)H * W;
    dim3 grid((unsigned)ceil_div_ll(npix, 128), B);
    phi_sum_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(Phi, out, npix, C);
    count_launch();
    return check_launch("phi_sum_kernel");
}

int scipnp_gap_project(const float* x_in, float* x_out, const float* y1_in, float* y1_out,
                       const float* y, const float* Phi, const float* Phi_sum, float lambda,
                       int accelerate, int B, int H, int W, int C, int phi_batched, void* stream) {
    if (int e = check_dims(B, H, W, C)) return e;
    SCIPNP_REQUIRE(x_in && x_out && y && Phi && Phi_sum, "null pointer");
    SCIPNP_REQUIRE(!accelerate || (y1_in && y1_out), "accelerated GAP needs y1");
    return launch_project(accelerate ? MODE_GAP_ACC : MODE_GAP_PLAIN, x_in, nullptr, x_out, nullptr,
                          y1_in, y1_out, y, Phi, Phi_sum, lambda, 0.f, B, H, W, C, phi_batched,
                          (cudaStream_t)stream);
}

int scipnp_admm_project(const float* theta, const float* b, float* x, float* f, const float* y,
                        const float* Phi, const float* Phi_sum, float lambda, float gamma, int B,
                        int H, int W, int C, int phi_batched, void* stream) {
    if (int e = check_dims(B, H, W, C)) return e;
    SCIPNP_REQUIRE(theta && b && x && f && y && Phi && Phi_sum, "null pointer");
    return launch_project(MODE_ADMM, theta, b, x, f, nullptr, nullptr, y, Phi, Phi_sum, lambda, gamma,
                          B, H, W, C, phi_batched, (cudaStream_t)stream);
}

int scipnp_admm_dual_update(float* b, const float* x, const float* theta, size_t n, void* stream) {
    SCIPNP_REQUIRE(b && x && theta, "null pointer");
    if (n == 0) return SCIPNP_OK;
    unsigned grid = (unsigned)min((long long)ceil_div_ll((long long)n, 256), 8LL * 148 * 4);
    admm_dual_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(b, x, theta, n);
    count_launch();
    return check_launch("admm_dual_kernel");
}

int scipnp_sq_err(const float* a, const float* b, size_t n, double* sum_dev, void* stream) {
    SCIPNP_REQUIRE(a && b && sum_dev, "null pointer");
    return launch_sq_err(a, b, n, 1, sum_dev, (cudaStream_t)stream);
}

// ---- per-frame image quality of the return tuples ---------------------------------------------
// compare_psnr / compare_ssim of scikit-image < 0.18 as the reference calls them
// (pnp_sci_algo.py:699-705, 857-863: per channel, data_range = 1): SSIM with a 7x7 uniform window,
// sample covariance (cov_norm = 49/48), K1 = 0.01, K2 = 0.03, averaged over the pixels whose window
// lies inside the image; squared error in float32, summed in double.  Arithmetic in double like
// skimage (inputs are converted to float64 there).
namespace {

constexpr int kIqaWin = 7, kIqaPad = 3, kIqaSeg = 64;

__global__ void __launch_bounds__(128) frames_iqa_kernel(const float* __restrict__ ref, const float* __restrict__ img,
                                                         int H, int W, int C, double* __restrict__ ssim_sum,
                                                         double* __restrict__ sqerr_sum) {
    extern __shared__ double sacc[];                   // [2][C]
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.0;
    __syncthreads();
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (w, c), c fastest
    const bool valid = j < (long long)W * C;
    const int w = valid ? (int)(j / C) : 0, c = valid ? (int)(j - (long long)w * C) : 0;
    const int h0 = blockIdx.y * kIqaSeg, h1 = min(H, h0 + kIqaSeg);
    double ss = 0.0, se = 0.0;
    if (valid) {
        for (int h = h0; h < h1; ++h) {                // squared error: every pixel
            const size_t o = ((size_t)h * W + w) * C + c;
            const float d = ref[o] - img[o];
            se += (double)(d * d);
        }
        const bool inner = w >= kIqaPad && w < W - kIqaPad;
        const int c0 = max(h0, kIqaPad), c1 = min(h1, H - kIqaPad);           // window centres of this segment
        if (inner && c1 > c0) {
            double rx[kIqaWin], ry[kIqaWin], rxx[kIqaWin], ryy[kIqaWin], rxy[kIqaWin];   // row sums of the last 7 rows
#pragma unroll
            for (int i = 0; i < kIqaWin; ++i) rx[i] = ry[i] = rxx[i] = ryy[i] = rxy[i] = 0.0;
            for (int r = c0 - kIqaPad; r < c1 + kIqaPad; ++r) {
                double hx = 0, hy = 0, hxx = 0, hyy = 0, hxy = 0;
#pragma unroll
                for (int d = -kIqaPad; d <= kIqaPad; ++d) {
                    const size_t o = ((size_t)r * W + (w + d)) * C + c;
                    const double a = (double)ref[o], b = (double)img[o];
                    hx += a; hy += b; hxx += a * a; hyy += b * b; hxy += a * b;
                }
#pragma unroll
                for (int i = 0; i < kIqaWin - 1; ++i) {
                    rx[i] = rx[i + 1]; ry[i] = ry[i + 1]; rxx[i] = rxx[i + 1]; ryy[i] = ryy[i + 1]; rxy[i] = rxy[i + 1];
                }
                rx[kIqaWin - 1] = hx; ry[kIqaWin - 1] = hy; rxx[kIqaWin - 1] = hxx; ryy[kIqaWin - 1] = hyy; rxy[kIqaWin - 1] = hxy;
                if (r >= c0 + kIqaPad) {               // the window centred on row r-3 is complete
                    double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
                    for (int i = 0; i < kIqaWin; ++i) { sx += rx[i]; sy += ry[i]; sxx += rxx[i]; syy += ryy[i]; sxy += rxy[i]; }
                    const double inv = 1.0 / (kIqaWin * kIqaWin), cn = 49.0 / 48.0;
                    const double ux = sx * inv, uy = sy * inv;
                    const double vx = cn * (sxx * inv - ux * ux), vy = cn * (syy * inv - uy * uy);
                    const double vxy = cn * (sxy * inv - ux * uy);
                    const double C1 = 0.01 * 0.01, C2 = 0.03 * 0.03;
                    ss += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
                }
            }
        }
        atomicAdd(&sacc[c], ss);
        atomicAdd(&sacc[C + c], se);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        if (sacc[i] != 0.0) atomicAdd(ssim_sum + i, sacc[i]);
        if (sacc[C + i] != 0.0) atomicAdd(sqerr_sum + i, sacc[C + i]);
    }
}

}  // namespace

int scipnp_frames_iqa(const float* ref, const float* img, int H, int W, int C, double* ssim_sum_dev,
                      double* sqerr_sum_dev, void* stream) {
    if (int e = check_dims(1, H, W, C)) return e;
    SCIPNP_REQUIRE(ref && img && ssim_sum_dev && sqerr_sum_dev, "null pointer");
    SCIPNP_REQUIRE(H >= kIqaWin && W >= kIqaWin, "win_size exceeds image extent (7x7 SSIM window)");
    SCIPNP_REQUIRE(C <= 4096, "too many channels");
    cudaStream_t st = (cudaStream_t)stream;
    SCIPNP_CUDA(cudaMemsetAsync(ssim_sum_dev, 0, C * sizeof(double), st));
    SCIPNP_CUDA(cudaMemsetAsync(sqerr_sum_dev, 0, C * sizeof(double), st));
    dim3 grid((unsigned)ceil_div_ll((long long)W * C, 128), (unsigned)ceil_div_ll(H, kIqaSeg), 1);
    frames_iqa_kernel<<<grid, 128, 2 * C * sizeof(double), st>>>(ref, img, H, W, C, ssim_sum_dev, sqerr_sum_dev);
    count_launch();
    return check_launch("frames_iqa_kernel");
}

int scipnp_bayer_split(const float* full, float* quad, int H, int W, int C, void* stream) {
    if (int e = check_dims(1, H, W, C)) return e;
    SCIPNP_REQUIRE(full && quad, "null pointer");
    SCIPNP_REQUIRE(H % 2 == 0 && W % 2 == 0, "Bayer mosaics need even H and W");
    size_t n = (size_t)H * W * C;
    bayer_kernel<<<(unsigned)ceil_div_ll((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(full, quad, H, W, C, 0);
    count_launch();
    return check_launch("bayer_kernel");
}

int scipnp_bayer_merge(const float* quad, float* full, int H, int W, int C, void* stream) {
    if (int e = check_dims(1, H, W, C)) return e;
    SCIPNP_REQUIRE(full && quad, "null pointer");
    SCIPNP_REQUIRE(H % 2 == 0 && W % 2 == 0, "Bayer mosaics need even H and W");
    size_t n = (size_t)H * W * C;
    bayer_kernel<<<(unsigned)ceil_div_ll((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(quad, full, H, W, C, 1);
    count_launch();
    return check_launch("bayer_kernel");
}

int scipnp_cassi_shift_mask(const float* mask2d, float* Phi, int H, int W, int C, int step, void* stream) {
    if (int e = check_dims(1, H, W, C)) return e;
    SCIPNP_REQUIRE(mask2d && Phi, "null pointer");
    SCIPNP_REQUIRE(step >= 0, "negative dispersion step");
    size_t n = (size_t)H * (W + (size_t)(C - 1) * step) * C;
    cassi_mask_kernel<<<(unsigned)ceil_div_ll((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(mask2d, Phi, H, W, C, step);
    count_launch();
    return check_launch("cassi_mask_kernel");
}

}  // extern "C"
