// The rest of the MATLAB twin's TV family (gapdenoise.m:86-108) on [B][H][W][C] stacks (C = MATLAB's third
// dimension, the frames):
//   ATV_ClipB   TV_denoising_clip_LB.m:25-36     iterative clipping, no averaging, clip level lambda
//   ATV_cham    tvdenoise_cham_ATV2D.m:72-87     Chambolle projection, anisotropic:  p = t / max(1, |t|)
//   ITV2D_cham  tvdenoise_cham_ITV2D.m:73-90     isotropic per frame:                p = t / (1 + dt |grad z|)
//   ITV3D_cham  tvdenoise_cham_ITV3D.m:72-90     isotropic, gradient norm summed over the frames
//   ATV_FGP / ITV2D_FGP / ITV3D_FGP  fgp_denoise_*.m:73-121   fast gradient projection (Beck & Teboulle)
// IEEE single precision in the statement order of the .m files (a float32 NumPy restatement, oracle/matlab_tv.py, is
// reproduced bit for bit; no FMA contraction, correctly rounded divide and sqrt).  Exact-path style: the dual fields
// live in HBM, two launches per iteration.  Not fused, not tuned: these are the alternative denoisers of the MATLAB
// driver, outside the north-star path.
#include "internal.cuh"

namespace scipnp {
namespace {

// NumPy's float32 add-reduce order over one contiguous run (as numpy_sum of ops.cu; n <= 128 here)
__device__ float np_sum(const float* a, int n) {
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
        return r;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
}

constexpr int kMaxC = 128;
enum { V_CLIPB = 0, V_CHAM_ATV2D = 1, V_CHAM_ITV2D = 2, V_CHAM_ITV3D = 3, V_FGP_ATV2D = 4, V_FGP_ITV2D = 5, V_FGP_ITV3D = 6 };

struct Geo { int H, W, C; size_t n; };

__device__ __forceinline__ void locate(const Geo& g, size_t i, int& h, int& w) {
    const size_t pix = i / g.C;
    w = (int)(pix % g.W);
    h = (int)((pix / g.W) % g.H);
}

// ---- ATV_ClipB: x0 = y0 - dht(zh) - dvt(zv);  z = clip(z + (1/alpha) d(x0), lambda) --------------------------------
__global__ void clipb_x_kernel(const float* __restrict__ y0, const float* __restrict__ zh, const float* __restrict__ zv,
                               float* __restrict__ x0, Geo g) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    int h, w;
    locate(g, i, h, w);
    const size_t sw = (size_t)g.C, sh = (size_t)g.W * g.C;
    float dht, dvt;                                    // dht_3d / dvt_3d, TV_denoising_clip_LB.m:66-71
    if (w == 0) dht = -zh[i];
    else if (w == g.W - 1) dht = zh[i - sw];
    else dht = -__fsub_rn(zh[i], zh[i - sw]);
    if (h == 0) dvt = -zv[i];
    else if (h == g.H - 1) dvt = zv[i - sh];
    else dvt = -__fsub_rn(zv[i], zv[i - sh]);
    x0[i] = __fsub_rn(__fsub_rn(y0[i], dht), dvt);     // :33
}

__device__ __forceinline__ float clipf(float v, float t) {      // sign(x).*min(abs(x), t)
    const float m = fminf(fabsf(v), t);
    return v > 0.f ? m : (v < 0.f ? -m : 0.f * m);
}

__global__ void clipb_z_kernel(const float* __restrict__ x0, float* __restrict__ zh, float* __restrict__ zv,
                               float inv_alpha, float lam, Geo g) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    int h, w;
    locate(g, i, h, w);
    const size_t sw = (size_t)g.C, sh = (size_t)g.W * g.C;
    const float v = x0[i];
    if (w < g.W - 1) zh[i] = clipf(__fadd_rn(zh[i], __fmul_rn(inv_alpha, __fsub_rn(x0[i + sw], v))), lam);   // :34
    if (h < g.H - 1) zv[i] = clipf(__fadd_rn(zv[i], __fmul_rn(inv_alpha, __fsub_rn(x0[i + sh], v))), lam);   // :35
}

// ---- Chambolle variants: one thread per pixel (all C frames: ITV3D couples them) -----------------------------------
// z = divp - f*lambda;  z1 = z(:,ir,:) - z;  z2 = z(id,:,:) - z;  p update;  (divp in the second kernel)
template <int VAR>
__global__ void cham_p_kernel(const float* __restrict__ f, const float* __restrict__ divp, float* __restrict__ p1,
                              float* __restrict__ p2, float lam, float dt, Geo g) {
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= g.n / g.C) return;
    const int w = (int)(pix % g.W), h = (int)((pix / g.W) % g.H);
    const size_t sw = (size_t)g.C, sh = (size_t)g.W * g.C, i0 = pix * g.C;
    const bool hasr = w < g.W - 1, hasd = h < g.H - 1;
    float acc[kMaxC];
    float denom3 = 1.f;
    if (VAR == V_CHAM_ITV3D) {
        for (int c = 0; c < g.C; ++c) {
            const size_t i = i0 + c;
            const float z = __fsub_rn(divp[i], __fmul_rn(f[i], lam));
            const float z1 = hasr ? __fsub_rn(__fsub_rn(divp[i + sw], __fmul_rn(f[i + sw], lam)), z) : __fsub_rn(z, z);
            const float z2 = hasd ? __fsub_rn(__fsub_rn(divp[i + sh], __fmul_rn(f[i + sh], lam)), z) : __fsub_rn(z, z);
            acc[c] = __fadd_rn(__fmul_rn(z1, z1), __fmul_rn(z2, z2));
        }
        denom3 = __fadd_rn(1.f, __fmul_rn(dt, __fsqrt_rn(np_sum(acc, g.C))));      // ITV3D :82
    }
    for (int c = 0; c < g.C; ++c) {
        const size_t i = i0 + c;
        const float z = __fsub_rn(divp[i], __fmul_rn(f[i], lam));
        const float z1 = hasr ? __fsub_rn(__fsub_rn(divp[i + sw], __fmul_rn(f[i + sw], lam)), z) : __fsub_rn(z, z);
        const float z2 = hasd ? __fsub_rn(__fsub_rn(divp[i + sh], __fmul_rn(f[i + sh], lam)), z) : __fsub_rn(z, z);
        const float t1 = __fadd_rn(p1[i], __fmul_rn(dt, z1)), t2 = __fadd_rn(p2[i], __fmul_rn(dt, z2));
        if (VAR == V_CHAM_ATV2D) {
            p1[i] = __fdiv_rn(t1, fmaxf(1.f, fabsf(t1)));                           // ATV2D :83-84
            p2[i] = __fdiv_rn(t2, fmaxf(1.f, fabsf(t2)));
        } else {
            const float denom = VAR == V_CHAM_ITV3D
                ? denom3 : __fadd_rn(1.f, __fmul_rn(dt, __fsqrt_rn(__fadd_rn(__fmul_rn(z1, z1), __fmul_rn(z2, z2)))));
            p1[i] = __fdiv_rn(t1, denom);
            p2[i] = __fdiv_rn(t2, denom);
        }
    }
}

// divp = p1 - p1(:,il,:) + p2 - p2(iu,:,:) with il / iu = [1, 1:N-1] (the first column / row repeats itself);
// last = 1: u = f - divp/lambda instead
__global__ void cham_div_kernel(const float* __restrict__ p1, const float* __restrict__ p2, float* __restrict__ divp,
                                const float* __restrict__ f, float* __restrict__ u, float lam, int last, Geo g) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    int h, w;
    locate(g, i, h, w);
    const size_t sw = (size_t)g.C, sh = (size_t)g.W * g.C;
    const float a = p1[i], b = p2[i];
    const float al = w > 0 ? p1[i - sw] : a, bu = h > 0 ? p2[i - sh] : b;
    const float d = __fsub_rn(__fadd_rn(__fsub_rn(a, al), b), bu);
    if (last) u[i] = __fsub_rn(f[i], __fdiv_rn(d, lam));
    else divp[i] = d;
}

// ---- FGP variants -----------------------------------------------------------------------------------------------------
// D = Xobs - lambda*Lforward_3d(R)   (fgp_denoise_*.m:85, :127-150)
__global__ void fgp_d_kernel(const float* __restrict__ xobs, const float* __restrict__ r1, const float* __restrict__ r2,
                             float* __restrict__ D, float lam, Geo g) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    int h, w;
    locate(g, i, h, w);
    const size_t sw = (size_t)g.C, sh = (size_t)g.W * g.C;
    float x = h < g.H - 1 ? r1[i] : 0.f;                       // X(1:m-1,:,:) = P{1}
    if (w < g.W - 1) x = __fadd_rn(x, r2[i]);                   // X(:,1:n-1,:) += P{2}
    if (h > 0) x = __fsub_rn(x, r1[i - sh]);                    // X(2:m,:,:) -= P{1}
    if (w > 0) x = __fsub_rn(x, r2[i - sw]);                    // X(:,2:n,:) -= P{2}
    D[i] = __fsub_rn(xobs[i], __fmul_rn(lam, x));
}

// Q = Ltrans_3d(D); P = R + c*Q; projection; R = P + wgt*(P - Pold)   (:87-109); one thread per pixel
template <int VAR>
__global__ void fgp_p_kernel(const float* __restrict__ D, float* __restrict__ P1, float* __restrict__ P2,
                             float* __restrict__ R1, float* __restrict__ R2, float c, float wgt, Geo g) {
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= g.n / g.C) return;
    const int w = (int)(pix % g.W), h = (int)((pix / g.W) % g.H);
    const size_t sw = (size_t)g.C, sh = (size_t)g.W * g.C, i0 = pix * g.C;
    const bool has1 = h < g.H - 1, has2 = w < g.W - 1;
    float acc[kMaxC];
    float a3 = 1.f;
    if (VAR == V_FGP_ITV3D) {
        for (int ch = 0; ch < g.C; ++ch) {
            const size_t i = i0 + ch;
            const float n1 = has1 ? __fadd_rn(R1[i], __fmul_rn(c, __fsub_rn(D[i], D[i + sh]))) : 0.f;
            const float n2 = has2 ? __fadd_rn(R2[i], __fmul_rn(c, __fsub_rn(D[i], D[i + sw]))) : 0.f;
            acc[ch] = __fadd_rn(__fmul_rn(n1, n1), __fmul_rn(n2, n2));
        }
        a3 = __fsqrt_rn(fmaxf(np_sum(acc, g.C), 1.f));                               // ITV3D :96
    }
    for (int ch = 0; ch < g.C; ++ch) {
        const size_t i = i0 + ch;
        float n1 = has1 ? __fadd_rn(R1[i], __fmul_rn(c, __fsub_rn(D[i], D[i + sh]))) : 0.f;
        float n2 = has2 ? __fadd_rn(R2[i], __fmul_rn(c, __fsub_rn(D[i], D[i + sw]))) : 0.f;
        if (VAR == V_FGP_ATV2D) {
            n1 = __fdiv_rn(n1, fmaxf(fabsf(n1), 1.f));
            n2 = __fdiv_rn(n2, fmaxf(fabsf(n2), 1.f));
        } else {
            const float A = VAR == V_FGP_ITV3D ? a3 : __fsqrt_rn(fmaxf(__fadd_rn(__fmul_rn(n1, n1), __fmul_rn(n2, n2)), 1.f));
            n1 = __fdiv_rn(n1, A);
            n2 = __fdiv_rn(n2, A);
        }
        if (has1) { const float o = P1[i]; P1[i] = n1; R1[i] = __fadd_rn(n1, __fmul_rn(wgt, __fsub_rn(n1, o))); }
        if (has2) { const float o = P2[i]; P2[i] = n2; R2[i] = __fadd_rn(n2, __fmul_rn(wgt, __fsub_rn(n2, o))); }
    }
}

}  // namespace
}  // namespace scipnp

extern "C" {

size_t scipnp_tv_matlab_workspace_bytes(int B, int H, int W, int C) {
    if (B < 1 || H < 1 || W < 1 || C < 1) return 0;
    return 4 * (size_t)B * H * W * C * sizeof(float);
}

int scipnp_tv_matlab(const float* in, float* out, int variant, float lambda, int iters, int B, int H, int W, int C,
                     void* workspace, size_t workspace_bytes, void* stream) {
    using namespace scipnp;
    SCIPNP_REQUIRE(B >= 1 && H >= 2 && W >= 2 && C >= 1 && C <= kMaxC, "bad dimensions (frames at least 2x2, C <= 128)");
    SCIPNP_REQUIRE(in && out && workspace, "null pointer");
    SCIPNP_REQUIRE(in != out, "in and out must not alias");
    SCIPNP_REQUIRE(iters >= 1, "iters must be >= 1");
    SCIPNP_REQUIRE(variant >= V_CLIPB && variant <= V_FGP_ITV3D, "unknown variant");
    SCIPNP_REQUIRE(lambda > 0.f, "lambda must be positive");
    Geo g{H, W, C, (size_t)B * H * W * C};
    if (workspace_bytes < 4 * g.n * sizeof(float)) { set_error("MATLAB-TV workspace too small"); return SCIPNP_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    float* w0 = reinterpret_cast<float*>(workspace);
    float *w1 = w0 + g.n, *w2 = w1 + g.n, *w3 = w2 + g.n;
    SCIPNP_CUDA(cudaMemsetAsync(w0, 0, 4 * g.n * sizeof(float), st));
    const unsigned be = (unsigned)((g.n + 255) / 256), bp = (unsigned)((g.n / C + 127) / 128);
    if (variant == V_CLIPB) {
        const float inv_alpha = (float)(1.0 / 5.0);
        for (int it = 0; it < iters; ++it) {
            clipb_x_kernel<<<be, 256, 0, st>>>(in, w0, w1, out, g);
            count_launch();
            if (it + 1 < iters) { clipb_z_kernel<<<be, 256, 0, st>>>(out, w0, w1, inv_alpha, lambda, g); count_launch(); }
        }
    } else if (variant <= V_CHAM_ITV3D) {
        const float dt = variant == V_CHAM_ITV3D ? 0.25f : 0.125f;
        float *p1 = w0, *p2 = w1, *divp = w2;
        for (int it = 0; it < iters; ++it) {
            if (variant == V_CHAM_ATV2D) cham_p_kernel<V_CHAM_ATV2D><<<bp, 128, 0, st>>>(in, divp, p1, p2, lambda, dt, g);
            else if (variant == V_CHAM_ITV2D) cham_p_kernel<V_CHAM_ITV2D><<<bp, 128, 0, st>>>(in, divp, p1, p2, lambda, dt, g);
            else cham_p_kernel<V_CHAM_ITV3D><<<bp, 128, 0, st>>>(in, divp, p1, p2, lambda, dt, g);
            cham_div_kernel<<<be, 256, 0, st>>>(p1, p2, divp, in, out, lambda, it + 1 == iters ? 1 : 0, g);
            count_launch(2);
        }
    } else {
        float *P1 = w0, *P2 = w1, *R1 = w2, *R2 = w3;
        const float c = 1.0f / (8.0f * lambda);
        double tkp1 = 1.0;
        for (int it = 0; it < iters; ++it) {
            const double tk = tkp1;
            fgp_d_kernel<<<be, 256, 0, st>>>(in, R1, R2, out, lambda, g);
            count_launch();
            tkp1 = (1.0 + sqrt(1.0 + 4.0 * tk * tk)) / 2.0;
            if (it + 1 == iters) break;                       // X_den = D of the last iteration: its dual update is dead
            const float wgt = (float)((tk - 1.0) / tkp1);
            if (variant == V_FGP_ATV2D) fgp_p_kernel<V_FGP_ATV2D><<<bp, 128, 0, st>>>(out, P1, P2, R1, R2, c, wgt, g);
            else if (variant == V_FGP_ITV2D) fgp_p_kernel<V_FGP_ITV2D><<<bp, 128, 0, st>>>(out, P1, P2, R1, R2, c, wgt, g);
            else fgp_p_kernel<V_FGP_ITV3D><<<bp, 128, 0, st>>>(out, P1, P2, R1, R2, c, wgt, g);
            count_launch();
        }
    }
    return check_launch("MATLAB TV kernels");
}

}  // extern "C"
