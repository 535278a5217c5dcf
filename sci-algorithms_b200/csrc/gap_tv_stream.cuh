// Kernel template of the one-pass fused GAP-TV iteration (see gap_tv_fused.cu for the
// design notes).  Included by the per-R instantiation units fused_inst_r*.cu.
#pragma once
#include "internal.cuh"

namespace scipnp {
namespace fusedk {


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
__host__ __device__ constexpr Smem smem_layout(int K, int NG) {
    Smem s{};
    s.tile_f4_per_row = NG * K * PADL;
    s.row_bytes = 2 * s.tile_f4_per_row * 16 + 3 * NG * 32 * 4;
    s.buf_bytes = RB * s.row_bytes;
    s.KP = (K + 3) & ~3;
    s.part_off = 2 * s.buf_bytes;
    s.total = s.part_off + RB * NG * 32 * s.KP * 4;
    return s;
}

// ---- packed single precision (Blackwell FFMA2/FADD2/FMUL2: two lanes per issue slot) -------
typedef float2 P2;
__device__ __forceinline__ P2 splat(float a) { return make_float2(a, a); }
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ P2 mul2(P2 a, P2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ P2 add2(P2 a, P2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ P2 sqrt2(P2 a) { return make_float2(fast_sqrt(a.x), fast_sqrt(a.y)); }
__device__ __forceinline__ P2 rcp2(P2 a) { return make_float2(fast_rcp(a.x), fast_rcp(a.y)); }
__device__ __forceinline__ P2 shfl_idx2(P2 a, int src) {
    return make_float2(__shfl_sync(0xffffffffu, a.x, src), __shfl_sync(0xffffffffu, a.y, src));
}
__device__ __forceinline__ P2 shfl_up2(P2 a) {
    return make_float2(__shfl_up_sync(0xffffffffu, a.x, 1), __shfl_up_sync(0xffffffffu, a.y, 1));
}

// CTA size per chunk count K = C/4: NG = 8/K pixel groups of K chunk-warps each
__host__ __device__ constexpr int fused_groups(int K) { return 8 / K < 1 ? 1 : 8 / K; }
__host__ __device__ constexpr int fused_threads(int K) { return fused_groups(K) * K * 32; }

template <int R, int MODE, bool CHECK, int K>
__global__ void __launch_bounds__(fused_threads(K), 2)
gap_tv_stream_kernel(const FusedParams p) {
    constexpr int NG = fused_groups(K);
    constexpr int NT = fused_threads(K);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = p.W, H = p.H, C = p.C;
    constexpr Smem L = smem_layout(K, NG);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int gi = warp / K, k = warp - gi * K;
    const int b = blockIdx.z;
    constexpr int OWN = 32 - 2 * R;              // owned pixels per group
    const int group0 = blockIdx.x * NG;          // first pixel group of this CTA
    const int grp = group0 + gi;
    const bool grp_live = grp < p.ngroups;
    const int px = grp * OWN - R + lane;         // this lane's pixel column
    const bool px_in = grp_live && px >= 0 && px < W;
    const bool own_px = px_in && lane >= R && lane < 32 - R;
    // the right neighbour of the last image column is the pixel itself (g1 = 0 there)
    const int src_right = (px < W - 1 && lane < 31) ? lane + 1 : lane;
    const float pxin_f = px_in ? 1.f : 0.f;

    const int r0 = blockIdx.y * p.seg_rows;
    const int r1 = min(H, r0 + p.seg_rows);
    const int rs = max(0, r0 - R), rend = r1 + R;       // steps rho in [rs, rend)
    const int load_end = min(H, rend);
    const int nblk = (rend - rs + RB - 1) / RB;

    const size_t frame_b = (size_t)b * H * W * C;        // batch offsets
    const size_t meas_b = (size_t)b * H * W;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);

    // ---- producer set-up: every thread owns one 16-byte chunk of x and of Phi per row -----------
    //      (NG*32*K chunks per row == NT threads) and threads < 3*NG*32 one y/y1/Phi_sum value
    const float* src_x; const float* src_phi; uint32_t dst_tile; int tile_bytes;
    {
        const int g2 = tid / (32 * K);
        const int rem = tid - g2 * 32 * K;
        const int ln = rem / K, kk = rem - ln * K;
        const int gpx = (group0 + g2) * OWN - R + ln;
        const bool ok = gpx >= 0 && gpx < W && (group0 + g2) < p.ngroups;
        const size_t off = (size_t)(ok ? gpx : 0) * C + 4 * kk;
        src_x = p.x_in + frame_b + off;
        src_phi = p.Phi + (size_t)b * p.phi_bstride + off;
        dst_tile = smem_base + ((g2 * K + kk) * PADL + ln) * 16;
        tile_bytes = ok ? 16 : 0;
    }
    const float* src_small[3]; uint32_t dst_small[3]; int small_bytes[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {                       // 3*NG*32 values per row, <= 3 per thread
        const int idx = tid + t * NT;
        const int which = idx / (NG * 32);              // 0: y, 1: y1, 2: Phi_sum
        const int c = idx - which * NG * 32;
        src_small[t] = nullptr; dst_small[t] = 0; small_bytes[t] = 0;
        if (which < 3) {
            const int g2 = c >> 5, ln = c & 31;
            const int gpx = (group0 + g2) * OWN - R + ln;
            const bool ok = gpx >= 0 && gpx < W && (group0 + g2) < p.ngroups;
            const float* base = which == 0 ? p.y + meas_b
                              : which == 1 ? (MODE == MODE_GAP_ACC ? p.y1_in + meas_b : nullptr)
                                           : p.Phi_sum + (size_t)b * p.ps_bstride;
            if (base) {
                src_small[t] = base + (ok ? gpx : 0);
                small_bytes[t] = ok ? 4 : 0;
                dst_small[t] = smem_base + 2 * L.tile_f4_per_row * 16 + (which * NG * 32 + c) * 4;
            }
        }
    }
    const size_t row_f = (size_t)W * C;
    auto issue = [&](int blk) {
        if (blk < nblk) {
            const uint32_t boff = (blk & 1) * L.buf_bytes;
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int row = rs + blk * RB + j;
                if (row < load_end) {
                    const uint32_t d = boff + j * L.row_bytes;
                    cp_async16(dst_tile + d, src_x + row * row_f, tile_bytes);
                    cp_async16(dst_tile + d + L.tile_f4_per_row * 16, src_phi + row * row_f, tile_bytes);
#pragma unroll
                    for (int t = 0; t < 3; ++t)
                        if (src_small[t]) cp_async4(dst_small[t] + d, src_small[t] + (size_t)row * W, small_bytes[t]);
                }
            }
        }
        cp_async_commit();
    };

    // ---- pipeline state (registers), channels as two packed pairs ------------------------------------
    P2 o_prev[R][2], g1_prev[R][2];
    P2 P0[R + 1][2], P1[R + 1][2];           // P[i] = p^i(rho-i-1); P[0] stays 0
    P2 fd[R][2];                             // fd[j] = f(rho-1-j)
    P2 en[R][2];                             // energy partials of iterations 0..R-1
    const P2 zero2 = splat(0.f);
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int c = 0; c < 2; ++c) { o_prev[i][c] = zero2; g1_prev[i][c] = zero2; fd[i][c] = zero2; en[i][c] = zero2; }
#pragma unroll
    for (int i = 0; i <= R; ++i)
#pragma unroll
        for (int c = 0; c < 2; ++c) { P0[i][c] = zero2; P1[i][c] = zero2; }

    const P2 mone2 = splat(-1.f), mtau2 = splat(-0.25f), tvc2 = splat(p.tv_c), one2 = splat(1.f);
    const float tvw = p.tv_w, lam = p.lambda;
    float* part = reinterpret_cast<float*>(smem_raw + L.part_off);
    float* xo = p.x_out + frame_b;
    float* y1o = (MODE == MODE_GAP_ACC) ? p.y1_out + meas_b : nullptr;

    issue(0);
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
        issue(blk + 1);
        cp_async_wait<1>();
        __syncthreads();
        const unsigned char* buf = smem_raw + (blk & 1) * L.buf_bytes;

        // ---- phase A: partial dot products of this warp's chunk --------------------------------------
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const int row = rs + blk * RB + j;
            if (row < load_end) {
                const float4* tx = reinterpret_cast<const float4*>(buf + j * L.row_bytes) + (gi * K + k) * PADL + lane;
                const float4 xv = tx[0], pv = tx[L.tile_f4_per_row];
                float d = xv.x * pv.x;
                d = fmaf(xv.y, pv.y, d);
                d = fmaf(xv.z, pv.z, d);
                d = fmaf(xv.w, pv.w, d);
                part[((j * NG + gi) * 32 + lane) * L.KP + k] = d;
                if (k == 0)
#pragma unroll
                    for (int kk = K; kk < L.KP; ++kk) part[((j * NG + gi) * 32 + lane) * L.KP + kk] = 0.f;
            }
        }
        __syncthreads();

        // ---- phase B: RB pipeline steps ------------------------------------------------------------------
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const int rho = rs + blk * RB + j;
            if (rho < rend) {
                P2 f_new[2] = {zero2, zero2};
                if (rho < H) {
                    // stage 0: Euclidean projection of row rho
                    const unsigned char* rowp = buf + j * L.row_bytes;
                    const float4* tx = reinterpret_cast<const float4*>(rowp) + (gi * K + k) * PADL + lane;
                    const float4 xv = tx[0], pv = tx[L.tile_f4_per_row];
                    const float4* pp = reinterpret_cast<const float4*>(part + ((j * NG + gi) * 32 + lane) * L.KP);
                    float yb = 0.f;
#pragma unroll
                    for (int q = 0; q < L.KP / 4; ++q) { const float4 t = pp[q]; yb += (t.x + t.y) + (t.z + t.w); }
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
                    const P2 s2 = splat(px_in ? s * lam : 0.f);
                    f_new[0] = fma2(s2, make_float2(pv.x, pv.y), make_float2(xv.x, xv.y));
                    f_new[1] = fma2(s2, make_float2(pv.z, pv.w), make_float2(xv.z, xv.w));
                }
                P2 o_new[2] = {f_new[0], f_new[1]};
                P2 pend0[2], pend1[2];
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const int row_new = rho - i;          // row of o_new = out_i(row_new)
                    const int u = row_new - 1;            // row whose dual variable advances
                    // masks as multipliers: the dual variable of a row outside the segment/image or of
                    // a pixel outside the image stays 0; g0 = 0 below the last image row
                    const P2 m2 = splat(((u >= rs) && (u < H)) ? pxin_f : 0.f);
                    const P2 md2 = splat(row_new < H ? 1.f : 0.f);
                    const float me = (CHECK && own_px && u >= r0 && u < r1) ? 1.f : 0.f;
                    const P2 me2 = splat(me), wm2 = splat(me * tvw);
                    P2 pi0[2], pi1[2];
#pragma unroll
                    for (int c = 0; c < 2; ++c) { pi0[c] = P0[i][c]; pi1[c] = P1[i][c]; }
                    if (i > 0) {
#pragma unroll
                        for (int c = 0; c < 2; ++c) { P0[i][c] = pend0[c]; P1[i][c] = pend1[c]; }
                    }
                    P2 o_next[2];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const P2 o_right = shfl_idx2(o_new[c], src_right);
                        const P2 g0 = mul2(fma2(o_prev[i][c], mone2, o_new[c]), md2);
                        const P2 g1 = g1_prev[i][c];
                        const P2 nrm = sqrt2(fma2(g0, g0, mul2(g1, g1)));
                        const P2 r = mul2(rcp2(fma2(nrm, tvc2, one2)), m2);
                        const P2 pn0 = mul2(fma2(g0, mtau2, pi0[c]), r);
                        const P2 pn1 = mul2(fma2(g1, mtau2, pi1[c]), r);
                        const P2 p1l = shfl_up2(pn1);
                        // D(p^{i+1})(u) = (p0(u-1) - p0(u)) + (p1(u, left) - p1(u))
                        const P2 d = add2(fma2(pn0, mone2, P0[i + 1][c]), fma2(pn1, mone2, p1l));
                        o_next[c] = add2(fd[i][c], d);
                        if (CHECK) {
                            en[i][c] = fma2(nrm, wm2, en[i][c]);                        // w*|grad out_i|(u)
                            if (i + 1 < R) en[i + 1][c] = fma2(mul2(d, me2), d, en[i + 1][c]);   // D(p^{i+1})(u)^2
                        }
                        g1_prev[i][c] = fma2(o_new[c], mone2, o_right);
                        o_prev[i][c] = o_new[c];
                        pend0[c] = pn0;
                        pend1[c] = pn1;
                    }
                    o_new[0] = o_next[0];
                    o_new[1] = o_next[1];
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) { P0[R][c] = pend0[c]; P1[R][c] = pend1[c]; }
                // f delay line
#pragma unroll
                for (int i = R - 1; i > 0; --i) { fd[i][0] = fd[i - 1][0]; fd[i][1] = fd[i - 1][1]; }
                fd[0][0] = f_new[0];
                fd[0][1] = f_new[1];
                // out_R(rho-R) leaves the pipeline
                const int orow = rho - R;
                if (own_px && orow >= r0 && orow < r1)
                    *reinterpret_cast<float4*>(xo + ((size_t)orow * W + px) * C + 4 * k) =
                        make_float4(o_new[0].x, o_new[0].y, o_new[1].x, o_new[1].y);
            }
        }
        __syncthreads();
    }

    if (CHECK) {
        // reduce the energy partials over the pixels of the warp, one atomic per (channel, i)
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v = (c & 1) ? en[i][c >> 1].y : en[i][c >> 1].x;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && grp_live) atomicAdd(p.energy + ((size_t)b * C + 4 * k + c) * R + i, (double)v);
            }
    }
}


// one launcher per R, defined in fused_inst_r{2,3,4}.cu
template <int R> int launch_stream_r(int mode, int K, const FusedParams& fp, dim3 grid, cudaStream_t st);

template <int R, int MODE, int K>
int launch_stream_k(const FusedParams& fp, dim3 grid, cudaStream_t st) {
    auto kfn = gap_tv_stream_kernel<R, MODE, true, K>;
    constexpr Smem L = smem_layout(K, fused_groups(K));
    SCIPNP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    kfn<<<grid, fused_threads(K), L.total, st>>>(fp);
    return SCIPNP_OK;
}

template <int R, int MODE>
int launch_stream_mode(int K, const FusedParams& fp, dim3 grid, cudaStream_t st) {
    switch (K) {
#ifndef SCIPNP_FUSED_FAST_BUILD
        case 1: return launch_stream_k<R, MODE, 1>(fp, grid, st);
        case 3: return launch_stream_k<R, MODE, 3>(fp, grid, st);
        case 4: return launch_stream_k<R, MODE, 4>(fp, grid, st);
        case 5: return launch_stream_k<R, MODE, 5>(fp, grid, st);
        case 7: return launch_stream_k<R, MODE, 7>(fp, grid, st);
        case 8: return launch_stream_k<R, MODE, 8>(fp, grid, st);
#endif
        case 2: return launch_stream_k<R, MODE, 2>(fp, grid, st);
        case 6: return launch_stream_k<R, MODE, 6>(fp, grid, st);
    }
    set_error("fused kernel not built for C = %d", 4 * K);
    return SCIPNP_EINVAL;
}

#define SCIPNP_INSTANTIATE_FUSED_R(RR)                                                                 \
    template <> int launch_stream_r<RR>(int mode, int K, const FusedParams& fp, dim3 grid,             \
                                        cudaStream_t st) {                                             \
        if (mode == MODE_GAP_ACC) return launch_stream_mode<RR, MODE_GAP_ACC>(K, fp, grid, st);       \
        return launch_stream_mode<RR, MODE_GAP_PLAIN>(K, fp, grid, st);                                \
    }

}  // namespace fusedk
}  // namespace scipnp
