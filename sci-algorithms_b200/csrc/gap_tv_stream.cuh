// Kernel template of the one-pass fused GAP-TV iteration (see gap_tv_fused.cu for the
// design notes).  Included by the per-R instantiation units fused_inst_r*.cu.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "internal.cuh"

namespace scipnp {
namespace fusedk {

constexpr int RB = 4;             // rows per staged block
constexpr int BOX_BYTES = RB * 32 * 16;   // RB rows x 32 pixels x one 4-channel chunk
// A TMA box of a frame tile is RB rows x 32 pixels x SUBK chunks, pixel-major in shared memory.
// SUBK = K moves whole pixels: a box row is then one contiguous 32*C*4-byte piece of global memory,
// which is what the TMA unit is fast at (measured on the config-5 scene, loads + stores only:
// 0.54 ms per pass against 0.68 ms with 16-byte pieces; 48-byte pieces are no better than 16).
// The price is the bank pattern of the lane-per-pixel LDS.128: conflict-free for odd K, two-way
// for K = 2, 6 (phase A avoids it by swapping chunk pairs on every other group of four lanes),
// worse for K = 4, 8 -- those keep one box per chunk (SUBK = 1, chunk-major, conflict-free).
__host__ __device__ constexpr int fused_subk(int K) { return (K & 3) == 0 ? 1 : K; }
constexpr int kMaxWarps = 8;
constexpr int kSegCost = 12;      // cost of starting a row segment, in rows (see the work split in the kernel)

struct FusedParams {
    const float* x_in; float* x_out;
    const float* y1_in; float* y1_out;
    const float* y; const float* Phi; const float* Phi_sum;
    double* energy;               // [B][C][R] partial sums of d^2 + w*|g|
    float lambda, tv_c, tv_w;     // tv_c = tau / weight
    int B, H, W, C, K, NG, ngroups;  // K = C/4 chunk-warps per pixel group, NG groups per CTA
    int small_tma;                // y / y1 / Phi_sum rows staged by TMA (needs W % 4 == 0), else cp.async
    int phi_batched;
    long long phi_bstride, ps_bstride;   // batch strides (0 when shared)
};

// CASSI (template flag): Phi[h, w, c] = mask2d[h, w - step*c] on a canvas of W columns.  Kept out of
// FusedParams on purpose: ptxas' register allocation of the main kernel is sensitive to that layout.
struct CassiParams {
    const float* mask2d; int step, mask_w;
    // ADMM (MODE_ADMM): multiplier in/out, projection output x (returned by admm_denoise), gamma
    const float* b_in; float* b_out; float* xproj; float gamma;
    int clip01;       // clip the TV output to [0,1] (joint_pnp_sci_algo.py:633)
};

// tensor maps of one launch: x_in and Phi as [rows][W][C] (box RB x 32 x 4*SUBK), y / y1_in / Phi_sum
// as [rows][W]
struct FusedMaps { CUtensorMap x, phi, y, y1, ps; };

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- TMA + mbarrier (inline PTX) ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

__device__ __forceinline__ float fast_sqrt(float v) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float fast_rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }

// shared-memory carve-up (per CTA), all offsets in bytes.  One staging slot holds RB rows:
//   x   tiles [NG][K/SUBK][RB][32][SUBK] float4   (one TMA box per (group, sub-box))
//   Phi tiles likewise
//   y, y1, Phi_sum  [3][NG][RB][32] float
struct Smem {
    int tile_bytes;          // all x (or all Phi) boxes of a slot: NG*K*BOX_BYTES
    int small_off;           // offset of the y/y1/Phi_sum rows inside a slot
    int buf_bytes;           // one slot
    int part_off;            // projection scale lambda*s per (row, pixel): two buffers of [RB][NG][32]
    int part_bytes;
    int bar_off;             // nslot TMA mbarriers + the phase-A mbarrier
    int nslot, nsbuf;        // staging ring depth; buffers of the scale plane
    int total;
};
// Ring depth.  With four slots (and three scale planes) the per-block CTA barrier becomes a split
// arrive/wait on an mbarrier: a warp may then run up to one block ahead of the slowest one.  Used
// when two CTAs with four slots still fit one SM (228 KB, 1 KB reserved per CTA); else three slots
// and a plain __syncthreads per block.
__host__ __device__ constexpr int fused_slots(int K, int NG) {
    const int buf = 2 * NG * K * BOX_BYTES + 3 * NG * RB * 32 * 4;
    const int ctas = NG * K > 6 ? 1 : 2;       // fused_ctas(K), defined below
    return ctas * (4 * buf + 3 * RB * NG * 32 * 4 + 64 + 1024) <= 232448 ? 4 : 3;
}
__host__ __device__ constexpr Smem smem_layout(int K, int NG) {
    Smem s{};
    s.nslot = fused_slots(K, NG);
    s.nsbuf = s.nslot - 1;
    s.tile_bytes = NG * K * BOX_BYTES;
    s.small_off = 2 * s.tile_bytes;
    s.buf_bytes = s.small_off + 3 * NG * RB * 32 * 4;
    s.part_off = s.nslot * s.buf_bytes;
    s.part_bytes = RB * NG * 32 * 4;
    s.bar_off = s.part_off + s.nsbuf * s.part_bytes;
    s.total = s.bar_off + 64;
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

// CTA size per chunk count K = C/4: NG = 6/K pixel groups of K chunk-warps each (<= ~90 KB of
// staging per CTA so that two CTAs share an SM)
__host__ __device__ constexpr int fused_groups(int K) { return 6 / K < 1 ? 1 : 6 / K; }
__host__ __device__ constexpr int fused_threads(int K) { return fused_groups(K) * K * 32; }
// resident CTAs per SM the kernel is compiled for: two CTAs of <= 6 warps (168 registers per
// thread); the 7- and 8-warp CTAs of C = 28, 32 would be squeezed to 128 registers and spill, so
// they get the whole register file and run one per SM
__host__ __device__ constexpr int fused_ctas(int K) { return fused_threads(K) > 192 ? 1 : 2; }

// Register state of one thread: 4 channels as two packed pairs, R pipeline stages.
template <int R>
struct Pipe {
    P2 o_prev[R][2];      // out_i(rho-i-1)
    P2 g1_prev[R][2];     // horizontal difference of out_i at row rho-i-1
    P2 P0[R + 1][2];      // P[i] = p^i(rho-i-1), vertical component; P[0] stays 0
    P2 P1[R + 1][2];      //                       horizontal component
    P2 fd[R][2];          // fd[j] = f(rho-1-j)
    P2 en[R][2];          // energy partials of dual iterations 0..R-1
};

struct StepConst {
    P2 mone2, mtau2, tvc2, one2, wm2;   // wm2 = w on owned pixels (fast path)
    float tvw, own_f, pxin_f;
    int src_right;
    int rs, r0, r1, H;
};

// One pipeline step: f_new = f(rho) enters, out_R(rho-R) is returned in o_new.
// FAST: every row touched by the R stages lies inside the segment and the image, so no
// row masks are needed; energy contributions are masked per lane at the very end.
template <int R, bool CHECK, bool FAST>
__device__ __forceinline__ void pipe_step(Pipe<R>& S, const StepConst& c, int rho, const P2 (&f_new)[2], P2 (&o_new)[2]) {
    o_new[0] = f_new[0];
    o_new[1] = f_new[1];
    P2 pend0[2], pend1[2];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int row_new = rho - i;          // row of o_new = out_i(row_new)
        const int u = row_new - 1;            // row whose dual variable advances
        // general path: masks as multipliers.  The dual variable of a row outside the
        // segment/image or of a pixel outside the image stays 0; g0 = 0 below the last row.
        P2 m2, md2, me2, wm2;
        if (!FAST) {
            m2 = splat(((u >= c.rs) && (u < c.H)) ? c.pxin_f : 0.f);
            md2 = splat(row_new < c.H ? 1.f : 0.f);
            const float me = (u >= c.r0 && u < c.r1) ? 1.f : 0.f;
            me2 = splat(me);
            wm2 = splat(me * c.tvw);
        } else {
            wm2 = c.wm2;
        }
        P2 pi0[2], pi1[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) { pi0[q] = S.P0[i][q]; pi1[q] = S.P1[i][q]; }
        if (i > 0) {
#pragma unroll
            for (int q = 0; q < 2; ++q) { S.P0[i][q] = pend0[q]; S.P1[i][q] = pend1[q]; }
        }
        P2 o_next[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const P2 o_right = shfl_idx2(o_new[q], c.src_right);
            P2 g0 = fma2(S.o_prev[i][q], c.mone2, o_new[q]);
            if (!FAST) g0 = mul2(g0, md2);
            const P2 g1 = S.g1_prev[i][q];
            const P2 nrm = sqrt2(fma2(g0, g0, mul2(g1, g1)));
            P2 r = rcp2(fma2(nrm, c.tvc2, c.one2));
            if (!FAST) r = mul2(r, m2);
            const P2 pn0 = mul2(fma2(g0, c.mtau2, pi0[q]), r);
            const P2 pn1 = mul2(fma2(g1, c.mtau2, pi1[q]), r);
            const P2 p1l = shfl_up2(pn1);
            // D(p^{i+1})(u) = (p0(u-1) - p0(u)) + (p1(u, left) - p1(u))
            const P2 d = add2(fma2(pn0, c.mone2, S.P0[i + 1][q]), fma2(pn1, c.mone2, p1l));
            o_next[q] = add2(S.fd[i][q], d);
            if (CHECK) {
                S.en[i][q] = fma2(nrm, wm2, S.en[i][q]);                              // w*|grad out_i|(u)
                if (i + 1 < R) {
                    if (FAST) S.en[i + 1][q] = fma2(d, d, S.en[i + 1][q]);            // D(p^{i+1})(u)^2
                    else S.en[i + 1][q] = fma2(mul2(d, me2), d, S.en[i + 1][q]);
                }
            }
            S.g1_prev[i][q] = fma2(o_new[q], c.mone2, o_right);
            S.o_prev[i][q] = o_new[q];
            pend0[q] = pn0;
            pend1[q] = pn1;
        }
        o_new[0] = o_next[0];
        o_new[1] = o_next[1];
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) { S.P0[R][q] = pend0[q]; S.P1[R][q] = pend1[q]; }
#pragma unroll
    for (int i = R - 1; i > 0; --i) { S.fd[i][0] = S.fd[i - 1][0]; S.fd[i][1] = S.fd[i - 1][1]; }
    S.fd[0][0] = f_new[0];
    S.fd[0][1] = f_new[1];
}

template <int R, int MODE, bool CHECK, int K, bool CASSI>
__global__ void __launch_bounds__(fused_threads(K), fused_ctas(K))
gap_tv_stream_kernel(const FusedParams p, const __grid_constant__ FusedMaps maps, const CassiParams cp) {
    constexpr int NG = fused_groups(K);
    constexpr int NT = fused_threads(K);
    constexpr Smem L = smem_layout(K, NG);
    constexpr int NSLOT = L.nslot, NSBUF = L.nsbuf;
    constexpr bool SPLIT = NSLOT == 4;           // split per-block barrier (see fused_slots)
    constexpr int SUBK = fused_subk(K), NSUB = K / SUBK;
    constexpr bool SWAP = SUBK > 1 && (SUBK & 1) == 0;    // phase A: chunk pairs swapped on lanes 4..7 of each 8
    // float4 index of chunk kk of (group g2, row j, pixel ln) inside a frame tile
    auto chunk_idx = [](int g2, int kk, int j, int ln) {
        return (((g2 * NSUB + kk / SUBK) * RB + j) * 32 + ln) * SUBK + kk % SUBK;
    };
    constexpr int OWN = 32 - 2 * R;              // owned pixels per group
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = p.W, H = p.H, C = p.C;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int gi = warp / K, k = warp - gi * K;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t bar_base = smem_base + L.bar_off;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NSLOT; ++i) mbar_init(bar_base + i * 8, 1);
        mbar_init(bar_base + NSLOT * 8, NT / 32);          // phase A done: one arrival per warp
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    // ---- balanced work split: the scene is (batch x pixel-group bundles) column strips of H rows;
    //      laid end to end they form `total` row units, and every CTA takes the same number of
    //      consecutive units, i.e. at most a few row segments of neighbouring strips.  No wave
    //      quantisation, and the warm-up rows of a segment are paid once or twice per CTA.
    const int nbundles = (p.ngroups + NG - 1) / NG;
    // Every strip is charged kSegCost extra units in front of its H rows: a CTA whose share crosses a
    // strip boundary starts a second segment there (warm-up and drain rows, a pipeline refill, masked
    // first and last blocks), so it is given that many rows less.
    const int Hv = H + kSegCost;
    const long long total = (long long)p.B * nbundles * Hv;
    const long long per_cta = (total + gridDim.x - 1) / gridDim.x;
    long long unit = (long long)blockIdx.x * per_cta;
    const long long unit_end = min(total, unit + per_cta);
    int gb = 0;                                  // running block counter: ring slot and mbarrier phase
    const uint32_t pa_bar = bar_base + NSLOT * 8;
    // the y / y1 / Phi_sum rows staged by cp.async need the full barrier for visibility
    const bool split = SPLIT && p.small_tma;
#pragma unroll 1
    while (unit < unit_end) {
    const int strip = (int)(unit / Hv);
    const int v0 = (int)(unit - (long long)strip * Hv);                    // position in the charged strip
    const int v1 = (int)min((long long)Hv, v0 + (unit_end - unit));
    unit += v1 - v0;
    const int r0 = max(v0 - kSegCost, 0), r1 = v1 - kSegCost;
    if (r1 <= r0) continue;                                                // only charge units: no rows here
    const int b = strip / nbundles;
    const int group0 = (strip - b * nbundles) * NG;      // first pixel group of this segment
    const int grp = group0 + gi;
    const bool grp_live = grp < p.ngroups;
    const int px = grp * OWN - R + lane;         // this lane's pixel column
    const bool px_in = grp_live && px >= 0 && px < W;
    const bool own_px = px_in && lane >= R && lane < 32 - R;

    const int rs = max(0, r0 - R), rend = r1 + R;       // steps rho in [rs, rend)
    const int load_end = min(H, rend);
    const int nblk = (rend - rs + RB - 1) / RB;
    // steps whose R stages all work on rows inside [r0, r1) and the image
    const int fast_lo = r0 + R, fast_hi = min(r1, H) - 1;

    const size_t frame_b = (size_t)b * H * W * C;        // batch offsets
    const size_t meas_b = (size_t)b * H * W;

    // ---- producer: one lane (of the warp whose turn it is, see vwarp below) programs the TMA unit,
    //      RB rows per block: 2*NG*NSUB boxes of x / Phi (+ 3*NG rows of y, y1, Phi_sum) land in slot
    //      blk % NSLOT and complete on that slot's mbarrier.  Pixels left/right of the image are
    //      zero-filled by the TMA unit; rows past the segment are loaded but never used.
    constexpr uint32_t kTileTx = (CASSI ? 1u : 2u) * NG * K * BOX_BYTES;   // CASSI: no Phi stack to load
    constexpr uint32_t kSmallTx = (MODE == MODE_GAP_ACC ? 3u : 2u) * NG * RB * 32 * 4;
    const int rowc0 = b * H;                              // row coordinate of the batch element
    const int phirow0 = p.phi_batched ? b * H : 0;
    // cp.async fallback for the y / y1 / Phi_sum rows when W % 4 != 0 (TMA needs 16-byte row pitch)
    constexpr int NSMALL = (3 * NG * RB * 32 + NT - 1) / NT;
    // Per-block duties rotate over the warps (virtual warp index vw = warp - rot mod NW): the first
    // NITEM/32 virtual warps run phase A, the last one programs the TMA unit.  With the split
    // barrier the warps may drift by a block, so the extra work averages out instead of making
    // one warp the pace setter.
    constexpr int NW = NT / 32;
    auto vwarp = [&](int rot) { const int v = warp - rot; return v < 0 ? v + NW : v; };
    auto issue = [&](int blk, int rot) {
        if (blk < nblk) {
            const int slot = (gb + blk) % NSLOT;
            const uint32_t dst = smem_base + slot * L.buf_bytes;
            const int row0 = rs + blk * RB;
            if (lane == 0 && vwarp(rot) == NW - 1) {
                const uint32_t bar = bar_base + slot * 8;
                mbar_expect_tx(bar, kTileTx + (p.small_tma ? kSmallTx : 0u));
#pragma unroll
                for (int g2 = 0; g2 < NG; ++g2) {
                    const int px0 = (group0 + g2) * OWN - R;
#pragma unroll
                    for (int sb = 0; sb < NSUB; ++sb) {
                        const uint32_t d = dst + (g2 * NSUB + sb) * SUBK * BOX_BYTES;
                        tma_load_3d(d, &maps.x, 4 * SUBK * sb, px0, rowc0 + row0, bar);
                        if constexpr (!CASSI) tma_load_3d(d + L.tile_bytes, &maps.phi, 4 * SUBK * sb, px0, phirow0 + row0, bar);
                    }
                    if (p.small_tma) {
                        const uint32_t ds = dst + L.small_off + g2 * RB * 128;
                        tma_load_2d(ds, &maps.y, px0, rowc0 + row0, bar);
                        if (MODE == MODE_GAP_ACC) tma_load_2d(ds + NG * RB * 128, &maps.y1, px0, rowc0 + row0, bar);
                        tma_load_2d(ds + 2 * NG * RB * 128, &maps.ps, px0, phirow0 + row0, bar);
                    }
                }
            }
            if (!p.small_tma) {
#pragma unroll
                for (int t = 0; t < NSMALL; ++t) {
                    const int idx = tid + t * NT;                 // [which][g2][j][lane]
                    const int which = idx / (NG * RB * 32);
                    const int rem = idx - which * NG * RB * 32;
                    const int g2 = rem / (RB * 32), j = (rem >> 5) % RB, ln = rem & 31;
                    const int gpx = (group0 + g2) * OWN - R + ln;
                    const int row = row0 + j;
                    const bool ok = which < 3 && gpx >= 0 && gpx < W && row < H && (group0 + g2) < p.ngroups &&
                                    !(which == 1 && MODE != MODE_GAP_ACC);
                    if (which < 3) {
                        const float* base = which == 0 ? p.y + meas_b
                                          : which == 1 ? (MODE == MODE_GAP_ACC ? p.y1_in + meas_b : p.y + meas_b)
                                                       : p.Phi_sum + (size_t)b * p.ps_bstride;
                        cp_async4(dst + L.small_off + idx * 4, base + (ok ? (size_t)row * W + gpx : 0), ok ? 4 : 0);
                    }
                }
            }
        }
        if (!p.small_tma) cp_async_commit();
    };
    auto wait_block = [&](int blk) {       // tiles (and TMA-staged rows) of block blk have landed
        if (blk < nblk) mbar_wait(bar_base + ((gb + blk) % NSLOT) * 8, ((gb + blk) / NSLOT) & 1);
    };
    // Phi of chunk kk (4 channels) at (row, gpx): from the staged tile, or -- CASSI -- the 2-D coded
    // aperture read at the per-band offset (the dispersion shift is an index offset, no stack in HBM)
    auto load_phi = [&](const float4* tx, int row, int gpx, bool in, int kk) -> float4 {
        if constexpr (!CASSI) {
            return tx[L.tile_bytes / 16];          // same slot layout as x, one tile further
        } else {
            float v[4];
            const float* mrow = cp.mask2d + ((size_t)(p.phi_batched ? b : 0) * H + min(row, H - 1)) * cp.mask_w;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int wm = gpx - cp.step * (4 * kk + i);
                v[i] = (in && row < H && wm >= 0 && wm < cp.mask_w) ? __ldg(mrow + wm) : 0.f;
            }
            return make_float4(v[0], v[1], v[2], v[3]);
        }
    };
    // Phase A of block `blk`: the Euclidean projection is pointwise in (row, px), so one thread per
    // (row, pixel) of the block forms the whole dot product over the C channels from the staged
    // tiles, updates y1 and leaves the scale lambda*s in shared memory for the K chunk-warps.
    const float lam = p.lambda;
    float* y1o = (MODE == MODE_GAP_ACC) ? p.y1_out + meas_b : nullptr;
    constexpr int NITEM = RB * NG * 32;
    auto phase_a = [&](int blk, int rot) {
        if (blk >= nblk) return;
        const int vtid = vwarp(rot) * 32 + lane;
        const unsigned char* buf = smem_raw + ((gb + blk) % NSLOT) * L.buf_bytes;
        float* sbuf = reinterpret_cast<float*>(smem_raw + L.part_off + ((gb + blk) % NSBUF) * L.part_bytes);
#pragma unroll
        for (int it0 = 0; it0 < NITEM; it0 += NT) {
            const int it = it0 + vtid;                      // [j][g2][ln]
            if (NITEM % NT != 0 && it >= NITEM) break;
            const int ln = it & 31, g2 = (it >> 5) % NG, j = it / (32 * NG);
            const int row = rs + blk * RB + j;
            const int gpx = (group0 + g2) * OWN - R + ln;
            const bool in = (group0 + g2) < p.ngroups && gpx >= 0 && gpx < W;
            const float4* tile = reinterpret_cast<const float4*>(buf);
            // SWAP: this lane visits the chunks as 1,0,3,2,.. so that a quarter-warp touches 8 bank groups
            const int sw = SWAP ? (ln >> 2) & 1 : 0;
            float acc = 0.f;
#pragma unroll
            for (int k0 = 0; k0 < K; ++k0) {
                const int kk = SWAP ? k0 + ((k0 & 1) ? -sw : sw) : k0;
                const float4* tx = tile + chunk_idx(g2, k0, j, ln) + (kk - k0);
                float4 xv = tx[0];
                const float4 pv = load_phi(tx, row, gpx, in, kk);
                if constexpr (MODE == MODE_ADMM) {          // the projection acts on u = theta + b
                    if (in && row < H) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(cp.b_in + frame_b + ((size_t)row * W + gpx) * C + 4 * kk));
                        xv.x += bv.x; xv.y += bv.y; xv.z += bv.z; xv.w += bv.w;
                    }
                }
                acc = fmaf(xv.x, pv.x, acc);
                acc = fmaf(xv.y, pv.y, acc);
                acc = fmaf(xv.z, pv.z, acc);
                acc = fmaf(xv.w, pv.w, acc);
            }
            const float* sm = reinterpret_cast<const float*>(buf + L.small_off) + (g2 * RB + j) * 32 + ln;
            const float yv = sm[0];
            const float psv = sm[2 * NG * RB * 32];
            float sv;
            if (MODE == MODE_GAP_ACC) {
                const float y1n = sm[NG * RB * 32] + (yv - acc);
                if (in && ln >= R && ln < 32 - R && row >= r0 && row < r1) y1o[(size_t)row * W + gpx] = y1n;
                sv = (y1n - acc) * fast_rcp(psv);
            } else if (MODE == MODE_ADMM) {
                sv = (yv - acc) * fast_rcp(psv + cp.gamma);      // pnp_sci_algo.py:809
            } else {
                sv = (yv - acc) * fast_rcp(psv);
            }
            sbuf[it] = in ? sv * lam : 0.f;
        }
    };

    Pipe<R> S;
    const P2 zero2 = splat(0.f);
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int q = 0; q < 2; ++q) { S.o_prev[i][q] = zero2; S.g1_prev[i][q] = zero2; S.fd[i][q] = zero2; S.en[i][q] = zero2; }
#pragma unroll
    for (int i = 0; i <= R; ++i)
#pragma unroll
        for (int q = 0; q < 2; ++q) { S.P0[i][q] = zero2; S.P1[i][q] = zero2; }

    StepConst sc;
    sc.mone2 = splat(-1.f); sc.mtau2 = splat(-0.25f); sc.tvc2 = splat(p.tv_c); sc.one2 = splat(1.f);
    sc.tvw = p.tv_w; sc.wm2 = splat(p.tv_w);
    sc.own_f = own_px ? 1.f : 0.f; sc.pxin_f = px_in ? 1.f : 0.f;
    // the right neighbour of the last image column is the pixel itself (g1 = 0 there)
    // (and of any lane outside the image, whose state then stays identically zero)
    sc.src_right = (px_in && px < W - 1 && lane < 31) ? lane + 1 : lane;
    sc.rs = rs; sc.r0 = r0; sc.r1 = r1; sc.H = H;
    // Output cursors of this thread: element offsets of (row 0, px, channel 4k) in a frame and of
    // (row 0, px) in a measurement plane.  Rows are added per block (FAST: once per RB rows), so the
    // steady state carries no 64-bit multiplies.  Never dereferenced for pixels outside the image.
    // (32-bit: launch_fused keeps H*W*C below 2^31; the batch offset sits in the base pointers)
    const int xstride = W * C;
    const int xoff0 = px * C + 4 * k;
    float* const xo_b = p.x_out + frame_b;

    // stage 0 of step rho (row j of the block in `buf`): f(rho) = x + (lambda*s) * Phi for this warp's chunk
    // ROWS_OK: the caller guarantees r0 <= rho < r1 (no row predicate on the stores)
    auto project_row = [&](const unsigned char* buf, const float* sbuf, int j, int rho, int xoff, auto rows_ok, P2 (&f_new)[2]) {
        [[maybe_unused]] constexpr bool ROWS_OK = decltype(rows_ok)::value;      // used by the ADMM store only
        const float4* tx = reinterpret_cast<const float4*>(buf) + chunk_idx(gi, k, j, lane);
        const float4 xv = tx[0], pv = load_phi(tx, rho, px, px_in, k);
        const P2 s2 = splat(sbuf[(j * NG + gi) * 32 + lane]);
        f_new[0] = fma2(s2, make_float2(pv.x, pv.y), make_float2(xv.x, xv.y));
        f_new[1] = fma2(s2, make_float2(pv.z, pv.w), make_float2(xv.z, xv.w));
        if constexpr (MODE == MODE_ADMM) {
            // f = x - b = theta + lambda*s*Phi is the TV input; x = f + b is what admm_denoise returns
            if (own_px && (ROWS_OK || (rho >= r0 && rho < r1))) {
                const size_t o = frame_b + (long long)xoff;
                const float4 bv = __ldg(reinterpret_cast<const float4*>(cp.b_in + o));
                *reinterpret_cast<float4*>(cp.xproj + o) =
                    make_float4(f_new[0].x + bv.x, f_new[0].y + bv.y, f_new[1].x + bv.z, f_new[1].y + bv.w);
            }
        }
    };
    // f_out = f(orow), the value that left the f delay line in this step: the ADMM multiplier update
    // b - (x - theta_new) equals theta_new - f  (pnp_sci_algo.py:836 with x = f + b)
    auto store_row = [&](int orow, int xoff, auto rows_ok, const P2 (&oo)[2], const P2 (&f_out)[2]) {
        constexpr bool ROWS_OK = decltype(rows_ok)::value;
        if (own_px && (ROWS_OK || (orow >= r0 && orow < r1))) {
            P2 o[2] = {oo[0], oo[1]};
            if constexpr (MODE == MODE_ADMM) {          // the joint ADMM variant clips theta
                if (cp.clip01) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) { o[q].x = fminf(fmaxf(o[q].x, 0.f), 1.f); o[q].y = fminf(fmaxf(o[q].y, 0.f), 1.f); }
                }
            }
            *reinterpret_cast<float4*>(xo_b + xoff) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
            if constexpr (MODE == MODE_ADMM)
                *reinterpret_cast<float4*>(cp.b_out + frame_b + (long long)xoff) =
                    make_float4(o[0].x - f_out[0].x, o[0].y - f_out[0].y, o[1].x - f_out[1].x, o[1].y - f_out[1].y);
        }
    };

    // phase A of a block is done (scale plane written) by this warp: one arrival per warp
    auto pa_arrive = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive(pa_bar);
    };
    {
    const int rot = gb % NW;
    issue(0, rot);
    issue(1, rot);
    wait_block(0);
    if (!p.small_tma) { cp_async_wait<0>(); __syncthreads(); }   // phase A reads the y / y1 / Phi_sum rows
    phase_a(0, rot);
    }
    if (split) pa_arrive();
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
        // Once per block: the scale plane of block blk is complete.  Split form: every warp arrived
        // after its phase A of block blk, i.e. before it started the rows of block blk-1, so passing
        // the wait also means everybody is done with block blk-2 -- whose slot (of four) the next TMA
        // loads overwrite and whose scale plane (of three) phase A of block blk+1 overwrites.
        // Barrier form (three slots): everybody is done with block blk-1.
        if (split) {
            mbar_wait(pa_bar, (gb + blk) & 1);       // one wait per block since the kernel started
        } else {
            if (!p.small_tma) cp_async_wait<0>();
            __syncthreads();
        }
        const int rot = (gb + blk + 1) % NW;
        issue(blk + 2, rot);
        wait_block(blk + 1);
        phase_a(blk + 1, rot);
        if (split && blk + 1 < nblk) pa_arrive();
        const unsigned char* buf = smem_raw + ((gb + blk) % NSLOT) * L.buf_bytes;
        const float* sbuf = reinterpret_cast<const float*>(smem_raw + L.part_off + ((gb + blk) % NSBUF) * L.part_bytes);
        const int rho0 = rs + blk * RB;
        const int xoff_in = xoff0 + rho0 * xstride;       // (rho0, px, 4k)
        const int xoff_out = xoff_in - R * xstride;        // (rho0 - R, px, 4k)
        if (rho0 >= fast_lo && rho0 + RB - 1 <= fast_hi) {
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int rho = rho0 + j;
                P2 f_new[2], o_new[2];
                project_row(buf, sbuf, j, rho, xoff_in + j * xstride, std::true_type{}, f_new);
                const P2 f_out[2] = {S.fd[R - 1][0], S.fd[R - 1][1]};
                pipe_step<R, CHECK, true>(S, sc, rho, f_new, o_new);
                store_row(rho - R, xoff_out + j * xstride, std::true_type{}, o_new, f_out);
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < RB; ++j) {
                const int rho = rho0 + j;
                if (rho < rend) {
                    P2 f_new[2] = {zero2, zero2}, o_new[2];
                    if (rho < H) project_row(buf, sbuf, j, rho, xoff_in + j * xstride, std::false_type{}, f_new);
                    const P2 f_out[2] = {S.fd[R - 1][0], S.fd[R - 1][1]};
                    pipe_step<R, CHECK, false>(S, sc, rho, f_new, o_new);
                    store_row(rho - R, xoff_out + j * xstride, std::false_type{}, o_new, f_out);
                }
            }
        }
    }

    if (CHECK) {
        // reduce the energy partials over the owned pixels of the warp, one atomic per (channel, i)
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float v = ((ch & 1) ? S.en[i][ch >> 1].y : S.en[i][ch >> 1].x) * sc.own_f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && grp_live) atomicAdd(p.energy + ((size_t)b * C + 4 * k + ch) * R + i, (double)v);
            }
    }
    // segment boundary: every warp is done with the staging ring and the scale planes
    gb += nblk;
    if (!p.small_tma) cp_async_wait<0>();
    __syncthreads();
    }   // while (unit < unit_end)
}

// one launcher per R, defined in fused_inst_r{2,3,4}.cu
template <int R> int launch_stream_r(int mode, int K, const FusedParams& fp, const FusedMaps& maps, const CassiParams& cp,
                                     dim3 grid, cudaStream_t st);
// CASSI variants (index-offset mask), built for R = 4 only: fused_inst_r4c.cu
int launch_stream_cassi_r4(int mode, int K, const FusedParams& fp, const FusedMaps& maps, const CassiParams& cp,
                           dim3 grid, cudaStream_t st);

template <int R, int MODE, int K, bool CASSI = false>
int launch_stream_k(FusedParams fp, const FusedMaps& maps, dim3 grid, cudaStream_t st, CassiParams cp = CassiParams{}) {
    auto kfn = gap_tv_stream_kernel<R, MODE, true, K, CASSI>;
    constexpr Smem L = smem_layout(K, fused_groups(K));
    static int ctas_per_sm = 0;            // resident CTAs of this instance, queried once
    if (!ctas_per_sm) {
        SCIPNP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        int n = 0;
        SCIPNP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kfn, fused_threads(K), L.total));
        ctas_per_sm = n > 0 ? n : 1;
    }
    // One resident wave: every CTA takes an equal share of the (batch x bundle x row) units
    // (see the kernel).  Large scenes fill all slots; small ones get segments of at least 8 rows.
    const long long slots = (long long)ctas_per_sm * num_sms();
    const long long nbundles = (fp.ngroups + fused_groups(K) - 1) / fused_groups(K);
    const long long total = (long long)fp.B * nbundles * (fp.H + kSegCost);
    long long n = total / 8;      // small scenes are latency-bound: short segments, more CTAs
    if (n > slots) n = slots;
    if (n < 1) n = 1;
    grid = dim3((unsigned)n, 1, 1);
    kfn<<<grid, fused_threads(K), L.total, st>>>(fp, maps, cp);
    return SCIPNP_OK;
}

template <int R, int MODE, bool CASSI = false>
int launch_stream_mode(int K, const FusedParams& fp, const FusedMaps& maps, dim3 grid, cudaStream_t st,
                       CassiParams cp = CassiParams{}) {
    switch (K) {
#ifndef SCIPNP_FUSED_FAST_BUILD
        case 1: return launch_stream_k<R, MODE, 1, CASSI>(fp, maps, grid, st, cp);
        case 3: return launch_stream_k<R, MODE, 3, CASSI>(fp, maps, grid, st, cp);
        case 4: return launch_stream_k<R, MODE, 4, CASSI>(fp, maps, grid, st, cp);
        case 5: return launch_stream_k<R, MODE, 5, CASSI>(fp, maps, grid, st, cp);
        case 7: return launch_stream_k<R, MODE, 7, CASSI>(fp, maps, grid, st, cp);
        case 8: return launch_stream_k<R, MODE, 8, CASSI>(fp, maps, grid, st, cp);
#endif
        case 2: return launch_stream_k<R, MODE, 2, CASSI>(fp, maps, grid, st, cp);
        case 6: return launch_stream_k<R, MODE, 6, CASSI>(fp, maps, grid, st, cp);
    }
    set_error("fused kernel not built for C = %d", 4 * K);
    return SCIPNP_EINVAL;
}

#define SCIPNP_INSTANTIATE_FUSED_R(RR)                                                                 \
    template <> int launch_stream_r<RR>(int mode, int K, const FusedParams& fp, const FusedMaps& maps, \
                                        const CassiParams& cp, dim3 grid, cudaStream_t st) {           \
        if (mode == MODE_GAP_ACC) return launch_stream_mode<RR, MODE_GAP_ACC>(K, fp, maps, grid, st, cp);   \
        if (mode == MODE_ADMM) return launch_stream_mode<RR, MODE_ADMM>(K, fp, maps, grid, st, cp);         \
        return launch_stream_mode<RR, MODE_GAP_PLAIN>(K, fp, maps, grid, st, cp);                          \
    }

}  // namespace fusedk
}  // namespace scipnp
