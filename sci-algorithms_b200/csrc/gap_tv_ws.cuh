// Warp-specialised one-pass GAP-TV iteration (second-generation fused kernel, sm_100a).
//
// Same mathematics and the same HBM traffic as gap_tv_stream.cuh (one launch = one outer iteration
// of pnp_sci_algo.py:640-650 with denoiser='tv'), different mapping onto the SM:
//
//   * The CTA is split by role.  PRODUCER warps turn TMA-staged rows of x / Phi / y / y1 / Phi_sum
//     into the TV input f = x + lambda*s*Phi (one thread per (row, pixel): the whole dot product over
//     the C channels, the y1 update, the scale, f for every channel) and leave f in shared memory in a
//     channel-pair-major layout.  CONSUMER warps run the Chambolle pipeline on f and never touch the
//     projection.  One TMA warp: lane 0 issues every load, lane 1 every store.  The roles are coupled only
//     through mbarrier rings (raw tiles, f tiles, output tiles), so no warp waits for a CTA-wide barrier.
//   * A consumer lane owns two horizontally adjacent pixels x one channel pair (4 values, two packed
//     float2 chains): a warp covers 64 pixels of one channel pair.  Half of the horizontal neighbours
//     are in the thread's own registers (4 SHFL per dual iteration instead of 8), and the halo of a
//     group is 4 pixels = 2 lanes per side, i.e. up to 56 of 64 pixels are owned (24 of 32 before).
//   * Output rows are assembled in a chunk-major shared-memory tile ([C/4][own][4]) and written with TMA
//     tensor stores of exactly the owned pixels; y1 is written by the producers (coalesced).
//   * One CTA per SM.  The scene is (batch x column strips) laid end to end in ticks (a row of an interior strip 16, a
//     row of the first / last strip 17: their blocks carry pixel masks; every strip charged 8 rows for starting a row
//     segment) and every CTA takes the same number of consecutive ticks (WsSegIter): one to three row segments per
//     CTA, no wave quantisation.  With several measurements whose groups do not fill the strips, the groups of all
//     measurements are laid end to end (ws_slot).  The two TMA lanes run alone in their warp: whatever does not depend
//     on the block is computed once per segment.
//   * Modes: accelerated / plain GAP, ADMM (the multiplier staged by TMA where shared memory allows, else read by the
//     projection threads; b_new = theta_new - f written by the consumers), the TV denoiser alone.  Ring depths per
//     channel count (ws_smem): 2-3 raw slots, 2-4 f slots, 2 output slots; 8 consumer warps for C <= 8, else 12.
//   * Row-tiled multi-GPU mode: output row window, seam rows stored a second time into the neighbours' buffers (TMA
//     stores / plain peer stores over NVLink), flags raised by the last CTA and awaited by the loader lane.
//
// Dual-iteration pipeline of a consumer (row streaming, one-row lag per dual iteration), per step t:
//   stage i (0..R-1) receives out_i(t-i), advances the dual variable of row u = t-i-1 and emits
//   out_{i+1}(u); out_R(t-R) leaves the pipeline.  State per stage: out_i(u), p^{i+1}(u-1) (both
//   components: the vertical one closes the divergence of this stage, both feed stage i+1), the
//   horizontal difference of the pair's right pixel, and the f delay line.
#pragma once
#include "gap_tv_stream.cuh"

namespace scipnp {
namespace wsk {

using namespace fusedk;     // PTX helpers, P2 arithmetic

// build-time knobs (tools/build_exp.sh)
#ifndef WS_RCP
#define WS_RCP 0          // 0: one MUFU.RCP per value, 1: one per pixel pair, 2: one per four values
#endif
#ifndef WS_CREGS
#define WS_CREGS 0        // > 0: setmaxnreg -- consumers WS_CREGS registers, producers WS_PREGS (12*C + 4*P <= 2048)
#define WS_PREGS 0
#endif
#ifndef WS_GEN_UNROLL
#define WS_GEN_UNROLL 0   // 1: the masked (first / last) blocks of a row segment are unrolled like the steady-state ones
#endif
#ifndef WS_PATH3
#define WS_PATH3 1        // 1: first / last blocks of a row segment whose rows all exist take an unrolled path with energy masks only
#endif
#ifndef WS_PROD_FIRST
#define WS_PROD_FIRST 0   // 1: the producer-side warps take warp ids 0..3 and the consumers 4..: the arbiter prefers high warp
#endif                    // ids, so the consumers (the bottleneck) win ties and the producers fill the gaps
#ifndef WS_BLK_UNROLL
#define WS_BLK_UNROLL 1   // unroll factor of the consumers' block loop (2: +0 %, not used)
#endif
#ifndef WS_SANITIZE
#define WS_SANITIZE 0     // 1: every lane arrives on the mbarriers (counts x 32) instead of one elected lane behind a __syncwarp:
#endif                    // same protocol, but visible thread by thread to compute-sanitizer's racecheck (tools/sanitize.sh)
#ifndef WS_EXP
#define WS_EXP 0          // timing experiments (bit flags; results are meaningless): 1 independent stages, 2 no out-tile
#endif                    // stores, 4 no energies, 8 no shuffles

constexpr int WRB = 4;            // rows per staged block
constexpr int GW = 64;            // pixels per group tile (lane = 2 pixels)
constexpr int HALO = 4;           // halo pixels per side of a group (2 lanes)
constexpr int OWN_MAX = GW - 2 * HALO;
constexpr int NPROD = 4;          // producer-side warps: one TMA warp (loader lane + storer lane) and NCOMP projection warps
constexpr int NCOMP = NPROD - 1;
constexpr int kNrawMax = 3;       // ring depth of the raw (TMA-staged) tiles: ws_nraw(Q, bstage), 2 or 3
constexpr int NOUT = 2;           // ring depth of the output tiles

struct WsParams {
    float* y1_out;                // accelerated GAP: y1 written by the producers
    // ADMM (MODE_ADMM): x_in = theta; the multiplier b is read by the projection threads straight from global memory
    // (L1/L2, no staging: the raw ring has no room for a third frame tile) and b_new = theta_new - f is written by the
    // consumers; x = f + b only where the caller wants it (xproj_out, may be null)
    const float* b_in; float* b_out; float* xproj_out;
    float gamma;
    double* energy;               // [B][C][R] partial sums of d^2 + w*|g|
    int* flag;                    // early-stop flag (may be null: no check)
    unsigned* ticket;             // last-CTA detection for the in-kernel energy check
    double tv_eps;
    float lambda, tv_c, tv_w;     // tv_c = tau / weight
    int B, H, W, C;
    int pack;                     // 1: the groups of all batch elements are laid end to end and a strip takes NGRP consecutive
                                  // ones, whatever element they belong to (ngroups % NGRP != 0 would leave group slots empty)
    int own;                      // owned pixels per group (multiple of 4, <= OWN_MAX)
    int ngroups;                  // ceil(W / own)
    int nstrips;                  // ceil(ngroups / NGRP): column strips
    int phi_batched;
    long long* prof;              // optional [grid][warps][4] cycle counters (SCIPNP_WS_PROF), else null
    // Output row window [out_lo, out_hi) inside the H local rows (the whole scene: [0, H)).  A rank of the row-tiled
    // multi-GPU mode produces its owned rows only; the rows outside the window are halo rows, input only.
    int out_lo, out_hi;
    int seg_cost;                 // rows a strip is charged in the work split for starting a row segment (kWsSegCost)
    int edge_cost;                // extra weight of a row of the first / last strip of a scene, in sixteenths of a row
    int use_pdl;                  // host side only: launch with programmatic stream serialization (short kernels)
    double* energy_log;           // this launch's [B][C][R] energies are also left here (null: not kept)
    // ---- halo push (row-tiled mode, one exchange per iteration, no exchange kernel): the R owned rows next to a
    // seam are stored a second time, into the neighbour's halo rows of ITS output buffers (CUDA-IPC mapped, NVLink);
    // the last CTA raises the neighbours' flags to sig_epoch; the loader lane waits for my flags to reach wait_epoch
    // before it touches a block that holds halo rows or the owned rows next to them.
    float* y1_up; float* y1_dn;   // the neighbours' y1_out, shifted so that my local (row, px) index applies
    float* b_up; float* b_dn;     // ADMM: the neighbours' b_out, shifted likewise
    int up_shift, dn_shift;       // my local row + shift = the neighbour's local row (TMA store coordinate)
    const int* wait_up; const int* wait_dn;
    int* sig_up; int* sig_dn;
    int wait_epoch, sig_epoch;
    int* timeout_flag;
};

struct WsMaps { CUtensorMap x, phi, y, y1, ps, out, out_up, out_dn, b; };

// Group slot gi of strip `strip` (of batch element b when the groups are not packed): batch element, group, live?
struct WsSlot { int b, grp; bool live; };
__device__ __forceinline__ WsSlot ws_slot(const WsParams& p, int b, int strip, int gi, int ngrp) {
    if (!p.pack) {
        const int grp = strip * ngrp + gi;
        return WsSlot{b, grp, grp < p.ngroups};
    }
    const int G = strip * ngrp + gi, bb = G / p.ngroups;
    if (bb >= p.B) return WsSlot{0, p.ngroups, false};            // past the last element: a dead slot (loads zero-fill)
    return WsSlot{bb, G - bb * p.ngroups, true};
}

// pixel groups per CTA for Q = C/2 channel pairs: about 12 consumer warps -- but 8 for C = 4 and C = 8: with 12 their
// staging planes (three per group) leave room for two f slots only, and since an f slot is handed back one block late
// the projection warps and the consumers then take turns instead of overlapping (measured at 28x256x256x8: both
// sides waiting half of the time)
__host__ __device__ constexpr int ws_groups(int Q) { return Q <= 4 ? 8 / Q : (12 / Q < 1 ? 1 : 12 / Q); }
__host__ __device__ constexpr int ws_consumers(int Q) { return ws_groups(Q) * Q; }
__host__ __device__ constexpr int ws_threads(int Q) { return (ws_consumers(Q) + NPROD) * 32; }

struct WsSmem {
    int nraw;           // raw ring depth
    int x_bytes;        // x tiles of one raw slot: [NGRP][WRB][GW][C]
    int small_off;      // y / y1 / Phi_sum rows inside a raw slot: [3][NGRP][WRB][GW]
    int raw_bytes;      // one raw slot
    int f_off, f_bytes, nf;
    int out_off, out_sub, out_bytes;   // out_sub: bytes of one (row, group) sub-tile (128-byte multiple)
    int bar_off;
    int total;
};
// `bstage`: a third frame tile per raw slot (ADMM: the multiplier b staged by TMA next to theta and Phi)
__host__ __device__ constexpr WsSmem ws_smem(int Q, bool bstage = false) {
    WsSmem s{};
    const int NGRP = ws_groups(Q), C = 2 * Q;
    s.x_bytes = NGRP * WRB * GW * C * 4;
    s.small_off = (bstage ? 3 : 2) * s.x_bytes;
    s.raw_bytes = s.small_off + 3 * NGRP * WRB * GW * 4;
    s.f_bytes = WRB * NGRP * Q * GW * 2 * 4;
    s.out_sub = (Q / 2) * OWN_MAX * 16;
    s.out_bytes = WRB * NGRP * s.out_sub;
    // Three raw slots where four f slots still fit beside them (C = 4, 8 without a staged multiplier: with eight
    // consumer warps a block of four rows is consumed faster than a TMA round trip, and two slots starve the
    // projection warps -- measured: they waited for data half of the time), else two.
    const int lim = 232448 - 1024;
    const int fix3 = 3 * s.raw_bytes + NOUT * s.out_bytes + 256;
    s.nraw = (fix3 + 4 * s.f_bytes <= lim) ? 3 : 2;
    s.f_off = s.nraw * s.raw_bytes;
    const int fixed = s.f_off + NOUT * s.out_bytes + 256;
    s.nf = (fixed + 4 * s.f_bytes <= lim) ? 4 : (fixed + 3 * s.f_bytes <= lim) ? 3 : 2;
    s.out_off = s.f_off + s.nf * s.f_bytes;
    s.bar_off = s.out_off + NOUT * s.out_bytes;
    s.total = s.bar_off + 256;
    return s;
}
// ADMM: is there room to stage the multiplier as well and still keep three f slots?  (C = 4, 8, 16, 20: yes;
// C = 12, 24: no -- there the projection threads read it from global memory)
__host__ __device__ constexpr bool ws_bstage(int Q) {
    return ws_smem(Q, true).nf >= 3 && ws_smem(Q, true).total <= 232448 - 1024;
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// non-blocking poll (test_wait returns at once; try_wait may suspend the thread for a system-dependent time)
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes or about `ns`
// nanoseconds have passed (no issue slots burnt while waiting)
__device__ __forceinline__ bool mbar_try_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// spin until a neighbour's flag reached `need` (8 s time-out, reported through *timeout), then order the TMA
// (async proxy) reads that follow behind the acquire
static __device__ __noinline__ void wait_peer_flag(const int* f, int need, int* timeout) {
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < need) {
        __nanosleep(64);
        if (clock64() - t0 > 16000000000LL) { if (timeout) atomicExch(timeout, 1); break; }
    }
    asm volatile("fence.proxy.async;\n" ::: "memory");
}

__device__ __forceinline__ P2 shfl_dn2(P2 a) {
    return make_float2(__shfl_down_sync(0xffffffffu, a.x, 1), __shfl_down_sync(0xffffffffu, a.y, 1));
}

// Register state of a consumer lane: [.][0] = left pixel (A), [.][1] = right pixel (B); a P2 holds the two
// channels of the pair.
template <int R>
struct WsPipe {
    P2 o_prev[R][2];     // out_i(u)
    P2 g1b[R];           // horizontal difference of out_i(u) at the right pixel (its neighbour is a SHFL)
    P2 P0[R][2];         // p^{i+1}(u-1), vertical component
    P2 P1[R][2];         //               horizontal component
    P2 fd[R - 1][2];     // fd[j] = f(t-1-j), j < R-1; the last stage's f(t-R) is read back from the f ring
    P2 en[R];            // energy partials of dual iterations 0..R-1 (both pixels)
};

struct WsConst {
    P2 mone2, mtau2, tvc2, one2, w2;
    float tvw, pair_in, right_in;    // pair inside the image; pixel right of B inside the image
    int rs, r0, r1, H;
};

// One pipeline step: f_new = f(t) enters, out_R(t-R) is returned in o_out.  f_old = f(t-R) closes the last stage
// (out_R = f + D p^R); it comes from shared memory, so that the register delay line holds R-1 rows: together with
// f(t) that is R = WRB live rows, a period the unrolled block of WRB rows maps onto fixed registers without moves.
// PATH 1 (fast): every row touched lies inside [r0, r1) and the image and the whole group lies inside the image.
// PATH 2 (edge): the rows as in 1, but the group hangs over the left or right image edge: pixel masks only.
// PATH 0 (general): row masks and pixel masks.
// PATH 3 (energy window): rows and pixels as in 1 -- every row touched exists and lies inside the image -- but rows outside
//        [r0, r1) (warm-up rows above the segment, drain rows below it) must stay out of the energy sums.
template <int R, int PATH>
__device__ __forceinline__ void ws_step(WsPipe<R>& S, const WsConst& c, int t, const P2 (&f_new)[2], const P2 (&f_old)[2],
                                        P2 (&o_out)[2]) {
    constexpr bool FAST = PATH != 0;          // no row masks
    constexpr bool PXM = PATH == 0 || PATH == 2;   // pixel masks
    constexpr bool EM = PATH == 3;            // energy masks only
    P2 o_new[2] = {f_new[0], f_new[1]};
    P2 o_last[2] = {f_new[0], f_new[1]};
    P2 pi0[2], pi1[2];
    const P2 z = splat(0.f);
    pi0[0] = z; pi0[1] = z; pi1[0] = z; pi1[1] = z;              // p^0 = 0
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int row_new = t - i, u = row_new - 1;
        P2 m2, md2, me2, wm2;
        if (FAST && PXM) m2 = splat(c.pair_in);
        if (!FAST) {
            m2 = splat(((u >= c.rs) && (u < c.H)) ? c.pair_in : 0.f);
            md2 = splat(row_new < c.H ? 1.f : 0.f);
            const float me = (u >= c.r0 && u < c.r1) ? 1.f : 0.f;
            me2 = splat(me);
            wm2 = splat(me * c.tvw);
        } else if (EM) {
            const float me = (u >= c.r0 && u < c.r1) ? 1.f : 0.f;
            me2 = splat(me);
            wm2 = splat(me * c.tvw);
        } else {
            wm2 = c.w2;
        }
        if (FAST && (WS_EXP & 1)) { o_new[0] = f_new[0]; o_new[1] = f_new[1]; }
        // right neighbour of B in the new row: the next lane's A
        const P2 o_rb = (WS_EXP & 8) ? o_new[0] : shfl_dn2(o_new[0]);
        P2 g1[2], g0[2];
        g1[0] = fma2(S.o_prev[i][0], c.mone2, S.o_prev[i][1]);     // out_i(u, B) - out_i(u, A)
        g1[1] = S.g1b[i];
        P2 pn0[2], pn1[2], nrm[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            g0[q] = fma2(S.o_prev[i][q], c.mone2, o_new[q]);
            if (!FAST) g0[q] = mul2(g0[q], md2);
            nrm[q] = sqrt2(fma2(g0[q], g0[q], mul2(g1[q], g1[q])));
        }
        // r = 1 / (1 + (tau/w)|g|) for the four values.  The XU pipe (MUFU: 4 lanes per clock and scheduler) is one of
        // the kernel's three co-limiters, so the reciprocals of pixel A and pixel B share one MUFU.RCP:
        // 1/a = b * (1/(ab)), 1/b = a * (1/(ab)).  The denominators are >= 1; ab stays finite up to 1.8e19 each.
        P2 r[2];
        {
            const P2 den0 = fma2(nrm[0], c.tvc2, c.one2), den1 = fma2(nrm[1], c.tvc2, c.one2);
#if WS_RCP == 1
            const P2 rc = rcp2(mul2(den0, den1));
            r[0] = mul2(rc, den1);
            r[1] = mul2(rc, den0);
#elif WS_RCP == 2
            const P2 m = mul2(den0, den1);
            const float rc = fast_rcp(m.x * m.y);
            const P2 rx = make_float2(rc * m.y, rc * m.x);
            r[0] = mul2(rx, den1);
            r[1] = mul2(rx, den0);
#else
            r[0] = rcp2(den0);
            r[1] = rcp2(den1);
#endif
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (PXM) r[q] = mul2(r[q], m2);
            if (i == 0) {                                  // p^0 = 0: p^1 = (-tau*r) * g
                const P2 rt = mul2(r[q], c.mtau2);
                pn0[q] = mul2(g0[q], rt);
                pn1[q] = mul2(g1[q], rt);
            } else {
                pn0[q] = mul2(fma2(g0[q], c.mtau2, pi0[q]), r[q]);
                pn1[q] = mul2(fma2(g1[q], c.mtau2, pi1[q]), r[q]);
            }
        }
        // D(p^{i+1})(u) = (p0(u-1) - p0(u)) + (p1(u, left) - p1(u)); the left neighbour of A is the previous lane's B
        const P2 p1l_a = (WS_EXP & 8) ? pn1[1] : shfl_up2(pn1[1]);
        P2 d[2];
        d[0] = add2(fma2(pn0[0], c.mone2, S.P0[i][0]), fma2(pn1[0], c.mone2, p1l_a));
        d[1] = add2(fma2(pn0[1], c.mone2, S.P0[i][1]), fma2(pn1[1], c.mone2, pn1[0]));
        // energies: w*|grad out_i|(u) belongs to iteration i, D(p^{i+1})(u)^2 to iteration i+1
        if (!(FAST && (WS_EXP & 4))) {
        S.en[i] = fma2(nrm[0], wm2, S.en[i]);
        S.en[i] = fma2(nrm[1], wm2, S.en[i]);
        }
        if (!(FAST && (WS_EXP & 4)))
        if (i + 1 < R) {
            if (FAST && !EM) {
                S.en[i + 1] = fma2(d[0], d[0], S.en[i + 1]);
                S.en[i + 1] = fma2(d[1], d[1], S.en[i + 1]);
            } else {
                S.en[i + 1] = fma2(mul2(d[0], me2), d[0], S.en[i + 1]);
                S.en[i + 1] = fma2(mul2(d[1], me2), d[1], S.en[i + 1]);
            }
        }
        // state of this stage for the next step
        P2 g1b_new = fma2(o_new[1], c.mone2, o_rb);
        if (PXM) g1b_new = mul2(g1b_new, splat(c.right_in));
        S.g1b[i] = g1b_new;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const P2 o_next = add2(i == R - 1 ? f_old[q] : S.fd[i < R - 1 ? i : 0][q], d[q]);
            // the dual variable stage i+1 needs now is the one this stage produced in the previous step
            pi0[q] = S.P0[i][q];
            pi1[q] = S.P1[i][q];
            S.P0[i][q] = pn0[q];
            S.P1[i][q] = pn1[q];
            S.o_prev[i][q] = o_new[q];
            o_new[q] = o_next;
            if (WS_EXP & 1) o_last[q] = add2(o_last[q], o_next);
        }
    }
    if (FAST && (WS_EXP & 1)) { o_new[0] = o_last[0]; o_new[1] = o_last[1]; }
#pragma unroll
    for (int i = R - 2; i > 0; --i) { S.fd[i][0] = S.fd[i - 1][0]; S.fd[i][1] = S.fd[i - 1][1]; }
    S.fd[0][0] = f_new[0];
    S.fd[0][1] = f_new[1];
    o_out[0] = o_new[0];
    o_out[1] = o_new[1];
}

// Work of one CTA: the scene is (batch x column strips) strips of H rows laid end to end, every strip charged
// kSegCost extra units in front of its rows; a CTA takes `per_cta` consecutive units, i.e. a few row segments of
// neighbouring strips (one or two on large scenes).  No wave quantisation: every SM gets the same share.
constexpr int kWsSegCost = 8;     // cost of starting a row segment, in rows (warm-up + drain rows); 278-row tiles: 8..12 best, 16 +2 %
constexpr int kWsTick = 16;       // work units ("ticks") per row of an interior strip
template <int R>
struct WsSegIter {
    long long unit, unit_end, Tb;
    int Li, Le, wi, we, chg, Hw, H, lo, nstrips;    // strip lengths and row weights in ticks (interior / edge)
    int b, strip, r0, r1, rs, t_end, nblk;          // current segment
    __device__ WsSegIter(const WsParams& p, int per_cta_unused = 0) {
        H = p.H; lo = p.out_lo; Hw = p.out_hi - p.out_lo; nstrips = p.nstrips;
        // The first and the last strip of a scene hang over the image edge: their blocks carry pixel masks (about 7 %
        // more instructions), so a row of theirs weighs kWsTick + edge_cost ticks instead of kWsTick.
        wi = kWsTick; we = kWsTick + p.edge_cost; chg = p.seg_cost * kWsTick;
        Li = chg + Hw * wi; Le = chg + Hw * we;
        Tb = nstrips == 1 ? (long long)Le : 2LL * Le + (long long)(nstrips - 2) * Li;
        // total = q * grid + rem: the first `rem` CTAs take q + 1 ticks, the others q
        const long long total = (long long)(p.pack ? 1 : p.B) * Tb;
        const long long q = total / gridDim.x, rem = total - q * gridDim.x, c = blockIdx.x;
        unit = c * q + (c < rem ? c : rem);
        unit_end = unit + q + (c < rem ? 1 : 0);
    }
    __device__ bool next() {
        while (unit < unit_end) {
            b = (int)(unit / Tb);
            long long v = unit - (long long)b * Tb;
            int s, len, w;
            if (v < Le) { s = 0; len = Le; w = we; }
            else {
                v -= Le;
                s = 1 + (int)(v / Li);
                if (s >= nstrips - 1) { s = nstrips - 1; v -= (long long)(nstrips - 2) * Li; len = Le; w = we; }
                else { v -= (long long)(s - 1) * Li; len = Li; w = wi; }
            }
            const int v0 = (int)v;
            const long long left = unit_end - unit;
            const int v1 = (long long)(len - v0) < left ? len : v0 + (int)left;
            unit += v1 - v0;
            // a row belongs to the range that holds its first tick
            const int a0 = v0 - chg > 0 ? v0 - chg : 0, a1 = v1 - chg > 0 ? v1 - chg : 0;
            r0 = lo + (a0 + w - 1) / w;
            r1 = lo + (a1 + w - 1) / w;
            if (r1 <= r0) continue;                                  // only charge ticks: no rows here
            strip = s;
            rs = r0 - R > 0 ? r0 - R : 0;
            t_end = r1 + R;                                          // steps t in [rs, t_end)
            nblk = (t_end - rs + WRB - 1) / WRB;
            return true;
        }
        return false;
    }
};

// MODE: MODE_GAP_ACC, MODE_GAP_PLAIN, MODE_ADMM or MODE_TV (the denoiser alone: f = x_in).  Q = C/2 channel pairs.
template <int R, int MODE, int Q>
__global__ void __launch_bounds__(ws_threads(Q), 1)
gap_tv_ws_kernel(const WsParams p, const __grid_constant__ WsMaps maps) {
    constexpr int NGRP = ws_groups(Q);
    constexpr int CW = ws_consumers(Q);
    constexpr int C = 2 * Q, K = Q / 2;
    constexpr bool BST = MODE == MODE_ADMM && ws_bstage(Q);       // the multiplier is staged by TMA like theta and Phi
    constexpr WsSmem L = ws_smem(Q, BST);
    constexpr int NF = L.nf, NRAW = L.nraw;
    static_assert(Q % 2 == 0, "C must be a multiple of 4");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    // mbarriers: raw_full[NRAW], raw_empty[NRAW], f_full[NF], f_empty[NF], out_full[NOUT], out_empty[NOUT]
    const uint32_t bar_raw = smem_base + L.bar_off;
    const uint32_t bar_rempty = bar_raw + 8 * NRAW;
    const uint32_t bar_ffull = bar_rempty + 8 * NRAW;
    const uint32_t bar_fempty = bar_ffull + 8 * NF;
    const uint32_t bar_ofull = bar_fempty + 8 * NF;
    const uint32_t bar_oempty = bar_ofull + 8 * NOUT;
    if (tid == 0) {
        constexpr int kArr = WS_SANITIZE ? 32 : 1;       // arrivals per warp
        for (int i = 0; i < NRAW; ++i) { mbar_init(bar_raw + 8 * i, 1); mbar_init(bar_rempty + 8 * i, NCOMP * kArr); }
        for (int i = 0; i < NF; ++i) { mbar_init(bar_ffull + 8 * i, NCOMP * kArr); mbar_init(bar_fempty + 8 * i, CW * kArr); }
        for (int i = 0; i < NOUT; ++i) { mbar_init(bar_ofull + 8 * i, CW * kArr); mbar_init(bar_oempty + 8 * i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        fence_async_smem();
    }
    __syncthreads();
    // Programmatic dependent launch (short kernels only, WsParams::use_pdl): this grid may have been started while the
    // previous kernel of the stream (the previous outer iteration) was still draining, so that the launch latency and
    // the prologue above overlap its tail.  Everything below reads or overwrites what that kernel reads or writes:
    // wait for it here.  The next launch may be scheduled at once (it waits at this same point).  Measured: -1.7 us
    // of 30.5 at 256x256x8, -2 us of 49 at 4x256x256x24, nothing at 278x3840x24, +6 % on the 2160-row scene.
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");

    const int H = p.H, W = p.W, own = p.own;
    // Every role walks the same sequence of (segment, block) pairs; `gb` counts the blocks of all segments so far and
    // gives the ring slot and the mbarrier phase.

    // Register reallocation between the roles (warpgroups of four warps): the producers give registers back, the
    // consumers -- whose pipeline state fills the 128 registers a 512-thread CTA starts with -- take them.
    constexpr bool kRealloc = WS_CREGS > 0 && CW % 4 == 0 && ws_threads(Q) == 512;
    // role of this warp: consumer index cwarp (0..CW-1) or producer-side index pwarp (0 = TMA warp, 1..NCOMP projection)
    const int cwarp = WS_PROD_FIRST ? warp - NPROD : (warp < CW ? warp : -1);
    const int pwarp = WS_PROD_FIRST ? warp : warp - CW;
    if (cwarp >= 0) {
        // =================================== consumers ===================================
        if (kRealloc) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(WS_CREGS > 0 ? WS_CREGS : 128));
        const int gi = cwarp / Q, q = cwarp - gi * Q;
        // f tile: [WRB][NGRP][Q][GW][2] floats; this lane reads 16 bytes (A.c0 A.c1 B.c0 B.c1)
        const uint32_t f_lane = smem_base + L.f_off + ((gi * Q + q) * GW + 2 * lane) * 8;
        constexpr int F_ROW = NGRP * Q * GW * 8;
        // out tile: [WRB][NGRP] sub-tiles of [K][own][4] floats; chunk k = q/2, half h = q%2
        const int kq = q >> 1, hq = q & 1;
        const uint32_t o_lane = smem_base + L.out_off + gi * L.out_sub + ((kq * own + 2 * (lane - HALO / 2)) * 4 + 2 * hq) * 4;
        constexpr int O_ROW = NGRP * L.out_sub;
        WsConst sc;
        sc.mone2 = splat(-1.f); sc.mtau2 = splat(-0.25f); sc.tvc2 = splat(p.tv_c); sc.one2 = splat(1.f);
        sc.w2 = splat(p.tv_w); sc.tvw = p.tv_w;
        sc.H = H;
        long long pw0 = 0, pw1 = 0;
        const long long pstart = p.prof ? clock64() : 0;
        int gb = 0;
        WsSegIter<R> it(p);
#pragma unroll 1
        while (it.next()) {
        const int r0 = it.r0, r1 = it.r1, rs = it.rs, t_end = it.t_end, nblk = it.nblk;
        const WsSlot sl = ws_slot(p, it.b, it.strip, gi, NGRP);
        const int b = sl.b, grp = sl.grp;
        const bool grp_live = sl.live;
        const int base = grp * own - HALO;                          // pixel of lane 0's A
        const int pxa = base + 2 * lane;
        const bool pair_in = grp_live && pxa >= 0 && pxa < W;       // W is even and base is even: pairs are atomic
        const bool own_lane = pair_in && lane >= HALO / 2 && lane < HALO / 2 + own / 2;
        const bool interior = grp_live && base >= 0 && base + GW <= W;
        WsPipe<R> S;
        const P2 z = splat(0.f);
#pragma unroll
        for (int i = 0; i < R; ++i) {
            S.g1b[i] = z; S.en[i] = z;
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2) { S.o_prev[i][k2] = z; S.P0[i][k2] = z; S.P1[i][k2] = z; if (i < R - 1) S.fd[i][k2] = z; }
        }
        sc.pair_in = pair_in ? 1.f : 0.f;
        sc.right_in = (pair_in && pxa + 2 < W) ? 1.f : 0.f;
        sc.rs = rs; sc.r0 = r0; sc.r1 = r1;
        const int fast_lo = r0 + R, fast_hi = min(r1, H) - 1;
#if WS_BLK_UNROLL == 2
#pragma unroll 2
#else
#pragma unroll 1
#endif
        for (int blk = 0; blk < nblk; ++blk, ++gb) {
            const int fs = gb % NF, os = gb % NOUT;
            long long c0 = 0, c1 = 0;
            if (p.prof) c0 = clock64();
            mbar_wait(bar_ffull + 8 * fs, (gb / NF) & 1);
            if (p.prof) { c1 = clock64(); pw0 += c1 - c0; }
            mbar_wait(bar_oempty + 8 * os, ((gb / NOUT) & 1) ^ 1);
            if (p.prof) pw1 += clock64() - c1;
            const uint32_t fsrc = f_lane + fs * L.f_bytes;
            const uint32_t fprev = f_lane + ((gb + NF - 1) % NF) * L.f_bytes;      // the previous block's rows (kept until this block is done)
            const uint32_t odst = o_lane + os * L.out_bytes;
            static_assert(R <= WRB, "f(t-R) must lie in this block or the previous one");
            // shared-memory address of f(t0 + j - R)
            auto f_old_addr = [&](int j) { return j >= R ? fsrc + (j - R) * F_ROW : fprev + (j - R + WRB) * F_ROW; };
            const int t0 = rs + blk * WRB;
            const bool rows_fast = t0 >= fast_lo && t0 + WRB - 1 <= fast_hi;
            // ADMM multiplier: x = f + b, so b - (x - theta_new) = theta_new - f, and f(t-R) is the row that closes the
            // last stage.  Written straight to global memory (two 8-byte stores per lane; the channel pairs of the other
            // consumer warps complete the sectors in L2).
            auto store_b = [&](int orow, const P2 (&o)[2], const P2 (&fo)[2], bool check_rows) {
                if (!own_lane || (check_rows && (orow < r0 || orow >= r1))) return;
                const size_t off = (((size_t)b * H + orow) * W + pxa) * C + 2 * q;
                const float2 ba = make_float2(o[0].x - fo[0].x, o[0].y - fo[0].y), bb = make_float2(o[1].x - fo[1].x, o[1].y - fo[1].y);
                *reinterpret_cast<float2*>(p.b_out + off) = ba;
                *reinterpret_cast<float2*>(p.b_out + off + C) = bb;
                if (p.b_up != nullptr && orow < p.out_lo + R) {          // halo push: the rows next to a seam, into the neighbours' halo rows
                    *reinterpret_cast<float2*>(p.b_up + off) = ba;
                    *reinterpret_cast<float2*>(p.b_up + off + C) = bb;
                }
                if (p.b_dn != nullptr && orow >= p.out_hi - R) {
                    *reinterpret_cast<float2*>(p.b_dn + off) = ba;
                    *reinterpret_cast<float2*>(p.b_dn + off + C) = bb;
                }
            };
            auto fast_block = [&](auto path) {
                constexpr int PATH = decltype(path)::value;
#pragma unroll
                for (int j = 0; j < WRB; ++j) {
                    float4 fv;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(fv.x), "=f"(fv.y), "=f"(fv.z), "=f"(fv.w) : "r"(fsrc + j * F_ROW));
                    const P2 f_new[2] = {make_float2(fv.x, fv.y), make_float2(fv.z, fv.w)};
                    float4 fo;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(fo.x), "=f"(fo.y), "=f"(fo.z), "=f"(fo.w) : "r"(f_old_addr(j)));
                    const P2 f_old[2] = {make_float2(fo.x, fo.y), make_float2(fo.z, fo.w)};
                    P2 o[2];
                    ws_step<R, PATH>(S, sc, t0 + j, f_new, f_old, o);
                    if (own_lane && (!(WS_EXP & 2) || o[0].x == 123.456f)) {
                        asm volatile("st.shared.v2.f32 [%0], {%1,%2};\n" ::"r"(odst + j * O_ROW), "f"(o[0].x), "f"(o[0].y) : "memory");
                        asm volatile("st.shared.v2.f32 [%0], {%1,%2};\n" ::"r"(odst + j * O_ROW + 16), "f"(o[1].x), "f"(o[1].y) : "memory");
                    }
                    if (MODE == MODE_ADMM) store_b(t0 + j - R, o, f_old, PATH == 3);
                }
            };
            // every row touched exists (warm-up done, f(t-R) in the ring) and lies inside the image; steps past t_end
            // are harmless there (their rows are neither stored nor counted)
            const bool rows_em = WS_PATH3 && t0 - R >= rs && t0 + WRB - 1 <= H - 1;
            if (rows_fast && interior) {
                fast_block(std::integral_constant<int, 1>{});
            } else if (rows_fast) {
                fast_block(std::integral_constant<int, 2>{});
            } else if (rows_em && interior) {
                fast_block(std::integral_constant<int, 3>{});
            } else {
#if WS_GEN_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
                for (int j = 0; j < WRB; ++j) {
                    const int t = t0 + j;
                    if (t < t_end) {
                        float4 fv;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(fv.x), "=f"(fv.y), "=f"(fv.z), "=f"(fv.w) : "r"(fsrc + j * F_ROW));
                        const P2 f_new[2] = {make_float2(fv.x, fv.y), make_float2(fv.z, fv.w)};
                        float4 fo = make_float4(0.f, 0.f, 0.f, 0.f);          // rows above the segment: never used unmasked
                        if (t - R >= rs)
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(fo.x), "=f"(fo.y), "=f"(fo.z), "=f"(fo.w) : "r"(f_old_addr(j)));
                        const P2 f_old[2] = {make_float2(fo.x, fo.y), make_float2(fo.z, fo.w)};
                        P2 o[2];
                        ws_step<R, 0>(S, sc, t, f_new, f_old, o);
                        if (own_lane) {
                            asm volatile("st.shared.v2.f32 [%0], {%1,%2};\n" ::"r"(odst + j * O_ROW), "f"(o[0].x), "f"(o[0].y) : "memory");
                            asm volatile("st.shared.v2.f32 [%0], {%1,%2};\n" ::"r"(odst + j * O_ROW + 16), "f"(o[1].x), "f"(o[1].y) : "memory");
                        }
                        if (MODE == MODE_ADMM) store_b(t - R, o, f_old, true);
                    }
                }
            }
            fence_async_smem();                  // the TMA store reads what this warp just wrote
            __syncwarp();
            if (WS_SANITIZE || lane == 0) {
                mbar_arrive(bar_ofull + 8 * os);
                // the last stage reads f(t-R) out of the previous block's slot: a slot is handed back one block late,
                // the last one of a segment together with its predecessor
                if (blk > 0) mbar_arrive(bar_fempty + 8 * ((gb + NF - 1) % NF));
                if (blk == nblk - 1) mbar_arrive(bar_fempty + 8 * fs);
            }
        }
        // energy partials of this segment: owned lanes only, one atomic per (channel, iteration)
        if (p.flag != nullptr || p.energy_log != nullptr) {
#pragma unroll
            for (int i = 0; i < R; ++i)
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    float v = own_lane ? (ch ? S.en[i].y : S.en[i].x) : 0.f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0 && grp_live) atomicAdd(p.energy + ((size_t)b * C + 2 * q + ch) * R + i, (double)v);
                }
        }
        }   // segments
        if (MODE == MODE_ADMM && (p.b_up != nullptr || p.b_dn != nullptr)) __threadfence_system();   // pushed b rows before the flag
        if (p.prof && lane == 0) {
            long long* pr = p.prof + ((size_t)blockIdx.x * (blockDim.x / 32) + warp) * 4;
            pr[0] = clock64() - pstart; pr[1] = pw0; pr[2] = pw1; pr[3] = 0;
        }
    } else {
        // =================================== producers ===================================
        if (kRealloc) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(WS_PREGS > 0 ? WS_PREGS : 128));
        constexpr int NPT = NCOMP * 32;
        constexpr int NITEM = WRB * NGRP * GW;                   // (row, group, pixel) items per block
        constexpr bool TVONLY = MODE == MODE_TV;                 // standalone denoiser: f is the input itself
        constexpr bool ADMM = MODE == MODE_ADMM;
        constexpr uint32_t kTx = TVONLY ? (uint32_t)L.x_bytes
                                        : (BST ? 3u : 2u) * L.x_bytes + (MODE == MODE_GAP_ACC ? 3u : 2u) * NGRP * WRB * GW * 4;
        if (pwarp == 0) {
            // ------------- the TMA warp: lane 0 loads, lane 1 stores, each walking the block sequence on its own.
            // Neither ever waits for the other, and the projection warps never wait for a store: a finished output
            // block leaves for HBM as soon as the last consumer warp has arrived.
            if (lane == 0) {
                WsSegIter<R> ld(p);
                int g = 0;
                bool up_ok = p.wait_up == nullptr, dn_ok = p.wait_dn == nullptr;
#pragma unroll 1
                while (ld.next()) {
                    // per group slot of this segment (this lane runs alone: keep its per-block instruction count small)
                    int l_px0[NGRP], l_row[NGRP], l_prow[NGRP];
#pragma unroll
                    for (int g2 = 0; g2 < NGRP; ++g2) {
                        const WsSlot sl = ws_slot(p, ld.b, ld.strip, g2, NGRP);
                        l_px0[g2] = sl.grp * own - HALO; l_row[g2] = sl.b * H; l_prow[g2] = p.phi_batched ? sl.b * H : 0;
                    }
#pragma unroll 1
                    for (int blk = 0; blk < ld.nblk; ++blk, ++g) {
                        const int slot = g % NRAW;
                        mbar_wait(bar_rempty + 8 * slot, ((g / NRAW) & 1) ^ 1);      // the projection warps are done with it
                        const uint32_t dst = smem_base + slot * L.raw_bytes;
                        const uint32_t bar = bar_raw + 8 * slot;
                        const int row0 = ld.rs + blk * WRB;
                        // Halo rows hold what the neighbour pushed in its previous iteration, and the owned rows next
                        // to a seam are pushed into buffers the neighbour read in that iteration: both need its flag.
                        if (!up_ok && row0 < p.out_lo + R) { wait_peer_flag(p.wait_up, p.wait_epoch, p.timeout_flag); up_ok = true; }
                        if (!dn_ok && row0 + WRB > p.out_hi - R) { wait_peer_flag(p.wait_dn, p.wait_epoch, p.timeout_flag); dn_ok = true; }
                        mbar_expect_tx(bar, kTx);
#pragma unroll
                        for (int g2 = 0; g2 < NGRP; ++g2) {
                            const int rowc = l_row[g2] + row0, prow = l_prow[g2] + row0, px0 = l_px0[g2];
                            tma_load_3d(dst + g2 * (WRB * GW * C * 4), &maps.x, 0, px0, rowc, bar);
                            if (TVONLY) continue;
                            tma_load_3d(dst + L.x_bytes + g2 * (WRB * GW * C * 4), &maps.phi, 0, px0, prow, bar);
                            if (BST) tma_load_3d(dst + 2 * L.x_bytes + g2 * (WRB * GW * C * 4), &maps.b, 0, px0, rowc, bar);
                            const uint32_t ds = dst + L.small_off + g2 * (WRB * GW * 4);
                            tma_load_2d(ds, &maps.y, px0, rowc, bar);
                            if (MODE == MODE_GAP_ACC) tma_load_2d(ds + NGRP * WRB * GW * 4, &maps.y1, px0, rowc, bar);
                            tma_load_2d(ds + 2 * NGRP * WRB * GW * 4, &maps.ps, px0, prow, bar);
                        }
                    }
                }
            } else if (lane == 1) {
                WsSegIter<R> st(p);
                int g = 0;
#pragma unroll 1
                while (st.next()) {
                    int s_px0[NGRP], s_row[NGRP]; bool s_live[NGRP];
#pragma unroll
                    for (int g2 = 0; g2 < NGRP; ++g2) {
                        const WsSlot sl = ws_slot(p, st.b, st.strip, g2, NGRP);
                        s_px0[g2] = sl.grp * own; s_row[g2] = sl.b * H; s_live[g2] = sl.live;
                    }
#pragma unroll 1
                    for (int blk = 0; blk < st.nblk; ++blk, ++g) {
                        const int os = g % NOUT;
                        mbar_wait(bar_ofull + 8 * os, (g / NOUT) & 1);
                        const uint32_t src = smem_base + L.out_off + os * L.out_bytes;
                        for (int j = 0; j < WRB; ++j) {
                            const int orow = st.rs + blk * WRB + j - R;
                            if (orow >= st.r0 && orow < st.r1) {
                                const bool to_up = p.sig_up != nullptr && orow < p.out_lo + R;
                                const bool to_dn = p.sig_dn != nullptr && orow >= p.out_hi - R;
#pragma unroll
                                for (int g2 = 0; g2 < NGRP; ++g2) {
                                    if (s_live[g2]) {
                                        const uint32_t sub = src + (j * NGRP + g2) * L.out_sub;
                                        const int px0 = s_px0[g2];
                                        tma_store_4d(&maps.out, sub, 0, px0, 0, s_row[g2] + orow);
                                        if (to_up) tma_store_4d(&maps.out_up, sub, 0, px0, 0, orow + p.up_shift);
                                        if (to_dn) tma_store_4d(&maps.out_dn, sub, 0, px0, 0, orow + p.dn_shift);
                                    }
                                }
                            }
                        }
                        bulk_commit();
                        bulk_wait_read0();
                        mbar_arrive(bar_oempty + 8 * os);
                    }
                }
                bulk_wait0();
                if (p.sig_up != nullptr || p.sig_dn != nullptr) __threadfence_system();     // pushed rows before the flag
            }
        } else {
        // ------------- projection warps
        const int ptid = (pwarp - 1) * 32 + lane;                // 0 .. NCOMP*32-1
        long long pw0 = 0, pw1 = 0, pw2 = 0;
        const long long pstart = p.prof ? clock64() : 0;
        const float lam = p.lambda;
        // bank pattern of the lane-per-pixel 16-byte reads: pixels are C*4 bytes apart.  For K = 2, 6 every
        // other group of four lanes visits the chunk pairs in swapped order, for K = 4 the chunk index is
        // xor-ed with (pixel / 2) % 4; odd K is conflict-free as it is.
        const int lsw = (K % 4 == 2) ? ((lane >> 2) & 1) : (K % 8 == 4) ? ((lane >> 1) & 3) : 0;
        int gb = 0;
        WsSegIter<R> it(p);
#pragma unroll 1
        while (it.next()) {
        const int r0 = it.r0, r1 = it.r1, rs = it.rs, nblk = it.nblk;
        // per group slot of this strip: first pixel of the tile, live?, element offset of its measurement plane
        int s_px0[NGRP]; bool s_live[NGRP]; size_t s_plane[NGRP];
#pragma unroll
        for (int g2 = 0; g2 < NGRP; ++g2) {
            const WsSlot sl = ws_slot(p, it.b, it.strip, g2, NGRP);
            s_px0[g2] = sl.grp * own - HALO; s_live[g2] = sl.live; s_plane[g2] = (size_t)sl.b * H * W;
        }
#pragma unroll 1
        for (int blk = 0; blk < nblk; ++blk, ++gb) {
            const int slot = gb % NRAW, fs = gb % NF;
            long long c0 = 0, c1 = 0;
            if (p.prof) c0 = clock64();
            mbar_wait(bar_raw + 8 * slot, (gb / NRAW) & 1);
            if (p.prof) { c1 = clock64(); pw0 += c1 - c0; }
            mbar_wait(bar_fempty + 8 * fs, ((gb / NF) & 1) ^ 1);
            if (p.prof) pw1 += clock64() - c1;
            const unsigned char* raw = smem_raw + slot * L.raw_bytes;
            unsigned char* fdst = smem_raw + L.f_off + fs * L.f_bytes;
#pragma unroll 1
            for (int itx = ptid; itx < NITEM; itx += NPT) {
                const int px = itx & (GW - 1), g = (itx / GW) % NGRP, j = itx / (GW * NGRP);
                const int row = rs + blk * WRB + j;
                int gpx = s_px0[0] + px; bool live = s_live[0]; size_t plane = s_plane[0];
#pragma unroll
                for (int g2 = 1; g2 < NGRP; ++g2) if (g == g2) { gpx = s_px0[g2] + px; live = s_live[g2]; plane = s_plane[g2]; }
                const bool in = live && gpx >= 0 && gpx < W && row < H;
                const float4* tx = reinterpret_cast<const float4*>(raw) + ((g * WRB + j) * GW + px) * K;
                // f tile: [WRB][NGRP][Q][GW][2]
                float2* frow = reinterpret_cast<float2*>(fdst) + ((j * NGRP + g) * Q) * GW + px;
                if constexpr (TVONLY) {
#pragma unroll
                    for (int k0 = 0; k0 < K; ++k0) {
                        const int kc = k0 ^ lsw;
                        const float4 v = tx[kc];                     // pixels outside the image: zero-filled by TMA
                        frow[(2 * kc) * GW] = make_float2(v.x, v.y);
                        frow[(2 * kc + 1) * GW] = make_float2(v.z, v.w);
                    }
                    continue;
                }
                const float4* tp = tx + L.x_bytes / 16;
                float4 xv[K], pv[K];
#pragma unroll
                for (int k0 = 0; k0 < K; ++k0) { xv[k0] = tx[k0 ^ lsw]; pv[k0] = tp[k0 ^ lsw]; }
                P2 acc2 = splat(0.f);
                // ADMM: this pixel's multiplier, C contiguous floats in global memory (zero outside the image)
                const size_t goff = ADMM ? (plane + (size_t)(row < H ? row : 0) * W + (in ? gpx : 0)) * C : 0;
                // staged: the tile next to Phi's (same layout as theta's, so the same chunk order k0 ^ lsw applies)
                const float4* bsrc = !ADMM ? nullptr : BST ? tx + 2 * (L.x_bytes / 16) : reinterpret_cast<const float4*>(p.b_in + goff);
#pragma unroll
                for (int k0 = 0; k0 < K; ++k0) {
                    float4 u = xv[k0];
                    if (ADMM && in) {                                // u = theta + b enters the dot product
                        const float4 bq = BST ? bsrc[k0 ^ lsw] : __ldcg(bsrc + (k0 ^ lsw));     // L2: halo rows are written by other GPUs
                        u.x += bq.x; u.y += bq.y; u.z += bq.z; u.w += bq.w;
                    }
                    acc2 = fma2(make_float2(u.x, u.y), make_float2(pv[k0].x, pv[k0].y), acc2);
                    acc2 = fma2(make_float2(u.z, u.w), make_float2(pv[k0].z, pv[k0].w), acc2);
                }
                const float acc = acc2.x + acc2.y;
                const float* sm = reinterpret_cast<const float*>(raw + L.small_off) + (g * WRB + j) * GW + px;
                const float yv = sm[0];
                const float psv = sm[2 * NGRP * WRB * GW];
                float sv;
                if (MODE == MODE_GAP_ACC) {
                    const float y1n = sm[NGRP * WRB * GW] + (yv - acc);
                    if (in && px >= HALO && px < HALO + own && row >= r0 && row < r1) {
                        const size_t idx = (size_t)row * W + gpx;
                        p.y1_out[plane + idx] = y1n;
                        if (p.y1_up != nullptr && row < p.out_lo + R) p.y1_up[idx] = y1n;
                        if (p.y1_dn != nullptr && row >= p.out_hi - R) p.y1_dn[idx] = y1n;
                    }
                    sv = (y1n - acc) * fast_rcp(psv);
                } else if (ADMM) {
                    sv = (yv - acc) * fast_rcp(psv + p.gamma);
                } else {
                    sv = (yv - acc) * fast_rcp(psv);
                }
                const P2 s2 = splat(in ? sv * lam : 0.f);
                const bool want_x = ADMM && p.xproj_out != nullptr && in && px >= HALO && px < HALO + own && row >= r0 && row < r1;
#pragma unroll
                for (int k0 = 0; k0 < K; ++k0) {
                    const int kc = k0 ^ lsw;
                    const P2 f01 = fma2(s2, make_float2(pv[k0].x, pv[k0].y), make_float2(xv[k0].x, xv[k0].y));
                    const P2 f23 = fma2(s2, make_float2(pv[k0].z, pv[k0].w), make_float2(xv[k0].z, xv[k0].w));
                    frow[(2 * kc) * GW] = f01;                       // TV input: f = theta + lambda*s*Phi = x - b
                    frow[(2 * kc + 1) * GW] = f23;
                    if (ADMM && want_x) {                            // x = f + b (the projection output the caller reads)
                        const float4 bq = BST ? bsrc[kc] : __ldcg(bsrc + kc);
                        reinterpret_cast<float4*>(p.xproj_out + goff)[kc] = make_float4(f01.x + bq.x, f01.y + bq.y, f23.x + bq.z, f23.y + bq.w);
                    }
                }
            }
            __syncwarp();
            if (WS_SANITIZE || lane == 0) {
                mbar_arrive(bar_ffull + 8 * fs);
                mbar_arrive(bar_rempty + 8 * slot);          // this warp is done reading the raw slot
            }
        }
        }   // segments
        if (p.y1_up != nullptr || p.y1_dn != nullptr) __threadfence_system();               // pushed y1 rows before the flag
        if (p.prof && lane == 0) {
            long long* pr = p.prof + ((size_t)blockIdx.x * (blockDim.x / 32) + warp) * 4;
            pr[0] = clock64() - pstart; pr[1] = pw0; pr[2] = pw1; pr[3] = pw2;
        }
        }   // projection warps
    }

    // ---- the last CTA: skimage's stopping rule replayed on the accumulated energies; the neighbours' flags
    const bool want_energy = p.flag != nullptr || p.energy_log != nullptr;
    const bool want_signal = p.sig_up != nullptr || p.sig_dn != nullptr;
    if (want_energy || want_signal) {
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (want_energy) {
                const int nslice = p.B * C;
                bool fired = false;
                for (int s = tid; s < nslice; s += blockDim.x) {
                    volatile double* e = p.energy + (size_t)s * R;
                    double ev[R];
#pragma unroll
                    for (int i = 0; i < R; ++i) { ev[i] = e[i]; e[i] = 0.0; }     // leave the accumulators clean
                    if (p.energy_log != nullptr) {
#pragma unroll
                        for (int i = 0; i < R; ++i) p.energy_log[(size_t)s * R + i] = ev[i];
                    }
                    double e_prev = ev[0];
#pragma unroll
                    for (int i = 1; i < R; ++i) {                                   // a stop at i = R changes nothing
                        if (fabs(e_prev - ev[i]) < p.tv_eps * ev[0]) fired = true;
                        e_prev = ev[i];
                    }
                }
                if (fired && p.flag != nullptr) atomicOr(p.flag, 1);
            }
            if (tid == 0) {
                *p.ticket = 0u;
                if (want_signal) {
                    // every CTA's pushed rows are behind its ticket (bulk_wait / __threadfence_system above)
                    __threadfence_system();
                    if (p.sig_up != nullptr) st_release_sys(p.sig_up, p.sig_epoch);
                    if (p.sig_dn != nullptr) st_release_sys(p.sig_dn, p.sig_epoch);
                }
            }
        }
    }
}

// one launcher per R, defined in ws_inst_r{2,3,4}.cu
template <int R> int ws_launch_r(int mode, int Q, const WsParams& p, const WsMaps& maps, int grid, cudaStream_t st);

template <int R, int MODE, int Q>
int ws_launch_q(const WsParams& p, const WsMaps& maps, int grid, cudaStream_t st) {
    auto kfn = gap_tv_ws_kernel<R, MODE, Q>;
    constexpr WsSmem L = ws_smem(Q, MODE == MODE_ADMM && ws_bstage(Q));
    static bool configured = false;
    if (!configured) {
        SCIPNP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        configured = true;
    }
    static const bool pdl = getenv("SCIPNP_WS_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)ws_threads(Q));
    cfg.dynamicSmemBytes = L.total;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && p.use_pdl) ? 1 : 0;
    SCIPNP_CUDA(cudaLaunchKernelEx(&cfg, kfn, p, maps));
    return SCIPNP_OK;
}

template <int R, int MODE>
int ws_launch_mode(int Q, const WsParams& p, const WsMaps& maps, int grid, cudaStream_t st) {
    switch (Q) {
        case 4: return ws_launch_q<R, MODE, 4>(p, maps, grid, st);
        case 12: return ws_launch_q<R, MODE, 12>(p, maps, grid, st);
#ifndef SCIPNP_FUSED_FAST_BUILD
        case 2: return ws_launch_q<R, MODE, 2>(p, maps, grid, st);
        case 6: return ws_launch_q<R, MODE, 6>(p, maps, grid, st);
        case 8: return ws_launch_q<R, MODE, 8>(p, maps, grid, st);
        case 10: return ws_launch_q<R, MODE, 10>(p, maps, grid, st);
#endif
    }
    set_error("warp-specialised fused kernel not built for C = %d", 2 * Q);
    return SCIPNP_EINVAL;
}

#define SCIPNP_INSTANTIATE_WS_R(RR)                                                                              \
    template <> int ws_launch_r<RR>(int mode, int Q, const WsParams& p, const WsMaps& maps, int grid, cudaStream_t st) { \
        if (mode == MODE_GAP_ACC) return ws_launch_mode<RR, MODE_GAP_ACC>(Q, p, maps, grid, st);                 \
        if (mode == MODE_TV) return ws_launch_mode<RR, MODE_TV>(Q, p, maps, grid, st);                           \
        if (mode == MODE_ADMM) return ws_launch_mode<RR, MODE_ADMM>(Q, p, maps, grid, st);                       \
        return ws_launch_mode<RR, MODE_GAP_PLAIN>(Q, p, maps, grid, st);                                         \
    }

}  // namespace wsk
}  // namespace scipnp
