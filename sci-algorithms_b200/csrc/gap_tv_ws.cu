// Host side of the warp-specialised fused GAP-TV kernel (gap_tv_ws.cuh): work split, TMA descriptors
// (cached per buffer set: x ping-pongs between two buffers, so a solver sees two sets), launch.
#include "gap_tv_ws.cuh"

#include <algorithm>
#include <mutex>
#include <vector>

namespace scipnp {

using namespace wsk;

int make_tensor_map_f32(CUtensorMap* tm, const float* base, int rank, const unsigned long long* dims,
                        const unsigned long long* strides_bytes, const unsigned* box, int l2_promotion_128, int swizzle128 = 0);

namespace {

int g_variant = -1;      // -1: from the environment (SCIPNP_FUSED_VARIANT), 0: auto, 1: stream kernel, 2: ws kernel

struct MapKey {
    const float *x_in, *x_out, *phi, *y, *y1, *ps;
    int B, H, W, C, own, phi_batched;
    const float* b_in = nullptr;                       // ADMM with a staged multiplier
    const float *x_up = nullptr, *x_dn = nullptr;      // halo push: the neighbours' output buffers and their row counts
    int up_rows = 0, dn_rows = 0;
    bool operator==(const MapKey& o) const {
        return x_in == o.x_in && x_out == o.x_out && phi == o.phi && y == o.y && y1 == o.y1 && ps == o.ps &&
               B == o.B && H == o.H && W == o.W && C == o.C && own == o.own && phi_batched == o.phi_batched &&
               x_up == o.x_up && x_dn == o.x_dn && up_rows == o.up_rows && dn_rows == o.dn_rows && b_in == o.b_in;
    }
};
struct MapEntry { MapKey key; WsMaps maps; bool valid = false; unsigned long long stamp = 0; };
constexpr int kMapCache = 16;
MapEntry g_cache[kMapCache];
unsigned long long g_stamp = 0;
std::mutex g_cache_mu;

int build_maps(const MapKey& k, WsMaps* m) {
    const unsigned long long rows = (unsigned long long)k.B * k.H, prows = k.phi_batched ? rows : (unsigned long long)k.H;
    {   // frames [rows][W][C]: box = WRB rows x GW pixels x whole pixels
        unsigned long long dims[3] = {(unsigned long long)k.C, (unsigned long long)k.W, rows};
        unsigned long long str[2] = {(unsigned long long)k.C * 4, (unsigned long long)k.W * k.C * 4};
        unsigned box[3] = {(unsigned)k.C, GW, WRB};
        if (int e = make_tensor_map_f32(&m->x, k.x_in, 3, dims, str, box, 1)) return e;
        m->b = m->x;
        if (k.b_in) if (int e = make_tensor_map_f32(&m->b, k.b_in, 3, dims, str, box, 1)) return e;
        dims[2] = prows;
        if (int e = make_tensor_map_f32(&m->phi, k.phi, 3, dims, str, box, 1)) return e;
    }
    {   // measurement planes [rows][W]
        unsigned long long dims[2] = {(unsigned long long)k.W, rows};
        unsigned long long str[1] = {(unsigned long long)k.W * 4};
        unsigned box[2] = {GW, WRB};
        if (int e = make_tensor_map_f32(&m->y, k.y, 2, dims, str, box, 0)) return e;
        if (int e = make_tensor_map_f32(&m->y1, k.y1 ? k.y1 : k.y, 2, dims, str, box, 0)) return e;
        dims[1] = prows;
        if (int e = make_tensor_map_f32(&m->ps, k.ps, 2, dims, str, box, 0)) return e;
    }
    {   // output frames as [rows][C/4][W][4]: a box is one row x all chunks x the owned pixels of a group,
        // chunk-major in shared memory
        unsigned long long dims[4] = {4, (unsigned long long)k.W, (unsigned long long)k.C / 4, rows};
        unsigned long long str[3] = {(unsigned long long)k.C * 4, 16, (unsigned long long)k.W * k.C * 4};
        unsigned box[4] = {4, (unsigned)k.own, (unsigned)k.C / 4, 1};
        if (int e = make_tensor_map_f32(&m->out, k.x_out, 4, dims, str, box, 0)) return e;
        m->out_up = m->out;
        m->out_dn = m->out;
        if (k.x_up) { dims[3] = (unsigned long long)k.up_rows; if (int e = make_tensor_map_f32(&m->out_up, k.x_up, 4, dims, str, box, 0)) return e; }
        if (k.x_dn) { dims[3] = (unsigned long long)k.dn_rows; if (int e = make_tensor_map_f32(&m->out_dn, k.x_dn, 4, dims, str, box, 0)) return e; }
    }
    return SCIPNP_OK;
}

int get_maps(const MapKey& k, WsMaps* out) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    int victim = 0;
    for (int i = 0; i < kMapCache; ++i) {
        if (g_cache[i].valid && g_cache[i].key == k) {
            g_cache[i].stamp = ++g_stamp;
            *out = g_cache[i].maps;
            return SCIPNP_OK;
        }
        if (!g_cache[i].valid) victim = i;
        else if (g_cache[victim].valid && g_cache[i].stamp < g_cache[victim].stamp) victim = i;
    }
    MapEntry& e = g_cache[victim];
    e.valid = false;
    if (int rc = build_maps(k, &e.maps)) return rc;
    e.key = k;
    e.valid = true;
    e.stamp = ++g_stamp;
    *out = e.maps;
    return SCIPNP_OK;
}

}  // namespace

int fused_variant() {
    if (g_variant < 0) {
        const char* e = getenv("SCIPNP_FUSED_VARIANT");
        g_variant = e ? atoi(e) : 0;
        if (g_variant < 0 || g_variant > 2) g_variant = 0;
    }
    return g_variant;
}

bool fused_ws_supported(const FusedArgs& a) {
    if (fused_variant() == 1) return false;
    if (a.mode != MODE_GAP_ACC && a.mode != MODE_GAP_PLAIN && a.mode != MODE_TV && a.mode != MODE_ADMM) return false;
    if (a.mode == MODE_ADMM && (!a.b_in || !a.b_out || a.b_in == a.b_out || !aligned16(a.b_in) || !aligned16(a.b_out) ||
                                (a.xproj_out && !aligned16(a.xproj_out)))) return false;
    // ADMM with few channels stays on the stream kernel: the projection threads read the multiplier from global memory
    // pixel by pixel, and at C = 8 that makes them the slower side (28x256x256x8: 0.20 ms against 0.15); at C = 24
    // (3840x2160) the warp-specialised kernel wins, 1.20 ms against 1.64
    if (a.mode == MODE_ADMM && a.C < 12 && !ws_bstage(a.C / 2)) return false;
    if (a.mask2d) return false;                                   // CASSI index-offset masks: stream kernel
    if (a.clip01) return false;
    const int Q = a.C / 2;
#ifdef SCIPNP_FUSED_FAST_BUILD
    if (a.C != 8 && a.C != 24) return false;
#else
    if (a.C % 4 != 0 || a.C < 4 || a.C > 24 || Q == 0) return false;
#endif
    if (a.tv_iter_max < 3 || a.tv_iter_max > 5) return false;
    if (a.W % 4 != 0 || a.H < 1 || a.B < 1) return false;         // TMA row pitch of the planes; atomic pixel pairs
    if ((long long)a.B * a.H >= (1LL << 31)) return false;
    if (!a.x_in || !a.x_out || a.x_in == a.x_out || !aligned16(a.x_in) || !aligned16(a.x_out)) return false;
    if (a.mode == MODE_TV) return true;
    if (!aligned16(a.Phi) || !aligned16(a.y) || !aligned16(a.Phi_sum)) return false;
    if (a.mode == MODE_GAP_ACC && (!a.y1_in || !a.y1_out || !aligned16(a.y1_in))) return false;
    return true;
}

// Owned pixels per group and grid size.  The kernel deals the (batch x strips x charged rows) units out evenly
// (WsSegIter), so the grid is one CTA per SM on large scenes; small scenes get segments of about 8 units.
static void ws_split(int B, int H, int W, int Q, int cost, int* own_out, int* grid_out) {
    const int NGRP = ws_groups(Q), nsm = num_sms();
    long long best = -1;
    int bown = OWN_MAX, bgrid = 1;
    for (int own = OWN_MAX; own >= 32; own -= 4) {
        const int ngroups = (W + own - 1) / own, nstrips = (ngroups + NGRP - 1) / NGRP;
        const long long total = (long long)B * nstrips * (H + cost);
        long long grid = total / 8;          // small scenes: 256x256x8 measured 47 us at 32 units per CTA, 28 us at 8, 30 us at 4
        if (grid > nsm) grid = nsm;
        if (grid < 1) grid = 1;
        const long long per_cta = (total + grid - 1) / grid;
        if (best < 0 || per_cta < best) { best = per_cta; bown = own; bgrid = (int)grid; }
    }
    *own_out = bown;
    *grid_out = bgrid;
}

int launch_fused_ws(const FusedArgs& a, cudaStream_t st) {
    const int R = a.tv_iter_max - 1, Q = a.C / 2;
    WsParams p{};
    p.y1_out = a.y1_out;
    p.b_in = a.b_in; p.b_out = a.b_out; p.xproj_out = a.xproj_out; p.gamma = a.gamma;
    p.energy = reinterpret_cast<double*>(a.workspace);
    p.ticket = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(a.workspace) + fused_workspace_bytes(a.B, a.H, a.W, a.C, a.tv_iter_max) - 16);
    p.flag = (a.flag && R > 1) ? a.flag : nullptr;
    p.tv_eps = a.tv_eps;
    p.lambda = a.lambda;
    p.tv_c = (float)(0.25 / a.tv_weight);
    p.tv_w = (float)a.tv_weight;
    p.B = a.B; p.H = a.H; p.W = a.W; p.C = a.C;
    p.phi_batched = a.phi_batched ? 1 : 0;
    p.out_lo = a.out_hi > 0 ? a.out_lo : 0;
    p.out_hi = a.out_hi > 0 ? a.out_hi : a.H;
    if (p.out_lo < 0 || p.out_hi > a.H || p.out_lo >= p.out_hi) { set_error("bad output row window"); return SCIPNP_EINVAL; }
    p.energy_log = a.energy_log;
    int own = OWN_MAX, grid = 1;
    p.seg_cost = kWsSegCost;
    if (const char* e = getenv("SCIPNP_WS_SEGCOST")) { int v = atoi(e); if (v >= 0 && v <= 256) p.seg_cost = v; }
    ws_split(a.B, p.out_hi - p.out_lo, a.W, Q, p.seg_cost, &own, &grid);
    {
        // edge strips cost about 7 % more per row (pixel masks): their rows weigh 17/16 in the work split
        int w = 1;
        if (const char* e = getenv("SCIPNP_WS_EDGE")) { int v = atoi(e); if (v >= 0 && v <= 16) w = v; }
        const int ngr = (a.W + own - 1) / own, nst = (ngr + ws_groups(Q) - 1) / ws_groups(Q);
        p.edge_cost = nst > 2 ? w : 0;
        // rows of one CTA's share: programmatic dependent launch pays below about a hundred
        const long long per = (long long)a.B * nst * (p.out_hi - p.out_lo + p.seg_cost) / (grid > 0 ? grid : 1);
        p.use_pdl = per < 96 ? 1 : 0;
    }
    p.own = own;
    p.ngroups = (a.W + own - 1) / own;
    p.nstrips = (p.ngroups + ws_groups(Q) - 1) / ws_groups(Q);
    // several measurements whose groups do not fill the strips: lay all groups end to end (B = 28, W = 256, two groups
    // per CTA: 70 strips instead of 84)
    p.pack = (a.B > 1 && p.ngroups % ws_groups(Q) != 0 && !a.push && getenv("SCIPNP_WS_NO_PACK") == nullptr) ? 1 : 0;
    if (p.pack) {
        p.nstrips = (int)(((long long)a.B * p.ngroups + ws_groups(Q) - 1) / ws_groups(Q));
        p.edge_cost = 0;
    }
    const long long ctas = grid;
    MapKey key{a.x_in, a.x_out, a.Phi, a.y, a.mode == MODE_GAP_ACC ? a.y1_in : nullptr, a.Phi_sum,
               a.B, a.H, a.W, a.C, own, p.phi_batched};
    if (a.mode == MODE_TV) {          // the denoiser alone stages its input only; the other descriptors are never used
        key.phi = a.x_in; key.y = a.x_in; key.ps = a.x_in; key.phi_batched = 1;
    }
    if (a.mode == MODE_ADMM && ws_bstage(Q)) key.b_in = a.b_in;
    if (a.push) {
        const TilePush& t = *a.push;
        if (a.B != 1 || (!t.x_up && !t.x_dn)) { set_error("halo push: one scene, at least one neighbour"); return SCIPNP_EINVAL; }
        // the window must leave R halo rows towards every neighbour and the neighbours' halo rows must exist
        if ((t.x_up && (p.out_lo < R || p.out_lo + t.up_shift < 0 || p.out_lo + R + t.up_shift > t.up_rows)) ||
            (t.x_dn && (a.H - p.out_hi < R || p.out_hi - R + t.dn_shift < 0 || p.out_hi + t.dn_shift > t.dn_rows)) ||
            p.out_hi - p.out_lo < R) {
            set_error("halo push: row window and neighbour geometry do not match");
            return SCIPNP_EINVAL;
        }
        key.x_up = t.x_up; key.x_dn = t.x_dn; key.up_rows = t.up_rows; key.dn_rows = t.dn_rows;
        p.up_shift = t.up_shift; p.dn_shift = t.dn_shift;
        // y1 rows: my local (row, px) index + shift rows
        if (a.mode == MODE_GAP_ACC) {
            p.y1_up = t.y1_up ? t.y1_up + (long long)t.up_shift * a.W : nullptr;
            p.y1_dn = t.y1_dn ? t.y1_dn + (long long)t.dn_shift * a.W : nullptr;
            if ((t.x_up && !t.y1_up) || (t.x_dn && !t.y1_dn)) { set_error("halo push: accelerated GAP needs the neighbours' y1"); return SCIPNP_EINVAL; }
        }
        if (a.mode == MODE_ADMM) {
            p.b_up = t.b_up ? t.b_up + (long long)t.up_shift * a.W * a.C : nullptr;
            p.b_dn = t.b_dn ? t.b_dn + (long long)t.dn_shift * a.W * a.C : nullptr;
            if ((t.x_up && !t.b_up) || (t.x_dn && !t.b_dn)) { set_error("halo push: ADMM needs the neighbours' multiplier"); return SCIPNP_EINVAL; }
        }
        p.wait_up = t.x_up ? t.wait_up : nullptr; p.wait_dn = t.x_dn ? t.wait_dn : nullptr;
        p.sig_up = t.x_up ? t.sig_up : nullptr; p.sig_dn = t.x_dn ? t.sig_dn : nullptr;
        p.wait_epoch = t.wait_epoch; p.sig_epoch = t.sig_epoch;
        p.timeout_flag = t.timeout_flag;
    }
    alignas(64) WsMaps maps;
    if (int e = get_maps(key, &maps)) return e;
    if (!(a.workspace_clean && a.flag && R > 1))
        SCIPNP_CUDA(cudaMemsetAsync(a.workspace, 0, fused_workspace_bytes(a.B, a.H, a.W, a.C, a.tv_iter_max), st));
    // SCIPNP_WS_PROF=1: per-warp cycle counters of CTA 0 and of the last CTA, printed after the launch (profiling only)
    static const bool prof_on = getenv("SCIPNP_WS_PROF") != nullptr;
    long long* prof = nullptr;
    const int nwarp = ws_threads(Q) / 32;
    if (prof_on) {
        SCIPNP_CUDA(cudaMalloc(&prof, (size_t)ctas * nwarp * 4 * sizeof(long long)));
        SCIPNP_CUDA(cudaMemsetAsync(prof, 0, (size_t)ctas * nwarp * 4 * sizeof(long long), st));
        p.prof = prof;
    }
    int rc;
    switch (R) {
        case 2: rc = ws_launch_r<2>(a.mode, Q, p, maps, (int)ctas, st); break;
        case 3: rc = ws_launch_r<3>(a.mode, Q, p, maps, (int)ctas, st); break;
        case 4: rc = ws_launch_r<4>(a.mode, Q, p, maps, (int)ctas, st); break;
        default: set_error("unsupported tv_iter_max"); return SCIPNP_EINVAL;
    }
    if (rc) return rc;
    count_launch();
    if (prof_on) {
        std::vector<long long> h((size_t)ctas * nwarp * 4);
        SCIPNP_CUDA(cudaStreamSynchronize(st));
        SCIPNP_CUDA(cudaMemcpy(h.data(), prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(prof);
        static int printed = 0;
        if (printed++ % 8 == 3) {
            const int ncw = ws_consumers(Q);
            std::vector<std::pair<long long, long long>> dur;
            for (long long c = 0; c < ctas; ++c) dur.push_back({h[((size_t)c * nwarp) * 4], c});
            std::sort(dur.begin(), dur.end());
            fprintf(stderr, "[ws prof] consumer-0 cycles over %lld CTAs: min %lld (cta %lld)  median %lld  max %lld (cta %lld); slowest:",
                    ctas, dur.front().first, dur.front().second, dur[dur.size() / 2].first, dur.back().first, dur.back().second);
            for (size_t i = dur.size() > 6 ? dur.size() - 6 : 0; i < dur.size(); ++i) fprintf(stderr, " %lld:%lld", dur[i].second, dur[i].first);
            fprintf(stderr, "\n");
            for (long long cta : {0LL, ctas / 2, ctas - 1}) {
                fprintf(stderr, "[ws prof] cta %lld own=%d grid=%lld\n", cta, own, ctas);
                for (int w = 0; w < nwarp; ++w) {
                    const long long* r = &h[((size_t)cta * nwarp + w) * 4];
                    if (WS_PROD_FIRST ? w >= NPROD : w < ncw) fprintf(stderr, "  consumer %2d: total %9lld  wait f_full %9lld  wait out_empty %9lld\n", w, r[0], r[1], r[2]);
                    else fprintf(stderr, "  producer %2d: total %9lld  wait raw %9lld  wait f_empty(+stores) %9lld  bar %9lld\n", w, r[0], r[1], r[2], r[3]);
                }
            }
        }
    }
    return check_launch("gap_tv_ws_kernel");
}

// buffers that were freed must not be found in the descriptor cache by a later allocation at the same address
// with different contents of the descriptor (same key => same descriptor, so this is only hygiene)
void fused_ws_forget(const void* base) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (auto& e : g_cache)
        if (e.valid && (e.key.x_in == base || e.key.x_out == base || e.key.phi == base || e.key.x_up == base || e.key.x_dn == base || e.key.b_in == base)) e.valid = false;
}

}  // namespace scipnp

extern "C" {

size_t scipnp_tv_fused_workspace_bytes(int B, int H, int W, int C, int n_iter_max) {
    if (B < 1 || H < 1 || W < 1 || C < 1 || n_iter_max < 1) return 0;
    return scipnp::fused_workspace_bytes(B, H, W, C, n_iter_max);
}

int scipnp_tv_fused_supported(int B, int H, int W, int C, int n_iter_max) {
    scipnp::FusedArgs a{};
    a.mode = scipnp::MODE_TV;
    a.B = B; a.H = H; a.W = W; a.C = C; a.tv_iter_max = n_iter_max;
    a.x_in = reinterpret_cast<const float*>(16); a.x_out = reinterpret_cast<float*>(32);      // alignment probes only
    return scipnp::fused_ws_supported(a) ? 1 : 0;
}

int scipnp_tv_chambolle_fused(const float* in, float* out, double weight, double eps, int n_iter_max, int B, int H,
                              int W, int C, void* workspace, size_t workspace_bytes, int* flag_dev, void* stream) {
    using namespace scipnp;
    SCIPNP_REQUIRE(B >= 1 && H >= 1 && W >= 1 && C >= 1 && B <= 65535, "bad dimensions");
    SCIPNP_REQUIRE(in && out && workspace, "null pointer");
    SCIPNP_REQUIRE(in != out, "in and out must not alias");
    SCIPNP_REQUIRE(weight > 0.0, "weight must be positive");
    FusedArgs a{};
    a.mode = MODE_TV;
    a.x_in = in; a.x_out = out;
    a.tv_weight = weight; a.tv_eps = eps; a.tv_iter_max = n_iter_max;
    a.B = B; a.H = H; a.W = W; a.C = C; a.phi_batched = 1;
    a.workspace = workspace; a.workspace_bytes = workspace_bytes;
    a.flag = flag_dev;
    if (!fused_ws_supported(a)) {
        set_error("one-pass TV: needs C %% 4 == 0, C <= 24, W %% 4 == 0, n_iter_max 3..5, 16-byte aligned arrays "
                  "(use scipnp_tv_chambolle)");
        return SCIPNP_EINVAL;
    }
    SCIPNP_REQUIRE(workspace_bytes >= fused_workspace_bytes(B, H, W, C, n_iter_max), "workspace too small");
    return launch_fused_ws(a, (cudaStream_t)stream);
}

// 0: automatic (warp-specialised kernel where it applies), 1: stream kernel only, 2: same as 0
int scipnp_set_fused_variant(int v) {
    if (v < 0 || v > 2) { scipnp::set_error("fused variant must be 0, 1 or 2"); return SCIPNP_EINVAL; }
    scipnp::g_variant = v;
    return SCIPNP_OK;
}

}  // extern "C"
