// Persistent GAP-TV / ADMM-TV solver (R4 gap_denoise, R5 admm_denoise with
// denoiser='tv'): every array of the reference's loop (pnp_sci_algo.py:638-650,
// 805-836) lives in HBM for the whole reconstruction.
//
// Two execution paths, chosen by scipnp_params.fused:
//   fused = 1  one kernel per outer iteration (gap_tv_fused.cu), state ping-pongs
//              between two buffers.  The kernel cannot apply skimage's energy
//              early stop (a global per-slice reduction per dual iteration), so
//              it evaluates the criterion and raises a flag; if any iteration of
//              a run raised it the run is rolled back to its snapshot and redone
//              on the exact path, so results never depend on the path taken.
//   fused = 0  exact path: projection kernel + per-iteration TV kernels.
#include <math.h>
#include <mutex>
#include <new>
#include <vector>

#include "internal.cuh"

using namespace scipnp;

namespace {
constexpr int kPsnrCap = 1 << 16;
constexpr int kElogCap = 1024;      // iterations per run whose TV energies are kept for the cross-rank stopping rule
constexpr int kSyncInts = 16;
}

struct scipnp_solver {
    scipnp_params p;
    size_t n_frame = 0;   // B*H*W*C
    size_t n_meas = 0;    // B*H*W
    size_t n_phi = 0, n_phisum = 0;
    // device buffers
    float *xa = nullptr, *xb = nullptr;       // GAP: x ping-pong.  ADMM: theta ping-pong
    float *y1a = nullptr, *y1b = nullptr;     // GAP accelerated
    float *ba = nullptr, *bb = nullptr;       // ADMM multiplier ping-pong
    float *xproj = nullptr, *fbuf = nullptr;  // ADMM x and TV input
    float *y = nullptr, *Phi = nullptr, *PhiSum = nullptr, *Xorig = nullptr;
    const float* phi = nullptr;               // the mask stack the iterations read: Phi (owned copy) or the caller's array
    bool sqerr_dirty = true;                  // the squared-error track holds values of an earlier reconstruction
    float *xsnap = nullptr, *y1snap = nullptr, *bsnap = nullptr;   // rollback copies (fused)
    void* tvws = nullptr; size_t tvws_bytes = 0;
    void* fws = nullptr; size_t fws_bytes = 0; bool fws_clean = false;
    double* sqerr = nullptr;   // [kPsnrCap]
    int* flags = nullptr;      // [1]
    // host state
    bool loaded = false, has_orig = false, use_fused = false;
    bool x0_given = false;     // load() got an initial guess (else x0 = At(y) can be recomputed)
    bool snap_valid = false;   // begin() took a snapshot (else the run started from the load state)
    int iters_done = 0, psnr_count = 0, refined = 0, begin_iter = 0;
    bool fused_possible = false;
    // CASSI: coded aperture [H][mask_w]; the fused kernel reads it at per-band offsets
    bool cassi = false; float* mask2d = nullptr; int cassi_step = 0, mask_w = 0;
    long long launches0 = 0;
    std::vector<void*> owned;
    // ---- row-tiled multi-GPU mode: direct peer access to the neighbours' buffers ----
    struct PeerLink {
        bool present = false;
        float* x[2] = {nullptr, nullptr};     // the neighbour's two x buffers (IPC-mapped)
        float* y1[2] = {nullptr, nullptr};
        int* sync = nullptr;                  // the neighbour's flag block
        int row_lo = 0;                       // global row of the neighbour's local row 0
        int rows = 0;                         // the neighbour's local row count (halo push)
        void* mapped[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    };
    bool tiled = false;
    int t_lo = 0, t_hi = 0, t_row_lo = 0, t_row_hi = 0;
    float* xbuf[2] = {nullptr, nullptr};      // identity of my own two x / y1 buffers
    float* y1buf[2] = {nullptr, nullptr};
    int* sync = nullptr;                      // [0] ready<-up [1] ready<-down [2] ack<-up [3] ack<-down [4] timeout
                                              // [5] pull counter [8] pushed<-up [9] pushed<-down
    int epoch = 0;
    bool ack_pending = false;
    // halo push (gap_tv_ws.cuh): one exchange per iteration inside the fused kernel, no exchange kernel
    bool push_enabled = false;
    int push_epoch = 0;                       // fused push steps so far = the value every rank's flags reach
    double* elog = nullptr;                   // [kElogCap][B*C*R] energies of the iterations since begin()
    int elog_count = 0;
    TvTiling tv_tiling;                       // owned rows + cross-rank energy sum for the exact path's stopping rule
    PeerLink up, dn;

    float* x_cur() { return xa; }
};

static int dmalloc(scipnp_solver* s, void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    s->owned.push_back(*p);
    return SCIPNP_OK;
}
#define DM(ptr, count, type)                                                   \
    do {                                                                       \
        if (int e_ = dmalloc(s, (void**)&(ptr), (size_t)(count) * sizeof(type))) { \
            scipnp_solver_destroy(s);                                          \
            return e_;                                                         \
        }                                                                      \
    } while (0)

extern "C" {

int scipnp_solver_destroy(scipnp_solver* s) {
    if (!s) return SCIPNP_OK;
    for (scipnp_solver::PeerLink* l : {&s->up, &s->dn})
        for (void* m : l->mapped)
            if (m) cudaIpcCloseMemHandle(m);
    for (void* p : s->owned) cudaFree(p);
    delete s;
    return SCIPNP_OK;
}

int scipnp_solver_create(const scipnp_params* pp, scipnp_solver** out) {
    SCIPNP_REQUIRE(pp && out, "null pointer");
    const scipnp_params& p = *pp;
    SCIPNP_REQUIRE(p.B >= 1 && p.H >= 1 && p.W >= 1 && p.C >= 1 && p.B <= 65535, "bad dimensions");
    SCIPNP_REQUIRE(p.method == 0 || p.method == 1, "method must be 0 (GAP) or 1 (ADMM)");
    SCIPNP_REQUIRE(p.tv_weight > 0.0, "tv_weight must be positive");
    SCIPNP_REQUIRE(p.tv_iter_max >= 1, "tv_iter_max must be >= 1");
    if (scipnp_device_count() < 1) {
        set_error("no CUDA device: libscipnp has no CPU fallback");
        return SCIPNP_ECUDA;
    }
    scipnp_solver* s = new (std::nothrow) scipnp_solver();
    if (!s) { set_error("out of host memory"); return SCIPNP_ENOMEM; }
    s->p = p;
    s->n_meas = (size_t)p.B * p.H * p.W;
    s->n_frame = s->n_meas * p.C;
    s->n_phisum = p.phi_batched ? s->n_meas : (size_t)p.H * p.W;
    s->n_phi = s->n_phisum * p.C;
    const int mode = p.method == 1 ? MODE_ADMM : (p.accelerate ? MODE_GAP_ACC : MODE_GAP_PLAIN);
    s->use_fused = p.fused && fused_supported(mode, p.B, p.H, p.W, p.C, p.tv_iter_max) &&
                   !(p.clip01 && p.method == 0);          // the fused kernel clips in ADMM mode only
    s->fused_possible = s->use_fused;
    DM(s->xa, s->n_frame, float);
    DM(s->xb, s->n_frame, float);
    DM(s->y, s->n_meas, float);
    DM(s->Phi, s->n_phi, float);
    DM(s->PhiSum, s->n_phisum, float);
    DM(s->Xorig, s->n_frame, float);
    DM(s->sqerr, kPsnrCap, double);
    DM(s->flags, 4, int);
    if (p.method == 0) {
        DM(s->y1a, s->n_meas, float);
        if (s->use_fused) DM(s->y1b, s->n_meas, float);
    } else {
        DM(s->ba, s->n_frame, float);
        DM(s->xproj, s->n_frame, float);
        if (s->use_fused) DM(s->bb, s->n_frame, float);
        else DM(s->fbuf, s->n_frame, float);
    }
    if (s->use_fused) {
        DM(s->xsnap, s->n_frame, float);
        if (p.method == 0) DM(s->y1snap, s->n_meas, float);
        else DM(s->bsnap, s->n_frame, float);
        s->fws_bytes = fused_workspace_bytes(p.B, p.H, p.W, p.C, p.tv_iter_max);
        DM(s->fws, s->fws_bytes, char);
    }
    s->xbuf[0] = s->xa; s->xbuf[1] = s->xb;
    // the second carried array of the iteration: y1 (accelerated GAP) or the multiplier b (ADMM)
    s->y1buf[0] = p.method == 0 ? s->y1a : s->ba; s->y1buf[1] = p.method == 0 ? s->y1b : s->bb;
    // exact-path workspace is allocated lazily (only the exact path or a rollback needs it)
    s->launches0 = g_launches.load();
    *out = s;
    return SCIPNP_OK;
}

static int ensure_exact_buffers(scipnp_solver* s) {
    const scipnp_params& p = s->p;
    if (!s->tvws) {
        s->tvws_bytes = tv_workspace_bytes(p.B, p.H, p.W, p.C);
        if (int e = dmalloc(s, &s->tvws, s->tvws_bytes)) return e;
    }
    if (p.method == 1 && !s->fbuf)
        if (int e = dmalloc(s, (void**)&s->fbuf, s->n_frame * sizeof(float))) return e;
    return SCIPNP_OK;
}

static int solver_load(scipnp_solver* s, const float* y, const float* Phi, const float* Phi_sum,
                       const float* x0, const float* X_orig, bool borrow_phi, void* stream) {
    SCIPNP_REQUIRE(s && y && Phi, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const scipnp_params& p = s->p;
    s->cassi = false;
    SCIPNP_CUDA(cudaMemcpyAsync(s->y, y, s->n_meas * sizeof(float), cudaMemcpyDefault, st));
    if (borrow_phi) {
        // the caller's device array is read in place for the whole reconstruction (it must stay valid and
        // unchanged until the results were read); saves one pass over the mask stack per reconstruction
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, Phi) != cudaSuccess || at.type != cudaMemoryTypeDevice || !aligned16(Phi)) {
            cudaGetLastError();
            set_error("borrowed Phi must be 16-byte aligned device memory");
            return SCIPNP_EINVAL;
        }
        s->phi = Phi;
    } else {
        if (Phi != s->Phi) SCIPNP_CUDA(cudaMemcpyAsync(s->Phi, Phi, s->n_phi * sizeof(float), cudaMemcpyDefault, st));
        s->phi = s->Phi;
    }
    bool need_sum = Phi_sum == nullptr, need_x0 = x0 == nullptr;
    if (Phi_sum) SCIPNP_CUDA(cudaMemcpyAsync(s->PhiSum, Phi_sum, s->n_phisum * sizeof(float), cudaMemcpyDefault, st));
    if (x0) SCIPNP_CUDA(cudaMemcpyAsync(s->xa, x0, s->n_frame * sizeof(float), cudaMemcpyDefault, st));
    if (need_sum && need_x0) {           // both in one pass over the masks where the shape allows it
        const int e = launch_init_x0_phisum(s->y, s->phi, s->xa, s->PhiSum, p.B, p.H, p.W, p.C, p.phi_batched, st);
        if (e == SCIPNP_OK) need_sum = need_x0 = false;
        else if (e != SCIPNP_EINVAL) return e;
    }
    if (need_sum)
        if (int e = scipnp_phi_sum(s->phi, s->PhiSum, p.phi_batched ? p.B : 1, p.H, p.W, p.C, stream)) return e;
    if (need_x0)
        if (int e = scipnp_At(s->y, s->phi, s->xa, p.B, p.H, p.W, p.C, p.phi_batched, stream)) return e;
    s->x0_given = x0 != nullptr;
    s->has_orig = X_orig != nullptr;
    if (X_orig) SCIPNP_CUDA(cudaMemcpyAsync(s->Xorig, X_orig, s->n_frame * sizeof(float), cudaMemcpyDefault, st));
    if (p.method == 0) {
        SCIPNP_CUDA(cudaMemsetAsync(s->y1a, 0, s->n_meas * sizeof(float), st));
    } else {
        SCIPNP_CUDA(cudaMemsetAsync(s->ba, 0, s->n_frame * sizeof(float), st));
        // x = x0 until the first projection (pnp_sci_algo.py:800)
        SCIPNP_CUDA(cudaMemcpyAsync(s->xproj, s->xa, s->n_frame * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (s->has_orig || s->sqerr_dirty) SCIPNP_CUDA(cudaMemsetAsync(s->sqerr, 0, kPsnrCap * sizeof(double), st));
    s->sqerr_dirty = s->has_orig;
    s->loaded = true;
    s->iters_done = 0;
    s->psnr_count = 0;
    s->refined = 0;
    return SCIPNP_OK;
}

int scipnp_solver_load(scipnp_solver* s, const float* y, const float* Phi, const float* Phi_sum,
                       const float* x0, const float* X_orig, void* stream) {
    return solver_load(s, y, Phi, Phi_sum, x0, X_orig, false, stream);
}

// The same with the mask stack BORROWED: `Phi_dev` (device memory, 16-byte aligned) is read in place by every
// iteration instead of being copied into the handle; it must stay valid and unchanged until the results were read.
int scipnp_solver_load_borrow_phi(scipnp_solver* s, const float* y, const float* Phi_dev, const float* Phi_sum,
                                  const float* x0, const float* X_orig, void* stream) {
    return solver_load(s, y, Phi_dev, Phi_sum, x0, X_orig, true, stream);
}

// R9: CASSI.  `mask2d` is the coded aperture [H][W-(C-1)*step]; the solver's W is the sheared canvas.
int scipnp_solver_load_cassi(scipnp_solver* s, const float* y, const float* mask2d, int step,
                             const float* x0, const float* X_orig, void* stream) {
    SCIPNP_REQUIRE(s && y && mask2d, "null pointer");
    const scipnp_params& p = s->p;
    SCIPNP_REQUIRE(p.B == 1 && !p.phi_batched, "CASSI mode takes one measurement (B = 1)");
    SCIPNP_REQUIRE(step >= 0 && p.W - (p.C - 1) * step >= 1, "canvas narrower than the dispersion");
    cudaStream_t st = (cudaStream_t)stream;
    const int mw = p.W - (p.C - 1) * step;
    if (!s->mask2d || s->mask_w != mw) {
        if (int e = dmalloc(s, (void**)&s->mask2d, (size_t)p.H * mw * sizeof(float))) return e;
    }
    s->mask_w = mw;
    s->cassi_step = step;
    SCIPNP_CUDA(cudaMemcpyAsync(s->mask2d, mask2d, (size_t)p.H * mw * sizeof(float), cudaMemcpyDefault, st));
    // the explicit stack is built once (initial guess, Phi_sum, exact-path fallback); the fused
    // iterations never read it
    if (int e = scipnp_cassi_shift_mask(s->mask2d, s->Phi, p.H, mw, p.C, step, stream)) return e;
    if (int e = scipnp_solver_load(s, y, s->Phi, nullptr, x0, X_orig, stream)) return e;
    const int mode = p.accelerate ? MODE_GAP_ACC : MODE_GAP_PLAIN;
    s->cassi = p.method == 0 && fused_cassi_supported(mode, p.B, p.H, p.W, p.C, p.tv_iter_max);
    return SCIPNP_OK;
}

// sum of squared errors of iteration k against X_orig, one per batch element
static int record_sqerr(scipnp_solver* s, int k, const float* x, cudaStream_t st) {
    if (!s->has_orig || (size_t)(k + 1) * s->p.B > (size_t)kPsnrCap) return SCIPNP_OK;
    return launch_sq_err(s->Xorig, x, s->n_frame / s->p.B, s->p.B, s->sqerr + (size_t)k * s->p.B, st);
}

// one outer iteration on the exact path
static int step_exact(scipnp_solver* s, int k, cudaStream_t st) {
    const scipnp_params& p = s->p;
    if (p.method == 0) {
        int mode = p.accelerate ? MODE_GAP_ACC : MODE_GAP_PLAIN;
        if (int e = launch_project(mode, s->xa, nullptr, s->xa, nullptr, s->y1a, s->y1a, s->y, s->phi,
                                   s->PhiSum, p.lambda, 0.f, p.B, p.H, p.W, p.C, p.phi_batched, st)) return e;
        if (int e = tv_chambolle_exact(s->xa, s->xb, p.tv_weight, p.tv_eps, p.tv_iter_max, p.B, p.H, p.W,
                                       p.C, s->tvws, s->tvws_bytes, nullptr, nullptr, 0, st,
                                       s->tiled && s->tv_tiling.reduce ? &s->tv_tiling : nullptr)) return e;
        std::swap(s->xa, s->xb);
        if (p.clip01) if (int e = launch_clip01(s->xa, s->n_frame, st)) return e;
        if (int e = record_sqerr(s, k, s->xa, st)) return e;
    } else {
        if (int e = launch_project(MODE_ADMM, s->xa, s->ba, s->xproj, s->fbuf, nullptr, nullptr, s->y,
                                   s->phi, s->PhiSum, p.lambda, p.gamma, p.B, p.H, p.W, p.C,
                                   p.phi_batched, st)) return e;
        if (int e = tv_chambolle_exact(s->fbuf, s->xa, p.tv_weight, p.tv_eps, p.tv_iter_max, p.B, p.H,
                                       p.W, p.C, s->tvws, s->tvws_bytes, nullptr, nullptr, 0, st,
                                       s->tiled && s->tv_tiling.reduce ? &s->tv_tiling : nullptr)) return e;
        if (p.clip01) if (int e = launch_clip01(s->xa, s->n_frame, st)) return e;
        if (int e = scipnp_admm_dual_update(s->ba, s->xproj, s->xa, s->n_frame, st)) return e;
        if (int e = record_sqerr(s, k, s->xproj, st)) return e;
    }
    return SCIPNP_OK;
}

static int step_fused(scipnp_solver* s, int k, bool last, cudaStream_t st) {
    const scipnp_params& p = s->p;
    FusedArgs a{};
    a.x_in = s->xa; a.x_out = s->xb;
    a.y = s->y; a.Phi = s->phi; a.Phi_sum = s->PhiSum;
    if (s->cassi) { a.Phi = nullptr; a.mask2d = s->mask2d; a.cassi_step = s->cassi_step; a.mask_w = s->mask_w; }
    a.lambda = p.lambda; a.gamma = p.gamma;
    a.tv_weight = p.tv_weight; a.tv_eps = p.tv_eps; a.tv_iter_max = p.tv_iter_max;
    a.B = p.B; a.H = p.H; a.W = p.W; a.C = p.C; a.phi_batched = p.phi_batched;
    a.workspace = s->fws; a.workspace_bytes = s->fws_bytes;
    a.flag = s->flags;
    a.clip01 = p.clip01;
    a.workspace_clean = s->fws_clean;
    if (p.method == 0) {
        a.mode = p.accelerate ? MODE_GAP_ACC : MODE_GAP_PLAIN;
        a.y1_in = s->y1a; a.y1_out = s->y1b;
    } else {
        a.mode = MODE_ADMM;
        a.b_in = s->ba; a.b_out = s->bb; a.xproj_out = s->xproj;
        // x (the projection output) is only read by the PSNR track and by the caller after the last step; the
        // warp-specialised kernel can leave it out (the stream kernel always writes it)
        if (!(s->has_orig || last)) {
            a.xproj_out = nullptr;
            if (!fused_ws_supported(a)) a.xproj_out = s->xproj;
        }
    }
    TilePush tp{};
    const bool push = s->tiled && s->push_enabled && (s->up.present || s->dn.present);
    if (s->tiled && s->push_enabled) {
        // owned rows only; the halo rows of the output buffers are written by the neighbours
        a.out_lo = s->t_lo - s->t_row_lo;
        a.out_hi = s->t_hi - s->t_row_lo;
        const int slot = k - s->begin_iter;
        if (s->elog && slot >= 0 && slot < kElogCap) {
            a.energy_log = s->elog + (size_t)slot * p.B * p.C * (p.tv_iter_max - 1);
            if (slot + 1 > s->elog_count) s->elog_count = slot + 1;
        }
    }
    if (push) {
        const bool admm = p.method == 1;
        const int xo = s->xb == s->xbuf[0] ? 0 : 1, yo = (admm ? s->bb : s->y1b) == s->y1buf[0] ? 0 : 1;
        if (s->up.present) {
            tp.x_up = s->up.x[xo];
            tp.y1_up = (!admm && p.accelerate) ? s->up.y1[yo] : nullptr;
            tp.b_up = admm ? s->up.y1[yo] : nullptr;
            tp.up_rows = s->up.rows; tp.up_shift = s->t_row_lo - s->up.row_lo;
            tp.wait_up = s->sync + 8; tp.sig_up = s->up.sync + 9;
        }
        if (s->dn.present) {
            tp.x_dn = s->dn.x[xo];
            tp.y1_dn = (!admm && p.accelerate) ? s->dn.y1[yo] : nullptr;
            tp.b_dn = admm ? s->dn.y1[yo] : nullptr;
            tp.dn_rows = s->dn.rows; tp.dn_shift = s->t_row_lo - s->dn.row_lo;
            tp.wait_dn = s->sync + 9; tp.sig_dn = s->dn.sync + 8;
        }
        tp.wait_epoch = s->push_epoch; tp.sig_epoch = s->push_epoch + 1;
        tp.timeout_flag = s->sync + 4;
        a.push = &tp;
    }
    if (int e = launch_fused(a, st)) return e;
    if (push) ++s->push_epoch;
    s->fws_clean = true;            // the check kernel re-zeroes the accumulators
    std::swap(s->xa, s->xb);
    if (p.method == 0) { if (p.accelerate) std::swap(s->y1a, s->y1b); }
    else std::swap(s->ba, s->bb);
    return record_sqerr(s, k, p.method == 0 ? s->xa : s->xproj, st);
}

// ---- run = begin + step_async + commit; the pieces are public so that a caller that
//      interleaves its own work between iterations (the row-tiled multi-GPU driver exchanges
//      halos) can keep everything asynchronous and still get the exact-path guarantee.

int scipnp_solver_begin(scipnp_solver* s, void* stream) {
    SCIPNP_REQUIRE(s, "null solver");
    if (!s->loaded) { set_error("scipnp_solver_begin before scipnp_solver_load"); return SCIPNP_ESTATE; }
    cudaStream_t st = (cudaStream_t)stream;
    s->begin_iter = s->iters_done;
    s->elog_count = 0;
    if (!s->xsnap) return SCIPNP_OK;          // handle created without the fused path: nothing to roll back
    // A run that starts right after load() with the default initial guess needs no snapshot: a
    // rollback recomputes x0 = At(y) and clears y1 / b (saves copying the state per reconstruction).
    s->snap_valid = !(s->iters_done == 0 && !s->x0_given && !s->cassi);
    if (s->snap_valid) {
        SCIPNP_CUDA(cudaMemcpyAsync(s->xsnap, s->xa, s->n_frame * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (s->p.method == 0) SCIPNP_CUDA(cudaMemcpyAsync(s->y1snap, s->y1a, s->n_meas * sizeof(float), cudaMemcpyDeviceToDevice, st));
        else SCIPNP_CUDA(cudaMemcpyAsync(s->bsnap, s->ba, s->n_frame * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    SCIPNP_CUDA(cudaMemsetAsync(s->flags, 0, 4 * sizeof(int), st));
    return SCIPNP_OK;
}

int scipnp_solver_step_async(scipnp_solver* s, int iters, void* stream) {
    SCIPNP_REQUIRE(s, "null solver");
    SCIPNP_REQUIRE(iters >= 0, "negative iteration count");
    if (!s->loaded) { set_error("scipnp_solver_step_async before scipnp_solver_load"); return SCIPNP_ESTATE; }
    cudaStream_t st = (cudaStream_t)stream;
    const scipnp_params& p = s->p;
    if (s->has_orig && (long long)(s->iters_done + iters) * p.B > kPsnrCap) {
        set_error("psnr_all holds %d values: %d iterations x %d measurements do not fit (run without X_orig)",
                  kPsnrCap, s->iters_done + iters, p.B);
        return SCIPNP_EINVAL;
    }
    if (!s->use_fused)
        if (int e = ensure_exact_buffers(s)) return e;
    for (int i = 0; i < iters; ++i)
        if (int e = s->use_fused ? step_fused(s, s->iters_done + i, i + 1 == iters, st) : step_exact(s, s->iters_done + i, st)) return e;
    s->iters_done += iters;
    if (s->has_orig) s->psnr_count = s->iters_done < kPsnrCap / p.B ? s->iters_done : kPsnrCap / p.B;
    return SCIPNP_OK;
}

int scipnp_solver_fired(scipnp_solver* s, int* fired, void* stream) {
    SCIPNP_REQUIRE(s && fired, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int flag = 0;
    if (s->xsnap) SCIPNP_CUDA(cudaMemcpyAsync(&flag, s->flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCIPNP_CUDA(cudaStreamSynchronize(st));
    *fired = flag;
    return SCIPNP_OK;
}

int scipnp_solver_rollback(scipnp_solver* s, void* stream) {
    SCIPNP_REQUIRE(s, "null solver");
    if (!s->xsnap) { set_error("this handle keeps no snapshot (created with fused = 0)"); return SCIPNP_ESTATE; }
    cudaStream_t st = (cudaStream_t)stream;
    const scipnp_params& p = s->p;
    if (s->snap_valid) {
        SCIPNP_CUDA(cudaMemcpyAsync(s->xa, s->xsnap, s->n_frame * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (p.method == 0) SCIPNP_CUDA(cudaMemcpyAsync(s->y1a, s->y1snap, s->n_meas * sizeof(float), cudaMemcpyDeviceToDevice, st));
        else SCIPNP_CUDA(cudaMemcpyAsync(s->ba, s->bsnap, s->n_frame * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {                                   // the load state (scipnp_solver_load)
        if (int e = scipnp_At(s->y, s->phi, s->xa, p.B, p.H, p.W, p.C, p.phi_batched, stream)) return e;
        if (p.method == 0) {
            SCIPNP_CUDA(cudaMemsetAsync(s->y1a, 0, s->n_meas * sizeof(float), st));
        } else {
            SCIPNP_CUDA(cudaMemsetAsync(s->ba, 0, s->n_frame * sizeof(float), st));
            SCIPNP_CUDA(cudaMemcpyAsync(s->xproj, s->xa, s->n_frame * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
    }
    const int k0 = s->begin_iter, capi = kPsnrCap / p.B;
    const int n = s->iters_done - k0 < capi - k0 ? s->iters_done - k0 : (capi - k0 > 0 ? capi - k0 : 0);
    if (n > 0) SCIPNP_CUDA(cudaMemsetAsync(s->sqerr + (size_t)k0 * p.B, 0, (size_t)n * p.B * sizeof(double), st));
    SCIPNP_CUDA(cudaMemsetAsync(s->flags, 0, 4 * sizeof(int), st));
    s->iters_done = k0;
    return SCIPNP_OK;
}

// TV weight and ADMM regulariser of the iterations that follow (ADMM_TV_rec, pnp_sci_algo.py:898-899, shrinks both
// every iteration); the other parameters stay as created.
int scipnp_solver_set_tv(scipnp_solver* s, double tv_weight, double gamma) {
    SCIPNP_REQUIRE(s, "null solver");
    SCIPNP_REQUIRE(tv_weight > 0.0, "tv_weight must be positive");
    s->p.tv_weight = tv_weight;
    s->p.gamma = (float)gamma;
    return SCIPNP_OK;
}

int scipnp_solver_set_path(scipnp_solver* s, int fused) {
    SCIPNP_REQUIRE(s, "null solver");
    if (fused && !s->fused_possible) { set_error("the fused path does not cover this configuration"); return SCIPNP_ESTATE; }
    s->use_fused = fused != 0;
    return SCIPNP_OK;
}

int scipnp_solver_run(scipnp_solver* s, int iters, void* stream) {
    SCIPNP_REQUIRE(s, "null solver");
    SCIPNP_REQUIRE(iters >= 0, "negative iteration count");
    if (!s->loaded) { set_error("scipnp_solver_run before scipnp_solver_load"); return SCIPNP_ESTATE; }
    if (!s->use_fused) return scipnp_solver_step_async(s, iters, stream);
    if (int e = scipnp_solver_begin(s, stream)) return e;
    if (int e = scipnp_solver_step_async(s, iters, stream)) return e;
    int fired = 0;
    if (int e = scipnp_solver_fired(s, &fired, stream)) return e;
    if (fired) {
        // the reference would have stopped a TV slice early somewhere in this run:
        // redo the run on the exact path from the snapshot
        if (int e = scipnp_solver_rollback(s, stream)) return e;
        s->use_fused = false;
        int e = scipnp_solver_step_async(s, iters, stream);
        s->use_fused = true;
        if (e) return e;
        s->refined += iters;
    }
    return SCIPNP_OK;
}

int scipnp_solver_get_x(scipnp_solver* s, float* x_out, void* stream) {
    SCIPNP_REQUIRE(s && x_out, "null pointer");
    if (!s->loaded) { set_error("solver has no inputs"); return SCIPNP_ESTATE; }
    cudaStream_t st = (cudaStream_t)stream;
    const float* src = s->p.method == 0 ? s->xa : s->xproj;
    SCIPNP_CUDA(cudaMemcpyAsync(x_out, src, s->n_frame * sizeof(float), cudaMemcpyDefault, st));
    SCIPNP_CUDA(cudaStreamSynchronize(st));
    return SCIPNP_OK;
}

int scipnp_solver_sqerr(scipnp_solver* s, double* sums, int cap, int* count, void* stream) {
    SCIPNP_REQUIRE(s && count, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int total = s->psnr_count * s->p.B;
    *count = total;
    int n = total < cap ? total : cap;
    if (n <= 0 || !sums) return SCIPNP_OK;
    SCIPNP_CUDA(cudaMemcpyAsync(sums, s->sqerr, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    SCIPNP_CUDA(cudaStreamSynchronize(st));
    return SCIPNP_OK;
}

int scipnp_solver_psnr(scipnp_solver* s, double* psnr_all, int cap, int* count, void* stream) {
    if (int e = scipnp_solver_sqerr(s, psnr_all, cap, count, stream)) return e;
    int n = *count < cap ? *count : cap;
    if (!psnr_all) return SCIPNP_OK;
    const double per = (double)(s->n_frame / s->p.B);
    for (int i = 0; i < n; ++i) {
        double mse = psnr_all[i] / per;
        psnr_all[i] = (mse == 0.0) ? 100.0 : 20.0 * log10(1.0 / sqrt(mse));   // utils.py:32-36
    }
    return SCIPNP_OK;
}

int scipnp_solver_add_refined(scipnp_solver* s, int iters) {
    SCIPNP_REQUIRE(s, "null solver");
    s->refined += iters;
    return SCIPNP_OK;
}

int scipnp_solver_refined_iters(scipnp_solver* s, int* count) {
    SCIPNP_REQUIRE(s && count, "null pointer");
    *count = s->refined;
    return SCIPNP_OK;
}

int scipnp_solver_state(scipnp_solver* s, float** x_cur, float** y1_cur) {
    SCIPNP_REQUIRE(s, "null solver");
    if (x_cur) *x_cur = s->xa;
    if (y1_cur) *y1_cur = s->p.method == 0 ? s->y1a : s->ba;
    return SCIPNP_OK;
}

int scipnp_solver_admm_state(scipnp_solver* s, float** theta, float** b, float** x) {
    SCIPNP_REQUIRE(s, "null solver");
    if (s->p.method != 1) { set_error("not an ADMM handle"); return SCIPNP_ESTATE; }
    if (theta) *theta = s->xa;
    if (b) *b = s->ba;
    if (x) *x = s->xproj;
    return SCIPNP_OK;
}

int scipnp_solver_uses_fused(scipnp_solver* s) { return s && s->use_fused ? 1 : 0; }

long long scipnp_solver_launch_count(scipnp_solver* s) { return s ? g_launches.load() - s->launches0 : 0; }

// ---------------------------------------------------------------------------------------------
// Row-tiled multi-GPU mode ("single UHD scene"): this handle holds rows
// [row_lo, row_hi) of a taller scene and owns [lo, hi); the rows in between are halo copies of the
// neighbours' owned rows.  Neighbours are other processes on the same node: their buffers are
// mapped with CUDA IPC and the halo rows are pulled straight over NVLink, ordered by flags that
// the ranks write into each other's memory.  No host round trip and no collective per exchange.
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void tile_signal_kernel(int* a, int* b, int value) {
    __threadfence_system();
    if (a) *reinterpret_cast<volatile int*>(a) = value;
    if (b) *reinterpret_cast<volatile int*>(b) = value;
    __threadfence_system();
}

// spin (with back-off) until both flags reached `need`; gives up after ~8 s and raises sync[4]
__global__ void tile_wait_kernel(volatile int* f0, volatile int* f1, int need, int* timeout_flag) {
    const long long t0 = clock64();
    while ((f0 && *f0 < need) || (f1 && *f1 < need)) {
        __nanosleep(200);
        if (clock64() - t0 > 16000000000LL) { atomicExch(timeout_flag, 1); break; }
    }
    __threadfence_system();
}

// One launch per halo refresh: announce my rows, wait for the neighbours', pull their rows over
// NVLink (128-bit peer loads), acknowledge.  Regions: up to four (x and y1 from above and below).
struct PullJob {
    const float4* src[4];
    float4* dst[4];
    long long n4[4];              // float4 count (regions are multiples of 4 floats)
    int* ready_up; int* ready_dn;     // neighbours' flags I write
    int* ack_up; int* ack_dn;
    volatile int* my_ready0; volatile int* my_ready1;   // my flags the neighbours write
    int* counter;                 // last-CTA detection
    int* timeout_flag;
    int epoch;
};

// Peer loads over NVLink are latency-bound (a round trip is a few microseconds), so the pull wants as
// many loads in flight as the GPU can hold: full-size CTAs, two per SM, about one float4 per thread.
constexpr int kPullThreads = 1024;
__global__ void __launch_bounds__(kPullThreads, 2) tile_exchange_kernel(const PullJob j) {
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            __threadfence_system();
            if (j.ready_up) *reinterpret_cast<volatile int*>(j.ready_up) = j.epoch;
            if (j.ready_dn) *reinterpret_cast<volatile int*>(j.ready_dn) = j.epoch;
        }
        const long long t0 = clock64();
        while ((j.my_ready0 && *j.my_ready0 < j.epoch) || (j.my_ready1 && *j.my_ready1 < j.epoch)) {
            __nanosleep(100);
            if (clock64() - t0 > 16000000000LL) { atomicExch(j.timeout_flag, 1); break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long nmax = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) nmax = j.n4[r] > nmax ? j.n4[r] : nmax;
    // all four regions per pass, loads first: one NVLink round trip per pass instead of four
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nmax; i += stride) {
        float4 v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (i < j.n4[r]) v[r] = j.src[r][i];
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (i < j.n4[r]) j.dst[r][i] = v[r];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(j.counter, 1) == (int)gridDim.x - 1) {      // last CTA: everything is pulled
            *j.counter = 0;
            __threadfence_system();
            if (j.ack_up) *reinterpret_cast<volatile int*>(j.ack_up) = j.epoch;
            if (j.ack_dn) *reinterpret_cast<volatile int*>(j.ack_dn) = j.epoch;
        }
    }
}

constexpr int kIpcBlob = 5 * 64;

}  // namespace

int scipnp_solver_ipc_blob_bytes(void) { return kIpcBlob; }

int scipnp_solver_tiling(scipnp_solver* s, int lo, int hi, int row_lo, int row_hi) {
    SCIPNP_REQUIRE(s, "null solver");
    SCIPNP_REQUIRE(s->p.B == 1, "tiling covers a single scene (B = 1)");
    SCIPNP_REQUIRE(row_lo <= lo && lo < hi && hi <= row_hi && row_hi - row_lo == s->p.H, "inconsistent row ranges");
    s->t_lo = lo; s->t_hi = hi; s->t_row_lo = row_lo; s->t_row_hi = row_hi;
    if (!s->sync) {
        if (int e = dmalloc(s, (void**)&s->sync, kSyncInts * sizeof(int))) return e;
        SCIPNP_CUDA(cudaMemset(s->sync, 0, kSyncInts * sizeof(int)));
    }
    s->tiled = true;
    s->epoch = 0;
    s->ack_pending = false;
    s->tv_tiling.e_lo = lo - row_lo;
    s->tv_tiling.e_hi = hi - row_lo;
    return SCIPNP_OK;
}

int scipnp_solver_set_energy_reduce(scipnp_solver* s, scipnp_energy_reduce_fn reduce, void* user, long long total_rows) {
    SCIPNP_REQUIRE(s, "null solver");
    if (!s->tiled) { set_error("call scipnp_solver_tiling first"); return SCIPNP_ESTATE; }
    SCIPNP_REQUIRE(total_rows >= s->p.H, "total_rows is smaller than this tile");
    s->tv_tiling.reduce = reduce;
    s->tv_tiling.user = user;
    s->tv_tiling.total_rows = total_rows;
    return SCIPNP_OK;
}

int scipnp_solver_ipc_export(scipnp_solver* s, unsigned char* blob) {
    SCIPNP_REQUIRE(s && blob, "null pointer");
    if (!s->tiled) { set_error("call scipnp_solver_tiling first"); return SCIPNP_ESTATE; }
    memset(blob, 0, kIpcBlob);
    void* ptrs[5] = {s->xbuf[0], s->xbuf[1], s->y1buf[0], s->y1buf[1], s->sync};
    for (int i = 0; i < 5; ++i) {
        if (!ptrs[i]) continue;
        cudaIpcMemHandle_t h;
        SCIPNP_CUDA(cudaIpcGetMemHandle(&h, ptrs[i]));
        static_assert(sizeof(h) == 64, "unexpected IPC handle size");
        memcpy(blob + 64 * i, &h, 64);
    }
    return SCIPNP_OK;
}

int scipnp_solver_ipc_attach(scipnp_solver* s, int side, const unsigned char* blob, int peer_row_lo) {
    SCIPNP_REQUIRE(s && blob, "null pointer");
    SCIPNP_REQUIRE(side == 0 || side == 1, "side must be 0 (up) or 1 (down)");
    if (!s->tiled) { set_error("call scipnp_solver_tiling first"); return SCIPNP_ESTATE; }
    scipnp_solver::PeerLink& l = side == 0 ? s->up : s->dn;
    static const unsigned char zero[64] = {0};
    for (int i = 0; i < 5; ++i) {
        if (!memcmp(blob + 64 * i, zero, 64)) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, blob + 64 * i, 64);
        SCIPNP_CUDA(cudaIpcOpenMemHandle(&l.mapped[i], h, cudaIpcMemLazyEnablePeerAccess));
    }
    l.x[0] = (float*)l.mapped[0]; l.x[1] = (float*)l.mapped[1];
    l.y1[0] = (float*)l.mapped[2]; l.y1[1] = (float*)l.mapped[3];
    l.sync = (int*)l.mapped[4];
    l.row_lo = peer_row_lo;
    SCIPNP_REQUIRE(l.x[0] && l.x[1] && l.sync, "incomplete IPC blob");
    l.present = true;
    return SCIPNP_OK;
}

// wait until both neighbours have finished reading my buffers for the last exchange
static int tile_wait_ack(scipnp_solver* s, cudaStream_t st) {
    if (!s->ack_pending) return SCIPNP_OK;
    tile_wait_kernel<<<1, 1, 0, st>>>(s->up.present ? s->sync + 2 : nullptr, s->dn.present ? s->sync + 3 : nullptr,
                                     s->epoch, s->sync + 4);
    count_launch();
    s->ack_pending = false;
    return check_launch("tile_wait_kernel");
}

int scipnp_solver_exchange(scipnp_solver* s, void* stream) {
    SCIPNP_REQUIRE(s, "null solver");
    if (!s->tiled) { set_error("not a tiled solver"); return SCIPNP_ESTATE; }
    if (!s->up.present && !s->dn.present) return SCIPNP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = tile_wait_ack(s, st)) return e;
    const int e = ++s->epoch;
    const int W = s->p.W, C = s->p.C;
    const int xi = s->xa == s->xbuf[0] ? 0 : 1;
    const bool admm = s->p.method == 1;
    float* aux = admm ? s->ba : s->y1a;                       // the second carried array: y1 (GAP) or b (ADMM)
    const bool has_aux = admm || s->p.accelerate;
    const int yi = aux == s->y1buf[0] ? 0 : 1;
    // one launch: announce my rows (I am the upper neighbour's "down" side), wait for theirs,
    // pull my halo rows out of their owned rows, acknowledge
    const size_t rowx = (size_t)W * C, rowy = admm ? rowx : (size_t)W;
    PullJob j{};
    int nr = 0;
    long long total4 = 0;
    auto add = [&](const float* src, float* dst, size_t nfloat) {
        j.src[nr] = reinterpret_cast<const float4*>(src);
        j.dst[nr] = reinterpret_cast<float4*>(dst);
        j.n4[nr] = (long long)(nfloat / 4);
        total4 += j.n4[nr];
        ++nr;
    };
    if ((rowy % 4) != 0 || (rowx % 4) != 0) { set_error("tiled exchange needs W %% 4 == 0"); return SCIPNP_EINVAL; }
    if (s->up.present) {
        const int n = s->t_lo - s->t_row_lo, src = s->t_row_lo - s->up.row_lo;
        if (n > 0) {
            add(s->up.x[xi] + src * rowx, s->xa, n * rowx);
            if (has_aux && s->up.y1[yi]) add(s->up.y1[yi] + src * rowy, aux, n * rowy);
        }
    }
    if (s->dn.present) {
        const int n = s->t_row_hi - s->t_hi, dst = s->t_hi - s->t_row_lo, src = s->t_hi - s->dn.row_lo;
        if (n > 0) {
            add(s->dn.x[xi] + src * rowx, s->xa + dst * rowx, n * rowx);
            if (has_aux && s->dn.y1[yi]) add(s->dn.y1[yi] + src * rowy, aux + dst * rowy, n * rowy);
        }
    }
    j.ready_up = s->up.present ? s->up.sync + 1 : nullptr;
    j.ready_dn = s->dn.present ? s->dn.sync + 0 : nullptr;
    j.ack_up = s->up.present ? s->up.sync + 3 : nullptr;
    j.ack_dn = s->dn.present ? s->dn.sync + 2 : nullptr;
    j.my_ready0 = s->up.present ? s->sync + 0 : nullptr;
    j.my_ready1 = s->dn.present ? s->sync + 1 : nullptr;
    j.counter = s->sync + 5;
    j.timeout_flag = s->sync + 4;
    j.epoch = e;
    long long ctas = (total4 + kPullThreads - 1) / kPullThreads;
    if (ctas < 1) ctas = 1;
    if (ctas > 2 * num_sms()) ctas = 2 * num_sms();
    tile_exchange_kernel<<<(unsigned)ctas, kPullThreads, 0, st>>>(j);
    count_launch();
    s->ack_pending = true;
    return check_launch("tile exchange");
}

// `iters` iterations with a halo exchange every `k` (and after the last one); asynchronous
int scipnp_solver_run_tiled(scipnp_solver* s, int iters, int k, void* stream) {
    SCIPNP_REQUIRE(s, "null solver");
    SCIPNP_REQUIRE(iters >= 0 && k >= 1, "bad iteration counts");
    if (!s->tiled) { set_error("not a tiled solver"); return SCIPNP_ESTATE; }
    cudaStream_t st = (cudaStream_t)stream;
    if (s->push_enabled && s->use_fused && k == 1) {
        // Halo push: every fused launch stores the rows next to a seam into the neighbours' halo rows and raises
        // their flags; the next launch's loader waits for mine.  Nothing but the iterations is enqueued.
        if (int e = scipnp_solver_step_async(s, iters, stream)) return e;
        if (s->up.present || s->dn.present) {
            // my halo rows are complete (and nobody writes my buffers any more) once both flags reached the step count
            tile_wait_kernel<<<1, 1, 0, st>>>(s->up.present ? s->sync + 8 : nullptr, s->dn.present ? s->sync + 9 : nullptr,
                                             s->push_epoch, s->sync + 4);
            count_launch();
            if (int e = check_launch("tile_wait_kernel")) return e;
        }
        return SCIPNP_OK;
    }
    for (int it = 0; it < iters; ++it) {
        // The neighbours pull from the buffer the state was in at the last exchange.  The fused step ping-pongs, so
        // it is the step after next that overwrites that buffer: their acknowledgement must be in by the second step
        // after an exchange (k = 1: the next exchange waits for it).  The exact path projects in place, so there
        // the very next step has to wait.
        if (s->ack_pending && (!s->use_fused || (k > 1 && (it % k) == 1)))
            if (int e = tile_wait_ack(s, st)) return e;
        if (int e = scipnp_solver_step_async(s, 1, stream)) return e;
        if ((it + 1) % k == 0 || it + 1 == iters)
            if (int e = scipnp_solver_exchange(s, stream)) return e;
    }
    // When the stream has drained the neighbours are done reading my buffers: a following load(), rollback or
    // owned() copy may touch them without another rendezvous.
    if (int e = tile_wait_ack(s, st)) return e;
    return SCIPNP_OK;
}

// Halo push for a tiled handle whose halo is exactly tv_iter_max-1 rows per neighbour (one exchange per iteration):
// `up_rows` / `dn_rows` are the neighbours' local row counts.  Takes effect in scipnp_solver_run_tiled(k = 1) while
// the handle is on the fused path; the exact path keeps the pull exchange.
int scipnp_solver_enable_push(scipnp_solver* s, int up_rows, int dn_rows) {
    SCIPNP_REQUIRE(s, "null solver");
    if (!s->tiled) { set_error("call scipnp_solver_tiling first"); return SCIPNP_ESTATE; }
    const scipnp_params& p = s->p;
    const int R = p.tv_iter_max - 1;
    FusedArgs a{};
    a.mode = p.method == 1 ? MODE_ADMM : (p.accelerate ? MODE_GAP_ACC : MODE_GAP_PLAIN);
    a.B = p.B; a.H = p.H; a.W = p.W; a.C = p.C; a.tv_iter_max = p.tv_iter_max; a.clip01 = p.clip01;
    a.x_in = s->xa; a.x_out = s->xb; a.Phi = s->Phi; a.y = s->y; a.Phi_sum = s->PhiSum; a.y1_in = s->y1a; a.y1_out = s->y1b;
    a.b_in = s->ba; a.b_out = s->bb; a.xproj_out = s->xproj;
    if (!s->fused_possible || !s->xb || !fused_ws_supported(a)) {
        set_error("halo push needs the warp-specialised fused kernel (GAP or ADMM without clip, C %% 4 == 0, C <= 24, W %% 4 == 0)");
        return SCIPNP_ESTATE;
    }
    if ((s->up.present && (s->t_lo - s->t_row_lo != R || up_rows < 1)) ||
        (s->dn.present && (s->t_row_hi - s->t_hi != R || dn_rows < 1)) || s->t_hi - s->t_lo < R) {
        set_error("halo push needs exactly tv_iter_max-1 halo rows per neighbour and at least as many owned rows");
        return SCIPNP_EINVAL;
    }
    s->up.rows = up_rows; s->dn.rows = dn_rows;
    if (!s->elog) {
        if (int e = dmalloc(s, (void**)&s->elog, (size_t)kElogCap * p.B * p.C * R * sizeof(double))) return e;
    }
    s->push_enabled = true;
    return SCIPNP_OK;
}

int scipnp_solver_uses_push(scipnp_solver* s) { return s && s->push_enabled ? 1 : 0; }

// TV energies [iterations][B*C][tv_iter_max-1] of the fused iterations since scipnp_solver_begin, summed over this
// rank's owned rows (device memory; valid once the stream has drained).  A tiled caller sums them over the ranks and
// applies skimage's stopping rule to the whole scene.
int scipnp_solver_energy_log(scipnp_solver* s, double** dev, int* iterations, int* per_iteration) {
    SCIPNP_REQUIRE(s && dev && iterations && per_iteration, "null pointer");
    *dev = s->elog;
    *iterations = s->elog ? s->elog_count : 0;
    *per_iteration = s->p.B * s->p.C * (s->p.tv_iter_max - 1);
    return SCIPNP_OK;
}

// device address of the time-out flag (an int, nonzero after a neighbour never signalled), for callers that fold it into
// a transfer of their own instead of paying the stream drain of scipnp_solver_sync_error
int scipnp_solver_sync_flag(scipnp_solver* s, int** dev) {
    SCIPNP_REQUIRE(s && dev, "null pointer");
    *dev = s->sync ? s->sync + 4 : nullptr;
    return SCIPNP_OK;
}

int scipnp_solver_sync_error(scipnp_solver* s, int* timed_out, void* stream) {
    SCIPNP_REQUIRE(s && timed_out, "null pointer");
    *timed_out = 0;
    if (!s->sync) return SCIPNP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SCIPNP_CUDA(cudaMemcpyAsync(timed_out, s->sync + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCIPNP_CUDA(cudaStreamSynchronize(st));
    return SCIPNP_OK;
}

// The host-buffer entries keep their last solver handle: a second call with the same parameters
// reuses its ~6 GB of device buffers instead of paying cudaMalloc/cudaFree again.
static scipnp_solver* g_host_solver = nullptr;
static std::mutex g_host_mu;      // the one-call entries share the cached handle: one caller at a time

static bool same_params(const scipnp_params& a, const scipnp_params& b) {      // field by field: the struct has padding
    return a.method == b.method && a.accelerate == b.accelerate && a.lambda == b.lambda && a.gamma == b.gamma &&
           a.tv_weight == b.tv_weight && a.tv_eps == b.tv_eps && a.tv_iter_max == b.tv_iter_max && a.fused == b.fused &&
           a.B == b.B && a.H == b.H && a.W == b.W && a.C == b.C && a.phi_batched == b.phi_batched && a.clip01 == b.clip01;
}

static int denoise_host(int method, const float* y, const float* Phi, const float* x0,
                        const float* X_orig, const scipnp_params* p, int iters, float* x_out,
                        double* psnr_all, int* psnr_count) {
    SCIPNP_REQUIRE(p && y && Phi && x_out, "null pointer");
    scipnp_params q = *p;
    q.method = method;
    std::lock_guard<std::mutex> lock(g_host_mu);
    scipnp_solver* s = g_host_solver;
    if (s && !same_params(s->p, q)) {
        scipnp_solver_destroy(s);
        s = g_host_solver = nullptr;
    }
    if (!s) {
        if (int e = scipnp_solver_create(&q, &s)) return e;
        g_host_solver = s;
    }
    int e = scipnp_solver_load(s, y, Phi, nullptr, x0, X_orig, nullptr);
    if (!e) e = scipnp_solver_run(s, iters, nullptr);
    if (!e) e = scipnp_solver_get_x(s, x_out, nullptr);
    int cnt = 0;
    if (!e) e = scipnp_solver_psnr(s, psnr_all, psnr_all ? iters * q.B : 0, &cnt, nullptr);
    if (psnr_count) *psnr_count = cnt;
    if (e) {                       // do not keep a handle in an unknown state
        scipnp_solver_destroy(s);
        g_host_solver = nullptr;
    }
    return e;
}

int scipnp_host_release(void) {
    std::lock_guard<std::mutex> lock(g_host_mu);
    if (g_host_solver) scipnp_solver_destroy(g_host_solver);
    g_host_solver = nullptr;
    return SCIPNP_OK;
}

int scipnp_gap_denoise_host(const float* y, const float* Phi, const float* x0, const float* X_orig,
                            const scipnp_params* p, int iters, float* x_out, double* psnr_all,
                            int* psnr_count) {
    return denoise_host(0, y, Phi, x0, X_orig, p, iters, x_out, psnr_all, psnr_count);
}

int scipnp_admm_denoise_host(const float* y, const float* Phi, const float* x0, const float* X_orig,
                             const scipnp_params* p, int iters, float* x_out, double* psnr_all,
                             int* psnr_count) {
    return denoise_host(1, y, Phi, x0, X_orig, p, iters, x_out, psnr_all, psnr_count);
}


// ---------------------------------------------------------------------------------------------
// Host-buffer pipeline: a stream of reconstructions with the same parameters (the frame loop of
// admmdenoise_cacti, pnp_sci_algo.py:498-529, or a camera feed).  `depth` solver handles, each with
// its own stream; a submitted reconstruction is  H2D copies -> solve -> D2H copy  on its slot's
// stream, so the copies of one reconstruction run on the copy engines under the kernels of its
// neighbours.  Host buffers must be page-locked for the copies to overlap (pageable memory works,
// serialised by the driver) and must stay valid until the ticket was waited for.
// ---------------------------------------------------------------------------------------------
struct scipnp_pipeline {
    struct Slot {
        scipnp_solver* s = nullptr;
        cudaStream_t st = nullptr;
        cudaEvent_t done = nullptr;
        int* fired_host = nullptr;         // pinned copy of the early-stop flag
        int ticket = -1;                   // ticket in flight (-1: free)
        int iters = 0;
        float* x_out = nullptr;
    };
    std::vector<Slot> slots;
    scipnp_params p{};
    int next_ticket = 0;
};

int scipnp_pipeline_destroy(scipnp_pipeline* pl) {
    if (!pl) return SCIPNP_OK;
    for (auto& sl : pl->slots) {
        if (sl.st) cudaStreamSynchronize(sl.st);
        if (sl.s) scipnp_solver_destroy(sl.s);
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.fired_host) cudaFreeHost(sl.fired_host);
        if (sl.st) cudaStreamDestroy(sl.st);
    }
    delete pl;
    return SCIPNP_OK;
}

int scipnp_pipeline_create(const scipnp_params* p, int depth, scipnp_pipeline** out) {
    SCIPNP_REQUIRE(p && out, "null pointer");
    SCIPNP_REQUIRE(depth >= 1 && depth <= 8, "pipeline depth must be 1..8");
    scipnp_pipeline* pl = new (std::nothrow) scipnp_pipeline();
    if (!pl) { set_error("out of host memory"); return SCIPNP_ENOMEM; }
    pl->p = *p;
    pl->slots.resize(depth);
    for (auto& sl : pl->slots) {
        int e = scipnp_solver_create(p, &sl.s);
        if (!e && cudaStreamCreateWithFlags(&sl.st, cudaStreamNonBlocking) != cudaSuccess) e = SCIPNP_ECUDA;
        if (!e && cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming) != cudaSuccess) e = SCIPNP_ECUDA;
        if (!e && cudaMallocHost((void**)&sl.fired_host, sizeof(int)) != cudaSuccess) e = SCIPNP_ECUDA;
        if (e) {
            if (e == SCIPNP_ECUDA) set_error("pipeline: stream / event / pinned allocation failed");
            scipnp_pipeline_destroy(pl);
            return e;
        }
    }
    *out = pl;
    return SCIPNP_OK;
}

// wait for the reconstruction in a slot; redo it on the exact path if the early stop fired
static int pipeline_finish(scipnp_pipeline::Slot& sl) {
    SCIPNP_CUDA(cudaEventSynchronize(sl.done));
    scipnp_solver* s = sl.s;
    if (s->xsnap && *sl.fired_host) {
        if (int e = scipnp_solver_rollback(s, sl.st)) return e;
        s->use_fused = false;
        int e = scipnp_solver_step_async(s, sl.iters, sl.st);
        s->use_fused = true;
        if (e) return e;
        s->refined += sl.iters;
        const float* src = s->p.method == 0 ? s->xa : s->xproj;
        SCIPNP_CUDA(cudaMemcpyAsync(sl.x_out, src, s->n_frame * sizeof(float), cudaMemcpyDeviceToHost, sl.st));
        SCIPNP_CUDA(cudaStreamSynchronize(sl.st));
    }
    return SCIPNP_OK;
}

int scipnp_pipeline_submit(scipnp_pipeline* pl, const float* y, const float* Phi, const float* x0,
                           const float* X_orig, int iters, float* x_out, int* ticket) {
    SCIPNP_REQUIRE(pl && y && Phi && x_out && ticket, "null pointer");
    SCIPNP_REQUIRE(iters >= 0, "negative iteration count");
    scipnp_pipeline::Slot& sl = pl->slots[pl->next_ticket % pl->slots.size()];
    if (sl.ticket >= 0) {
        set_error("pipeline slot still holds ticket %d: wait for it before submitting %d more", sl.ticket,
                  (int)pl->slots.size());
        return SCIPNP_ESTATE;
    }
    scipnp_solver* s = sl.s;
    if (int e = scipnp_solver_load(s, y, Phi, nullptr, x0, X_orig, sl.st)) return e;
    *sl.fired_host = 0;
    if (s->use_fused) {
        if (int e = scipnp_solver_begin(s, sl.st)) return e;
    }
    if (int e = scipnp_solver_step_async(s, iters, sl.st)) return e;
    const float* src = s->p.method == 0 ? s->xa : s->xproj;
    SCIPNP_CUDA(cudaMemcpyAsync(x_out, src, s->n_frame * sizeof(float), cudaMemcpyDeviceToHost, sl.st));
    if (s->xsnap && s->use_fused)
        SCIPNP_CUDA(cudaMemcpyAsync(sl.fired_host, s->flags, sizeof(int), cudaMemcpyDeviceToHost, sl.st));
    SCIPNP_CUDA(cudaEventRecord(sl.done, sl.st));
    sl.ticket = pl->next_ticket++;
    sl.iters = iters;
    sl.x_out = x_out;
    *ticket = sl.ticket;
    return SCIPNP_OK;
}

int scipnp_pipeline_wait(scipnp_pipeline* pl, int ticket, double* psnr_all, int psnr_cap, int* psnr_count) {
    SCIPNP_REQUIRE(pl, "null pipeline");
    if (psnr_count) *psnr_count = 0;
    for (auto& sl : pl->slots) {
        if (sl.ticket != ticket) continue;
        int e = pipeline_finish(sl);
        int cnt = 0;
        if (!e && psnr_all) e = scipnp_solver_psnr(sl.s, psnr_all, psnr_cap, &cnt, sl.st);
        if (psnr_count) *psnr_count = cnt;
        sl.ticket = -1;
        return e;
    }
    set_error("pipeline: unknown or already collected ticket %d", ticket);
    return SCIPNP_EINVAL;
}

int scipnp_pipeline_refined_iters(scipnp_pipeline* pl, int* count) {
    SCIPNP_REQUIRE(pl && count, "null pointer");
    int n = 0;
    for (auto& sl : pl->slots) n += sl.s ? sl.s->refined : 0;
    *count = n;
    return SCIPNP_OK;
}

}  // extern "C"
