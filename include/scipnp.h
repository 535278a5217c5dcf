/*
 * scipnp.h -- C ABI of libscipnp.so, the B200 (sm_100a) GAP/ADMM-TV engine for
 * snapshot compressive imaging.
 *
 * This is the drop-in boundary for the iterative hot path of the reference's
 * PnP_SCI/python.  Every entry point states the reference
 * interface it replaces (file:line under /root/reference/PnP_SCI/python).
 *
 * Conventions
 *   - plain C: pointers, ints, floats; no C++ or torch types.
 *   - every function returns 0 on success or a negative SCIPNP_E* code; the
 *     message of the last failure on the calling thread is scipnp_last_error().
 *   - "dev" pointers are device pointers on the current CUDA device; "host"
 *     pointers are ordinary (pageable or pinned) host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls
 *     are asynchronous on that stream unless stated otherwise.
 *   - array layout is the reference's logical layout, C-contiguous float32:
 *       frames / masks  x, Phi : [B][H][W][C]   (channel-last, C = Cr)
 *       measurements    y, y1, Phi_sum : [B][H][W]
 *     `B` batches independent measurements.  `phi_batched` = 0 shares one
 *     Phi / Phi_sum ([H][W][C] / [H][W]) between all B measurements, 1 gives
 *     every measurement its own.
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     returns SCIPNP_ECUDA.
 */
#ifndef SCIPNP_H
#define SCIPNP_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCIPNP_OK        0
#define SCIPNP_EINVAL   -1   /* bad argument                                   */
#define SCIPNP_ECUDA    -2   /* CUDA runtime error (message has the detail)    */
#define SCIPNP_ENOMEM   -3   /* device or host allocation failed               */
#define SCIPNP_ESTATE   -4   /* call not valid in the handle's current state   */

/* library ------------------------------------------------------------------ */
int         scipnp_version(void);            /* 10000*major + 100*minor + patch */
const char *scipnp_last_error(void);
int         scipnp_device_count(void);       /* 0 when no usable CUDA device    */
long long   scipnp_launch_count(void);       /* kernels launched by this process */

/* R1  utils.A_  (utils.py:10-15):  y[b,h,w] = sum_c x[b,h,w,c]*Phi[.,h,w,c]     */
int scipnp_A(const float *x_dev, const float *Phi_dev, float *y_dev,
             int B, int H, int W, int C, int phi_batched, void *stream);

/* R2  utils.At_ (utils.py:17-26):  x[b,h,w,c] = y[b,h,w]*Phi[.,h,w,c]           */
int scipnp_At(const float *y_dev, const float *Phi_dev, float *x_dev,
              int B, int H, int W, int C, int phi_batched, void *stream);

/* R3  mask_sum  (pnp_sci_algo.py:491-492): sum_c Phi, zeros replaced by 1       */
int scipnp_phi_sum(const float *Phi_dev, float *Phi_sum_dev,
                   int B, int H, int W, int C, void *stream);

/* R4  Euclidean projection of gap_denoise (pnp_sci_algo.py:640-645):
 *        yb = A(x);  accelerate: y1 += y-yb; x += lambda*At((y1-yb)/Phi_sum)
 *                    else      :             x += lambda*At((y -yb)/Phi_sum)
 *     x_in/x_out and y1_in/y1_out may alias (pointwise).  y1 pointers may be
 *     NULL when accelerate == 0.                                                */
int scipnp_gap_project(const float *x_in, float *x_out,
                       const float *y1_in, float *y1_out,
                       const float *y, const float *Phi, const float *Phi_sum,
                       float lambda, int accelerate,
                       int B, int H, int W, int C, int phi_batched, void *stream);

/* R5  Euclidean projection of admm_denoise (pnp_sci_algo.py:808-809,812):
 *        u = theta+b; x = u + lambda*At((y-A(u))/(Phi_sum+gamma)); f = x-b
 *     writes x (the solver's return value) and f (the TV input).               */
int scipnp_admm_project(const float *theta, const float *b, float *x, float *f,
                        const float *y, const float *Phi, const float *Phi_sum,
                        float lambda, float gamma,
                        int B, int H, int W, int C, int phi_batched, void *stream);

/* R5  multiplier update (pnp_sci_algo.py:836):  b = b - (x - theta)            */
int scipnp_admm_dual_update(float *b, const float *x, const float *theta,
                            size_t n, void *stream);

/* R6  skimage.restoration.denoise_tv_chambolle(image, weight, eps, n_iter_max,
 *     multichannel=True)  (call sites pnp_sci_algo.py:164,409,650,812): every
 *     (b, c) slice is an independent 2-D ROF problem, tau = 1/4, energy-based
 *     early stop evaluated on the device per slice exactly as the reference.
 *     `in` and `out` must not alias.  `workspace` holds the dual field
 *     (scipnp_tv_workspace_bytes).  Optional outputs (may be NULL):
 *       n_exec_dev  int   [B*C]            iterations executed per slice
 *       energy_dev  double[B*C][energy_cap] E_i per slice (unused entries NaN)  */
size_t scipnp_tv_workspace_bytes(int B, int H, int W, int C);
int scipnp_tv_chambolle(const float *in, float *out, double weight, double eps,
                        int n_iter_max, int B, int H, int W, int C,
                        void *workspace, size_t workspace_bytes,
                        int *n_exec_dev, double *energy_dev, int energy_cap,
                        void *stream);

/* R6, one pass over HBM: the same denoiser with all n_iter_max - 1 effective dual updates fused
 *     into a single launch (n_iter_max 3..5, C % 4 == 0, C <= 24, W % 4 == 0; the intermediates
 *     live in registers and TMA-staged shared-memory tiles).  The kernel cannot wait
 *     for skimage's global energy criterion; it evaluates it on the side and sets *flag_dev
 *     (int, device, caller zeroes it; may be NULL) when some slice would have stopped early --
 *     the caller then repeats the call with scipnp_tv_chambolle.  Within 1e-6 of it otherwise.
 *     scipnp_tv_fused_supported: 1 when the shape is covered.                                  */
size_t scipnp_tv_fused_workspace_bytes(int B, int H, int W, int C, int n_iter_max);
int scipnp_tv_fused_supported(int B, int H, int W, int C, int n_iter_max);
int scipnp_tv_chambolle_fused(const float *in, float *out, double weight, double eps,
                              int n_iter_max, int B, int H, int W, int C,
                              void *workspace, size_t workspace_bytes, int *flag_dev,
                              void *stream);

/* MATLAB twin's default TV denoiser (PnP_SCI/matlab/algorithms/tvdenoisers/TV_denoising.m:1-44,
 *     the 'ATV_ClipA' branch of gapdenoise.m:93-94): anisotropic TV by iterative clipping,
 *     alpha = 5, per 2-D frame of a [B][H][W][C] stack; `iters` iterations, the x of the last one
 *     is returned.  workspace: scipnp_tv_atv_clip_workspace_bytes (the two dual fields).       */
/* The rest of the MATLAB twin's TV family (gapdenoise.m:86-108) on [B][H][W][C] stacks, C = the frames (<= 128):
 *   variant 0  ATV_ClipB   TV_denoising_clip_LB.m:25-36      lambda = tvweight
 *           1  ATV_cham    tvdenoise_cham_ATV2D.m:72-87      lambda = 1/tvweight (as gapdenoise.m:98 passes it)
 *           2  ITV2D_cham  tvdenoise_cham_ITV2D.m:73-90      lambda = 1/tvweight
 *           3  ITV3D_cham  tvdenoise_cham_ITV3D.m:72-90      lambda = 1/tvweight
 *           4  ATV_FGP     fgp_denoise_ATV2D.m:73-121        lambda = tvweight, iters = MAXITER
 *           5  ITV2D_FGP   fgp_denoise_ITV2D.m:73-121
 *           6  ITV3D_FGP   fgp_denoise_ITV3D.m:73-121
 * IEEE single precision in the statement order of the .m files.  workspace: scipnp_tv_matlab_workspace_bytes.  */
size_t scipnp_tv_matlab_workspace_bytes(int B, int H, int W, int C);
int scipnp_tv_matlab(const float *in, float *out, int variant, float lambda, int iters, int B, int H, int W, int C,
                     void *workspace, size_t workspace_bytes, void *stream);
size_t scipnp_tv_atv_clip_workspace_bytes(int B, int H, int W, int C);
int scipnp_tv_atv_clip(const float *in, float *out, float lambda, int iters, int B, int H, int W,
                       int C, void *workspace, size_t workspace_bytes, void *stream);

/* Per-frame quality numbers of the solvers' return tuples: compare_psnr / compare_ssim of
 *     scikit-image < 0.18 as called at pnp_sci_algo.py:699-705 and :857-863 (per channel,
 *     data_range = 1; SSIM: 7x7 uniform window, sample covariance, K1 = 0.01, K2 = 0.03, mean over
 *     the pixels whose window lies inside the image).  ref, img: [H][W][C] device arrays.  Writes
 *     per channel (device, double, C values each) the SUM of the SSIM map over its
 *     (H-6)*(W-6) inner pixels and the sum of squared errors over H*W pixels:
 *     ssim = ssim_sum / ((H-6)*(W-6)),  psnr = 10*log10(H*W / sqerr_sum).                 */
int scipnp_frames_iqa(const float *ref, const float *img, int H, int W, int C,
                      double *ssim_sum_dev, double *sqerr_sum_dev, void *stream);

/* R10 utils.psnr (utils.py:28-36): accumulates sum((a-b)^2) into *sum_dev
 *     (double, device; the caller zeroes it).  psnr = 10*log10(n/sum).         */
int scipnp_sq_err(const float *a, const float *b, size_t n, double *sum_dev,
                  void *stream);

/* One whole outer GAP-TV iteration, projection + TV in a single pass over HBM
 * (replaces pnp_sci_algo.py:640-650 for denoiser='tv', tvm='tv_chambolle').
 * Ping-pong: reads x_in / y1_in, writes x_out / y1_out (must not alias).
 * `flags_dev` (int, >= 1 entry, may be NULL): set to nonzero when the energy
 * criterion of the reference would have stopped some slice before n_iter_max;
 * the caller then repeats the iteration with scipnp_gap_project +
 * scipnp_tv_chambolle (the solver below does this by itself).                  */
size_t scipnp_gap_tv_workspace_bytes(int B, int H, int W, int C, int tv_iter_max);
int scipnp_gap_tv_fused(const float *x_in, float *x_out,
                        const float *y1_in, float *y1_out,
                        const float *y, const float *Phi, const float *Phi_sum,
                        float lambda, int accelerate, double tv_weight, double tv_eps,
                        int tv_iter_max, int B, int H, int W, int C, int phi_batched,
                        void *workspace, size_t workspace_bytes, int *flags_dev,
                        void *stream);

/* Which fused kernel scipnp_gap_tv_fused and the solver use: 0 = automatic (the warp-specialised
 * kernel csrc/gap_tv_ws.cuh where it applies: GAP, C % 4 == 0, C <= 24, W % 4 == 0; the stream kernel
 * csrc/gap_tv_stream.cuh otherwise), 1 = stream kernel only, 2 = same as 0.  Process-wide; meant for
 * tests and profiling (environment: SCIPNP_FUSED_VARIANT).  No reference counterpart.          */
int scipnp_set_fused_variant(int variant);

/* R8  Bayer sub-lattice (pnp_sci_algo.py:99,116-137,255-257): split a
 *     [H][W][C] mosaic stack into 4 half-resolution stacks [4][H/2][W/2][C]
 *     (order (0,0),(0,1),(1,0),(1,1)) and back.  C = 1 handles [H][W] planes.   */
int scipnp_bayer_split(const float *full, float *quad, int H, int W, int C, void *stream);
int scipnp_bayer_merge(const float *quad, float *full, int H, int W, int C, void *stream);

/* R9  CASSI dispersion (DeSCI/test_desci_cassi.m:53-62): expand a 2-D coded
 *     aperture M[H][W] into the shifted stack Phi[H][W+(C-1)*step][C],
 *     Phi[h, w+step*k, k] = M[h, w].                                            */
int scipnp_cassi_shift_mask(const float *mask2d, float *Phi, int H, int W, int C,
                            int step, void *stream);

/* --------------------------------------------------------------------------
 * Persistent solver: R4 gap_denoise / R5 admm_denoise with denoiser='tv'
 * (pnp_sci_algo.py:536-706, 708-864).  All state stays in HBM for the whole
 * run; only psnr_all comes back per iteration.
 * -------------------------------------------------------------------------- */
typedef struct scipnp_solver scipnp_solver;

typedef struct scipnp_params {
    int   method;        /* 0 = GAP, 1 = ADMM                                   */
    int   accelerate;    /* GAP only                                            */
    float lambda;        /* _lambda                                             */
    float gamma;         /* ADMM only                                           */
    double tv_weight;    /* double: tau/weight is rounded once, as NumPy does   */
    double tv_eps;       /* skimage default 2e-4                                */
    int   tv_iter_max;
    int   fused;         /* 1: one-pass fused iteration where available         */
    int   B, H, W, C;
    int   phi_batched;
    int   clip01;        /* 1: clip the TV output to [0,1] every iteration, as the joint
                            variant does (joint_pnp_sci_algo.py:633); 0 otherwise */
} scipnp_params;

int scipnp_solver_create(const scipnp_params *p, scipnp_solver **out);
int scipnp_solver_destroy(scipnp_solver *s);

/* Load inputs (host or device pointers, cudaMemcpyDefault).  Phi_sum == NULL
 * derives it from Phi (R3).  x0 == NULL starts from At(y) (pnp_sci_algo.py:
 * 625-627, 793-794).  X_orig == NULL disables psnr_all.                        */
int scipnp_solver_load(scipnp_solver *s, const float *y, const float *Phi,
                       const float *Phi_sum, const float *x0, const float *X_orig,
                       void *stream);

/* R9  CASSI (DeSCI/test_desci_cassi.m:53-75): load a single-disperser measurement.  The handle's W
 * is the sheared canvas width; `mask2d` is the coded aperture [H][W-(C-1)*step] (host or device).
 * The fused iterations read the aperture at per-band index offsets, Phi[h,w,c] = M[h, w-step*c],
 * instead of streaming a shifted mask stack from HBM.                                       */
/* scipnp_solver_load with the mask stack BORROWED: `Phi_dev` (device memory, 16-byte aligned) is read in place by
 * every iteration instead of being copied into the handle; it must stay valid and unchanged until the results were
 * read.  Saves one pass over the mask stack per reconstruction (the masks of a CACTI camera are constant over the
 * frame loop, pnp_sci_algo.py:498-529). */
int scipnp_solver_load_borrow_phi(scipnp_solver *s, const float *y, const float *Phi_dev, const float *Phi_sum,
                                  const float *x0, const float *X_orig, void *stream);
int scipnp_solver_load_cassi(scipnp_solver *s, const float *y, const float *mask2d, int step,
                             const float *x0, const float *X_orig, void *stream);

/* Run `iters` outer iterations on `stream`.  On the fused path the call returns
 * after the stream has drained (it has to look at the early-stop flag and, if it
 * is raised, redoes the run on the exact path); on the exact path it is
 * asynchronous.  psnr_all entries are appended to an internal device array read
 * by scipnp_solver_psnr.                                                        */
int scipnp_solver_run(scipnp_solver *s, int iters, void *stream);

/* The pieces of scipnp_solver_run, for callers that interleave their own work
 * between iterations (the row-tiled multi-GPU driver refreshes halos):
 *   _begin       snapshot the state and clear the early-stop flag   (asynchronous)
 *   _step_async  enqueue `iters` iterations on the current path     (asynchronous)
 *   _fired       drain the stream; *fired != 0 if the reference would have
 *                stopped some TV slice early since _begin
 *   _rollback    restore the snapshot taken by _begin               (asynchronous)
 *   _set_path    1 = fused kernel, 0 = exact kernels, for the following steps
 *   _add_refined account iterations the caller redid on the exact path          */
int scipnp_solver_begin(scipnp_solver *s, void *stream);
int scipnp_solver_step_async(scipnp_solver *s, int iters, void *stream);
int scipnp_solver_fired(scipnp_solver *s, int *fired, void *stream);
int scipnp_solver_rollback(scipnp_solver *s, void *stream);
int scipnp_solver_set_path(scipnp_solver *s, int fused);
/* TV weight and ADMM regulariser of the iterations that follow (ADMM_TV_rec, pnp_sci_algo.py:898-899, shrinks
 * both every iteration). */
int scipnp_solver_set_tv(scipnp_solver *s, double tv_weight, double gamma);
int scipnp_solver_add_refined(scipnp_solver *s, int iters);

/* Copy the current estimate (GAP: x after TV; ADMM: x before TV, as the
 * reference returns) to a host or device buffer and synchronise the stream.    */
int scipnp_solver_get_x(scipnp_solver *s, float *x_out, void *stream);
/* psnr_all (utils.psnr of every iteration against X_orig), one value per
 * (iteration, batch element): psnr_all_host[k*B + b].  *count receives the
 * number of values available; at most `cap` are written.  _sqerr returns the
 * underlying sums of squared errors instead (callers that pool batch elements,
 * e.g. the four Bayer sub-lattices).                                           */
int scipnp_solver_psnr(scipnp_solver *s, double *psnr_all_host, int cap, int *count,
                       void *stream);
int scipnp_solver_sqerr(scipnp_solver *s, double *sums_host, int cap, int *count,
                        void *stream);
/* Number of outer iterations that had to be redone on the exact path because
 * the TV energy criterion fired (0 in the reference's parameter range).        */
int scipnp_solver_refined_iters(scipnp_solver *s, int *count);
/* Device pointers of the state, for callers that manage halos themselves.      */
int scipnp_solver_state(scipnp_solver *s, float **x_cur, float **y1_cur);
/* ADMM handles: device pointers of theta (the TV output), of the multiplier b and of x, the
 * projection output that admm_denoise returns (pnp_sci_algo.py:809).  For callers that put a
 * second denoiser between the TV step and the multiplier update
 * (joint_pnp_sci_algo.py:118-306, admm_multistep_denoise).                       */
int scipnp_solver_admm_state(scipnp_solver *s, float **theta, float **b, float **x);
/* Kernel launches issued since this handle was created.                        */
long long scipnp_solver_launch_count(scipnp_solver *s);
/* 1 when the handle runs the one-pass fused iteration, 0 on the exact path.    */
int scipnp_solver_uses_fused(scipnp_solver *s);

/* --------------------------------------------------------------------------
 * Row-tiled multi-GPU mode (one process per GPU on one node; no counterpart in
 * the single-process reference).  The handle holds rows [row_lo, row_hi) of a
 * taller scene (so p.H = row_hi - row_lo, B = 1; GAP or ADMM) and owns [lo, hi); the
 * other rows are halo copies of rows owned by the neighbouring ranks.  The
 * neighbours' buffers are mapped with CUDA IPC; an exchange pulls the halo rows
 * of x and y1 (ADMM: theta and the multiplier b) straight over NVLink, ordered by flags the ranks write into each
 * other's memory (no host round trip, no collective):
 *   _tiling        declare the row ranges of this rank
 *   _ipc_export    blob (scipnp_solver_ipc_blob_bytes() bytes) describing my buffers
 *   _ipc_attach    map a neighbour's blob; side 0 = rank above, 1 = rank below;
 *                  peer_row_lo = global row of that rank's local row 0
 *   _exchange      enqueue one halo refresh on `stream`
 *   _run_tiled     `iters` iterations, an exchange every `k` and after the last
 *   _sync_error    drains the stream; *timed_out != 0 if a neighbour never showed up
 * -------------------------------------------------------------------------- */
int scipnp_solver_ipc_blob_bytes(void);
int scipnp_solver_tiling(scipnp_solver *s, int lo, int hi, int row_lo, int row_hi);
int scipnp_solver_ipc_export(scipnp_solver *s, unsigned char *blob);
int scipnp_solver_ipc_attach(scipnp_solver *s, int side, const unsigned char *blob, int peer_row_lo);
int scipnp_solver_exchange(scipnp_solver *s, void *stream);
int scipnp_solver_run_tiled(scipnp_solver *s, int iters, int k, void *stream);
int scipnp_solver_sync_error(scipnp_solver *s, int *timed_out, void *stream);
int scipnp_solver_sync_flag(scipnp_solver *s, int **dev);   /* device address of that flag (NULL before _tiling) */
/* skimage's stopping rule of the TV step (pnp_sci_algo.py:650 -> denoise_tv_chambolle) compares energies summed
 * over the WHOLE image.  On the exact path a tiled handle sums its owned rows only and hands the 2*C partial
 * energies of every dual iteration (device doubles) to `reduce`, which must sum them over all ranks in place,
 * ordered on `stream` (e.g. an NCCL all-reduce); `total_rows` is the height of the whole scene.  With it the
 * exact path of the tiled mode takes the stopping decisions of the single-GPU solve.  The one-pass kernel keeps
 * its per-tile side check (rows of the tile): it only decides whether the run is redone on the exact path.    */
/* Halo push (tiles whose halo is exactly tv_iter_max-1 rows per neighbour, i.e. one exchange per iteration): the
 * fused kernel produces the owned rows only, stores the rows next to a seam a second time into the neighbour's halo
 * rows of its output buffers (TMA stores / plain stores over NVLink into the IPC-mapped buffers) and its last CTA
 * raises the neighbours' flags; the next launch's loader waits for this rank's flags before it touches halo rows.
 * No exchange kernel, no acknowledgement (nobody reads remote memory).  up_rows / dn_rows = local row counts of the
 * neighbours.  _run_tiled(k = 1) uses it while the handle is on the fused path; the exact path keeps the pull.
 * _energy_log: TV energies [iterations since _begin][B*C][tv_iter_max-1] over this rank's owned rows (device
 * doubles, valid once the stream drained): summed over the ranks they give skimage's stopping rule for the whole
 * scene (pnp_sci_algo.py:650 -> denoise_tv_chambolle's eps test).                                             */
int scipnp_solver_enable_push(scipnp_solver *s, int up_rows, int dn_rows);
int scipnp_solver_uses_push(scipnp_solver *s);
int scipnp_solver_energy_log(scipnp_solver *s, double **dev, int *iterations, int *per_iteration);
typedef int (*scipnp_energy_reduce_fn)(double *partials_dev, int n, void *stream, void *user);
int scipnp_solver_set_energy_reduce(scipnp_solver *s, scipnp_energy_reduce_fn reduce, void *user,
                                    long long total_rows);

/* Host-buffer one-call entries (what a ctypes / cffi binding of the reference
 * would call in place of gap_denoise / admm_denoise).  Synchronous.  psnr_all
 * must hold iters*B doubles when X_orig is given (may be NULL otherwise).      */
int scipnp_gap_denoise_host(const float *y, const float *Phi, const float *x0,
                            const float *X_orig, const scipnp_params *p, int iters,
                            float *x_out, double *psnr_all, int *psnr_count);
int scipnp_admm_denoise_host(const float *y, const float *Phi, const float *x0,
                             const float *X_orig, const scipnp_params *p, int iters,
                             float *x_out, double *psnr_all, int *psnr_count);
/* The two entries above keep their last solver handle (device buffers) for reuse by a call with
 * identical parameters; this frees it.                                          */
int scipnp_host_release(void);

/* Host-buffer pipeline: a stream of reconstructions with one parameter set -- the frame loop of
 * admmdenoise_cacti (PnP_SCI/python/pnp_sci_algo.py:498-529: one gap_denoise / admm_denoise per
 * coded frame) or a camera feed.  `depth` solver handles with a stream each; submit() enqueues
 * H2D copies -> solve -> D2H copy and returns at once, wait() blocks until x_out of that ticket is
 * complete (redoing the solve on the exact path if the TV early stop fired).  With page-locked
 * host buffers the copies of one reconstruction run under the kernels of the next.  At most
 * `depth` tickets may be in flight; host buffers stay untouched until their ticket was waited
 * for.  psnr_all (with X_orig): up to psnr_cap values, iters*B of them are written.  refined_iters
 * sums the iterations redone on the exact path (it resets when a slot is reused).              */
typedef struct scipnp_pipeline scipnp_pipeline;
int scipnp_pipeline_create(const scipnp_params *p, int depth, scipnp_pipeline **out);
int scipnp_pipeline_destroy(scipnp_pipeline *pl);
int scipnp_pipeline_submit(scipnp_pipeline *pl, const float *y, const float *Phi, const float *x0,
                           const float *X_orig, int iters, float *x_out, int *ticket);
int scipnp_pipeline_wait(scipnp_pipeline *pl, int ticket, double *psnr_all, int psnr_cap,
                         int *psnr_count);
int scipnp_pipeline_refined_iters(scipnp_pipeline *pl, int *count);

#ifdef __cplusplus
}
#endif
#endif /* SCIPNP_H */
