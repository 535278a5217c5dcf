"""Drop-in for the TV branches of the reference's
``PnP_SCI/python/pnp_sci_algo.py``: same function names, argument meaning,
return tuples and error behaviour, with the whole iteration running on the GPU
through libscipnp.so (no CPU fallback).

    gap_denoise          pnp_sci_algo.py:536-706
    admm_denoise         pnp_sci_algo.py:708-864
    admmdenoise_cacti    pnp_sci_algo.py:479-534
    gap_denoise_bayer    pnp_sci_algo.py:20-265
    admm_denoise_bayer   pnp_sci_algo.py:268-475 (dead code in the reference; built from admm_denoise's semantics)
    GAP_TV_rec           pnp_sci_algo.py:866-882
    ADMM_TV_rec          pnp_sci_algo.py:884-907
    denoise_tv_chambolle skimage.restoration (imported at pnp_sci_algo.py:5)

Differences from the reference, all at the boundary:
  * arithmetic is float32 (float64 inputs are narrowed once on entry);
  * the sensing operators must be the mask operators ``A_``/``At_``: the solver
    needs ``Phi`` itself, so it takes it from ``Phi=`` when given, otherwise it
    recovers it as ``At(ones)`` and verifies that ``A``/``At`` act like
    ``A_(., Phi)`` / ``At_(., Phi)``; anything else raises ``ValueError``;
  * only ``denoiser='tv'`` with ``tvm='tv_chambolle'`` is on the hot path; other
    denoisers raise ``ValueError('Unsupported denoiser ...')`` like the
    reference does for unknown names (pnp_sci_algo.py:678);
  * ``admmdenoise_cacti(orig=None)`` works (the reference raises
    UnboundLocalError at :513 because ``orig_k`` is never bound).
"""
import ctypes as C
import time

import numpy as np
import torch

from ._lib import lib, check, ScipnpError
from .engine import Solver, to_device, stream_ptr, is_torch, dptr, f32c
from .iqa import frames_iqa, frame_psnr, frame_ssim
from .utils import A_, At_, psnr

__all__ = ["gap_denoise", "admm_denoise", "admmdenoise_cacti", "gap_denoise_bayer",
           "denoise_tv_chambolle", "gap_denoise_cassi", "A_", "At_", "psnr", "GAP_TV_rec", "ADMM_TV_rec",
           "admm_denoise_bayer"]

VERBOSE = False          # print the reference's progress lines (every 5th iteration)
USE_FUSED = True         # one-pass fused iteration where the library supports it

_BAYER = ((0, 0), (0, 1), (1, 0), (1, 1))       # pnp_sci_algo.py:99


# -- helpers -------------------------------------------------------------------

def _total_iters(sigma, iter_max):
    # pnp_sci_algo.py:628-631: sigma and iter_max may be lists (TV ignores sigma)
    if not isinstance(sigma, list):
        sigma = [sigma]
    if not isinstance(iter_max, list):
        iter_max = [iter_max] * len(sigma)
    return int(sum(int(v) for v in iter_max[:len(sigma)]))


def _check_tv(denoiser, tvm, multichannel):
    if str(denoiser).lower() != 'tv' or tvm != 'tv_chambolle':
        raise ValueError('Unsupported denoiser {}!'.format(denoiser))
    if not multichannel:
        raise NotImplementedError("multichannel=False (3-D coupled TV) is not on the "
                                  "reference's hot path and has no CPU fallback here")


def _host(a):
    return a.detach().cpu().numpy() if is_torch(a) else np.asarray(a)


def _recover_phi(A, At, y, Phi):
    """Phi from the caller, or from the opaque At callable (exact for At_)."""
    if Phi is not None:
        return f32c(_host(Phi))
    if A is None or At is None:
        raise ValueError("need either Phi= or the A/At mask operators")
    yh = f32c(_host(y))
    Phi = f32c(_host(At(np.ones_like(yh))))
    if Phi.ndim != 3 or Phi.shape[:2] != yh.shape:
        raise ValueError("At(ones) is not an [H, W, C] mask stack: A/At are not mask operators")
    rng = np.random.default_rng(0)
    r = rng.random(yh.shape, dtype=np.float32)
    z = rng.random(Phi.shape, dtype=np.float32)
    ok = np.allclose(_host(At(r)), r[:, :, None] * Phi, rtol=1e-5, atol=1e-6) and \
        np.allclose(_host(A(z)), np.sum(z * Phi, axis=2), rtol=1e-4, atol=1e-4)
    if not ok:
        raise ValueError("A/At do not act like A_(., Phi)/At_(., Phi); pass Phi= "
                         "(scipnp accelerates the mask operators only)")
    return Phi


def _progress(tag, psnr_all):
    if VERBOSE:
        for k in range(4, len(psnr_all), 5):
            print('  {0}-TV iteration {1: 3d}, PSNR {2:2.2f} dB.'.format(tag, k + 1, psnr_all[k]))


def _solve(method, y, Phi, Phi_sum, x0, X_orig, show_iqa, iters, phi_batched=False, **kw):
    """Batched core: y [B,H,W]; Phi [H,W,C] or [B,H,W,C]; returns x [B,H,W,C] and
    psnr_all as one list per batch element."""
    B, H, W = y.shape
    Cc = Phi.shape[-1]
    with Solver(B, H, W, Cc, method=method, phi_batched=phi_batched, fused=USE_FUSED, **kw) as s:
        s.load(y, Phi, Phi_sum=Phi_sum, x0=x0, X_orig=X_orig if show_iqa else None)
        s.run(iters)
        x = s.get_x()
        pa = s.psnr_all()
    return x, [[float(v) for v in pa[:, b]] for b in range(B)]


# -- R4 ------------------------------------------------------------------------

def gap_denoise(y, Phi_sum, A=None, At=None, _lambda=1, accelerate=True,
                denoiser='tv', iter_max=50, noise_estimate=False, sigma=None,
                tv_weight=0.1, tv_iter_max=5, multichannel=True, x0=None,
                X_orig=None, model=None, show_iqa=True, tvm='tv_chambolle', Phi=None):
    """GAP-TV (pnp_sci_algo.py:536-706).  Returns ``(x, psnr_, ssim_, psnr_all)``."""
    _check_tv(denoiser, tvm, multichannel)
    Phi = _recover_phi(A, At, y, Phi)
    yh = f32c(_host(y))
    Xo = None if X_orig is None else f32c(_host(X_orig))
    x, pa = _solve("gap", yh[None], Phi, f32c(_host(Phi_sum)),
                   None if x0 is None else f32c(_host(x0))[None],
                   None if Xo is None else Xo[None], show_iqa,
                   _total_iters(sigma, iter_max), accelerate=accelerate, _lambda=_lambda,
                   tv_weight=tv_weight, tv_iter_max=tv_iter_max)
    x, pa = x[0], pa[0]
    _progress('GAP', pa)
    ps, ss = frames_iqa(Xo, x)
    return x, ps, ss, pa


# -- R5 ------------------------------------------------------------------------

def admm_denoise(y, Phi_sum, A=None, At=None, _lambda=1, gamma=0.01,
                 denoiser='tv', iter_max=50, noise_estimate=False, sigma=None,
                 tv_weight=0.1, tv_iter_max=5, multichannel=True, x0=None, model=None,
                 X_orig=None, show_iqa=True, Phi=None):
    """ADMM-TV (pnp_sci_algo.py:708-864).  Returns ``x`` (the projection output)
    and the PSNR of ``x``, like the reference."""
    _check_tv(denoiser, 'tv_chambolle', multichannel)
    Phi = _recover_phi(A, At, y, Phi)
    yh = f32c(_host(y))
    Xo = None if X_orig is None else f32c(_host(X_orig))
    x, pa = _solve("admm", yh[None], Phi, f32c(_host(Phi_sum)),
                   None if x0 is None else f32c(_host(x0))[None],
                   None if Xo is None else Xo[None], show_iqa,
                   _total_iters(sigma, iter_max), _lambda=_lambda, gamma=gamma,
                   tv_weight=tv_weight, tv_iter_max=tv_iter_max)
    x, pa = x[0], pa[0]
    _progress('ADMM', pa)
    ps, ss = frames_iqa(Xo, x)
    return x, ps, ss, pa


# -- R7 ------------------------------------------------------------------------

def admmdenoise_cacti(meas, mask, A=None, At=None, projmeth='admm', v0=None, orig=None,
                      iframe=0, nframe=1, MAXB=1., maskdirection='plain', **args):
    """Coded-frame loop (pnp_sci_algo.py:479-534).  The ``nframe`` measurements
    share the mask and are independent, so they are reconstructed as one batch.
    Returns ``(x_, t_, psnr_, ssim_, psnrall_)``."""
    pm = str(projmeth).lower()
    if pm not in ('admm', 'gap'):
        raise ValueError('Unsupported projection method %s' % str(projmeth).upper())
    kw = dict(args)
    denoiser = kw.pop('denoiser', 'tv')
    _check_tv(denoiser, kw.pop('tvm', 'tv_chambolle'), kw.pop('multichannel', True))
    iters = _total_iters(kw.pop('sigma', None), kw.pop('iter_max', 50))
    show_iqa = kw.pop('show_iqa', True)
    for k in ('noise_estimate', 'model'):
        kw.pop(k, None)
    if pm == 'admm':
        kw.pop('accelerate', None)
    else:
        kw.pop('gamma', None)

    mask = f32c(_host(mask))
    meas = _host(meas)
    H, W, nmask = mask.shape
    md = str(maskdirection).lower()
    t0 = time.time()
    flips = [(md == 'updown' and (kf + iframe) % 2 == 1) or
             (md == 'downup' and (kf + iframe) % 2 == 0) for kf in range(nframe)]
    y = np.stack([f32c(meas[..., kf + iframe]) / np.float32(MAXB) for kf in range(nframe)])
    Xo = None
    if orig is not None:
        orig = _host(orig)
        Xo = np.stack([f32c(orig[..., (kf + iframe) * nmask:(kf + iframe + 1) * nmask])
                       / np.float32(MAXB) for kf in range(nframe)])
    x0 = None
    if v0 is not None:
        v0 = f32c(_host(v0))
        x0 = np.stack([v0[:, :, kf * nmask:(kf + 1) * nmask][..., ::-1] if flips[kf]
                       else v0[:, :, kf * nmask:(kf + 1) * nmask] for kf in range(nframe)])
    xs, pas = _solve(pm, y, mask, None, x0, Xo, show_iqa, iters, **kw)
    x_ = np.zeros((H, W, nmask * nframe), dtype=np.float32)
    psnr_, ssim_, psnrall_ = [], [], []
    for kf in range(nframe):
        _progress(pm.upper(), pas[kf])
        ps, ss = frames_iqa(None if Xo is None else Xo[kf], xs[kf])
        xk, pa = xs[kf], pas[kf]
        if flips[kf]:
            xk, ps, ss, pa = xk[..., ::-1], ps[::-1], ss[::-1], pa[::-1]
        x_[..., kf * nmask:(kf + 1) * nmask] = xk
        psnr_.extend(ps)
        ssim_.extend(ss)
        psnrall_.append(pa)
    t_ = time.time() - t0
    return x_, t_, psnr_, ssim_, psnrall_


# -- R8 ------------------------------------------------------------------------

def _bayer_split(a, Cc):
    """[H,W,(C)] host array -> device tensor [4,H/2,W/2,(C)] via scipnp_bayer_split."""
    d = to_device(a)
    H, W = d.shape[:2]
    out = torch.empty((4, H // 2, W // 2) + ((Cc,) if d.dim() == 3 else ()),
                      dtype=torch.float32, device=d.device)
    check(lib.scipnp_bayer_split(dptr(d), dptr(out), H, W, Cc if d.dim() == 3 else 1, stream_ptr()))
    return out


def _bayer_solve(method, y_bayer, Phi_bayer, x0_bayer, X_orig, show_iqa, iters, **kw):
    """The four RGGB sub-lattices as one batch of four measurements with their own masks; returns the merged
    mosaic (host), the device-side per-iteration squared errors and the float32 X_orig."""
    Phi = f32c(_host(Phi_bayer))
    H, W, Cc = Phi.shape
    if H % 2 or W % 2:
        raise ValueError("Bayer mosaics need even H and W")
    yq = _bayer_split(f32c(_host(y_bayer)), 1)
    Pq = _bayer_split(Phi, Cc)
    x0q = None if x0_bayer is None else _bayer_split(f32c(_host(x0_bayer)), Cc)
    Xo = None if X_orig is None else f32c(_host(X_orig))
    Xq = None if (Xo is None or not show_iqa) else _bayer_split(Xo, Cc)
    with Solver(4, H // 2, W // 2, Cc, method=method, phi_batched=True, fused=USE_FUSED, **kw) as s:
        s.load(yq, Pq, x0=x0q, X_orig=Xq)
        s.run(iters)
        xq = torch.empty((4, H // 2, W // 2, Cc), dtype=torch.float32, device=yq.device)
        s.get_x(xq)
        se = s.sqerr_all()
    # compare_psnr(X_orig, x_bayer, data_range=1.) over the whole mosaic (:240)
    pa = [float(10 * np.log10(float(H * W * Cc) / v)) for v in se.sum(axis=1)]
    xd = torch.empty((H, W, Cc), dtype=torch.float32, device=yq.device)
    check(lib.scipnp_bayer_merge(dptr(xq), dptr(xd), H, W, Cc, stream_ptr()))
    return xd.cpu().numpy(), pa, Xo


def gap_denoise_bayer(y_bayer, Phi_bayer, _lambda=1, accelerate=True, denoiser='tv',
                      iter_max=50, noise_estimate=True, sigma=None, tv_weight=0.1,
                      tv_iter_max=5, multichannel=True, x0_bayer=None, X_orig=None,
                      model=None, show_iqa=True):
    """Bayer GAP-TV (pnp_sci_algo.py:20-265): the four RGGB sub-lattices are four
    independent measurements with their own masks; the joint TV call over the
    [H/2, W/2, 4*Cr] stack (:163-166) is per-channel, i.e. the same batch."""
    _check_tv(denoiser, 'tv_chambolle', multichannel)
    x, pa, Xo = _bayer_solve("gap", y_bayer, Phi_bayer, x0_bayer, X_orig, show_iqa, _total_iters(sigma, iter_max),
                             accelerate=accelerate, _lambda=_lambda, tv_weight=tv_weight, tv_iter_max=tv_iter_max)
    _progress('GAP', pa)
    ps, ss = frames_iqa(Xo, x)
    return x, ps, ss, pa


def admm_denoise_bayer(y_bayer, Phi_bayer, _lambda=1, gamma=0.01, denoiser='tv', iter_max=50,
                       noise_estimate=True, sigma=None, tv_weight=0.1, tv_iter_max=5, multichannel=True,
                       x0_bayer=None, X_orig=None, model=None, show_iqa=True):
    """Bayer ADMM-TV (pnp_sci_algo.py:268-475).  The reference's body cannot run (``ball`` is never bound, :399);
    this is the loop it evidently means -- ``admm_denoise`` (:805-836) on the four sub-lattices with their own
    masks, multiplier ``ball`` starting at zero, one TV call over the [H/2, W/2, 4*Cr] stack.  Returns
    ``(x_bayer, psnr_all)`` like :475."""
    _check_tv(denoiser, 'tv_chambolle', multichannel)
    x, pa, _ = _bayer_solve("admm", y_bayer, Phi_bayer, x0_bayer, X_orig, show_iqa, _total_iters(sigma, iter_max),
                            _lambda=_lambda, gamma=gamma, tv_weight=tv_weight, tv_iter_max=tv_iter_max)
    _progress('ADMM', pa)
    return x, pa


# -- R9 ------------------------------------------------------------------------

def gap_denoise_cassi(y, mask2d, nband, step, _lambda=1, accelerate=True, denoiser='tv',
                      iter_max=50, noise_estimate=False, sigma=None, tv_weight=0.1, tv_iter_max=5,
                      multichannel=True, x0=None, X_orig=None, model=None, show_iqa=True,
                      tvm='tv_chambolle'):
    """GAP-TV for single-disperser CASSI (DeSCI/test_desci_cassi.m:53-75).  The coded aperture
    ``mask2d`` [H, W] is dispersed by ``step`` pixels per band (``Phi[h, w+step*k, k] = M[h, w]``)
    and the reconstruction runs on the sheared canvas [H, W+(nband-1)*step, nband], as the
    reference's data are.  The iterations read the aperture at per-band index offsets; the shifted
    mask stack is only built once on the device for the initial guess and ``Phi_sum``.
    Returns ``(x, psnr_, ssim_, psnr_all)`` like ``gap_denoise``."""
    _check_tv(denoiser, tvm, multichannel)
    m = f32c(_host(mask2d))
    H, W = m.shape
    Wc = W + (nband - 1) * step
    yh = f32c(_host(y))
    if yh.shape != (H, Wc):
        raise ValueError("y has shape %s, expected the sheared canvas %s" % (yh.shape, (H, Wc)))
    Xo = None if X_orig is None else f32c(_host(X_orig))
    with Solver(1, H, Wc, nband, method="gap", accelerate=accelerate, _lambda=_lambda,
                tv_weight=tv_weight, tv_iter_max=tv_iter_max, fused=USE_FUSED) as s:
        s.load_cassi(yh[None], m, step, x0=None if x0 is None else f32c(_host(x0))[None],
                     X_orig=None if (Xo is None or not show_iqa) else Xo[None])
        s.run(_total_iters(sigma, iter_max))
        x = s.get_x()[0]
        pa = [float(v) for v in s.psnr_all()[:, 0]]
    _progress('GAP', pa)
    ps, ss = frames_iqa(Xo, x)
    return x, ps, ss, pa


# -- R6 ------------------------------------------------------------------------

def denoise_tv_chambolle(image, weight=0.1, eps=2.e-4, n_iter_max=200, multichannel=False,
                         return_stats=False):
    """Chambolle TV denoising with the surface of scikit-image < 0.19
    (``skimage.restoration.denoise_tv_chambolle``).  ``multichannel=True`` treats
    every ``image[..., c]`` as an independent 2-D problem; a 2-D image is one
    problem.  ``return_stats=True`` also returns ``(n_exec[C], energy[C, T])``."""
    src = image
    d = to_device(image)
    if multichannel:
        if d.dim() != 3:
            raise ValueError("multichannel=True expects [H, W, C]")
        H, W, Cc = d.shape
    else:
        if d.dim() != 2:
            raise NotImplementedError("n-D coupled TV (multichannel=False on a 3-D array) is "
                                      "not on the reference's hot path")
        H, W = d.shape
        Cc = 1
    out = torch.empty_like(d)
    if USE_FUSED and multichannel and not return_stats and lib.scipnp_tv_fused_supported(1, H, W, Cc, int(n_iter_max)) \
            and d.data_ptr() % 16 == 0:
        # one pass over HBM (all dual updates in one launch); the exact kernels below take over when the
        # energy criterion of skimage would have stopped a slice early
        fwb = lib.scipnp_tv_fused_workspace_bytes(1, H, W, Cc, int(n_iter_max))
        fws = torch.empty(fwb, dtype=torch.uint8, device=d.device)
        flag = torch.zeros(1, dtype=torch.int32, device=d.device)
        check(lib.scipnp_tv_chambolle_fused(dptr(d), dptr(out), float(weight), float(eps), int(n_iter_max),
                                            1, H, W, Cc, dptr(fws), fwb, dptr(flag), stream_ptr()))
        if int(flag.item()) == 0:
            return out if is_torch(src) else out.cpu().numpy()
    wsb = lib.scipnp_tv_workspace_bytes(1, H, W, Cc)
    ws = torch.empty(wsb, dtype=torch.uint8, device=d.device)
    n_exec = energy = None
    cap = 0
    if return_stats:
        cap = int(n_iter_max)
        n_exec = torch.zeros(Cc, dtype=torch.int32, device=d.device)
        energy = torch.empty((Cc, max(cap, 1)), dtype=torch.float64, device=d.device)
    check(lib.scipnp_tv_chambolle(dptr(d), dptr(out), float(weight), float(eps), int(n_iter_max),
                                  1, H, W, Cc, dptr(ws), wsb, dptr(n_exec), dptr(energy), cap,
                                  stream_ptr()))
    res = out if is_torch(src) else out.cpu().numpy()
    if return_stats:
        return res, n_exec.cpu().numpy(), energy.cpu().numpy()
    return res


# -- the stand-alone TV loops (pnp_sci_algo.py:866-907) ---------------------------

def _rec_progress(tag, ni, t0, pa):
    # the reference prints every fifth iteration (:877-881, :901-906)
    if VERBOSE and (ni + 1) % 5 == 0 and pa:
        print("%s: Iteration %3d, PSNR = %2.2f dB, time = %3.1fs." % (tag, ni + 1, pa[-1], time.time() - t0))


def GAP_TV_rec(y, Phi, A, At, Phi_sum, maxiter, step_size, weight, row, col, ColT, X_ori):
    """Accelerated GAP with ``denoise_tv_chambolle(n_iter_max=30)`` per step (pnp_sci_algo.py:866-882).  ``A`` and
    ``At`` are the reference's two-argument mask operators and are not called: ``Phi`` is given.  The 30 dual
    iterations run on the exact path (two launches per dual iteration, skimage's early stop on the device).
    Returns ``f`` [row, col, ColT] (float32; the reference's loop is float64)."""
    Phi = f32c(_host(Phi))
    yh = f32c(_host(y))
    if yh.shape != (row, col) or Phi.shape != (row, col, ColT):
        raise ValueError("y / Phi do not match (row, col, ColT)")
    Xo = None if X_ori is None else f32c(_host(X_ori))
    t0 = time.time()
    with Solver(1, row, col, ColT, method="gap", accelerate=True, _lambda=float(step_size),
                tv_weight=float(weight), tv_iter_max=30, fused=USE_FUSED) as s:
        s.load(yh[None], Phi, Phi_sum=f32c(_host(Phi_sum)), X_orig=None if Xo is None else Xo[None])
        s.run(int(maxiter))
        x = s.get_x()[0]
        if VERBOSE and Xo is not None:
            pa = [float(v) for v in s.psnr_all()[:, 0]]
            for ni in range(int(maxiter)):
                _rec_progress("GAP-TV", ni, t0, pa[:ni + 1])
    return x


def ADMM_TV_rec(y, Phi, A, At, Phi_sum, maxiter, step_size, weight, row, col, ColT, eta, X_ori):
    """ADMM with 30 Chambolle iterations per step, TV weight x0.999 and eta x0.998 per iteration
    (pnp_sci_algo.py:884-907).  Returns ``v``, the projection output."""
    Phi = f32c(_host(Phi))
    yh = f32c(_host(y))
    if yh.shape != (row, col) or Phi.shape != (row, col, ColT):
        raise ValueError("y / Phi do not match (row, col, ColT)")
    Xo = None if X_ori is None else f32c(_host(X_ori))
    weight, eta = float(weight), float(eta)
    with Solver(1, row, col, ColT, method="admm", _lambda=float(step_size), gamma=eta,
                tv_weight=weight, tv_iter_max=30, fused=False) as s:
        s.load(yh[None], Phi, Phi_sum=f32c(_host(Phi_sum)), X_orig=None if Xo is None else Xo[None])
        for ni in range(int(maxiter)):
            s.set_tv(weight, eta)
            s.step_async(1)
            weight *= 0.999                      # :898
            eta *= 0.998                         # :899
        x = s.get_x()[0]
    return x
