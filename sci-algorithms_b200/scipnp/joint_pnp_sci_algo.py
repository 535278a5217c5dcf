"""Drop-in for the TV paths of the reference's ``PnP_SCI/python/joint_pnp_sci_algo.py``
(SURVEY.md section 8f-1): the module most ``pnp_sci_test_*`` drivers import.

    joint_admmdenoise_cacti  :20-79  coded-frame loop around the two-period drivers below
    admm_denoise   joint_pnp_sci_algo.py:502-665   ADMM-TV with ``theta = clip(theta, 0, 1)`` (:633)
    gap_denoise    joint_pnp_sci_algo.py:666-...    same loop as pnp_sci_algo.gap_denoise

    gap_multistep_denoise  :309-500   TV + learned-denoiser period: projection and TV on the
                                      device, hand-off of the device tensor to the caller's denoiser
    gap_joint_denoise      :100-116   GAP-TV period, then the period above from its result
    admm_multistep_denoise :118-306   the ADMM twin (second denoiser between TV and the multiplier)
    admm_joint_denoise     :81-98     ADMM-TV period, then the period above from its result

The learned denoisers themselves (``packages/ffdnet``, ``packages/fastdvdnet``) are outside this
engine: the TV+CNN period takes the second denoiser as a callable that receives the CUDA tensor.  ``tvm`` may be 'tv_chambolle', 'ITV3D_FGP' or 'ITV2D_cham' in ``admm_denoise``:
the reference calls ``denoise_tv_chambolle`` for all three (:606-611).
"""
import numpy as np
import torch

from . import pnp_sci_algo as _base
from .pnp_sci_algo import _check_tv, _recover_phi, _host, _total_iters, _progress
from .engine import Solver, f32c
from .iqa import frames_iqa
from .tiled import _wrap
from .utils import A_, At_, psnr  # noqa: F401

__all__ = ["joint_admmdenoise_cacti", "admm_denoise", "gap_denoise", "gap_multistep_denoise", "gap_joint_denoise",
           "admm_multistep_denoise", "admm_joint_denoise", "A_", "At_", "psnr"]


def joint_admmdenoise_cacti(meas, mask, A=None, At=None, projmeth='admm', v0=None, orig=None,
                            iframe=0, nframe=1, MAXB=1., maskdirection='plain', denoiser='tv',
                            iter_max1=50, iter_max2=50, sigma1=None, sigma2=None, **args):
    """Coded-frame loop of the joint module (joint_pnp_sci_algo.py:20-79): every coded frame runs the
    two-period driver (``admm_joint_denoise`` / ``gap_joint_denoise``); ``v0`` and the results are
    reversed along the frame axis for the down-going masks of 'updown' / 'downup'.  The mask is handed
    to the solvers as ``Phi=`` (the wrapper knows it; ``A``/``At`` may stay ``None``).  Returns
    ``(x_, t_, psnr_, ssim_, psnrall_)``."""
    import time
    pm = str(projmeth).lower()
    if pm not in ('admm', 'gap'):
        raise ValueError('Unsupported projection method %s' % str(projmeth).upper())
    mask = f32c(_host(mask))
    meas = _host(meas)
    nrow, ncol, nmask = mask.shape
    x_ = np.zeros((nrow, ncol, nmask * nframe), dtype=np.float32)
    psnr_, ssim_, psnrall_ = [], [], []
    t0 = time.time()
    mask_sum = np.sum(mask, axis=2)
    mask_sum[mask_sum == 0] = 1
    md = str(maskdirection).lower()
    args.setdefault('Phi', mask)
    t_ = 0.
    for kf in range(nframe):
        orig_k = None if orig is None else _host(orig)[:, :, (kf + iframe) * nmask:(kf + iframe + 1) * nmask] / MAXB
        meas_k = meas[:, :, kf + iframe] / MAXB
        down = (md == 'updown' and (kf + iframe) % 2 == 1) or (md == 'downup' and (kf + iframe) % 2 == 0)
        v0_k = None
        if v0 is not None:
            v0_k = _host(v0)[:, :, kf * nmask:(kf + 1) * nmask]
            if down:
                v0_k = v0_k[:, :, ::-1]
        joint = admm_joint_denoise if pm == 'admm' else gap_joint_denoise
        x_k, psnr_k, ssim_k, psnrall_k = joint(meas_k, mask_sum, A, At, x0=v0_k, X_orig=orig_k, denoiser=denoiser,
                                               iter_max1=iter_max1, iter_max2=iter_max2, sigma1=sigma1,
                                               sigma2=sigma2, **args)
        if down:
            x_k, psnr_k, ssim_k, psnrall_k = x_k[:, :, ::-1], psnr_k[::-1], ssim_k[::-1], psnrall_k[::-1]
        t_ = time.time() - t0
        x_[:, :, kf * nmask:(kf + 1) * nmask] = x_k
        psnr_.extend(psnr_k)
        ssim_.extend(ssim_k)
        psnrall_.append(psnrall_k)
    return x_, t_, psnr_, ssim_, psnrall_


def admm_denoise(y, Phi_sum, A=None, At=None, _lambda=1, gamma=0.0, accelerate=None,
                 denoiser='tv', iter_max=50, noise_estimate=False, sigma=None, tv_weight=0.1,
                 tv_iter_max=5, multichannel=True, x0=None, model=None, X_orig=None, show_iqa=True,
                 tvm='tv_chambolle', Phi=None):
    """ADMM-TV of the joint module: ``theta`` is clipped to [0, 1] after the TV step."""
    if tvm in ('ITV3D_FGP', 'ITV2D_cham'):
        tvm = 'tv_chambolle'                     # joint_pnp_sci_algo.py:608-611
    _check_tv(denoiser, tvm, multichannel)
    Phi = _recover_phi(A, At, y, Phi)
    yh = f32c(_host(y))
    Xo = None if X_orig is None else f32c(_host(X_orig))
    H, W, Cc = Phi.shape
    with Solver(1, H, W, Cc, method="admm", _lambda=_lambda, gamma=gamma, tv_weight=tv_weight,
                tv_iter_max=tv_iter_max, fused=_base.USE_FUSED, clip=True) as s:
        s.load(yh[None], Phi, Phi_sum=f32c(_host(Phi_sum)),
               x0=None if x0 is None else f32c(_host(x0))[None],
               X_orig=None if (Xo is None or not show_iqa) else Xo[None])
        s.run(_total_iters(sigma, iter_max))
        x = s.get_x()[0]
        pa = [float(v) for v in s.psnr_all()[:, 0]]
    _progress('ADMM', pa)
    ps, ss = frames_iqa(Xo, x)
    return x, ps, ss, pa


def gap_denoise(y, Phi_sum, A=None, At=None, _lambda=1, gamma=None, accelerate=True,
                denoiser='tv', iter_max=50, noise_estimate=False, sigma=None, tv_weight=0.1,
                tv_iter_max=5, multichannel=True, x0=None, X_orig=None, model=None, show_iqa=True,
                tvm='tv_chambolle', Phi=None):
    """GAP-TV of the joint module: the loop of ``pnp_sci_algo.gap_denoise`` (``gamma`` is unused
    there as well)."""
    return _base.gap_denoise(y, Phi_sum, A, At, _lambda=_lambda, accelerate=accelerate,
                             denoiser=denoiser, iter_max=iter_max, noise_estimate=noise_estimate,
                             sigma=sigma, tv_weight=tv_weight, tv_iter_max=tv_iter_max,
                             multichannel=multichannel, x0=x0, X_orig=X_orig, model=model,
                             show_iqa=show_iqa, tvm=tvm, Phi=Phi)


def gap_multistep_denoise(y, Phi_sum, A=None, At=None, _lambda=1, accelerate=True,
                          denoiser='tv+ffdnet', iter_max=50, noise_estimate=False, sigma=None,
                          tv_weight=0.1, tv_iter_max=5, multichannel=True, x0=None, X_orig=None,
                          model=None, show_iqa=True, tvm='tv_chambolle', Phi=None,
                          second_denoiser=None):
    """TV + learned-denoiser period (joint_pnp_sci_algo.py:309-500).  Every iteration runs the GAP
    projection and the Chambolle TV step as one fused launch, then hands the current estimate --
    a float32 CUDA tensor ``[H, W, C]`` aliasing the solver's buffer -- to
    ``second_denoiser(x_dev, nsig, model)``, which changes it in place or returns a tensor of the
    same shape (the slot of ``ffdnet_vdenoiser`` :441 / ``fastdvdnet_denoiser`` :466).  Nothing
    leaves the device between the two steps.  PSNR is taken after both (:473), like the reference.
    """
    if str(denoiser).lower() not in ('tv+ffdnet', 'tv+fastdvdnet'):
        raise ValueError('Unsupported denoiser {}!'.format(denoiser))
    if tvm != 'tv_chambolle':                  # the other names call functions the reference never defines
        raise ValueError('Unsupported TV denoiser {}!'.format(tvm))
    if not multichannel:
        raise NotImplementedError("multichannel=False is not on the reference's hot path")
    if second_denoiser is None:
        raise NotImplementedError(
            "the learned denoisers of the reference (packages/ffdnet, packages/fastdvdnet) are not part "
            "of this engine: pass second_denoiser=callable(x_dev, nsig, model)")
    Phi = _recover_phi(A, At, y, Phi)
    yh = f32c(_host(y))
    Xo = None if X_orig is None else f32c(_host(X_orig))
    H, W, Cc = Phi.shape
    if not isinstance(sigma, list):
        sigma = [sigma]
    if not isinstance(iter_max, list):
        iter_max = [iter_max] * len(sigma)
    psnr_all = []
    with Solver(1, H, W, Cc, method="gap", accelerate=accelerate, _lambda=_lambda, tv_weight=tv_weight,
                tv_iter_max=tv_iter_max, fused=_base.USE_FUSED) as s:
        s.load(yh[None], Phi, Phi_sum=f32c(_host(Phi_sum)),
               x0=None if x0 is None else f32c(_host(x0))[None])
        dev = torch.device("cuda", torch.cuda.current_device())
        Xd = None if (Xo is None or not show_iqa) else torch.from_numpy(Xo).to(dev)
        for idx, nsig in enumerate(sigma):
            for _ in range(int(iter_max[idx])):
                s.run(1)                                           # projection + TV (:412-429)
                xv = _wrap(s.state_ptrs()[0], (H, W, Cc), dev)     # the estimate, in place
                out = second_denoiser(xv, nsig, model)             # :441 / :466
                if out is not None and out is not xv:
                    xv.copy_(out.to(dev, torch.float32).reshape(H, W, Cc))
                if Xd is not None:
                    psnr_all.append(psnr(Xd, xv))                  # :473
        x = s.get_x()[0]
    ps, ss = frames_iqa(Xo, x)
    return x, ps, ss, psnr_all


def gap_joint_denoise(y, Phi_sum, A=None, At=None, x0=None, X_orig=None, denoiser='tv+ffdnet',
                      iter_max1=50, iter_max2=50, sigma1=None, sigma2=None, **args):
    """Two periods (joint_pnp_sci_algo.py:100-116): GAP-TV, then ``gap_multistep_denoise`` from its
    result; returns what the second period returns.  ``second_denoiser=`` (and ``Phi=``) travel in
    ``args`` like the reference's other keyword arguments."""
    second = args.pop("second_denoiser", None)
    x, _, _, _ = gap_denoise(y, Phi_sum, A, At, x0=x0, X_orig=X_orig, denoiser='tv',
                             iter_max=iter_max1, sigma=sigma1, **args)
    return gap_multistep_denoise(y, Phi_sum, A, At, x0=x, X_orig=X_orig, denoiser=denoiser,
                                 iter_max=iter_max2, sigma=sigma2, second_denoiser=second, **args)


def admm_multistep_denoise(y, Phi_sum, A=None, At=None, _lambda=1, gamma=0.0, accelerate=None,
                           denoiser='tv+ffdnet', iter_max=50, noise_estimate=False, sigma=None,
                           tv_weight=0.1, tv_iter_max=5, multichannel=True, x0=None, model=None,
                           X_orig=None, show_iqa=True, tvm='tv_chambolle', Phi=None,
                           second_denoiser=None):
    """ADMM twin of ``gap_multistep_denoise`` (joint_pnp_sci_algo.py:118-306).  One launch per
    iteration does the projection and ``theta = TV(x - b)`` (:213-230); ``theta`` then goes to
    ``second_denoiser(theta_dev, nsig, model)`` (:242 / :263) as a CUDA tensor, is clipped to
    [0, 1] (:268), and the multiplier is formed from it, ``b = b - (x - theta)`` (:270), with
    the reference's operation order.  Returns ``x`` and the PSNR of ``x`` (:273), like the
    reference."""
    if str(denoiser).lower() not in ('tv+ffdnet', 'tv+fastdvdnet'):
        raise ValueError('Unsupported denoiser {}!'.format(denoiser))
    if tvm not in ('tv_chambolle', 'ITV3D_FGP', 'ITV2D_cham'):      # all three call denoise_tv_chambolle here
        raise ValueError('Unsupported TV denoiser {}!'.format(tvm))
    if not multichannel:
        raise NotImplementedError("multichannel=False is not on the reference's hot path")
    if second_denoiser is None:
        raise NotImplementedError(
            "the learned denoisers of the reference (packages/ffdnet, packages/fastdvdnet) are not part "
            "of this engine: pass second_denoiser=callable(theta_dev, nsig, model)")
    Phi = _recover_phi(A, At, y, Phi)
    yh = f32c(_host(y))
    Xo = None if X_orig is None else f32c(_host(X_orig))
    H, W, Cc = Phi.shape
    if not isinstance(sigma, list):
        sigma = [sigma]
    if not isinstance(iter_max, list):
        iter_max = [iter_max] * len(sigma)
    psnr_all = []
    with Solver(1, H, W, Cc, method="admm", _lambda=_lambda, gamma=gamma, tv_weight=tv_weight,
                tv_iter_max=tv_iter_max, fused=_base.USE_FUSED, clip=False) as s:
        s.load(yh[None], Phi, Phi_sum=f32c(_host(Phi_sum)),
               x0=None if x0 is None else f32c(_host(x0))[None])
        dev = torch.device("cuda", torch.cuda.current_device())
        Xd = None if (Xo is None or not show_iqa) else torch.from_numpy(Xo).to(dev)
        shape = (H, W, Cc)
        for idx, nsig in enumerate(sigma):
            for _ in range(int(iter_max[idx])):
                b_old = _wrap(s.admm_state_ptrs()[1], shape, dev).clone()
                s.run(1)                                   # x, theta = TV(x - b)   (:213-230)
                tp, bp, xp = s.admm_state_ptrs()           # the buffers ping-pong: ask again
                theta, b, x = (_wrap(p, shape, dev) for p in (tp, bp, xp))
                out = second_denoiser(theta, nsig, model)  # :242 / :263
                if out is not None and out is not theta:
                    theta.copy_(out.to(dev, torch.float32).reshape(shape))
                theta.clamp_(0, 1)                          # :268
                b.copy_(b_old - (x - theta))                # :270 (the launch formed b from the TV output)
                if Xd is not None:
                    psnr_all.append(psnr(Xd, x))            # :273
        x = s.get_x()[0]
    ps, ss = frames_iqa(Xo, x)
    return x, ps, ss, psnr_all


def admm_joint_denoise(y, Phi_sum, A=None, At=None, x0=None, X_orig=None, denoiser='tv+ffdnet',
                       iter_max1=50, iter_max2=50, sigma1=None, sigma2=None, **args):
    """Two periods (joint_pnp_sci_algo.py:81-98): the ADMM-TV above (theta clipped), then
    ``admm_multistep_denoise`` from its result; returns what the second period returns."""
    second = args.pop("second_denoiser", None)
    x, _, _, _ = admm_denoise(y, Phi_sum, A, At, x0=x0, X_orig=X_orig, denoiser='tv',
                              iter_max=iter_max1, sigma=sigma1, **args)
    return admm_multistep_denoise(y, Phi_sum, A, At, x0=x, X_orig=X_orig, denoiser=denoiser,
                                  iter_max=iter_max2, sigma=sigma2, second_denoiser=second, **args)
