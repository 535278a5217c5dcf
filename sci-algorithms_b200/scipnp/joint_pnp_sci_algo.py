"""Drop-in for the TV paths of the reference's ``PnP_SCI/python/joint_pnp_sci_algo.py``
(SURVEY.md section 8f-1): the module most ``pnp_sci_test_*`` drivers import.

    admm_denoise   joint_pnp_sci_algo.py:502-665   ADMM-TV with ``theta = clip(theta, 0, 1)`` (:633)
    gap_denoise    joint_pnp_sci_algo.py:666-...    same loop as pnp_sci_algo.gap_denoise

The two-period drivers (``gap_joint_denoise`` / ``admm_joint_denoise``, :81-116) run a TV period
and then a TV+CNN period; their first period is exactly the functions below, the CNN period is
outside the hot path.  ``tvm`` may be 'tv_chambolle', 'ITV3D_FGP' or 'ITV2D_cham' in ``admm_denoise``:
the reference calls ``denoise_tv_chambolle`` for all three (:606-611).
"""
import numpy as np

from . import pnp_sci_algo as _base
from .pnp_sci_algo import _check_tv, _recover_phi, _host, _total_iters, _progress
from .engine import Solver, f32c
from .iqa import frames_iqa
from .utils import A_, At_, psnr  # noqa: F401

__all__ = ["admm_denoise", "gap_denoise", "A_", "At_", "psnr"]


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
