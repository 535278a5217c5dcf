"""MATLAB twin's default TV denoiser on the device:
``TV_denoising(y0, lambda, iter)`` of ``PnP_SCI/matlab/algorithms/tvdenoisers/TV_denoising.m``
(the 'ATV_ClipA' branch of ``gapdenoise.m:93-94``), and the GAP loop of ``gapdenoise.m:76-94``
around it.  Everything runs through ``libscipnp.so`` (``scipnp_tv_atv_clip``,
``scipnp_gap_project``); float32.
"""
import numpy as np
import torch

from ._lib import lib, check
from .engine import to_device, stream_ptr, dptr, is_torch
from .utils import psnr

__all__ = ["TV_denoising", "gapdenoise", "TV_denoising_clip_LB", "tvdenoise_cham_ATV2D", "tvdenoise_cham_ITV2D",
           "tvdenoise_cham_ITV3D", "fgp_denoise_ATV2D", "fgp_denoise_ITV2D", "fgp_denoise_ITV3D"]


def _ret(t, like):
    return t if is_torch(like) else t.cpu().numpy()


def _tv_dev(xd, lam, iters, out=None, ws=None):
    H, W, Cc = xd.shape
    if out is None:
        out = torch.empty_like(xd)
    nbytes = lib.scipnp_tv_atv_clip_workspace_bytes(1, H, W, Cc)
    if ws is None:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=xd.device)
    check(lib.scipnp_tv_atv_clip(dptr(xd), dptr(out), float(lam), int(iters), 1, H, W, Cc,
                                 dptr(ws), nbytes, stream_ptr()))
    return out


def TV_denoising(y0, lam, iter=100):
    """``[H, W]`` image or ``[H, W, F]`` stack (TV per frame).  NumPy in -> NumPy out, CUDA
    tensor in -> CUDA tensor out."""
    xd = to_device(y0)
    two_d = xd.dim() == 2
    if two_d:
        xd = xd[..., None]
    if xd.dim() != 3 or xd.shape[0] < 2 or xd.shape[1] < 2:
        raise ValueError("TV_denoising expects [H, W] or [H, W, F] with H, W >= 2")
    out = _tv_dev(xd.contiguous(), lam, iter)
    return _ret(out[..., 0] if two_d else out, y0)


def _family(variant, x, lam, iters, out=None, ws=None):
    """One member of the family through scipnp_tv_matlab; [H, W, F] stacks (the branch gapdenoise.m uses)."""
    xd = to_device(x)
    if xd.dim() != 3 or xd.shape[0] < 2 or xd.shape[1] < 2:
        raise ValueError("expects an [H, W, F] stack with H, W >= 2")
    xd = xd.contiguous()
    H, W, Cc = xd.shape
    if out is None:
        out = torch.empty_like(xd)
    nbytes = lib.scipnp_tv_matlab_workspace_bytes(1, H, W, Cc)
    if ws is None:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=xd.device)
    check(lib.scipnp_tv_matlab(dptr(xd), dptr(out), int(variant), float(lam), int(iters), 1, H, W, Cc,
                               dptr(ws), nbytes, stream_ptr()))
    return out


def TV_denoising_clip_LB(y0, lam, iter=20):
    """TV_denoising_clip_LB.m (3-D branch :25-36): iterative clipping without the averaging, clip level lambda."""
    return _ret(_family(0, y0, lam, iter), y0)


def tvdenoise_cham_ATV2D(f, lam, iters):
    """tvdenoise_cham_ATV2D.m:72-87 (anisotropic Chambolle projection, dt = 1/8)."""
    return _ret(_family(1, f, lam, iters), f)


def tvdenoise_cham_ITV2D(f, lam, iters):
    """tvdenoise_cham_ITV2D.m:73-90 (isotropic per frame, dt = 1/8)."""
    return _ret(_family(2, f, lam, iters), f)


def tvdenoise_cham_ITV3D(f, lam, iters):
    """tvdenoise_cham_ITV3D.m:72-90 (gradient norm summed over the frames, dt = 1/4)."""
    return _ret(_family(3, f, lam, iters), f)


def fgp_denoise_ATV2D(Xobs, lam, MAXITER):
    """fgp_denoise_ATV2D.m:73-121 (fast gradient projection, anisotropic)."""
    return _ret(_family(4, Xobs, lam, MAXITER), Xobs)


def fgp_denoise_ITV2D(Xobs, lam, MAXITER):
    return _ret(_family(5, Xobs, lam, MAXITER), Xobs)


def fgp_denoise_ITV3D(Xobs, lam, MAXITER):
    return _ret(_family(6, Xobs, lam, MAXITER), Xobs)


# gapdenoise.m:92-108: tvm -> (variant, lambda as passed there, iterations as written there)
_TVM = {'ATV_ClipB': (0, lambda w: w, 5), 'ATV_cham': (1, lambda w: 1. / w, 5), 'ATV_FGP': (4, lambda w: w, 2),
        'ITV2D_cham': (2, lambda w: 1. / w, 5), 'ITV2D_FGP': (5, lambda w: w, 2),
        'ITV3D_cham': (3, lambda w: 1. / w, 5), 'ITV3D_FGP': (6, lambda w: w, 2)}


def gapdenoise(y, Phi, Phisum=None, lambda_=0.2, maxiter=100, acc=True, tvweight=0.07, tviter=5,
               tvm='ATV_ClipA', v0=None, orig=None):
    """GAP with the MATLAB default TV (gapdenoise.m:62-94; defaults of :30-37).  ``Phisum``
    defaults to ``sum(Phi.^2, 3)`` with zeros set to one, as the MATLAB drivers form it
    (tests/test_pnpsci_benchmark_full.m:57).  Returns ``(v, psnrall)``."""
    if tvm != 'ATV_ClipA' and tvm not in _TVM:
        raise ValueError("no such tvdenoiser")                      # gapdenoise.m:110-111
    Pd = to_device(Phi).contiguous()
    yd = to_device(y).contiguous()
    H, W, Cc = Pd.shape
    if Phisum is None:
        ps = (Pd * Pd).sum(2)
        ps[ps == 0] = 1
    else:
        ps = to_device(Phisum).contiguous()
    v = torch.empty((H, W, Cc), dtype=torch.float32, device=Pd.device)
    if v0 is None:
        check(lib.scipnp_At(dptr(yd), dptr(Pd), dptr(v), 1, H, W, Cc, 0, stream_ptr()))      # :49
    else:
        v.copy_(to_device(v0))
    y1 = torch.zeros_like(yd)                                                               # :58
    f = torch.empty_like(v)
    ws = torch.empty(lib.scipnp_tv_matlab_workspace_bytes(1, H, W, Cc), dtype=torch.uint8, device=v.device)
    od = None if orig is None else to_device(orig)
    psnrall = []
    for _ in range(int(maxiter)):
        # Euclidean projection (:68-75): f = v + lambda*At((y1 - yb)/Phisum), y1 updated in place
        check(lib.scipnp_gap_project(dptr(v), dptr(f), dptr(y1), dptr(y1), dptr(yd), dptr(Pd), dptr(ps),
                                     float(lambda_), 1 if acc else 0, 1, H, W, Cc, 0, stream_ptr()))
        if tvm == 'ATV_ClipA':
            _tv_dev(f, tvweight, tviter, out=v, ws=ws)                                       # :93-94
        else:
            var, lam_of, nit = _TVM[tvm]
            _family(var, f, lam_of(float(tvweight)), nit, out=v, ws=ws)                      # :95-108
        if od is not None:
            psnrall.append(psnr(od, v))
    return _ret(v, y), psnrall
