"""Per-frame image-quality numbers of the solvers' return tuples
(``compare_psnr`` / ``compare_ssim`` of scikit-image < 0.18, called at
pnp_sci_algo.py:699-705, 857-863), computed on the device by
``scipnp_frames_iqa``: once per reconstruction on the
final frames, off the timed path.  Inputs may be host arrays (uploaded) or CUDA
tensors.
"""
import math

import numpy as np
import torch

from ._lib import lib, check
from .engine import to_device, stream_ptr, dptr

__all__ = ["frame_psnr", "frame_ssim", "frames_iqa"]


def _iqa(ref, img):
    a, b = to_device(ref), to_device(img)
    if a.shape != b.shape or a.dim() not in (2, 3):
        raise ValueError("frames_iqa expects two [H, W] or [H, W, C] arrays of equal shape")
    if a.dim() == 2:
        a, b = a[..., None], b[..., None]
    H, W, Cc = a.shape
    if H < 7 or W < 7:
        raise ValueError("win_size exceeds image extent.")          # skimage's message for the 7x7 window
    a, b = a.contiguous(), b.contiguous()
    acc = torch.empty((2, Cc), dtype=torch.float64, device=a.device)
    check(lib.scipnp_frames_iqa(dptr(a), dptr(b), H, W, Cc, dptr(acc[0]), dptr(acc[1]), stream_ptr()))
    s = acc.cpu().numpy()
    ssim = s[0] / float((H - 6) * (W - 6))
    mse = s[1] / float(H * W)
    with np.errstate(divide="ignore"):
        psnr = 10 * np.log10(1.0 / mse)                             # data_range = 1
    return [float(v) for v in psnr], [float(v) for v in ssim]


def frame_psnr(ref, img, data_range=1.):
    if data_range != 1.:
        raise NotImplementedError("the reference calls compare_psnr with data_range=1.")
    return _iqa(ref, img)[0][0]


def frame_ssim(ref, img, data_range=1., win=7):
    if data_range != 1. or win != 7:
        raise NotImplementedError("the reference calls compare_ssim with data_range=1. and the default window")
    return _iqa(ref, img)[1][0]


def frames_iqa(X_orig, x):
    """(psnr_, ssim_) lists over the last axis, or two empty lists."""
    if X_orig is None:
        return [], []
    return _iqa(X_orig, x)
