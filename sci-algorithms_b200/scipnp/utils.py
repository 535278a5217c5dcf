"""Drop-in for the hot-path part of the reference's ``PnP_SCI/python/utils.py``
(``A_``, ``At_``, ``psnr``; utils.py:10-36), computed on the GPU through
libscipnp.so.  NumPy in -> NumPy out; CUDA tensors in -> CUDA tensors out.
The plotting / saving helpers of the reference module (utils.py:41-184) are
outside the hot path and not provided.
"""
import ctypes as C
import math

import numpy as np
import torch

from ._lib import lib, check
from .engine import to_device, stream_ptr, is_torch, dptr

__all__ = ["A_", "At_", "psnr", "phi_sum"]


def _ret(t, like):
    return t if is_torch(like) else t.cpu().numpy()


def A_(x, Phi):
    """Forward model ``y = sum_c x[:,:,c]*Phi[:,:,c]`` (utils.py:10-15)."""
    xd, pd = to_device(x), to_device(Phi)
    if xd.shape != pd.shape or xd.dim() != 3:
        raise ValueError("A_ expects x and Phi of equal shape [H, W, C]")
    H, W, Cc = xd.shape
    y = torch.empty((H, W), dtype=torch.float32, device=xd.device)
    check(lib.scipnp_A(dptr(xd), dptr(pd), dptr(y), 1, H, W, Cc, 0, stream_ptr()))
    return _ret(y, x)


def At_(y, Phi):
    """Adjoint ``x[:,:,c] = y*Phi[:,:,c]`` (utils.py:17-26)."""
    yd, pd = to_device(y), to_device(Phi)
    if pd.dim() != 3 or tuple(yd.shape) != tuple(pd.shape[:2]):
        raise ValueError("At_ expects y [H, W] and Phi [H, W, C]")
    H, W, Cc = pd.shape
    x = torch.empty((H, W, Cc), dtype=torch.float32, device=pd.device)
    check(lib.scipnp_At(dptr(yd), dptr(pd), dptr(x), 1, H, W, Cc, 0, stream_ptr()))
    return _ret(x, y)


def phi_sum(Phi):
    """``sum_c Phi`` with zeros replaced by one (pnp_sci_algo.py:491-492)."""
    pd = to_device(Phi)
    H, W, Cc = pd.shape
    s = torch.empty((H, W), dtype=torch.float32, device=pd.device)
    check(lib.scipnp_phi_sum(dptr(pd), dptr(s), 1, H, W, Cc, stream_ptr()))
    return _ret(s, Phi)


def psnr(ref, img):
    """PSNR on [0,1] data, 100 when identical (utils.py:28-36)."""
    a, b = to_device(ref), to_device(img)
    if a.shape != b.shape:
        raise ValueError("psnr expects arrays of equal shape")
    acc = torch.zeros(1, dtype=torch.float64, device=a.device)
    check(lib.scipnp_sq_err(dptr(a), dptr(b), a.numel(), dptr(acc), stream_ptr()))
    mse = float(acc.item()) / a.numel()
    if mse == 0:
        return 100
    return 20 * math.log10(1. / math.sqrt(mse))
