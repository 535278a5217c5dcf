"""Data side of the reference's drivers, kept on the device:

    load_mat            pnp_sci_test_orig.py:94-110   ``orig`` / ``mask`` from MATLAB files
    synth_measurements  pnp_sci_test_orig.py:112-140  meas = sum_c mask*orig per coded frame,
                                                      Gaussian / Poisson noise, mask normalisation
    binary_mask         [dataset]/#code/binary_mask.m:63-68   Bernoulli(p) coded apertures
    shift_mask          [dataset]/#code/shift_mask.m, DeSCI/test_desci_cassi.m:53-62  CASSI stack

Measurement synthesis runs through ``scipnp_A`` (one batched launch for all coded frames) and
``scipnp_cassi_shift_mask``; random numbers come from torch's device generator, so noisy
measurements are reproducible per seed here but not bit-equal to NumPy's ``randn`` stream.
"""
import numpy as np
import torch

from ._lib import lib, check
from .engine import to_device, stream_ptr, dptr

__all__ = ["load_mat", "synth_measurements", "binary_mask", "shift_mask"]


def load_mat(path, names=("orig", "mask")):
    """Arrays ``names`` of a MATLAB file as float32, in MATLAB's index order ``[H, W, F]``.
    v5-v7.2 files go through ``scipy.io.loadmat``; v7.3 (HDF5) files need ``h5py`` and are
    transposed ``(2, 1, 0)`` like pnp_sci_test_orig.py:103-110."""
    with open(path, "rb") as f:
        head = f.read(128)
    out = {}
    if head.startswith(b"\x89HDF") or head.startswith(b"MATLAB 7.3"):
        try:
            import h5py
        except ImportError as e:                      # same dependency as the reference (:103)
            raise ImportError("MATLAB v7.3 files need h5py (not installed here)") from e
        with h5py.File(path, "r") as f:
            for n in names:
                if n in f:
                    a = np.float32(np.array(f[n]))
                    out[n] = a.transpose(tuple(range(a.ndim))[::-1])
    else:
        import scipy.io as sio
        m = sio.loadmat(path)
        for n in names:
            if n in m:
                out[n] = np.float32(np.array(m[n]))
    missing = [n for n in names if n not in out]
    if len(missing) == len(names):
        raise KeyError("none of %s found in %s" % (list(names), path))
    return out


def binary_mask(H, W, C, p=0.5, seed=0, device=None):
    """Bernoulli(p) 0/1 masks ``[H, W, C]`` as a float32 CUDA tensor."""
    dev = torch.device(device or "cuda")
    g = torch.Generator(device=dev).manual_seed(int(seed))
    return (torch.rand((H, W, C), device=dev, generator=g) <= p).to(torch.float32)


def shift_mask(mask2d, nband, step):
    """CASSI mask stack ``Phi[h, w + step*k, k] = mask2d[h, w]`` on the sheared canvas
    ``[H, W + (nband-1)*step, nband]`` (device)."""
    m = to_device(mask2d).contiguous()
    if m.dim() != 2:
        raise ValueError("shift_mask expects a 2-D coded aperture")
    H, W = m.shape
    Phi = torch.empty((H, W + (int(nband) - 1) * int(step), int(nband)), dtype=torch.float32, device=m.device)
    check(lib.scipnp_cassi_shift_mask(dptr(m), dptr(Phi), H, W, int(nband), int(step), stream_ptr()))
    return Phi


def synth_measurements(orig, mask, gaussian_noise_level=0.0, poisson_noise=False, seed=None):
    """``(meas [H,W,F], mask [H,W,C])`` as CUDA tensors from ``orig [H, W, F*C]`` and ``mask``:
    ``meas[:,:,i] = sum_c orig[:,:,i*C+c]*mask[:,:,c]`` (:113-118), plus Gaussian noise of the given
    level and optional Poisson noise (:127-130), then both divided by ``max(mask)`` (:134-136)."""
    od, md = to_device(orig), to_device(mask).contiguous()
    if od.dim() != 3 or md.dim() != 3 or od.shape[:2] != md.shape[:2] or od.shape[2] % md.shape[2]:
        raise ValueError("orig must be [H, W, F*C] for a mask [H, W, C]")
    H, W, Cc = md.shape
    F = od.shape[2] // Cc
    frames = od.reshape(H, W, F, Cc).permute(2, 0, 1, 3).contiguous()            # [F][H][W][C]
    meas = torch.empty((F, H, W), dtype=torch.float32, device=md.device)
    check(lib.scipnp_A(dptr(frames), dptr(md), dptr(meas), F, H, W, Cc, 0, stream_ptr()))
    meas = meas.permute(1, 2, 0).contiguous()                                    # [H][W][F]
    if gaussian_noise_level or poisson_noise:
        g = torch.Generator(device=md.device)
        g.manual_seed(0 if seed is None else int(seed))
        if gaussian_noise_level:
            meas = meas + float(gaussian_noise_level) * torch.randn(meas.shape, device=md.device, generator=g)
        if poisson_noise:
            meas = torch.poisson(meas.clamp_min(0), generator=g)
    mmax = md.max()
    return meas / mmax, md / mmax
