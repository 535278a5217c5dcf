"""Oracle (test infrastructure -- only tests/, smoke() and bench.py's cpu legs may import this):
NumPy restatement of the MATLAB twin's default TV denoiser, SURVEY.md section 8f-2.

    TV_denoising   PnP_SCI/matlab/algorithms/tvdenoisers/TV_denoising.m:1-44  (helpers :51-68)

Parity unpinned: there is no MATLAB or Octave in this image, so this file was never run against
the .m source; it follows it line by line (cited below) and is checked by properties
(tests/test_oracle_tv.py): mean preservation, constants are fixed points, lambda -> 0 is the
identity, the adjoint pairs (dh, dht) / (dv, dvt), frames are independent.
Arithmetic follows the input dtype (float32 in -> float32 throughout, as MATLAB does for single).
"""
import numpy as np

__all__ = ["TV_denoising", "dh", "dv", "dht", "dvt"]

ALPHA = 5                                            # TV_denoising.m:8


def dv(x):                                           # :51-52   diff(x)
    return x[1:] - x[:-1]


def dh(x):                                           # :55-56   diff(x,1,2)
    return x[:, 1:] - x[:, :-1]


def dvt(z):                                          # :59-60 / :67-68 (3-D)   [-z(1,:); -diff(z); z(end,:)]
    return np.concatenate([-z[:1], -(z[1:] - z[:-1]), z[-1:]], axis=0)


def dht(z):                                          # :63-64 / :71-72 (3-D)   [-z(:,1) -diff(z,1,2) z(:,end)]
    return np.concatenate([-z[:, :1], -(z[:, 1:] - z[:, :-1]), z[:, -1:]], axis=1)


def _clip(x, lam):                                   # :83-84   sign(x).*min(abs(x),lambda)
    return np.sign(x) * np.minimum(np.abs(x), lam)


def TV_denoising(y0, lam, iters=100):
    """2-D image [H, W] or stack of frames [H, W, F] (TV per frame, :20-30)."""
    y0 = np.asarray(y0)
    if y0.ndim not in (2, 3) or y0.shape[0] < 2 or y0.shape[1] < 2:
        raise ValueError("TV_denoising restated for [H, W] and [H, W, F] inputs with H, W >= 2")
    ft = y0.dtype if y0.dtype in (np.float32, np.float64) else np.float64
    y0 = y0.astype(ft, copy=False)
    c = ft.type(1.0 / ALPHA)
    half = ft.type(lam / 2.0)
    zh = np.zeros((y0.shape[0], y0.shape[1] - 1) + y0.shape[2:], ft)       # :14-15 / :22-23
    zv = np.zeros((y0.shape[0] - 1, y0.shape[1]) + y0.shape[2:], ft)
    x0 = y0
    for _ in range(int(iters)):
        x0h = y0 - dht(zh)                                                  # :17 / :25
        x0v = y0 - dvt(zv)                                                  # :18 / :26
        x0 = (x0h + x0v) / ft.type(2)                                       # :19 / :27
        zh = _clip(zh + c * dh(x0), half)                                   # :20 / :28
        zv = _clip(zv + c * dv(x0), half)                                   # :21 / :29
    return x0
