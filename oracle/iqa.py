"""Oracle (test infrastructure): image-quality helpers the reference imports
from scikit-image < 0.18 (``pnp_sci_algo.py:13-14``): ``compare_psnr`` and
``compare_ssim`` as published in scikit-image 0.17.2
(``skimage/metrics/simple_metrics.py``, ``_structural_similarity.py``).
Off the hot path; restated so that the reference's
solver loops can run here and so the host-side IQA of the product can be
checked.
"""
import numpy as np
from scipy.ndimage import uniform_filter

__all__ = ["compare_psnr", "compare_ssim"]


def compare_psnr(im_true, im_test, data_range=None):
    a = np.asarray(im_true)
    b = np.asarray(im_test)
    if data_range is None:
        raise ValueError("oracle requires data_range (the reference passes 1.)")
    ft = np.result_type(a.dtype, b.dtype, np.float32)
    err = np.mean((a.astype(ft) - b.astype(ft)) ** 2, dtype=np.float64)
    return 10 * np.log10((data_range ** 2) / err)


def compare_ssim(X, Y, win_size=7, data_range=None, multichannel=False):
    X = np.asarray(X)
    Y = np.asarray(Y)
    if multichannel:
        vals = [compare_ssim(X[..., c], Y[..., c], win_size, data_range)
                for c in range(X.shape[-1])]
        return float(np.mean(vals))
    if data_range is None:
        raise ValueError("oracle requires data_range (the reference passes 1.)")
    K1, K2 = 0.01, 0.03
    X = X.astype(np.float64)
    Y = Y.astype(np.float64)
    NP = win_size ** X.ndim
    cov_norm = NP / (NP - 1)
    ux = uniform_filter(X, size=win_size)
    uy = uniform_filter(Y, size=win_size)
    uxx = uniform_filter(X * X, size=win_size)
    uyy = uniform_filter(Y * Y, size=win_size)
    uxy = uniform_filter(X * Y, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    C1 = (K1 * data_range) ** 2
    C2 = (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / \
        ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    core = S[tuple(slice(pad, s - pad) for s in S.shape)]
    return float(core.mean())
