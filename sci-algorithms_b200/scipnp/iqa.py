"""Per-frame image-quality numbers of the solvers' return tuples
(``compare_psnr`` / ``compare_ssim`` of scikit-image < 0.18, called at
pnp_sci_algo.py:699-705, 857-863).  Computed once per reconstruction on the
final frames, on the host: off the hot path.
"""
import numpy as np
from scipy.ndimage import uniform_filter

__all__ = ["frame_psnr", "frame_ssim", "frames_iqa"]


def frame_psnr(ref, img, data_range=1.):
    ref = np.asarray(ref, dtype=np.float32)
    img = np.asarray(img, dtype=np.float32)
    mse = np.mean((ref - img) ** 2, dtype=np.float64)
    return 10 * np.log10((data_range ** 2) / mse)


def frame_ssim(ref, img, data_range=1., win=7):
    X = np.asarray(ref, dtype=np.float64)
    Y = np.asarray(img, dtype=np.float64)
    n = win ** X.ndim
    cn = n / (n - 1.)
    mx, my = uniform_filter(X, win), uniform_filter(Y, win)
    vx = cn * (uniform_filter(X * X, win) - mx * mx)
    vy = cn * (uniform_filter(Y * Y, win) - my * my)
    vxy = cn * (uniform_filter(X * Y, win) - mx * my)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    S = ((2 * mx * my + c1) * (2 * vxy + c2)) / ((mx * mx + my * my + c1) * (vx + vy + c2))
    pad = (win - 1) // 2
    return float(S[tuple(slice(pad, d - pad) for d in S.shape)].mean())


def frames_iqa(X_orig, x):
    """(psnr_, ssim_) lists over the last axis, or two empty lists."""
    if X_orig is None:
        return [], []
    ps = [frame_psnr(X_orig[..., c], x[..., c]) for c in range(x.shape[-1])]
    ss = [frame_ssim(X_orig[..., c], x[..., c]) for c in range(x.shape[-1])]
    return ps, ss
