"""The IQA helpers of the oracle (oracle/iqa.py) restate scikit-image 0.17's compare_psnr / compare_ssim, a third-party
dependency that is absent here (parity unpinned).  Corroboration from the definitions: SSIM of Wang et al. with a
7x7 uniform window over every fully covered position, sample covariances (N/(N-1)), K1 = 0.01, K2 = 0.03 -- computed
window by window with explicit loops -- and PSNR = 10 log10(range^2 / MSE)."""
import numpy as np

from oracle.iqa import compare_psnr, compare_ssim


def _ssim_by_definition(X, Y, win=7, data_range=1.0):
    X = X.astype(np.float64)
    Y = Y.astype(np.float64)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    n = win * win
    vals = []
    for r in range(X.shape[0] - win + 1):
        for c in range(X.shape[1] - win + 1):
            a = X[r:r + win, c:c + win].ravel()
            b = Y[r:r + win, c:c + win].ravel()
            ma, mb = a.mean(), b.mean()
            va = ((a - ma) ** 2).sum() / (n - 1)
            vb = ((b - mb) ** 2).sum() / (n - 1)
            cab = ((a - ma) * (b - mb)).sum() / (n - 1)
            vals.append(((2 * ma * mb + C1) * (2 * cab + C2)) / ((ma ** 2 + mb ** 2 + C1) * (va + vb + C2)))
    return float(np.mean(vals))


def test_ssim_restatement_equals_the_definition():
    rng = np.random.default_rng(2)
    for shape in ((16, 19), (9, 31), (7, 7)):
        X = rng.random(shape).astype(np.float32)
        Y = np.clip(X + 0.1 * rng.standard_normal(shape), 0, 1).astype(np.float32)
        assert abs(compare_ssim(X, Y, data_range=1.) - _ssim_by_definition(X, Y)) < 1e-12
    assert abs(compare_ssim(X, X, data_range=1.) - 1.0) < 1e-12


def test_psnr_restatement_equals_the_definition():
    rng = np.random.default_rng(3)
    X = rng.random((12, 10)).astype(np.float32)
    Y = (X + np.float32(0.05) * rng.standard_normal((12, 10)).astype(np.float32)).astype(np.float32)
    mse = np.mean((X.astype(np.float64) - Y.astype(np.float64)) ** 2)
    assert abs(compare_psnr(X, Y, data_range=1.) - 10 * np.log10(1.0 / mse)) < 1e-5
