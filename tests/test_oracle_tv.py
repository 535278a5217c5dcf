"""R6 oracle checks (CPU): regression vectors, the in-tree MATLAB cross-check and
the mathematical properties of the Chambolle projection."""
import numpy as np
import pytest

from oracle.tv_chambolle import denoise_tv_chambolle, tv_chambolle_2d


@pytest.mark.parametrize("T", [1, 2, 5, 200])
def test_regression_vectors(golden, T):
    g = golden("tv_T%d" % T)
    en = []
    out = denoise_tv_chambolle(g["image"], float(g["weight"]), n_iter_max=int(g["n_iter_max"]),
                               multichannel=True, energy_out=en)
    np.testing.assert_array_equal(out, g["out"])
    np.testing.assert_array_equal([len(e) for e in en], g["n_exec"])


def test_early_stop_fires_in_vector(golden):
    g = golden("tv_T200")
    assert (g["n_exec"] < 200).all() and (g["n_exec"] > 2).all()


def _matlab_itv2d(f, lam, iters, dt):
    """Transliteration of /root/reference/PnP_SCI/matlab/algorithms/tvdenoisers/
    tvdenoise_cham_ITV2D.m:49-72,93 (2-D branch) with dt as a parameter
    (the file hard-codes dt = 1/8 at :49)."""
    H, W = f.shape
    idn = np.r_[1:H, H - 1]
    iu = np.r_[0, 0:H - 1]
    ir = np.r_[1:W, W - 1]
    il = np.r_[0, 0:W - 1]
    p1 = np.zeros_like(f)
    p2 = np.zeros_like(f)
    divp = np.zeros_like(f)
    for _ in range(iters):
        z = divp - f * lam
        z1 = z[:, ir] - z
        z2 = z[idn, :] - z
        den = 1 + dt * np.sqrt(z1 ** 2 + z2 ** 2)
        p1 = (p1 + dt * z1) / den
        p2 = (p2 + dt * z2) / den
        divp = p1 - p1[:, il] + p2 - p2[iu, :]
    return f - divp / lam


@pytest.mark.parametrize("T", [2, 5, 10])
def test_matches_in_tree_matlab_chambolle_interior(T):
    """The only TV source inside the reference tree is MATLAB.  With dt=1/4,
    lambda=1/w and iters=T-1 its isotropic Chambolle iteration equals the
    restated skimage recurrence away from the top/left border (the two differ in
    how the divergence treats index 0), which corroborates tau, the update rule
    and the T-1 effective-update count."""
    rng = np.random.default_rng(3)
    f = rng.random((40, 36))
    w = 0.3
    ours = tv_chambolle_2d(f, w, eps=0.0, n_iter_max=T)
    theirs = _matlab_itv2d(f, 1.0 / w, T - 1, 0.25)
    np.testing.assert_allclose(ours[T:, T:], theirs[T:, T:], rtol=0, atol=1e-13)
    # an off-by-one in the iteration count is detectable
    wrong = _matlab_itv2d(f, 1.0 / w, T, 0.25)
    assert np.abs(ours[T:, T:] - wrong[T:, T:]).max() > 1e-3


def test_constant_image_is_fixed_point():
    f = np.full((9, 11, 2), 0.37, np.float32)
    np.testing.assert_array_equal(denoise_tv_chambolle(f, 0.3, n_iter_max=5, multichannel=True), f)


def test_n_iter_one_returns_input():
    rng = np.random.default_rng(0)
    f = rng.random((8, 7, 3)).astype(np.float32)
    np.testing.assert_array_equal(denoise_tv_chambolle(f, 0.3, n_iter_max=1, multichannel=True), f)


def test_channels_independent_and_mean_preserving():
    rng = np.random.default_rng(1)
    f = rng.random((16, 12, 4))
    out = denoise_tv_chambolle(f, 0.2, n_iter_max=6, multichannel=True)
    for c in range(4):
        np.testing.assert_array_equal(out[..., c], tv_chambolle_2d(f[..., c], 0.2, n_iter_max=6))
    # f + div p has the mean of f (the divergence telescopes to zero)
    np.testing.assert_allclose(out.mean(axis=(0, 1)), f.mean(axis=(0, 1)), atol=1e-14)


def test_rof_objective_decreases():
    rng = np.random.default_rng(2)
    f = rng.random((24, 24))
    w = 0.25

    def rof(u):
        g0 = np.zeros_like(u); g1 = np.zeros_like(u)
        g0[:-1] = u[1:] - u[:-1]; g1[:, :-1] = u[:, 1:] - u[:, :-1]
        return 0.5 * ((u - f) ** 2).sum() + w * np.sqrt(g0 ** 2 + g1 ** 2).sum()

    vals = [rof(tv_chambolle_2d(f, w, eps=0.0, n_iter_max=T)) for T in (1, 20, 80, 300)]
    assert vals[0] > vals[1] > vals[2] > vals[3]


def test_float32_dtype_preserved():
    f = np.random.default_rng(4).random((6, 6, 2)).astype(np.float32)
    assert denoise_tv_chambolle(f, 0.1, n_iter_max=3, multichannel=True).dtype == np.float32


# -- MATLAB twin's default TV: oracle/matlab_tv.py, property checks -----------------

def _frames(shape, seed=3, dtype=np.float32):
    rng = np.random.default_rng(seed)
    return rng.random(shape).astype(dtype)


def test_matlab_tv_adjoint_pairs():
    from oracle import matlab_tv as M
    x = _frames((9, 11, 3), dtype=np.float64)
    zh = _frames((9, 10, 3), 4, np.float64)
    zv = _frames((8, 11, 3), 5, np.float64)
    assert abs(np.sum(M.dh(x) * zh) - np.sum(x * M.dht(zh))) < 1e-10       # TV_denoising.m:55-64
    assert abs(np.sum(M.dv(x) * zv) - np.sum(x * M.dvt(zv))) < 1e-10


@pytest.mark.parametrize("iters", [1, 5, 30])
def test_matlab_tv_properties(iters):
    from oracle import matlab_tv as M
    y = _frames((13, 10, 4), dtype=np.float64)
    out = M.TV_denoising(y, 0.2, iters)
    assert out.shape == y.shape and out.dtype == y.dtype
    np.testing.assert_allclose(out.mean(axis=(0, 1)), y.mean(axis=(0, 1)), atol=1e-12)   # sum(dht z) = 0
    const = np.full((6, 7, 2), 0.3)
    np.testing.assert_array_equal(M.TV_denoising(const, 0.5, iters), const)             # fixed point
    np.testing.assert_allclose(M.TV_denoising(y, 0.0, iters), y, atol=0)                 # lambda = 0
    one = M.TV_denoising(y[:, :, 2], 0.2, iters)                                         # frames are independent
    np.testing.assert_array_equal(one, out[:, :, 2])
    if iters > 1:                                                                        # it does smooth
        tv = lambda a: np.abs(np.diff(a, axis=0)).sum() + np.abs(np.diff(a, axis=1)).sum()
        assert tv(out) < tv(y)


def test_matlab_tv_keeps_single_precision():
    from oracle import matlab_tv as M
    y = _frames((8, 8, 2))
    assert M.TV_denoising(y, 0.1, 5).dtype == np.float32
    with pytest.raises(ValueError):
        M.TV_denoising(y[:1], 0.1, 5)
