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


# -- the rest of the MATLAB TV family (oracle/matlab_tv.py; parity unpinned: checked by properties) ----------------

_FAMILY = ["TV_denoising_clip_LB", "tvdenoise_cham_ATV2D", "tvdenoise_cham_ITV2D", "tvdenoise_cham_ITV3D",
           "fgp_denoise_ATV2D", "fgp_denoise_ITV2D", "fgp_denoise_ITV3D"]


def _lam(name):
    return 12.0 if "cham" in name else 0.08        # the Chambolle variants take 1/weight (gapdenoise.m:98)


@pytest.mark.parametrize("name", _FAMILY)
def test_matlab_family_properties(name):
    from oracle import matlab_tv as M
    fn = getattr(M, name)
    y = _frames((18, 22, 4))
    u = fn(y, _lam(name), 4)
    assert u.dtype == np.float32 and u.shape == y.shape
    # a constant stack is a fixed point
    c = np.full((8, 9, 3), 0.4, np.float32)
    np.testing.assert_allclose(fn(c, _lam(name), 3), c, atol=1e-6)
    # denoising reduces the total variation of the frames
    tv = lambda a: float(np.abs(np.diff(a, axis=0)).sum() + np.abs(np.diff(a, axis=1)).sum())
    assert tv(u) < tv(y)
    # divergence-form members keep the mean of every frame (the Getreuer variants repeat the first row / column in
    # the backward difference, tvdenoise_cham_*.m:54-57, and do not)
    if "cham" not in name:
        np.testing.assert_allclose(u.mean(axis=(0, 1)), y.mean(axis=(0, 1)), atol=2e-6)
    # per-frame members treat the frames independently; the ITV3D members couple them
    one = fn(y[:, :, 1:2].copy(), _lam(name), 4)[:, :, 0]
    if "ITV3D" in name:
        assert np.abs(one - u[:, :, 1]).max() > 1e-4
    else:
        np.testing.assert_array_equal(one, u[:, :, 1])


def test_matlab_cham_itv2d_equals_the_2d_transliteration():
    """The 3-D branch of tvdenoise_cham_ITV2D.m (:73-90) applied to one frame is its 2-D branch (:59-72),
    transliterated independently above (_matlab_itv2d)."""
    from oracle import matlab_tv as M
    y = _frames((15, 19, 1), dtype=np.float64)
    want = _matlab_itv2d(y[:, :, 0], 9.0, 6, 1.0 / 8)
    np.testing.assert_allclose(M.tvdenoise_cham_ITV2D(y, 9.0, 6)[:, :, 0], want, rtol=0, atol=1e-14)


def test_matlab_fgp_first_iterate_is_the_input_and_weights_follow_the_t_sequence():
    """fgp_denoise_*.m:85: with R = 0 the first D is Xobs itself; MAXITER = 1 therefore returns the input."""
    from oracle import matlab_tv as M
    y = _frames((10, 12, 3))
    for name in ("fgp_denoise_ATV2D", "fgp_denoise_ITV2D", "fgp_denoise_ITV3D"):
        np.testing.assert_array_equal(getattr(M, name)(y, 0.1, 1), y)


def test_independent_restatements_agree_on_the_minimisers():
    """Mutual corroboration of the unpinned restatements: run to convergence, algorithms of different families that
    minimise the same functional must meet.  Anisotropic ROF (1/2|u-f|^2 + w*TV_aniso): the iterative-clipping
    ATV_ClipB (TV_denoising_clip_LB.m) and the fast gradient projection ATV_FGP (fgp_denoise_ATV2D.m).  Isotropic
    ROF with the same weight: scikit-image's Chambolle iteration (oracle/tv_chambolle.py, R6) and ITV2D_FGP
    (fgp_denoise_ITV2D.m) -- two code bases, two boundary treatments, one minimiser."""
    from oracle import matlab_tv as M
    from oracle.tv_chambolle import denoise_tv_chambolle
    rng = np.random.default_rng(1)
    f = rng.random((24, 28, 2))
    w = 0.15
    atv_fgp = M.fgp_denoise_ATV2D(f, w, 400)
    assert np.abs(M.TV_denoising_clip_LB(f, w, 2000) - atv_fgp).max() < 1e-4
    itv_fgp = M.fgp_denoise_ITV2D(f, w, 400)
    sk = denoise_tv_chambolle(f, weight=w, eps=0.0, n_iter_max=3000, multichannel=True)
    assert np.abs(sk - itv_fgp).max() < 2e-3
    # and the two functionals are different problems: the check above is not vacuous
    assert np.abs(atv_fgp - itv_fgp).max() > 1e-2
