"""Pin the oracle restatement (oracle/pnp_sci.py) to the reference's own outputs
(tests/golden/*.npz, produced by tests/golden/make_golden.py from the unmodified
reference code).  CPU only.  Bit-exact: same NumPy, same statement order."""
import numpy as np
import pytest

from oracle import pnp_sci as O
from oracle import reference_loader


def _ops(mask):
    return (lambda x: O.A_(x, mask)), (lambda y: O.At_(y, mask))


@pytest.mark.parametrize("tag", ["ops_8", "ops_5"])
def test_operators(golden, tag):
    g = golden(tag)
    np.testing.assert_array_equal(O.A_(g["x"], g["Phi"]), g["A"])
    np.testing.assert_array_equal(O.At_(g["y"], g["Phi"]), g["At"])
    np.testing.assert_array_equal(O.phi_sum(g["Phi"]), g["Phi_sum"])
    assert O.psnr(g["x"], g["x2"]) == float(g["psnr"])
    assert O.psnr(g["x"], g["x"]) == 100 == float(g["psnr_same"])


def test_gap_accelerated(golden):
    g = golden("gap_acc")
    A, At = _ops(g["mask"])
    x, ps, ss, pa = O.gap_denoise(g["y"], O.phi_sum(g["mask"]), A, At, _lambda=1,
                                  accelerate=True, denoiser='tv', iter_max=12,
                                  tv_weight=0.3, tv_iter_max=5, X_orig=g["X_orig"])
    np.testing.assert_array_equal(x, g["x"])
    np.testing.assert_array_equal(np.array(pa), g["psnr_all"])
    np.testing.assert_array_equal(np.array(ps), g["psnr"])
    np.testing.assert_allclose(np.array(ss), g["ssim"], rtol=0, atol=0)


def test_gap_plain_schedule(golden):
    g = golden("gap_plain")
    A, At = _ops(g["mask"])
    x, ps, ss, pa = O.gap_denoise(g["y"], O.phi_sum(g["mask"]), A, At, _lambda=0.75,
                                  accelerate=False, denoiser='tv', iter_max=[3, 4],
                                  sigma=[0.2, 0.1], tv_weight=0.1, tv_iter_max=3,
                                  X_orig=g["X_orig"])
    assert len(pa) == 7
    np.testing.assert_array_equal(x, g["x"])
    np.testing.assert_array_equal(np.array(pa), g["psnr_all"])


def test_admm(golden):
    g = golden("admm")
    A, At = _ops(g["mask"])
    x, ps, ss, pa = O.admm_denoise(g["y"], O.phi_sum(g["mask"]), A, At, _lambda=1,
                                   gamma=0.01, denoiser='tv', iter_max=12,
                                   tv_weight=0.3, tv_iter_max=5, X_orig=g["X_orig"])
    np.testing.assert_array_equal(x, g["x"])
    np.testing.assert_array_equal(np.array(pa), g["psnr_all"])
    np.testing.assert_array_equal(np.array(ps), g["psnr"])


def test_gap_warm_start_ragged(golden):
    g = golden("gap_c5_warm")
    A, At = _ops(g["mask"])
    x, _, _, pa = O.gap_denoise(g["y"], O.phi_sum(g["mask"]), A, At, iter_max=6,
                                tv_weight=0.2, tv_iter_max=4, x0=g["x0"],
                                X_orig=g["X_orig"])
    np.testing.assert_array_equal(x, g["x"])
    np.testing.assert_array_equal(np.array(pa), g["psnr_all"])


@pytest.mark.parametrize("pm", ["gap", "admm"])
@pytest.mark.parametrize("md", ["plain", "updown"])
def test_cacti_wrapper(golden, pm, md):
    g = golden("cacti_%s_%s" % (pm, md))
    A, At = _ops(g["mask"])
    kw = dict(_lambda=1, denoiser='tv', iter_max=5, tv_weight=0.3, tv_iter_max=5)
    kw.update({"accelerate": True} if pm == "gap" else {"gamma": 0.01})
    x_, t_, ps, ss, pa = O.admmdenoise_cacti(g["meas"], g["mask"], A, At, projmeth=pm,
                                             orig=g["orig"], nframe=2, MAXB=255.,
                                             maskdirection=md, **kw)
    np.testing.assert_array_equal(x_, g["x"])
    np.testing.assert_array_equal(np.array(pa), g["psnr_all"])
    np.testing.assert_array_equal(np.array(ps), g["psnr"])


def test_bayer(golden):
    g = golden("bayer")
    x, ps, ss, pa = O.gap_denoise_bayer(g["y_bayer"], g["Phi_bayer"], _lambda=1,
                                        accelerate=True, denoiser='tv', iter_max=8,
                                        tv_weight=0.1, tv_iter_max=5, X_orig=g["X_orig"])
    np.testing.assert_array_equal(x, g["x"])
    np.testing.assert_array_equal(np.array(pa), g["psnr_all"])
    np.testing.assert_array_equal(np.array(ps), g["psnr"])


def test_cassi(golden):
    g = golden("cassi")
    Phi = O.cassi_shift_mask(g["mask2d"], int(g["nband"]), int(g["step"]))
    np.testing.assert_array_equal(Phi, g["Phi"])
    A, At = _ops(Phi)
    x, _, _, pa = O.gap_denoise(g["y"], O.phi_sum(Phi), A, At, iter_max=8, tv_weight=0.1,
                                tv_iter_max=5, X_orig=g["X_orig"])
    np.testing.assert_array_equal(x, g["x"])
    np.testing.assert_array_equal(np.array(pa), g["psnr_all"])


def test_unsupported_denoiser_raises():
    y = np.zeros((4, 4), np.float32)
    m = np.ones((4, 4, 2), np.float32)
    A, At = _ops(m)
    with pytest.raises(ValueError):
        O.gap_denoise(y, O.phi_sum(m), A, At, denoiser='bm3d', iter_max=1)
    with pytest.raises(ValueError):
        O.admm_denoise(y, O.phi_sum(m), A, At, denoiser='bm3d', iter_max=1)


@pytest.mark.skipif(not reference_loader.available(), reason="reference tree not mounted")
def test_live_reference_matches_oracle():
    """Where /root/reference is mounted, run the reference itself on a fresh seed
    (not in the fixtures) and compare bit for bit."""
    ref_utils, ref_algo = reference_loader.load()
    rng = np.random.default_rng(99)
    H, W, C = 21, 19, 6
    mask = (rng.random((H, W, C)) <= 0.5).astype(np.float32)
    orig = rng.random((H, W, C), dtype=np.float32)
    y = np.sum(mask * orig, axis=2)
    A = lambda x: ref_utils.A_(x, mask)
    At = lambda v: ref_utils.At_(v, mask)
    Ao, Ato = _ops(mask)
    ms = O.phi_sum(mask)
    r = ref_algo.gap_denoise(y, ms, A, At, iter_max=7, tv_weight=0.2, tv_iter_max=5, X_orig=orig)
    o = O.gap_denoise(y, ms, Ao, Ato, iter_max=7, tv_weight=0.2, tv_iter_max=5, X_orig=orig)
    np.testing.assert_array_equal(r[0], o[0])
    assert r[3] == o[3]
    r = ref_algo.admm_denoise(y, ms, A, At, iter_max=7, tv_weight=0.2, tv_iter_max=5, X_orig=orig)
    o = O.admm_denoise(y, ms, Ao, Ato, iter_max=7, tv_weight=0.2, tv_iter_max=5, X_orig=orig)
    np.testing.assert_array_equal(r[0], o[0])
    assert r[3] == o[3]


def test_joint_admm_clip(golden):
    """joint_pnp_sci_algo.admm_denoise (theta clipped to [0,1]) against the reference's joint module."""
    g = golden("joint_admm")
    A, At = _ops(g["mask"])
    x, ps, ss, pa = O.joint_admm_denoise(g["y"], O.phi_sum(g["mask"]), A, At, _lambda=1, gamma=0.0,
                                         denoiser='tv', iter_max=12, tv_weight=0.3, tv_iter_max=5,
                                         X_orig=g["X_orig"])
    np.testing.assert_array_equal(x, g["x"])
    np.testing.assert_array_equal(np.array(pa), g["psnr_all"])
    np.testing.assert_array_equal(np.array(ps), g["psnr"])


def _second_standin(x, nsig, model=None):
    # the stand-in of tests/golden/make_golden.py for the learned denoiser
    a = np.float32(1.0 - 0.1 * float(nsig))
    return np.clip(x * a + np.float32(0.01), 0, 1).astype(np.float32)


def test_joint_multistep_and_two_period(golden):
    """TV + second-denoiser period and the two-period driver of the joint module, against the
    reference's own loops run with the same stand-in in FFDNet's place."""
    g = golden("joint_multistep")
    A, At = _ops(g["mask"])
    ms = O.phi_sum(g["mask"])
    x, ps, ss, pa = O.gap_multistep_denoise(g["y"], ms, A, At, _second_standin, iter_max=[3, 3],
                                            sigma=[0.2, 0.1], tv_weight=0.3, tv_iter_max=5,
                                            X_orig=g["X_orig"])
    np.testing.assert_array_equal(x, g["x"])
    np.testing.assert_array_equal(np.array(pa), g["psnr_all"])
    np.testing.assert_array_equal(np.array(ps), g["psnr"])
    g = golden("joint_two_period")
    x, ps, ss, pa = O.gap_joint_denoise(g["y"], ms, A, At, _second_standin, X_orig=g["X_orig"],
                                        iter_max1=4, iter_max2=[2, 2], sigma1=None, sigma2=[0.2, 0.1],
                                        _lambda=1, accelerate=True, tv_weight=0.3, tv_iter_max=5)
    assert len(pa) == 4                      # the PSNR track of the second period only
    np.testing.assert_array_equal(x, g["x"])
    np.testing.assert_array_equal(np.array(pa), g["psnr_all"])
    with pytest.raises(ValueError):
        O.gap_multistep_denoise(g["y"], ms, A, At, _second_standin, denoiser='tv', iter_max=1)


def test_joint_admm_multistep_and_two_period(golden):
    g = golden("joint_admm_multistep")
    A, At = _ops(g["mask"])
    ms = O.phi_sum(g["mask"])
    x, ps, ss, pa = O.admm_multistep_denoise(g["y"], ms, A, At, _second_standin, gamma=0.01,
                                             iter_max=[3, 3], sigma=[0.2, 0.1], tv_weight=0.3,
                                             tv_iter_max=5, X_orig=g["X_orig"])
    np.testing.assert_array_equal(x, g["x"])
    np.testing.assert_array_equal(np.array(pa), g["psnr_all"])
    g = golden("joint_admm_two_period")
    x, ps, ss, pa = O.admm_joint_denoise(g["y"], ms, A, At, _second_standin, X_orig=g["X_orig"],
                                         iter_max1=4, iter_max2=[2, 2], sigma1=None, sigma2=[0.2, 0.1],
                                         _lambda=1, gamma=0.01, tv_weight=0.3, tv_iter_max=5)
    assert len(pa) == 4
    np.testing.assert_array_equal(x, g["x"])
    np.testing.assert_array_equal(np.array(pa), g["psnr_all"])


def test_tv_rec_loops(golden):
    """GAP_TV_rec / ADMM_TV_rec (pnp_sci_algo.py:866-907; float64 loops, 30 dual iterations per step) against the
    reference's own functions (tests/golden/make_golden_rec.py)."""
    g = golden("tv_rec")
    H, W, Cc = g["mask"].shape
    out = O.GAP_TV_rec(g["y"], g["mask"], O.A_, O.At_, g["Phi_sum"], int(g["maxiter"]), float(g["step_size"]),
                       float(g["weight"]), H, W, Cc, g["X_orig"])
    assert out.dtype == np.float64
    np.testing.assert_array_equal(out, g["gap"])
    out = O.ADMM_TV_rec(g["y"], g["mask"], O.A_, O.At_, g["Phi_sum"], int(g["maxiter"]), float(g["step_size"]),
                        float(g["weight"]), H, W, Cc, float(g["eta"]), g["X_orig"])
    np.testing.assert_array_equal(out, g["admm"])


def test_admm_denoise_bayer_is_four_pinned_admm_solves():
    """admm_denoise_bayer (pnp_sci_algo.py:268-475) is dead code in the reference (NameError at :399), so its
    restatement cannot be pinned directly; it must equal the pinned admm_denoise (R5) run on each sub-lattice."""
    rng = np.random.default_rng(5)
    Phi = (rng.random((24, 28, 4)) <= 0.5).astype(np.float32)
    orig = rng.random((24, 28, 4), dtype=np.float32)
    y = np.sum(Phi * orig, axis=2)
    x, pa = O.admm_denoise_bayer(y, Phi, _lambda=1, gamma=0.02, denoiser='tv', iter_max=5, tv_weight=0.2,
                                 tv_iter_max=4, X_orig=orig)
    assert len(pa) == 5
    for (i, j) in ((0, 0), (0, 1), (1, 0), (1, 1)):
        P = np.ascontiguousarray(Phi[i::2, j::2])
        A, At = _ops(P)
        xs = O.admm_denoise(np.ascontiguousarray(y[i::2, j::2]), O.phi_sum(P), A, At, _lambda=1, gamma=0.02,
                            iter_max=5, tv_weight=0.2, tv_iter_max=4)[0]
        np.testing.assert_array_equal(x[i::2, j::2], xs)
