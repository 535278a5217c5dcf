"""GPU tests of the warp-specialised fused GAP-TV kernel (csrc/gap_tv_ws.cuh) through the C ABI:
against the exact path (statement-order replica of the reference), against the stream kernel and
against the CPU oracle, on ragged shapes (widths that are not a multiple of the owned pixels of a
group, short images, batches, per-measurement masks, every supported tv_iter_max and channel count).

Tolerance (BASELINE.json north_star): max abs error <= 1e-4 on [0,1] frames.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_X = 1e-4


@pytest.fixture(scope="module")
def sp():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import scipnp
    return scipnp


@pytest.fixture()
def variant():
    from scipnp._lib import lib, check

    def _set(v):
        check(lib.scipnp_set_fused_variant(v))
    yield _set
    _set(0)


def _scene(H, W, C, B=1, seed=0, phi_batched=False):
    rng = np.random.default_rng(seed)
    shape = (B, H, W, C) if phi_batched else (H, W, C)
    Phi = (rng.random(shape) <= 0.5).astype(np.float32)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    orig = np.stack([[0.5 + 0.3 * np.sin((xx + 3 * c + 7 * b) / 9.0) * np.cos((yy - 2 * c) / 7.0)
                      + 0.1 * ((xx // 11 + yy // 13 + c) % 2) for c in range(C)] for b in range(B)])
    orig = np.moveaxis(orig, 1, -1).astype(np.float32)                  # [B,H,W,C]
    y = (Phi * orig).sum(-1).astype(np.float32)
    return y, Phi, orig


def _run(Solver, y, Phi, iters, fused, accelerate=True, tv_iter_max=5, tv_weight=0.3, phi_batched=False, lam=1.0, tv_eps=0.0):
    """tv_eps = 0 disables skimage's early stop on both paths, so the fused kernels themselves are compared (with
    the rule active a fired flag sends the run to the exact path and the comparison would be vacuous)."""
    B, H, W = y.shape
    C = Phi.shape[-1]
    with Solver(B, H, W, C, method="gap", accelerate=accelerate, _lambda=lam, tv_weight=tv_weight,
                tv_iter_max=tv_iter_max, phi_batched=phi_batched, fused=fused, tv_eps=tv_eps) as s:
        s.load(y, Phi)
        s.run(iters)
        assert s.uses_fused == fused
        x = s.get_x()
        # the accelerated variant's y1 is part of the state: compare it too
        import torch
        xp, y1p = s.state_ptrs()
        from scipnp.tiled import _wrap
        y1 = _wrap(y1p, (B, H, W), torch.device("cuda", torch.cuda.current_device())).cpu().numpy() if accelerate else None
        return x, y1, s.refined_iters


SHAPES = [
    # (H, W, C, B)                        what it exercises
    (40, 64, 8, 1),                     # one group wide, three groups per CTA mostly dead
    (37, 100, 8, 1),                    # ragged width, odd height
    (64, 256, 8, 4),                    # config-1-like, batch of coded frames
    (23, 60, 24, 1),                    # narrower than one group tile
    (50, 232, 24, 1),                   # several strips, last one partial
    (9, 128, 24, 2),                    # fewer rows than the pipeline is deep
    (130, 176, 24, 1),                  # more rows than one wave of segments needs
    (33, 120, 4, 1), (33, 120, 12, 1), (33, 120, 16, 1), (33, 120, 20, 1),
]


@pytest.mark.parametrize("H,W,C,B", SHAPES)
@pytest.mark.parametrize("accelerate", [True, False], ids=["acc", "plain"])
def test_ws_kernel_matches_exact_path(sp, variant, H, W, C, B, accelerate):
    from scipnp.engine import Solver
    y, Phi, _ = _scene(H, W, C, B, seed=H + W + C)
    xe, y1e, _ = _run(Solver, y, Phi, 3, fused=False, accelerate=accelerate)
    variant(0)
    xw, y1w, refined = _run(Solver, y, Phi, 3, fused=True, accelerate=accelerate)
    assert refined == 0
    assert float(np.abs(xw - xe).max()) <= TOL_X
    if accelerate:
        assert float(np.abs(y1w - y1e).max()) <= TOL_X
    variant(1)
    xs, y1s, _ = _run(Solver, y, Phi, 3, fused=True, accelerate=accelerate)
    assert float(np.abs(xw - xs).max()) <= 2e-5           # the two fused kernels share the approximations


@pytest.mark.parametrize("T", [3, 4, 5])
def test_ws_kernel_tv_iter_max(sp, variant, T):
    from scipnp.engine import Solver
    y, Phi, _ = _scene(45, 184, 24, 1, seed=T)
    xe, y1e, _ = _run(Solver, y, Phi, 4, fused=False, tv_iter_max=T, tv_weight=0.1, lam=0.8)
    variant(0)
    xw, y1w, refined = _run(Solver, y, Phi, 4, fused=True, tv_iter_max=T, tv_weight=0.1, lam=0.8)
    assert refined == 0
    assert float(np.abs(xw - xe).max()) <= TOL_X
    assert float(np.abs(y1w - y1e).max()) <= TOL_X


def test_ws_kernel_per_measurement_masks(sp, variant):
    """phi_batched = 1: every batch element has its own mask (the four Bayer sub-lattices)."""
    from scipnp.engine import Solver
    y, Phi, _ = _scene(48, 136, 24, 4, seed=9, phi_batched=True)
    xe, _, _ = _run(Solver, y, Phi, 3, fused=False, phi_batched=True)
    variant(0)
    xw, _, _ = _run(Solver, y, Phi, 3, fused=True, phi_batched=True)
    assert float(np.abs(xw - xe).max()) <= TOL_X


def test_ws_kernel_against_oracle_40_iterations(sp, variant):
    """config-1 size, the reference's parameter set, 40 outer iterations, against the CPU oracle."""
    from oracle import pnp_sci as O
    from scipnp import synth
    from scipnp.engine import Solver
    meas, mask, orig = synth.make_cacti(256, 256, 8, 1, cfg=1)
    y = meas[:, :, 0] / np.float32(255.)
    o = orig[:, :, :8] / np.float32(255.)
    A = lambda x: O.A_(x, mask)
    At = lambda v: O.At_(v, mask)
    xo, _, _, pao = O.gap_denoise(y, O.phi_sum(mask), A, At, _lambda=1, accelerate=True, denoiser='tv', iter_max=40,
                                  tv_weight=0.3, tv_iter_max=5, X_orig=o, show_iqa=True)
    variant(0)
    with Solver(1, 256, 256, 8, method="gap", tv_weight=0.3, tv_iter_max=5) as s:
        s.load(y[None], mask, X_orig=o[None])
        s.run(40)
        xg = s.get_x()[0]
        pag = s.psnr_all()[:, 0]
        assert s.refined_iters == 0
    assert float(np.abs(xg - xo).max()) <= TOL_X
    assert float(np.abs(pag - np.array(pao)).max()) <= 0.01


def test_ws_kernel_early_stop_flag(sp, variant):
    """A huge eps makes skimage's stopping rule fire: the in-kernel replay of the rule (last CTA) must raise
    the flag, and the solver then redoes the run on the exact path -- results never depend on the path."""
    from scipnp.engine import Solver
    y, Phi, _ = _scene(40, 128, 8, 1, seed=3)
    variant(0)
    with Solver(1, 40, 128, 8, method="gap", tv_weight=0.3, tv_iter_max=5, tv_eps=0.5) as s:
        s.load(y, Phi)
        s.run(2)
        assert s.refined_iters == 2
        xg = s.get_x()
        s.run(2)                                   # the accumulators and the ticket were left clean
        assert s.refined_iters == 4
    with Solver(1, 40, 128, 8, method="gap", tv_weight=0.3, tv_iter_max=5, tv_eps=0.5, fused=False) as s:
        s.load(y, Phi)
        s.run(2)
        xe = s.get_x()
    np.testing.assert_array_equal(xg, xe)


# -- the denoiser alone (north_star subsystem 2): MODE_TV of the same kernel ------------------------

@pytest.mark.parametrize("H,W,C,T", [(64, 128, 8, 5), (37, 60, 24, 5), (50, 52, 4, 3), (9, 300, 12, 4), (130, 64, 16, 5)])
def test_fused_tv_alone_matches_oracle(sp, H, W, C, T):
    """scipnp_tv_chambolle_fused (all dual updates in one launch) against the oracle's skimage restatement,
    through the C ABI; eps = 0 keeps the stopping rule out of both."""
    import torch
    from scipnp._lib import lib, check
    from scipnp.engine import dptr, stream_ptr
    from oracle.tv_chambolle import denoise_tv_chambolle as tv_oracle
    assert lib.scipnp_tv_fused_supported(1, H, W, C, T) == 1
    rng = np.random.default_rng(H * W + C)
    img = (0.5 + 0.25 * rng.standard_normal((H, W, C))).astype(np.float32)
    ref = tv_oracle(img, 0.3, eps=0.0, n_iter_max=T, multichannel=True)
    d = torch.from_numpy(img).cuda()
    out = torch.empty_like(d)
    wsb = lib.scipnp_tv_fused_workspace_bytes(1, H, W, C, T)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    check(lib.scipnp_tv_chambolle_fused(dptr(d), dptr(out), 0.3, 0.0, T, 1, H, W, C, dptr(ws), wsb, dptr(flag), stream_ptr()))
    assert int(flag.item()) == 0
    assert float(np.abs(out.cpu().numpy() - ref).max()) <= 2e-6


def test_fused_tv_alone_batched_through_the_abi(sp):
    """B > 1: every (b, c) slice is its own problem; rows of neighbouring batch elements never mix."""
    import torch
    from scipnp._lib import lib, check
    from scipnp.engine import dptr, stream_ptr
    from oracle.tv_chambolle import denoise_tv_chambolle as tv_oracle
    B, H, W, C, T = 3, 21, 64, 8, 5
    rng = np.random.default_rng(5)
    img = rng.random((B, H, W, C)).astype(np.float32)
    d = torch.from_numpy(img).cuda()
    out = torch.empty_like(d)
    wsb = lib.scipnp_tv_fused_workspace_bytes(B, H, W, C, T)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    check(lib.scipnp_tv_chambolle_fused(dptr(d), dptr(out), 0.2, 0.0, T, B, H, W, C, dptr(ws), wsb, None, stream_ptr()))
    got = out.cpu().numpy()
    for b in range(B):
        assert float(np.abs(got[b] - tv_oracle(img[b], 0.2, eps=0.0, n_iter_max=T, multichannel=True)).max()) <= 2e-6


def test_denoise_tv_chambolle_uses_the_one_pass_kernel_and_keeps_the_early_stop(sp):
    """The public R6 entry: one launch when the rule does not fire, the exact kernels (and skimage's
    result) when it does."""
    from scipnp._lib import lib
    from oracle.tv_chambolle import denoise_tv_chambolle as tv_oracle
    rng = np.random.default_rng(11)
    img = rng.random((48, 64, 8)).astype(np.float32)
    l0 = lib.scipnp_launch_count()
    got = sp.denoise_tv_chambolle(img, 0.3, n_iter_max=5, multichannel=True)
    assert lib.scipnp_launch_count() - l0 == 1
    assert float(np.abs(got - tv_oracle(img, 0.3, n_iter_max=5, multichannel=True)).max()) <= 2e-6
    # a nearly flat image with a large eps: the rule fires at the first check
    flat = (0.5 + 1e-3 * rng.random((48, 64, 8))).astype(np.float32)
    l0 = lib.scipnp_launch_count()
    got = sp.denoise_tv_chambolle(flat, 0.3, eps=0.5, n_iter_max=5, multichannel=True)
    assert lib.scipnp_launch_count() - l0 > 1
    np.testing.assert_array_equal(got, tv_oracle(flat, 0.3, eps=0.5, n_iter_max=5, multichannel=True))
    assert lib.scipnp_tv_fused_supported(1, 48, 64, 8, 30) == 0 and lib.scipnp_tv_fused_supported(1, 48, 63, 8, 5) == 0
