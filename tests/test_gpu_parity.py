"""GPU parity tests: the CUDA path (through the C ABI) against the committed golden
fixtures of the reference and against the CPU oracle on fresh seeded inputs.

Tolerance (BASELINE.json north_star): max abs error <= 1e-4 on [0,1] frames and
PSNR within 0.01 dB at equal iteration count.  The exact path reproduces the
reference's float32 arithmetic statement by statement and is held to 2e-6.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_X = 1e-4          # north_star: max abs on [0,1] frames
TOL_DB = 0.01         # north_star: PSNR delta
TOL_EXACT = 2e-6      # exact path (IEEE statement-order replica)


@pytest.fixture(scope="module")
def sp():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import scipnp
    from scipnp import _lib
    assert _lib.lib.scipnp_device_count() >= 1
    return scipnp


@pytest.fixture(params=[True, False], ids=["fused", "exact"])
def path(request, sp):
    from scipnp import pnp_sci_algo as M
    old = M.USE_FUSED
    M.USE_FUSED = request.param
    yield request.param
    M.USE_FUSED = old


def _tol(fused):
    return TOL_X if fused else TOL_EXACT


def _ops(mask):
    from oracle import pnp_sci as O
    return (lambda x: O.A_(x, mask)), (lambda y: O.At_(y, mask))


def _psum(mask):
    s = mask.sum(2)
    s[s == 0] = 1
    return s


def _cmp(x, gx, pa, gpa, fused):
    err = float(np.abs(x - gx).max())
    assert err <= _tol(fused), "max abs %.3g" % err
    if gpa is not None and len(gpa):
        assert len(pa) == len(gpa)
        assert np.abs(np.array(pa) - np.array(gpa)).max() <= TOL_DB


# -- operators -----------------------------------------------------------------

@pytest.mark.parametrize("tag", ["ops_8", "ops_5"])
def test_operators_bit_exact(sp, golden, tag):
    g = golden(tag)
    np.testing.assert_array_equal(sp.A_(g["x"], g["Phi"]), g["A"])
    np.testing.assert_array_equal(sp.At_(g["y"], g["Phi"]), g["At"])
    np.testing.assert_array_equal(sp.phi_sum(g["Phi"]), g["Phi_sum"])
    assert abs(sp.psnr(g["x"], g["x2"]) - float(g["psnr"])) < 1e-4
    assert sp.psnr(g["x"], g["x"]) == 100


def test_operator_properties_large(sp):
    """Size-independent properties at the UHD size of config 5."""
    import torch
    H, W, Cc = 2160, 3840, 24
    gen = torch.Generator(device="cuda").manual_seed(5)
    Phi = (torch.rand((H, W, Cc), device="cuda", generator=gen) <= 0.5).float()
    x = torch.rand((H, W, Cc), device="cuda", generator=gen)
    y = torch.rand((H, W), device="cuda", generator=gen)
    Ax = sp.A_(x, Phi)
    Aty = sp.At_(y, Phi)
    # adjoint identity <A x, y> = <x, At y>
    lhs = float((Ax.double() * y.double()).sum())
    rhs = float((x.double() * Aty.double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * abs(lhs)
    # A(At(y)/Phi_sum) = y wherever some mask is open (binary mask: A At = diag(Phi_sum))
    ps = sp.phi_sum(Phi)
    back = sp.A_(sp.At_(y / ps, Phi), Phi)
    open_ = Phi.sum(2) > 0
    assert float((back - y)[open_].abs().max()) <= 2e-6
    assert float(ps.min()) >= 1.0


# -- TV --------------------------------------------------------------------------

@pytest.mark.parametrize("T", [1, 2, 5, 200])
def test_tv_regression_vectors(sp, golden, T):
    g = golden("tv_T%d" % T)
    out, n_exec, energy = sp.denoise_tv_chambolle(g["image"], float(g["weight"]),
                                                  n_iter_max=int(g["n_iter_max"]),
                                                  multichannel=True, return_stats=True)
    np.testing.assert_array_equal(n_exec, g["n_exec"])      # same early-stop decisions
    assert np.abs(out - g["out"]).max() <= TOL_EXACT
    ne = int(g["n_exec"].max())
    np.testing.assert_allclose(energy[:, :ne], g["energy"][:, :ne], rtol=1e-5)


def test_tv_single_channel_2d(sp):
    from oracle.tv_chambolle import denoise_tv_chambolle as otv
    rng = np.random.default_rng(11)
    f = rng.random((33, 47)).astype(np.float32)
    out = sp.denoise_tv_chambolle(f, 0.2, n_iter_max=7)
    assert np.abs(out - otv(f, 0.2, n_iter_max=7)).max() <= TOL_EXACT


def test_tv_properties_large(sp):
    import torch
    H, W, Cc = 1080, 1920, 24
    gen = torch.Generator(device="cuda").manual_seed(6)
    f = torch.rand((H, W, Cc), device="cuda", generator=gen)
    out = sp.denoise_tv_chambolle(f, 0.3, n_iter_max=5, multichannel=True)
    # f + div p keeps the mean of every channel
    assert float((out.double().mean((0, 1)) - f.double().mean((0, 1))).abs().max()) < 1e-6
    # constant image is a fixed point
    c = torch.full((64, 64, 8), 0.25, device="cuda")
    assert torch.equal(sp.denoise_tv_chambolle(c, 0.3, n_iter_max=5, multichannel=True), c)
    # channels are independent problems
    # (one channel runs the exact path, 24 channels the one-pass kernel: FMA contraction and MUFU rounding differ)
    one = sp.denoise_tv_chambolle(f[:, :, 3].contiguous(), 0.3, n_iter_max=5)
    assert float((one - out[:, :, 3]).abs().max()) <= 1e-5
    # total variation does not increase
    def tv(u):
        return float((u[1:] - u[:-1]).abs().sum() + (u[:, 1:] - u[:, :-1]).abs().sum())
    assert tv(out) < tv(f)


# -- solver loops against the reference's golden outputs ---------------------------

def test_gap_accelerated_golden(sp, golden, path):
    g = golden("gap_acc")
    A, At = _ops(g["mask"])
    x, ps, ss, pa = sp.gap_denoise(g["y"], _psum(g["mask"]), A, At, _lambda=1, accelerate=True,
                                   denoiser='tv', iter_max=12, tv_weight=0.3, tv_iter_max=5,
                                   X_orig=g["X_orig"])
    _cmp(x, g["x"], pa, g["psnr_all"], path)
    assert x.dtype == np.float32
    assert np.abs(np.array(ps) - g["psnr"]).max() <= TOL_DB
    assert np.abs(np.array(ss) - g["ssim"]).max() <= 1e-4


def test_gap_plain_schedule_golden(sp, golden, path):
    g = golden("gap_plain")
    x, _, _, pa = sp.gap_denoise(g["y"], _psum(g["mask"]), Phi=g["mask"], _lambda=0.75,
                                 accelerate=False, denoiser='tv', iter_max=[3, 4],
                                 sigma=[0.2, 0.1], tv_weight=0.1, tv_iter_max=3,
                                 X_orig=g["X_orig"])
    _cmp(x, g["x"], pa, g["psnr_all"], path)


def test_admm_golden(sp, golden, path):
    g = golden("admm")
    A, At = _ops(g["mask"])
    x, ps, ss, pa = sp.admm_denoise(g["y"], _psum(g["mask"]), A, At, _lambda=1, gamma=0.01,
                                    denoiser='tv', iter_max=12, tv_weight=0.3, tv_iter_max=5,
                                    X_orig=g["X_orig"])
    _cmp(x, g["x"], pa, g["psnr_all"], path)
    assert np.abs(np.array(ps) - g["psnr"]).max() <= TOL_DB


def test_gap_warm_start_ragged_golden(sp, golden, path):
    g = golden("gap_c5_warm")        # C = 5: the scalar (non-vectorised) kernels
    ms = _psum(g["mask"])
    ms[ms == 0] = 1
    x, _, _, pa = sp.gap_denoise(g["y"], ms, Phi=g["mask"], iter_max=6, tv_weight=0.2,
                                 tv_iter_max=4, x0=g["x0"], X_orig=g["X_orig"])
    _cmp(x, g["x"], pa, g["psnr_all"], path)


@pytest.mark.parametrize("pm", ["gap", "admm"])
@pytest.mark.parametrize("md", ["plain", "updown"])
def test_cacti_wrapper_golden(sp, golden, path, pm, md):
    g = golden("cacti_%s_%s" % (pm, md))
    A, At = _ops(g["mask"])
    kw = dict(_lambda=1, denoiser='tv', iter_max=5, tv_weight=0.3, tv_iter_max=5)
    kw.update({"accelerate": True} if pm == "gap" else {"gamma": 0.01})
    x_, t_, ps, ss, pa = sp.admmdenoise_cacti(g["meas"], g["mask"], A, At, projmeth=pm,
                                              orig=g["orig"], nframe=2, MAXB=255.,
                                              maskdirection=md, **kw)
    assert x_.shape == g["x"].shape and x_.dtype == np.float32
    assert np.abs(x_ - g["x"]).max() <= _tol(path)
    assert np.abs(np.array(pa) - g["psnr_all"]).max() <= TOL_DB
    assert np.abs(np.array(ps) - g["psnr"]).max() <= TOL_DB
    assert t_ > 0


def test_cacti_wrapper_without_orig(sp, golden):
    g = golden("cacti_gap_plain")
    x_, t_, ps, ss, pa = sp.admmdenoise_cacti(g["meas"], g["mask"], None, None, projmeth='gap',
                                              nframe=2, MAXB=255., denoiser='tv', iter_max=5,
                                              tv_weight=0.3, tv_iter_max=5)
    assert np.abs(x_ - g["x"]).max() <= TOL_X
    assert ps == [] and ss == [] and pa == [[], []]


def test_bayer_golden(sp, golden, path):
    g = golden("bayer")
    x, ps, ss, pa = sp.gap_denoise_bayer(g["y_bayer"], g["Phi_bayer"], _lambda=1, accelerate=True,
                                         denoiser='tv', iter_max=8, tv_weight=0.1, tv_iter_max=5,
                                         X_orig=g["X_orig"])
    _cmp(x, g["x"], pa, g["psnr_all"], path)
    assert np.abs(np.array(ps) - g["psnr"]).max() <= TOL_DB


def test_cassi_golden(sp, golden, path):
    g = golden("cassi")
    x, _, _, pa = sp.gap_denoise_cassi(g["y"], g["mask2d"], int(g["nband"]), int(g["step"]),
                                       iter_max=8, tv_weight=0.1, tv_iter_max=5,
                                       X_orig=g["X_orig"])
    _cmp(x, g["x"], pa, g["psnr_all"], path)


# -- fresh inputs against the oracle at the reference's demo configuration -----------

def test_config1_gap_tv_40_iterations(sp, path):
    """BASELINE config 1 (256x256x8, 40 it, tv_weight 0.3, tv_iter_max 5)."""
    from oracle import pnp_sci as O
    from scipnp import synth
    meas, mask, orig = synth.make_cacti(256, 256, 8, 1, cfg=1)
    A, At = _ops(mask)
    kw = dict(projmeth='gap', orig=orig, nframe=1, MAXB=255., _lambda=1, accelerate=True,
              denoiser='tv', iter_max=40, tv_weight=0.3, tv_iter_max=5)
    xo, _, pso, sso, pao = O.admmdenoise_cacti(meas, mask, A, At, **kw)
    xg, _, psg, ssg, pag = sp.admmdenoise_cacti(meas, mask, A, At, **kw)
    assert np.abs(xg - xo).max() <= _tol(path)
    assert np.abs(np.array(pag) - np.array(pao)).max() <= TOL_DB
    assert np.abs(np.array(psg) - np.array(pso)).max() <= TOL_DB
    assert np.abs(np.array(ssg) - np.array(sso)).max() <= 1e-4
    assert pag[0][-1] > 24.0          # the reconstruction actually converges


def test_config2_admm_batch(sp, path):
    """BASELINE config 2 shape: several 256x256x8 measurements solved as one batch."""
    from oracle import pnp_sci as O
    from scipnp import synth
    meas, mask, orig = synth.make_cacti(256, 256, 8, 3, cfg=2)
    A, At = _ops(mask)
    kw = dict(projmeth='admm', orig=orig, nframe=3, MAXB=255., _lambda=1, gamma=0.01,
              denoiser='tv', iter_max=10, tv_weight=0.3, tv_iter_max=5)
    xo, _, pso, _, pao = O.admmdenoise_cacti(meas, mask, A, At, **kw)
    xg, _, psg, _, pag = sp.admmdenoise_cacti(meas, mask, A, At, **kw)
    assert np.abs(xg - xo).max() <= _tol(path)
    assert np.abs(np.array(pag) - np.array(pao)).max() <= TOL_DB


def test_uhd_crop_matches_oracle(sp, path):
    """A 96-row, full-width (3840) crop of the config-5 scene, 3 iterations."""
    from oracle import pnp_sci as O
    from scipnp import synth
    meas, mask, orig = synth.make_cacti(96, 3840, 24, 1, cfg=5)
    A, At = _ops(mask)
    y = meas[:, :, 0] / np.float32(255.)
    ms = O.phi_sum(mask)
    xo, _, _, pao = O.gap_denoise(y, ms, A, At, iter_max=3, tv_weight=0.3, tv_iter_max=5,
                                  X_orig=orig / np.float32(255.))
    xg, _, _, pag = sp.gap_denoise(y, ms, Phi=mask, iter_max=3, tv_weight=0.3, tv_iter_max=5,
                                   X_orig=orig / np.float32(255.))
    assert np.abs(xg - xo).max() <= _tol(path)
    assert np.abs(np.array(pag) - np.array(pao)).max() <= TOL_DB


def test_uhd_full_size_properties(sp):
    """Config 5 at full size (3840x2160x24): properties the domain offers, since
    the oracle needs ~40 s per iteration there."""
    import torch
    from scipnp import Solver
    H, W, Cc = 2160, 3840, 24
    gen = torch.Generator(device="cuda").manual_seed(7)
    Phi = (torch.rand((H, W, Cc), device="cuda", generator=gen) <= 0.5).float()
    yy = torch.arange(H, device="cuda").float()[:, None, None]
    xx = torch.arange(W, device="cuda").float()[None, :, None]
    tt = torch.arange(Cc, device="cuda").float()[None, None, :]
    orig = 0.5 + 0.3 * torch.sin((xx + 3 * tt) / 97.) * torch.cos(yy / 131.)
    y = (Phi * orig).sum(2)
    with Solver(1, H, W, Cc, method="gap", tv_weight=0.3, tv_iter_max=5) as s:
        s.load(y[None], Phi, X_orig=orig[None])
        s.run(6)
        pa = s.psnr_all()[:, 0]
        x_full = torch.empty((1, H, W, Cc), device="cuda")
        s.get_x(x_full)
        fused = s.uses_fused
    assert np.all(np.diff(pa) > 0) and pa[-1] > 15.0     # PSNR climbs monotonically
    # a horizontal band solved alone agrees with the full solve away from the cut:
    # information travels <= (tv_iter_max-1) rows per outer iteration
    r0, r1, it = 1000, 1128, 6
    with Solver(1, r1 - r0, W, Cc, method="gap", tv_weight=0.3, tv_iter_max=5) as s:
        s.load(y[None, r0:r1].contiguous(), Phi[r0:r1].contiguous())
        s.run(it)
        x_band = torch.empty((1, r1 - r0, W, Cc), device="cuda")
        s.get_x(x_band)
    m = 4 * it
    d = (x_band[0, m:-m] - x_full[0, r0 + m:r1 - m]).abs().max()
    assert float(d) <= (1e-5 if fused else 0.0)


# -- the C ABI called directly ----------------------------------------------------------

def test_c_abi_host_entry_matches_python_surface(sp, golden):
    from scipnp._lib import lib, Params, check
    g = golden("gap_acc")
    H, W, Cc = g["mask"].shape
    p = Params()
    p.method, p.accelerate, p.lambda_, p.gamma = 0, 1, 1.0, 0.0
    p.tv_weight, p.tv_eps, p.tv_iter_max, p.fused = 0.3, 2e-4, 5, 0
    p.B, p.H, p.W, p.C, p.phi_batched, p.clip01 = 1, H, W, Cc, 0, 0
    y = np.ascontiguousarray(g["y"], np.float32)
    Phi = np.ascontiguousarray(g["mask"], np.float32)
    Xo = np.ascontiguousarray(g["X_orig"], np.float32)
    x = np.empty((H, W, Cc), np.float32)
    pa = (C.c_double * 12)()
    n = C.c_int(0)
    check(lib.scipnp_gap_denoise_host(y.ctypes.data, Phi.ctypes.data, None, Xo.ctypes.data,
                                      C.byref(p), 12, x.ctypes.data, pa, C.byref(n)))
    assert n.value == 12
    assert np.abs(x - g["x"]).max() <= TOL_EXACT
    assert np.abs(np.array(pa[:]) - g["psnr_all"]).max() <= TOL_DB
    # ADMM through the same door
    g = golden("admm")
    p.gamma = 0.01
    check(lib.scipnp_admm_denoise_host(y.ctypes.data, Phi.ctypes.data, None, Xo.ctypes.data,
                                       C.byref(p), 12, x.ctypes.data, pa, C.byref(n)))
    assert np.abs(x - g["x"]).max() <= TOL_EXACT
    assert lib.scipnp_launch_count() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_host_pipeline_matches_one_call_entries(sp, golden, fused):
    """scipnp_pipeline_*: three different measurements through two slots give what the
    synchronous solver gives, in submission order, PSNR track included; a full pipeline refuses
    a further submit."""
    import torch
    from scipnp import HostPipeline, Solver, ScipnpError
    g = golden("gap_acc")
    H, W, Cc = g["mask"].shape
    Phi = np.ascontiguousarray(g["mask"], np.float32)
    Xo = np.ascontiguousarray(g["X_orig"], np.float32)
    ys = [np.ascontiguousarray(g["y"] * np.float32(s), np.float32) for s in (1.0, 0.5, 0.25)]
    kw = dict(method="gap", tv_weight=0.3, tv_iter_max=5, fused=fused)
    want = []
    with Solver(1, H, W, Cc, **kw) as so:
        for y in ys:
            so.load(y[None], Phi, X_orig=Xo[None])
            so.run(12)
            want.append((so.get_x()[0].copy(), np.array(so.psnr_all())))
    outs = [torch.empty((1, H, W, Cc), dtype=torch.float32).pin_memory() for _ in ys]
    with HostPipeline(1, H, W, Cc, depth=2, **kw) as pl:
        t0 = pl.submit(torch.from_numpy(ys[0][None]).pin_memory(), Phi, 12, outs[0], X_orig=Xo[None])
        t1 = pl.submit(ys[1][None], Phi, 12, outs[1], X_orig=Xo[None])
        with pytest.raises(ScipnpError):
            pl.submit(ys[2][None], Phi, 12, outs[2])
        x0, p0 = pl.wait(t0)
        t2 = pl.submit(ys[2][None], Phi, 12, outs[2], X_orig=Xo[None])
        res = [(x0, p0), pl.wait(t1), pl.wait(t2)]
        with pytest.raises(ScipnpError):
            pl.wait(t2)                        # already collected
    for (x, ps), (xw, pw) in zip(res, want):
        np.testing.assert_array_equal(x.numpy()[0], xw)
        # squared errors are summed with atomics: equal up to the summation order
        np.testing.assert_allclose(ps.reshape(-1), pw.reshape(-1), rtol=0, atol=1e-9)
    if not fused:                              # the exact path is the reference bit for bit
        assert np.abs(res[0][0].numpy()[0] - g["x"]).max() <= TOL_EXACT


def test_c_abi_rejects_bad_arguments(sp):
    from scipnp._lib import lib
    assert lib.scipnp_A(None, None, None, 1, 4, 4, 4, 0, None) == -1
    assert b"null" in lib.scipnp_last_error()
    assert lib.scipnp_tv_chambolle(None, None, 0.1, 2e-4, 5, 1, 0, 4, 4, None, 0, None, None, 0, None) == -1


def test_early_stop_rolls_back_to_exact_path(sp):
    """Huge eps forces the energy criterion to fire: the solver must give the
    reference's (early-stopped) answer whichever path it started on."""
    from oracle import pnp_sci as O
    from oracle.tv_chambolle import denoise_tv_chambolle as otv
    from scipnp import synth, Solver
    meas, mask, orig = synth.make_cacti(64, 64, 8, 1, cfg=9)
    y = meas[:, :, 0] / np.float32(255.)
    ms = O.phi_sum(mask)
    eps = 0.5
    # oracle loop with the big eps
    x = O.At_(y, mask)
    y1 = np.zeros_like(y)
    for _ in range(4):
        yb = O.A_(x, mask)
        y1 = y1 + (y - yb)
        x = x + 1 * O.At_((y1 - yb) / ms, mask)
        x = otv(x, 0.3, eps=eps, n_iter_max=5, multichannel=True)
    with Solver(1, 64, 64, 8, method="gap", tv_weight=0.3, tv_iter_max=5, tv_eps=eps) as s:
        s.load(y[None], mask)
        s.run(4)
        xg = s.get_x()[0]
        if s.uses_fused:
            assert s.refined_iters == 4
    assert np.abs(xg - x).max() <= TOL_EXACT


# -- row-tiled solve (two ranks sharing cuda:0, gloo rendezvous) ---------------------------------------

def _tiled_worker(rank, world, port, H, W, C, iters, k, out_dir):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        from scipnp.tiled import TiledSolver
        from scipnp import synth
        meas, mask, _ = synth.make_cacti(H, W, C, 1, cfg=21)
        y = meas[:, :, 0] / np.float32(255.)
        ts = TiledSolver(H, W, C, rank, world, tv_weight=0.3, tv_iter_max=5, exchange_every=k)
        ts.load(torch.from_numpy(y[ts.row_lo:ts.row_hi]).cuda(), torch.from_numpy(mask[ts.row_lo:ts.row_hi]).cuda())
        ts.run(iters)
        np.save(os.path.join(out_dir, "x_%d.npy" % rank), ts.owned().cpu().numpy())
        ts.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k", [1, 2])
def test_row_tiled_equals_single_gpu(sp, tmp_path, k):
    import socket
    import torch.multiprocessing as mp
    from scipnp import synth, Solver
    H, W, C, iters, world = 96, 80, 8, 6, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_tiled_worker, args=(world, port, H, W, C, iters, k, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / ("x_%d.npy" % r)) for r in range(world)], axis=0)
    meas, mask, _ = synth.make_cacti(H, W, C, 1, cfg=21)
    y = meas[:, :, 0] / np.float32(255.)
    with Solver(1, H, W, C, method="gap", tv_weight=0.3, tv_iter_max=5) as so:
        so.load(y[None], mask)
        so.run(iters)
        ref = so.get_x()[0]
        fused = so.uses_fused
    # owned rows never see a seam: the tiled result is the single-GPU result
    assert np.abs(got - ref).max() <= (1e-6 if fused else 0.0)


def test_cassi_index_offset_config4(sp, path):
    """BASELINE config 4: 256x256x28 bands, dispersion 2 px/band (canvas 256x310).  The fused path reads
    the coded aperture at per-band offsets; the oracle uses the explicit shifted stack."""
    from oracle import pnp_sci as O
    from scipnp import synth
    nband, step = 28, 2
    y, m2, cube = synth.make_cassi(256, 256, nband, step=step, cfg=4)
    Phi = O.cassi_shift_mask(m2, nband, step)
    A, At = _ops(Phi)
    xo, _, _, pao = O.gap_denoise(y, O.phi_sum(Phi), A, At, iter_max=6, tv_weight=0.1, tv_iter_max=5,
                                  X_orig=cube)
    xg, psg, ssg, pag = sp.gap_denoise_cassi(y, m2, nband, step, iter_max=6, tv_weight=0.1, tv_iter_max=5,
                                             X_orig=cube)
    assert xg.shape == (256, 310, nband)
    assert np.abs(xg - xo).max() <= _tol(path)
    assert np.abs(np.array(pag) - np.array(pao)).max() <= TOL_DB
    assert len(psg) == nband


def test_fused_path_coverage(sp):
    """Which configurations run the one-pass kernel (the rest use the exact kernels)."""
    from scipnp import Solver
    cases = [(dict(method="gap", C=8, T=5), True), (dict(method="gap", C=24, T=5), True),
             (dict(method="admm", C=8, T=5), True), (dict(method="gap", C=28, T=5), True),
             (dict(method="gap", C=8, T=3), True), (dict(method="gap", C=5, T=5), False),
             (dict(method="gap", C=8, T=9), False), (dict(method="gap", C=40, T=5), False)]
    for kw, want in cases:
        with Solver(1, 32, 32, kw["C"], method=kw["method"], tv_iter_max=kw["T"]) as s:
            assert s.uses_fused == want, kw
    with Solver(1, 32, 32, 8, method="gap", fused=False) as s:
        assert not s.uses_fused


@pytest.mark.parametrize("shape", [
    # (B, H, W, C, tv_iter_max, method, accelerate, phi_batched)
    (1, 37, 45, 8, 5, "gap", True, False),       # odd sizes: plane rows staged without TMA
    (2, 3, 29, 4, 5, "gap", True, False),        # fewer rows than dual iterations
    (1, 1, 64, 8, 5, "gap", False, False),       # a single row
    (1, 50, 5, 12, 4, "gap", True, False),       # narrower than one pixel group, tv_iter_max 4
    (3, 41, 70, 16, 3, "gap", True, True),       # per-measurement masks, tv_iter_max 3
    (2, 33, 52, 28, 5, "admm", True, False),     # ADMM, 7 chunk-warps
    (1, 130, 300, 32, 5, "admm", True, False),   # widest supported channel count
    (5, 64, 64, 20, 5, "gap", True, False),      # more measurements than row bands
])
def test_fused_equals_exact_on_ragged_shapes(sp, shape):
    """The one-pass kernel against the statement-order kernels on awkward shapes (image edges inside
    a pixel group, segments shorter than the pipeline, batches, per-measurement masks)."""
    import torch
    from scipnp import Solver
    B, H, W, C, T, method, acc, pb = shape
    g = torch.Generator(device="cuda").manual_seed(H * 1000 + W)
    Phi = (torch.rand((B, H, W, C) if pb else (H, W, C), device="cuda", generator=g) <= 0.5).float()
    orig = torch.rand((B, H, W, C), device="cuda", generator=g)
    y = (Phi * orig).sum(-1)
    outs = []
    for fused in (True, False):
        with Solver(B, H, W, C, method=method, accelerate=acc, tv_weight=0.2, tv_iter_max=T,
                    phi_batched=pb, fused=fused, gamma=0.02) as s:
            assert s.uses_fused == fused
            s.load(y, Phi, X_orig=orig)
            s.run(5)
            outs.append((s.get_x(), s.psnr_all()))
    (xf, pf), (xe, pe) = outs
    assert np.isfinite(xf).all()
    assert np.abs(xf - xe).max() <= 2e-5
    assert np.abs(pf - pe).max() <= 1e-3


def test_joint_admm_clip_golden(sp, golden, path):
    """the joint module's ADMM-TV (theta clipped to [0,1], gamma = 0)."""
    from scipnp import joint_pnp_sci_algo as J
    g = golden("joint_admm")
    A, At = _ops(g["mask"])
    x, ps, ss, pa = J.admm_denoise(g["y"], _psum(g["mask"]), A, At, _lambda=1, gamma=0.0, denoiser='tv',
                                   iter_max=12, tv_weight=0.3, tv_iter_max=5, X_orig=g["X_orig"],
                                   tvm='ITV2D_cham')
    _cmp(x, g["x"], pa, g["psnr_all"], path)
    assert np.abs(np.array(ps) - g["psnr"]).max() <= TOL_DB
    with pytest.raises(ValueError):
        J.admm_denoise(g["y"], _psum(g["mask"]), A, At, denoiser='ffdnet', iter_max=1)


def test_joint_multistep_and_two_period_golden(sp, golden, path):
    """the TV + learned-denoiser period (projection and TV fused on the device, the
    estimate handed to the second denoiser as a CUDA tensor) and the two-period driver, against
    the reference's own loops run with the same stand-in in FFDNet's place."""
    import torch
    from scipnp import joint_pnp_sci_algo as J
    seen = []

    def second(x_dev, nsig, model=None):          # make_golden.py: second_standin, in torch
        assert x_dev.is_cuda and x_dev.dtype == torch.float32
        seen.append(float(nsig))
        a = float(np.float32(1.0 - 0.1 * float(nsig)))
        x_dev.mul_(a).add_(float(np.float32(0.01))).clamp_(0, 1)

    g = golden("joint_multistep")
    A, At = _ops(g["mask"])
    ms = _psum(g["mask"])
    x, ps, ss, pa = J.gap_multistep_denoise(g["y"], ms, A, At, denoiser='tv+ffdnet', iter_max=[3, 3],
                                            sigma=[0.2, 0.1], tv_weight=0.3, tv_iter_max=5,
                                            X_orig=g["X_orig"], second_denoiser=second)
    assert seen == [0.2] * 3 + [0.1] * 3
    _cmp(x, g["x"], pa, g["psnr_all"], path)
    assert np.abs(np.array(ps) - g["psnr"]).max() <= TOL_DB
    g = golden("joint_two_period")
    x, ps, ss, pa = J.gap_joint_denoise(g["y"], ms, A, At, X_orig=g["X_orig"], denoiser='tv+ffdnet',
                                        iter_max1=4, iter_max2=[2, 2], sigma1=None, sigma2=[0.2, 0.1],
                                        _lambda=1, accelerate=True, tv_weight=0.3, tv_iter_max=5,
                                        second_denoiser=lambda xd, ns, m: torch.clamp(
                                            xd * float(np.float32(1.0 - 0.1 * ns)) + float(np.float32(0.01)), 0, 1))
    assert len(pa) == 4
    _cmp(x, g["x"], pa, g["psnr_all"], path)
    with pytest.raises(NotImplementedError):       # the CNNs themselves are not part of the engine
        J.gap_multistep_denoise(g["y"], ms, A, At, iter_max=1, sigma=0.1)
    with pytest.raises(ValueError):
        J.gap_multistep_denoise(g["y"], ms, A, At, denoiser='tv', iter_max=1, second_denoiser=second)


def test_joint_admm_multistep_and_two_period_golden(sp, golden, path):
    """The ADMM twin: the second denoiser sits between the TV step and the multiplier update."""
    import torch
    from scipnp import joint_pnp_sci_algo as J

    def second(theta_dev, nsig, model=None):
        assert theta_dev.is_cuda
        return torch.clamp(theta_dev * float(np.float32(1.0 - 0.1 * float(nsig))) + float(np.float32(0.01)), 0, 1)

    g = golden("joint_admm_multistep")
    A, At = _ops(g["mask"])
    ms = _psum(g["mask"])
    x, ps, ss, pa = J.admm_multistep_denoise(g["y"], ms, A, At, gamma=0.01, iter_max=[3, 3],
                                             sigma=[0.2, 0.1], tv_weight=0.3, tv_iter_max=5,
                                             X_orig=g["X_orig"], second_denoiser=second, tvm='ITV3D_FGP')
    _cmp(x, g["x"], pa, g["psnr_all"], path)
    g = golden("joint_admm_two_period")
    x, ps, ss, pa = J.admm_joint_denoise(g["y"], ms, A, At, X_orig=g["X_orig"], iter_max1=4,
                                         iter_max2=[2, 2], sigma1=None, sigma2=[0.2, 0.1], _lambda=1,
                                         gamma=0.01, tv_weight=0.3, tv_iter_max=5, second_denoiser=second)
    assert len(pa) == 4
    _cmp(x, g["x"], pa, g["psnr_all"], path)
    with pytest.raises(NotImplementedError):
        J.admm_multistep_denoise(g["y"], ms, A, At, iter_max=1, sigma=0.1)


@pytest.mark.parametrize("shape", [(40, 48, 8), (7, 7, 1), (9, 133, 3), (130, 11, 5)])
def test_frames_iqa_matches_skimage_restatement(sp, shape):
    """per-frame PSNR / SSIM of the return tuples on the device
    (scipnp_frames_iqa) against the oracle's compare_psnr / compare_ssim (skimage 0.17.2)."""
    from oracle import iqa as OI
    from scipnp.iqa import frames_iqa, frame_ssim
    rng = np.random.default_rng(5)
    H, W, Cc = shape
    ref = rng.random(shape, dtype=np.float32)
    img = np.clip(ref + np.float32(0.05) * rng.standard_normal(shape).astype(np.float32), 0, 1)
    img[..., 0] = ref[..., 0] * np.float32(0.5)                  # a very different frame too
    ps, ss = frames_iqa(ref, img)
    for c in range(Cc):
        assert abs(ps[c] - OI.compare_psnr(ref[..., c], img[..., c], data_range=1.)) <= 1e-9
        assert abs(ss[c] - OI.compare_ssim(ref[..., c], img[..., c], data_range=1.)) <= 1e-10
    assert frame_ssim(ref[..., 0], ref[..., 0]) == pytest.approx(1.0, abs=1e-12)
    assert frames_iqa(None, img) == ([], [])
    with pytest.raises(ValueError):
        frames_iqa(ref[:6], img[:6])                              # smaller than the 7x7 window


@pytest.mark.parametrize("shape,iters", [((2, 2, 1), 3), ((17, 23, 5), 1), ((17, 23, 5), 5), ((40, 48, 8), 20)])
def test_matlab_tv_atv_clip_bit_exact(sp, shape, iters):
    """TV_denoising.m (the MATLAB twin's default TV, anisotropic, iterative clipping)
    on the device against its float32 NumPy restatement, bit for bit."""
    from oracle import matlab_tv as M
    from scipnp import matlab_tv as G
    rng = np.random.default_rng(11)
    y = rng.random(shape, dtype=np.float32)
    for lam in (0.07, 0.5):
        np.testing.assert_array_equal(G.TV_denoising(y, lam, iters), M.TV_denoising(y, lam, iters))
    np.testing.assert_array_equal(G.TV_denoising(y[:, :, 0], 0.2, iters), M.TV_denoising(y[:, :, 0], 0.2, iters))
    with pytest.raises(ValueError):
        G.TV_denoising(y[:1], 0.1, 2)


def test_matlab_gapdenoise_loop(sp, golden):
    """gapdenoise.m:62-94 with tvm = 'ATV_ClipA': the device loop against the same loop written
    with the oracle's operators and TV restatement (Phisum = sum(mask.^2), as the MATLAB drivers do)."""
    from oracle import matlab_tv as M
    from oracle import pnp_sci as O
    from scipnp import matlab_tv as G
    g = golden("gap_acc")
    mask, y, Xo = g["mask"], g["y"], g["X_orig"]
    Phisum = np.sum(mask * mask, axis=2)
    Phisum[Phisum == 0] = 1
    f32 = np.float32
    for acc in (True, False):
        v = O.At_(y, mask)
        y1 = np.zeros_like(y)
        want_psnr = []
        for _ in range(6):
            yb = O.A_(v, mask)
            if acc:
                y1 = y1 + (y - yb)
                v = v + f32(0.2) * O.At_((y1 - yb) / Phisum, mask)
            else:
                v = v + f32(0.2) * O.At_((y - yb) / Phisum, mask)
            v = M.TV_denoising(v, 0.07, 5)
            want_psnr.append(O.psnr(Xo, v))
        got, pa = G.gapdenoise(y, mask, lambda_=0.2, maxiter=6, acc=acc, tvweight=0.07, tviter=5, orig=Xo)
        assert np.abs(got - v).max() <= TOL_EXACT
        assert np.abs(np.array(pa) - np.array(want_psnr)).max() <= TOL_DB
    with pytest.raises(ValueError):
        G.gapdenoise(y, mask, tvm='no_such_tv', maxiter=1)


def test_data_side_on_device(sp):
    """measurement synthesis, mask generators and the CASSI stack on the device
    against the reference's NumPy statements (pnp_sci_test_orig.py:112-136)."""
    import torch
    from oracle import pnp_sci as O
    from scipnp import data as D
    rng = np.random.default_rng(8)
    H, W, Cc, F = 20, 26, 4, 3
    orig = (rng.random((H, W, F * Cc)) * 255).astype(np.float32)
    mask = (rng.random((H, W, Cc)) * 2).astype(np.float32)              # not normalised: max < 2
    meas, mk = D.synth_measurements(orig, mask)
    want = np.zeros((H, W, F), np.float32)
    for i in range(F):
        want[:, :, i] = np.sum(orig[:, :, i * Cc:(i + 1) * Cc] * mask, 2)
    mmax = np.max(mask)
    np.testing.assert_array_equal(meas.cpu().numpy(), want / mmax)
    np.testing.assert_array_equal(mk.cpu().numpy(), mask / mmax)
    noisy, _ = D.synth_measurements(orig, mask, gaussian_noise_level=5, seed=1)
    again, _ = D.synth_measurements(orig, mask, gaussian_noise_level=5, seed=1)
    assert torch.equal(noisy, again)
    d = (noisy - meas) * float(mmax)
    assert 4.0 < float(d.std()) < 6.0 and abs(float(d.mean())) < 1.0
    pois, _ = D.synth_measurements(orig, mask, poisson_noise=True, seed=2)
    assert float((pois * float(mmax) - torch.round(pois * float(mmax))).abs().max()) < 1e-3
    with pytest.raises(ValueError):
        D.synth_measurements(orig[:, :, :5], mask)
    m = D.binary_mask(64, 48, 8, p=0.5, seed=3)
    assert m.shape == (64, 48, 8) and m.is_cuda and set(np.unique(m.cpu().numpy())) <= {0.0, 1.0}
    assert 0.45 < float(m.mean()) < 0.55
    assert torch.equal(m, D.binary_mask(64, 48, 8, p=0.5, seed=3))
    m2 = (rng.random((12, 9)) > 0.5).astype(np.float32)
    np.testing.assert_array_equal(D.shift_mask(m2, 5, 2).cpu().numpy(), O.cassi_shift_mask(m2, 5, 2))


def test_c_abi_kernel_entries_directly(sp):
    """The stateless C entries called with raw device pointers: one fused iteration equals
    scipnp_gap_project + scipnp_tv_chambolle, and the ADMM pieces compose to the reference update."""
    import ctypes as ct
    import torch
    from scipnp._lib import lib, check
    B, H, W, Cc, T = 2, 48, 56, 8, 5
    g = torch.Generator(device="cuda").manual_seed(3)
    Phi = (torch.rand((H, W, Cc), device="cuda", generator=g) <= 0.5).float()
    x = torch.rand((B, H, W, Cc), device="cuda", generator=g)
    y1 = 0.1 * torch.rand((B, H, W), device="cuda", generator=g)
    y = torch.rand((B, H, W), device="cuda", generator=g) * Cc / 2
    ps = Phi.sum(2)
    ps[ps == 0] = 1
    st = ct.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: ct.c_void_p(t.data_ptr())
    # exact: projection then TV
    xp, y1p = torch.empty_like(x), torch.empty_like(y1)
    check(lib.scipnp_gap_project(ptr(x), ptr(xp), ptr(y1), ptr(y1p), ptr(y), ptr(Phi), ptr(ps), 1.0, 1,
                                 B, H, W, Cc, 0, st))
    ws = torch.empty(lib.scipnp_tv_workspace_bytes(B, H, W, Cc), dtype=torch.uint8, device="cuda")
    xe = torch.empty_like(x)
    check(lib.scipnp_tv_chambolle(ptr(xp), ptr(xe), 0.3, 2e-4, T, B, H, W, Cc, ptr(ws), ws.numel(),
                                  None, None, 0, st))
    # fused: one call
    fws = torch.empty(max(256, lib.scipnp_gap_tv_workspace_bytes(B, H, W, Cc, T)), dtype=torch.uint8, device="cuda")
    xf, y1f = torch.empty_like(x), torch.empty_like(y1)
    flag = torch.zeros(4, dtype=torch.int32, device="cuda")
    check(lib.scipnp_gap_tv_fused(ptr(x), ptr(xf), ptr(y1), ptr(y1f), ptr(y), ptr(Phi), ptr(ps), 1.0, 1,
                                  0.3, 2e-4, T, B, H, W, Cc, 0, ptr(fws), fws.numel(), ptr(flag), st))
    torch.cuda.synchronize()
    assert int(flag[0]) == 0
    assert float((xf - xe).abs().max()) <= 2e-5
    assert float((y1f - y1p).abs().max()) <= 2e-5
    # in == out is refused (ping-pong contract)
    assert lib.scipnp_gap_tv_fused(ptr(x), ptr(x), ptr(y1), ptr(y1f), ptr(y), ptr(Phi), ptr(ps), 1.0, 1,
                                   0.3, 2e-4, T, B, H, W, Cc, 0, ptr(fws), fws.numel(), None, st) == -1
    # ADMM pieces: x = (theta+b) + At((y - A(theta+b))/(Phi_sum+gamma)); f = x - b; b -= x - theta
    theta, b = x, 0.05 * torch.rand_like(x)
    xo, fo = torch.empty_like(x), torch.empty_like(x)
    check(lib.scipnp_admm_project(ptr(theta), ptr(b), ptr(xo), ptr(fo), ptr(y), ptr(Phi), ptr(ps), 1.0, 0.01,
                                  B, H, W, Cc, 0, st))
    u = theta + b
    ref_x = u + ((y - (u * Phi).sum(-1)) / (ps + 0.01))[..., None] * Phi
    assert float((xo - ref_x).abs().max()) <= 1e-5
    assert float((fo - (ref_x - b)).abs().max()) <= 1e-5
    b2 = b.clone()
    check(lib.scipnp_admm_dual_update(ptr(b2), ptr(xo), ptr(theta), b2.numel(), st))
    assert float((b2 - (b - (xo - theta))).abs().max()) <= 1e-6


def test_tv_rec_loops_match_the_reference(sp, golden):
    """GAP_TV_rec / ADMM_TV_rec (pnp_sci_algo.py:866-907): 30 dual iterations per step on the exact path, ADMM with the
    per-iteration decay of the TV weight and eta; the reference's loops are float64, the engine float32."""
    from scipnp import pnp_sci_algo as P
    g = golden("tv_rec")
    H, W, Cc = g["mask"].shape
    out = P.GAP_TV_rec(g["y"], g["mask"], None, None, g["Phi_sum"], int(g["maxiter"]), float(g["step_size"]),
                       float(g["weight"]), H, W, Cc, g["X_orig"])
    assert out.dtype == np.float32 and float(np.abs(out - g["gap"]).max()) <= TOL_X
    out = P.ADMM_TV_rec(g["y"], g["mask"], None, None, g["Phi_sum"], int(g["maxiter"]), float(g["step_size"]),
                        float(g["weight"]), H, W, Cc, float(g["eta"]), g["X_orig"])
    assert float(np.abs(out - g["admm"]).max()) <= TOL_X


def test_admm_denoise_bayer_matches_the_oracle(sp, path):
    """Bayer ADMM-TV (pnp_sci_algo.py:268-475, built from admm_denoise's semantics: the reference's body is dead
    code) against the oracle's restatement, which equals four pinned admm_denoise solves."""
    from oracle import pnp_sci as O
    from scipnp import synth, pnp_sci_algo as P
    y, Phi, orig = synth.make_bayer(48, 64, 8, cfg=3)
    kw = dict(_lambda=1, gamma=0.01, denoiser='tv', iter_max=6, tv_weight=0.1, tv_iter_max=5, X_orig=orig)
    xo, pao = O.admm_denoise_bayer(y, Phi, **kw)
    xg, pag = P.admm_denoise_bayer(y, Phi, **kw)
    _cmp(xg, xo, pag, pao, path)


@pytest.mark.parametrize("B,H,W,Cc,pb", [(1, 37, 52, 8, False), (3, 20, 36, 24, True), (2, 33, 40, 12, False), (1, 19, 23, 5, False)])
def test_load_one_pass_init_and_borrowed_masks(sp, B, H, W, Cc, pb):
    """load(): Phi_sum and x0 = At(y) come out of one pass over the masks (C % 4 == 0) and must equal the separate
    operators bit for bit; a borrowed device mask stack (read in place) gives the same reconstruction as a copy."""
    import torch
    from scipnp.engine import Solver
    rng = np.random.default_rng(3)
    Phi = (rng.random((B, H, W, Cc) if pb else (H, W, Cc)) <= 0.5).astype(np.float32)
    Phi[..., 0, 0, :] = 0                                       # a pixel no mask opens: Phi_sum -> 1
    y = rng.random((B, H, W), dtype=np.float32)
    out = []
    for borrow in (False, True):
        Pd = torch.from_numpy(Phi).cuda()
        with Solver(B, H, W, Cc, method="gap", tv_weight=0.2, tv_iter_max=5, phi_batched=pb) as s:
            s.load(torch.from_numpy(y).cuda(), Pd, borrow_phi=borrow)
            x0 = s.get_x().copy()
            s.run(3)
            out.append((x0, s.get_x()))
    Pb = Phi if pb else np.broadcast_to(Phi, (B, H, W, Cc))
    np.testing.assert_array_equal(out[0][0], y[..., None] * Pb)          # x0 = At(y), utils.py:17-26
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])
    # Phi_sum through the public operator equals what the solver used: same iterate as an explicit Phi_sum
    ms = sp.phi_sum(Phi[0] if pb else Phi)
    ref = Phi[0].sum(axis=2) if pb else Phi.sum(axis=2)
    ref[ref == 0] = 1
    np.testing.assert_array_equal(np.asarray(ms), ref)


_FAMILY = [("TV_denoising_clip_LB", 0.07, 5), ("tvdenoise_cham_ATV2D", 1 / 0.07, 5), ("tvdenoise_cham_ITV2D", 1 / 0.07, 5),
           ("tvdenoise_cham_ITV3D", 1 / 0.07, 5), ("fgp_denoise_ATV2D", 0.07, 2), ("fgp_denoise_ITV2D", 0.07, 4),
           ("fgp_denoise_ITV3D", 0.07, 3)]


@pytest.mark.parametrize("name,lam,iters", _FAMILY)
@pytest.mark.parametrize("shape", [(24, 20, 8), (17, 23, 5), (9, 12, 24)])
def test_matlab_tv_family_bit_exact(sp, name, lam, iters, shape):
    """the rest of the MATLAB twin's TV family (gapdenoise.m:86-108) against the float32 NumPy
    restatement of the .m files (oracle/matlab_tv.py; parity unpinned: no MATLAB here), statement order kept."""
    from oracle import matlab_tv as M
    from scipnp import matlab_tv as G
    rng = np.random.default_rng(17)
    y = rng.random(shape, dtype=np.float32)
    got = getattr(G, name)(y, lam, iters)
    want = getattr(M, name)(y, lam, iters)
    assert got.dtype == np.float32 and want.dtype == np.float32
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("tvm", ["ATV_ClipB", "ATV_cham", "ATV_FGP", "ITV2D_cham", "ITV2D_FGP", "ITV3D_cham", "ITV3D_FGP"])
def test_matlab_gapdenoise_tvm_switch(sp, golden, tvm):
    """gapdenoise.m:92-108: every `tvm` branch with the weight and the iteration count as written there."""
    from oracle import matlab_tv as M
    from oracle import pnp_sci as O
    from scipnp import matlab_tv as G
    g = golden("gap_acc")
    mask, y = g["mask"], g["y"]
    Phisum = np.sum(mask * mask, axis=2)
    Phisum[Phisum == 0] = 1
    w = 0.07
    tv = {"ATV_ClipB": lambda v: M.TV_denoising_clip_LB(v, w, 5), "ATV_cham": lambda v: M.tvdenoise_cham_ATV2D(v, 1 / w, 5),
          "ATV_FGP": lambda v: M.fgp_denoise_ATV2D(v, w, 2), "ITV2D_cham": lambda v: M.tvdenoise_cham_ITV2D(v, 1 / w, 5),
          "ITV2D_FGP": lambda v: M.fgp_denoise_ITV2D(v, w, 2), "ITV3D_cham": lambda v: M.tvdenoise_cham_ITV3D(v, 1 / w, 5),
          "ITV3D_FGP": lambda v: M.fgp_denoise_ITV3D(v, w, 2)}[tvm]
    v = O.At_(y, mask)
    y1 = np.zeros_like(y)
    for _ in range(4):
        yb = O.A_(v, mask)
        y1 = y1 + (y - yb)
        v = tv(v + np.float32(0.2) * O.At_((y1 - yb) / Phisum, mask))
    got, _ = G.gapdenoise(y, mask, lambda_=0.2, maxiter=4, acc=True, tvweight=w, tvm=tvm)
    assert np.abs(got - v).max() <= TOL_EXACT
