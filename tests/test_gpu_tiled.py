"""GPU tests of the row-tiled multi-GPU mode with the transport bench.py measures: "p2p" -- CUDA-IPC mapped
neighbour buffers, halo rows pulled by tile_exchange_kernel, device-side flags (csrc/solver.cu).

Several ranks share cuda:0 here (one process per rank, gloo rendezvous, CUDA IPC between the processes), so the
whole flag protocol, the lazy acknowledgements and the pull kernel run exactly as on a multi-GPU box; only the
wire is missing.  Every test fails if the solver fell back to another transport.

Checked: tiled result == single-GPU solve of the same scene (owned rows never see a seam) for refresh periods
k = 1, 2, 3, two and three ranks, two runs back to back (halos must be fresh at a run boundary), the exact path
(in-place projection: the acknowledgement must be awaited before the very next step), and a forced early stop
(collective rollback to the exact path in the middle of a p2p run).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import scipnp
    return scipnp


def _worker(rank, world, port, H, W, C, iters, k, fused, tv_eps, method, out_dir, T=5):
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
        ts = TiledSolver(H, W, C, rank, world, tv_weight=0.3, tv_iter_max=T, exchange_every=k, transport="p2p",
                         fused=fused, tv_eps=tv_eps, method=method)
        assert ts.transport == "p2p", "fell back to %s" % ts.transport
        ts.load(torch.from_numpy(y[ts.row_lo:ts.row_hi]).cuda(), torch.from_numpy(mask[ts.row_lo:ts.row_hi]).cuda())
        ts.run(iters // 2)
        ts.run(iters - iters // 2)             # two runs: halos must be fresh at a run boundary
        np.save(os.path.join(out_dir, "x_%d.npy" % rank), ts.result().cpu().numpy())
        np.save(os.path.join(out_dir, "meta_%d.npy" % rank), np.array([ts.refined_iters, int(ts.uses_fused), int(ts.push)]))
        ts.close()
    finally:
        dist.destroy_process_group()


def _tiled(tmp_path, world, H, W, C, iters, k, fused=True, tv_eps=2e-4, method="gap", T=5):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, H, W, C, iters, k, fused, tv_eps, method, str(tmp_path), T), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / ("x_%d.npy" % r)) for r in range(world)], axis=0)
    meta = np.stack([np.load(tmp_path / ("meta_%d.npy" % r)) for r in range(world)])
    return got, meta


def _single(H, W, C, iters, fused=True, tv_eps=2e-4, method="gap", T=5):
    from scipnp import synth, Solver
    meas, mask, _ = synth.make_cacti(H, W, C, 1, cfg=21)
    y = meas[:, :, 0] / np.float32(255.)
    with Solver(1, H, W, C, method=method, tv_weight=0.3, tv_iter_max=T, fused=fused, tv_eps=tv_eps) as so:
        so.load(y[None], mask)
        so.run(iters // 2)
        so.run(iters - iters // 2)
        return so.get_x()[0], so.refined_iters


@pytest.mark.parametrize("world,k", [(2, 1), (2, 2), (2, 3), (3, 1), (3, 2)])
def test_p2p_tiled_equals_single_gpu(sp, tmp_path, world, k):
    H, W, C, iters = 64 * world + 8, 128, 8, 7
    got, meta = _tiled(tmp_path, world, H, W, C, iters, k)
    ref, refined = _single(H, W, C, iters)
    assert refined == 0 and (meta[:, 0] == 0).all() and (meta[:, 1] == 1).all()
    # one exchange per iteration = the halo push inside the fused kernel (what bench.py measures at N > 1)
    assert (meta[:, 2] == (1 if k == 1 else 0)).all()
    assert float(np.abs(got - ref).max()) <= 1e-6


@pytest.mark.parametrize("world,H,W,C,T", [(2, 96, 3840, 24, 5), (3, 150, 256, 8, 4), (4, 64, 512, 12, 5), (2, 41, 64, 4, 3)])
def test_push_tiled_shapes(sp, tmp_path, world, H, W, C, T):
    """Halo push (k = 1) on the bench's tile width and on ragged row counts / other channel counts and dual
    iteration counts: the seam rows written by the neighbours' TMA stores, the y1 rows by plain peer stores."""
    got, meta = _tiled(tmp_path, world, H, W, C, 6, 1, T=T)
    ref, refined = _single(H, W, C, 6, T=T)
    assert refined == 0 and (meta[:, 1] == 1).all() and (meta[:, 2] == 1).all()
    assert float(np.abs(got - ref).max()) <= 1e-6


@pytest.mark.parametrize("world,H,W,C,k,fused", [(2, 72, 128, 8, 1, True), (3, 100, 192, 24, 1, True), (2, 72, 128, 8, 2, True),
                                                  (2, 72, 128, 12, 1, False)])
def test_tiled_admm_equals_single_gpu(sp, tmp_path, world, H, W, C, k, fused):
    """ADMM-TV over row tiles: theta and the multiplier b are the carried arrays (the seam rows of both are pushed by
    the fused kernel at k = 1 -- b by plain peer stores of the consumers, staged by TMA at C = 8, read from global
    memory at C = 24 -- and pulled by the exchange kernel otherwise); the result is x, the projection output."""
    iters = 6
    got, meta = _tiled(tmp_path, world, H, W, C, iters, k, fused=fused, method="admm")
    ref, refined = _single(H, W, C, iters, fused=fused, method="admm")
    assert refined == 0 and (meta[:, 1] == (1 if fused else 0)).all()
    assert (meta[:, 2] == (1 if (fused and k == 1) else 0)).all()
    assert float(np.abs(got - ref).max()) <= (1e-6 if fused else 0.0)


def test_push_tiled_rollback_whole_scene_rule(sp, tmp_path):
    """k = 1 (push) with an eps at which skimage's rule fires: the decision is taken on the energies of the whole
    scene (the per-iteration log summed over the ranks), every rank rolls back, the exact path (pull exchange)
    redoes the run with the single-GPU stopping decisions."""
    H, W, C, iters = 136, 128, 8, 4
    got, meta = _tiled(tmp_path, 2, H, W, C, iters, 1, tv_eps=0.5)
    ref, refined = _single(H, W, C, iters, tv_eps=0.5)
    assert refined == iters and (meta[:, 0] == iters).all() and (meta[:, 2] == 1).all()
    assert float(np.abs(got - ref).max()) <= 2e-6


@pytest.mark.parametrize("k", [1, 2])
def test_p2p_tiled_exact_path(sp, tmp_path, k):
    """fused = False: the projection runs in place, so the step right after an exchange must wait for the
    neighbours' acknowledgement (it used to overwrite rows that were still being pulled)."""
    H, W, C, iters = 136, 128, 8, 5
    got, meta = _tiled(tmp_path, 2, H, W, C, iters, k, fused=False)
    ref, _ = _single(H, W, C, iters, fused=False)
    assert (meta[:, 1] == 0).all()
    np.testing.assert_array_equal(got, ref)


def test_p2p_tiled_rollback_in_the_middle_of_a_run(sp, tmp_path):
    """A huge eps makes the early stop fire: every rank rolls back and redoes the run on the exact path over the
    same p2p links.  With the [C][T] energies summed over the owned rows of all ranks the stopping decisions are
    those of the single-GPU solve, so the results agree."""
    H, W, C, iters = 136, 128, 8, 4
    got, meta = _tiled(tmp_path, 2, H, W, C, iters, 2, tv_eps=0.5)
    ref, refined = _single(H, W, C, iters, tv_eps=0.5)
    assert refined == iters and (meta[:, 0] == iters).all()
    assert float(np.abs(got - ref).max()) <= 2e-6


def test_p2p_tiled_config5_width(sp, tmp_path):
    """Two ranks on a 3840-wide, 24-channel band (the bench's tile shape, fewer rows), k = 2."""
    H, W, C, iters = 96, 3840, 24, 4
    got, meta = _tiled(tmp_path, 2, H, W, C, iters, 2)
    ref, _ = _single(H, W, C, iters)
    assert (meta[:, 1] == 1).all()
    assert float(np.abs(got - ref).max()) <= 1e-6


def _stream_worker(rank, world, port, H, W, C, iters, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        from scipnp.tiled import TiledSolver
        from scipnp import synth
        ts = TiledSolver(H, W, C, rank, world, tv_weight=0.3, tv_iter_max=5, exchange_every=2, transport="p2p")
        jobs = []
        for j in range(3):
            meas, mask, _ = synth.make_cacti(H, W, C, 1, cfg=70 + j)
            y = meas[:, :, 0] / np.float32(255.)
            jobs.append((torch.from_numpy(np.ascontiguousarray(y[ts.row_lo:ts.row_hi])).pin_memory(),
                         torch.from_numpy(np.ascontiguousarray(mask[ts.row_lo:ts.row_hi])).pin_memory(),
                         torch.empty((ts.hi - ts.lo, W, C), dtype=torch.float32).pin_memory()))
        ts.run_host_stream(jobs, iters)
        for j in range(3):
            np.save(os.path.join(out_dir, "s_%d_%d.npy" % (j, rank)), jobs[j][2].numpy())
        ts.close()
    finally:
        dist.destroy_process_group()


def test_host_stream_of_reconstructions(sp, tmp_path):
    """run_host_stream: three different scenes from pinned host buffers, copies overlapped with the iterations
    of the neighbouring jobs (two staging sets): every job equals the single-GPU solve of its scene."""
    import socket
    import torch.multiprocessing as mp
    from scipnp import synth, Solver
    world, H, W, C, iters = 2, 136, 128, 8, 5
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_stream_worker, args=(world, port, H, W, C, iters, str(tmp_path)), nprocs=world, join=True)
    for j in range(3):
        meas, mask, _ = synth.make_cacti(H, W, C, 1, cfg=70 + j)
        y = meas[:, :, 0] / np.float32(255.)
        with Solver(1, H, W, C, method="gap", tv_weight=0.3, tv_iter_max=5) as so:
            so.load(y[None], mask)
            so.run(iters)
            ref = so.get_x()[0]
        got = np.concatenate([np.load(tmp_path / ("s_%d_%d.npy" % (j, r))) for r in range(world)], axis=0)
        assert float(np.abs(got - ref).max()) <= 1e-6, "job %d" % j


def test_p2p_tiled_stopping_rule_sums_energies_over_all_ranks(sp, tmp_path):
    """tv_iter_max = 20 with an eps at which slices stop at different dual iterations: the exact path all-reduces the
    per-slice energies of the owned rows (scipnp_solver_set_energy_reduce), so every tile stops where the
    single-GPU solve stops.  The same scene with eps = 0 differs visibly, i.e. the rule did fire."""
    H, W, C, iters = 160, 128, 8, 3
    got, meta = _tiled(tmp_path, 2, H, W, C, iters, 1, fused=False, tv_eps=2e-3, T=20)
    ref, _ = _single(H, W, C, iters, fused=False, tv_eps=2e-3, T=20)
    full, _ = _single(H, W, C, iters, fused=False, tv_eps=0.0, T=20)
    assert float(np.abs(ref - full).max()) > 1e-4, "the stopping rule never fired: the test would be vacuous"
    assert float(np.abs(got - ref).max()) <= 1e-6


def _onecall_worker(rank, world, port, H, W, C, iters, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        from scipnp.tiled import gap_denoise_tiled
        from scipnp import synth
        meas, mask, orig = synth.make_cacti(H, W, C, 1, cfg=21)
        y = meas[:, :, 0] / np.float32(255.)
        x, ps, ss, pa = gap_denoise_tiled(y, None, Phi=mask, iter_max=iters, tv_weight=0.3, tv_iter_max=5,
                                          X_orig=orig[:, :, :C] / np.float32(255.), transport="p2p")
        np.save(os.path.join(out_dir, "one_%d.npy" % rank), x)
        np.save(os.path.join(out_dir, "one_ps_%d.npy" % rank), np.array(ps))
    finally:
        dist.destroy_process_group()


def test_gap_denoise_tiled_one_call_entry(sp, tmp_path):
    """The reference's gap_denoise signature over two ranks: every rank passes the whole scene and gets the whole
    reconstruction back, equal to the single-GPU gap_denoise."""
    import socket
    import torch.multiprocessing as mp
    H, W, C, iters = 72, 128, 8, 6
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_onecall_worker, args=(2, port, H, W, C, iters, str(tmp_path)), nprocs=2, join=True)
    from scipnp import synth
    meas, mask, orig = synth.make_cacti(H, W, C, 1, cfg=21)
    y = meas[:, :, 0] / np.float32(255.)
    ms = mask.sum(axis=2)
    ms[ms == 0] = 1
    ref, ps, _, _ = sp.gap_denoise(y, ms, Phi=mask, iter_max=iters, tv_weight=0.3, tv_iter_max=5,
                                   X_orig=orig[:, :, :C] / np.float32(255.))
    for r in range(2):
        x = np.load(tmp_path / ("one_%d.npy" % r))
        assert x.shape == (H, W, C) and float(np.abs(x - ref).max()) <= 1e-6
        assert np.abs(np.load(tmp_path / ("one_ps_%d.npy" % r)) - np.array(ps)).max() <= 1e-3
