"""Multi-process (gloo, CPU) tests of the row-tiling protocol used for a single large
scene on several GPUs: partitioning, halo bookkeeping and the neighbour exchange.  The
local solver is the CPU oracle here, so the test isolates the host-side logic: the
tiled result must equal the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scipnp.tiled import partition_rows, exchange_halos, tiled_reference_run


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene(H, W, C, seed=3):
    rng = np.random.default_rng(seed)
    mask = (rng.random((H, W, C)) <= 0.5).astype(np.float32)
    orig = rng.random((H, W, C), dtype=np.float32)
    y = np.sum(mask * orig, axis=2)
    return y, mask


def _oracle_full(y, mask, iters, T, w):
    from oracle import pnp_sci as O
    A = lambda x: O.A_(x, mask)
    At = lambda v: O.At_(v, mask)
    return O.gap_denoise(y, O.phi_sum(mask), A, At, iter_max=iters, tv_weight=w, tv_iter_max=T)[0]


def _worker(rank, world, port, H, W, C, iters, T, k, w, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pnp_sci as O
        from oracle.tv_chambolle import denoise_tv_chambolle
        y, mask = _scene(H, W, C)
        halo = k * (T - 1)
        lo, hi, rlo, rhi = partition_rows(H, world, rank, halo)
        yl, ml = y[rlo:rhi], mask[rlo:rhi]
        ms = O.phi_sum(ml)
        x = torch.from_numpy(O.At_(yl, ml).copy())
        y1 = torch.zeros((rhi - rlo, W), dtype=torch.float32)

        def step():
            xn, y1n = x.numpy(), y1.numpy()
            yb = O.A_(xn, ml)
            y1n[...] = y1n + (yl - yb)
            xn[...] = xn + 1 * O.At_((y1n - yb) / ms, ml)
            xn[...] = denoise_tv_chambolle(xn, w, n_iter_max=T, multichannel=True)

        tiled_reference_run(step, lambda: [x, y1], H, halo, k, iters, rank, world)
        np.save(os.path.join(out_dir, "x_%d.npy" % rank), x.numpy()[lo - rlo:hi - rlo])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,k", [(2, 37, 1), (3, 41, 2), (2, 12, 2)])
def test_tiled_equals_single_process(tmp_path, world, H, k):
    W, C, iters, T, w = 18, 4, 5, 5, 0.3
    mp.spawn(_worker, args=(world, _free_port(), H, W, C, iters, T, k, w, str(tmp_path)),
             nprocs=world, join=True)
    y, mask = _scene(H, W, C)
    ref = _oracle_full(y, mask, iters, T, w)
    got = np.concatenate([np.load(tmp_path / ("x_%d.npy" % r)) for r in range(world)], axis=0)
    assert got.shape == ref.shape
    np.testing.assert_array_equal(got, ref)


def test_partition_covers_rows_exactly():
    for H, world, halo in [(2160, 8, 4), (2160, 7, 16), (10, 4, 4), (5, 5, 2)]:
        parts = [partition_rows(H, world, r, halo) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == H
        for a, b in zip(parts, parts[1:]):
            assert a[1] == b[0]
        for lo, hi, rlo, rhi in parts:
            assert rlo == max(0, lo - halo) and rhi == min(H, hi + halo)


def test_single_rank_exchange_is_a_no_op():
    f = torch.arange(12.).reshape(4, 3)
    exchange_halos([f], 4, 2, 0, 1)
    assert torch.equal(f, torch.arange(12.).reshape(4, 3))


# -- independent measurements sharded over ranks (gloo) ------------------------------------------------

def _shard_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pnp_sci as O
        from scipnp import synth
        from scipnp.sharded import admmdenoise_cacti_sharded
        meas, mask, orig = synth.make_cacti(24, 20, 4, 5, cfg=41)
        A = lambda x: O.A_(x, mask)
        At = lambda v: O.At_(v, mask)
        res = admmdenoise_cacti_sharded(meas, mask, A, At, projmeth='gap', orig=orig, nframe=5, MAXB=255.,
                                        maskdirection='updown', solve=O.admmdenoise_cacti, _lambda=1,
                                        accelerate=True, denoiser='tv', iter_max=4, tv_weight=0.3,
                                        tv_iter_max=5)
        np.save(os.path.join(out_dir, "x_%d.npy" % rank), res[0])
        np.save(os.path.join(out_dir, "p_%d.npy" % rank), np.array(res[2]))
        np.save(os.path.join(out_dir, "pa_%d.npy" % rank), np.array(res[4]))
    finally:
        dist.destroy_process_group()


def test_sharded_measurements_equal_single_process(tmp_path):
    from oracle import pnp_sci as O
    from scipnp import synth
    from scipnp.sharded import shard_indices
    world = 2
    mp.spawn(_shard_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    meas, mask, orig = synth.make_cacti(24, 20, 4, 5, cfg=41)
    A = lambda x: O.A_(x, mask)
    At = lambda v: O.At_(v, mask)
    ref = O.admmdenoise_cacti(meas, mask, A, At, projmeth='gap', orig=orig, nframe=5, MAXB=255.,
                              maskdirection='updown', _lambda=1, accelerate=True, denoiser='tv',
                              iter_max=4, tv_weight=0.3, tv_iter_max=5)
    for r in range(world):                 # every rank holds the complete, identical result
        np.testing.assert_array_equal(np.load(tmp_path / ("x_%d.npy" % r)), ref[0])
        np.testing.assert_array_equal(np.load(tmp_path / ("p_%d.npy" % r)), np.array(ref[2]))
        np.testing.assert_array_equal(np.load(tmp_path / ("pa_%d.npy" % r)), np.array(ref[4]))
    assert shard_indices(5, 2, 0) == [0, 2, 4] and shard_indices(5, 2, 1) == [1, 3]
    assert sorted(shard_indices(28, 8, 3)) == [3, 11, 19, 27]


def test_whole_scene_stop_rule_on_logged_energies():
    """The tiled driver sums the per-rank energy logs and replays skimage's eps test (oracle/tv_chambolle.py: the
    loop breaks at dual iteration i >= 1 when |E_{i-1} - E_i| < eps * E_0).  The replay must take the oracle's
    decision: against the oracle's own executed-iteration count, on scenes where the rule fires and where it does
    not.  (A stop at the last iteration returns the same image and is not reported.)"""
    import torch
    from scipnp.tiled import stop_rule_hits
    from oracle.tv_chambolle import denoise_tv_chambolle
    rng = np.random.default_rng(4)
    img = rng.random((24, 20, 3)).astype(np.float32)
    img[:, :, 2] = 0.5 + 0.01 * img[:, :, 2]              # a nearly flat slice: its energies settle at once
    T = 5
    full = []
    denoise_tv_chambolle(img, 0.3, eps=0.0, n_iter_max=T, multichannel=True, energy_out=full)
    assert all(len(e) == T for e in full)
    logged = torch.tensor(np.array(full)[:, :T - 1], dtype=torch.float64)      # what the fused kernel logs: E_0 .. E_{T-2}
    seen = set()
    for eps in (0.9, 0.3, 0.05, 2e-4, 1e-7):
        got = []
        denoise_tv_chambolle(img, 0.3, eps=eps, n_iter_max=T, multichannel=True, energy_out=got)
        stopped_early = [len(e) <= T - 1 for e in got]                          # broke at some i <= T-2
        fired = stop_rule_hits(logged, eps).tolist()
        assert fired == stopped_early, (eps, fired, stopped_early)
        seen.update(stopped_early)
    assert seen == {True, False}
