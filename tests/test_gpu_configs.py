"""GPU parity tests at the sizes of BASELINE.json's configurations, every one of them
against the CPU oracle on the same seeded inputs:

  c2  ADMM-TV, the 28 coded frames of the six grayscale benchmarks (4+6+5+5+4+4), 256x256x8, 40 iterations
  c3  GAP-TV Bayer 512x512x24 (four sub-lattices of 256x256x24), 5 iterations
  c5  GAP-TV 3840x2160x24: the full scene for 2 iterations and a 512-row full-width crop for 40 iterations

Tolerance (north_star): max abs <= 1e-4 on [0,1] frames, |dPSNR| <= 0.01 dB.

The oracle's Chambolle TV treats every channel as an independent 2-D problem (oracle/tv_chambolle.py); for the
UHD cases the test maps those per-channel calls over a process pool so that the CPU side finishes in about a
minute instead of six.  The arithmetic of every channel is untouched.
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_X, TOL_DB = 1e-4, 0.01


@pytest.fixture(scope="module")
def sp():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import scipnp
    return scipnp


def _ops(mask):
    from oracle import pnp_sci as O
    return (lambda x: O.A_(x, mask)), (lambda y: O.At_(y, mask))


def _tv_channel(args):
    from oracle.tv_chambolle import tv_chambolle_2d
    img, weight, eps, n_iter_max = args
    return tv_chambolle_2d(img, weight, eps, n_iter_max)


class _ParallelTV:
    """oracle.tv_chambolle.denoise_tv_chambolle(multichannel=True) with the independent channels spread over
    worker processes (forked before they could ever touch CUDA; they only run NumPy)."""

    def __init__(self, procs):
        self.pool = mp.get_context("fork").Pool(procs)

    def __call__(self, image, weight=0.1, eps=2.e-4, n_iter_max=200, multichannel=False, energy_out=None):
        assert multichannel and image.ndim == 3 and energy_out is None
        chans = [(np.ascontiguousarray(image[..., c]), weight, eps, n_iter_max) for c in range(image.shape[-1])]
        out = np.empty_like(image)
        for c, r in enumerate(self.pool.map(_tv_channel, chans)):
            out[..., c] = r
        return out

    def close(self):
        self.pool.close()
        self.pool.join()


@pytest.fixture()
def parallel_oracle(monkeypatch):
    from oracle import pnp_sci as O
    tv = _ParallelTV(min(24, os.cpu_count() or 1))
    monkeypatch.setattr(O, "denoise_tv_chambolle", tv)
    yield O
    tv.close()


def test_parallel_oracle_is_the_oracle(parallel_oracle):
    """The process-pool TV is bit-identical to the serial oracle."""
    from oracle.tv_chambolle import denoise_tv_chambolle
    rng = np.random.default_rng(0)
    x = rng.random((40, 56, 6)).astype(np.float32)
    np.testing.assert_array_equal(parallel_oracle.denoise_tv_chambolle(x, 0.3, n_iter_max=5, multichannel=True),
                                  denoise_tv_chambolle(x, 0.3, n_iter_max=5, multichannel=True))


# -- config 2 ---------------------------------------------------------------------------------------------

C2_SCENES = [("kobe", 4), ("traffic", 6), ("runner", 5), ("drop", 5), ("crash", 4), ("aerial", 4)]   # 28 coded frames


def test_config2_admm_28_measurements_40_iterations(sp):
    """BASELINE config 2 in full: ADMM-TV (gamma 0.01, tv_weight 0.3, tv_iter_max 5, 40 iterations,
    pnp_sci_demo_kobe.py:107-113) on the 28 coded frames of six synthetic scenes, each scene through
    admmdenoise_cacti (its frames are one batched solve on the device)."""
    from oracle import pnp_sci as O
    from scipnp import synth
    worst_x = worst_db = 0.0
    nmeas = 0
    for i, (name, F) in enumerate(C2_SCENES):
        meas, mask, orig = synth.make_cacti(256, 256, 8, F, cfg=20 + i)
        A, At = _ops(mask)
        kw = dict(projmeth='admm', orig=orig, nframe=F, MAXB=255., _lambda=1, gamma=0.01, denoiser='tv',
                  iter_max=40, tv_weight=0.3, tv_iter_max=5)
        xo, _, pso, _, pao = O.admmdenoise_cacti(meas, mask, A, At, **kw)
        xg, _, psg, _, pag = sp.admmdenoise_cacti(meas, mask, A, At, **kw)
        worst_x = max(worst_x, float(np.abs(xg - xo).max()))
        worst_db = max(worst_db, float(np.abs(np.array(pag) - np.array(pao)).max()),
                       float(np.abs(np.array(psg) - np.array(pso)).max()))
        nmeas += F
    assert nmeas == 28
    assert worst_x <= TOL_X, "max abs %.3g" % worst_x
    assert worst_db <= TOL_DB


# -- config 3 ---------------------------------------------------------------------------------------------

def test_config3_bayer_512(sp):
    """BASELINE config 3: GAP-TV on a 512x512x24 Bayer mosaic (pnp_sci_demo_bayer.py:143-148: tv_weight 0.1,
    tv_iter_max 5), 5 outer iterations: four 256x256x24 sub-lattices with their own masks."""
    from oracle import pnp_sci as O
    from scipnp import synth
    y, Phi, orig = synth.make_bayer(512, 512, 24, cfg=3)
    kw = dict(_lambda=1, accelerate=True, denoiser='tv', iter_max=5, tv_weight=0.1, tv_iter_max=5, X_orig=orig)
    xo, pso, _, pao = O.gap_denoise_bayer(y, Phi, **kw)
    xg, psg, _, pag = sp.gap_denoise_bayer(y, Phi, **kw)
    assert xg.shape == (512, 512, 24)
    assert float(np.abs(xg - xo).max()) <= TOL_X
    assert float(np.abs(np.array(pag) - np.array(pao)).max()) <= TOL_DB
    assert float(np.abs(np.array(psg) - np.array(pso)).max()) <= TOL_DB


# -- config 5 ---------------------------------------------------------------------------------------------

def _c5_inputs(rows):
    from scipnp import synth
    meas, mask, orig = synth.make_cacti(rows, 3840, 24, 1, cfg=5)
    return meas[:, :, 0] / np.float32(255.), mask, orig / np.float32(255.)


def test_config5_full_scene_two_iterations(sp, parallel_oracle):
    """BASELINE config 5 at full size (3840x2160x24), 2 outer iterations against the oracle."""
    O = parallel_oracle
    y, mask, orig = _c5_inputs(2160)
    A, At = _ops(mask)
    ms = O.phi_sum(mask)
    kw = dict(_lambda=1, accelerate=True, denoiser='tv', iter_max=2, tv_weight=0.3, tv_iter_max=5, X_orig=orig)
    xo, _, _, pao = O.gap_denoise(y, ms, A, At, **kw)
    xg, _, _, pag = sp.gap_denoise(y, ms, Phi=mask, **kw)
    assert float(np.abs(xg - xo).max()) <= TOL_X
    assert float(np.abs(np.array(pag) - np.array(pao)).max()) <= TOL_DB


def test_config5_crop_512_rows_40_iterations(sp, parallel_oracle):
    """A 512-row, full-width crop of the config-5 scene for the full 40 iterations: the accumulated
    difference of the single-precision fused path after a whole reconstruction."""
    O = parallel_oracle
    y, mask, orig = _c5_inputs(512)
    A, At = _ops(mask)
    ms = O.phi_sum(mask)
    kw = dict(_lambda=1, accelerate=True, denoiser='tv', iter_max=40, tv_weight=0.3, tv_iter_max=5, X_orig=orig)
    xo, _, _, pao = O.gap_denoise(y, ms, A, At, **kw)
    xg, _, _, pag = sp.gap_denoise(y, ms, Phi=mask, **kw)
    assert float(np.abs(xg - xo).max()) <= TOL_X
    assert float(np.abs(np.array(pag) - np.array(pao)).max()) <= TOL_DB
    assert pag[-1] > pag[0] + 3.0                     # the reconstruction converges


# -- independent measurements sharded over ranks, with the real solver -----------------------

def _shard_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        from scipnp import synth
        from scipnp.sharded import admmdenoise_cacti_sharded
        meas, mask, orig = synth.make_cacti(64, 96, 8, 5, cfg=41)
        res = admmdenoise_cacti_sharded(meas, mask, None, None, projmeth='admm', orig=orig, nframe=5, MAXB=255.,
                                        maskdirection='updown', _lambda=1, gamma=0.01, denoiser='tv', iter_max=8,
                                        tv_weight=0.3, tv_iter_max=5)
        np.save(os.path.join(out_dir, "x_%d.npy" % rank), res[0])
        np.save(os.path.join(out_dir, "pa_%d.npy" % rank), np.array(res[4]))
    finally:
        dist.destroy_process_group()


def test_sharded_measurements_with_the_cuda_solver(sp, tmp_path):
    """admmdenoise_cacti_sharded (frames round-robin over the ranks, no collective on the data path) with the
    CUDA solver on every rank, against the single-process oracle (ref loop: pnp_sci_algo.py:498-529)."""
    import socket
    import torch.multiprocessing as tmp
    from oracle import pnp_sci as O
    from scipnp import synth
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    tmp.spawn(_shard_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    meas, mask, orig = synth.make_cacti(64, 96, 8, 5, cfg=41)
    A, At = _ops(mask)
    ref = O.admmdenoise_cacti(meas, mask, A, At, projmeth='admm', orig=orig, nframe=5, MAXB=255.,
                              maskdirection='updown', _lambda=1, gamma=0.01, denoiser='tv', iter_max=8,
                              tv_weight=0.3, tv_iter_max=5)
    for r in range(world):
        assert float(np.abs(np.load(tmp_path / ("x_%d.npy" % r)) - ref[0]).max()) <= TOL_X
        assert float(np.abs(np.load(tmp_path / ("pa_%d.npy" % r)) - np.array(ref[4])).max()) <= TOL_DB
