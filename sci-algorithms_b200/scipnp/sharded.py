"""Independent measurements over several GPUs (one process per GPU, ``torch.distributed``).

The coded frames of ``admmdenoise_cacti`` (the ``nframe`` loop, pnp_sci_algo.py:498) and whole
scenes of a benchmark set are independent problems: rank r reconstructs the frames
``r, r+world, r+2*world, ...`` on its own GPU as one batched solve and the results are gathered.
There is no collective on the data path; the only communication is the final gather of the
reconstructions ("independent measurements").
"""
import time

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_indices", "gather_frames", "admmdenoise_cacti_sharded"]


def shard_indices(n, world, rank):
    """Frame indices of ``rank``: round-robin, so that uneven counts differ by at most one."""
    return list(range(rank, n, world))


def gather_frames(local, idx, n, group=None):
    """All ranks contribute ``local[j]`` = result of frame ``idx[j]``; every rank gets the list of all
    ``n`` results in frame order.  Works with gloo (CPU tests) and NCCL."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        out = [None] * n
        for j, k in enumerate(idx):
            out[k] = local[j]
        return out
    parts = [None] * world
    dist.all_gather_object(parts, (list(idx), list(local)), group=group)
    out = [None] * n
    for ids, vals in parts:
        for k, v in zip(ids, vals):
            out[k] = v
    return out


def admmdenoise_cacti_sharded(meas, mask, A=None, At=None, projmeth='admm', v0=None, orig=None,
                              iframe=0, nframe=1, MAXB=1., maskdirection='plain', group=None,
                              solve=None, **args):
    """``admmdenoise_cacti`` with the coded frames spread over the ranks of ``group``.

    Same arguments and return tuple as ``pnp_sci_algo.admmdenoise_cacti``; every rank returns the
    complete result.  ``solve`` (tests only) replaces the local solver; it must have the
    signature of ``admmdenoise_cacti``."""
    if solve is None:
        from .pnp_sci_algo import admmdenoise_cacti as solve
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    nmask = mask.shape[-1]
    t0 = time.time()
    mine = shard_indices(nframe, world, rank)
    local = []
    for kf in mine:
        # one frame at a time keeps `iframe`-dependent behaviour (mask direction flips) exact;
        # the frames of a rank still share one device and run back to back
        v0_k = None if v0 is None else v0[:, :, kf * nmask:(kf + 1) * nmask]
        x_k, _, ps, ss, pa = solve(meas, mask, A, At, projmeth=projmeth, v0=v0_k, orig=orig,
                                   iframe=iframe + kf, nframe=1, MAXB=MAXB,
                                   maskdirection=maskdirection, **args)
        local.append((np.asarray(x_k), list(ps), list(ss), list(pa[0]) if pa else []))
    allr = gather_frames(local, mine, nframe, group)
    H, W = mask.shape[:2]
    x_ = np.zeros((H, W, nmask * nframe), dtype=np.float32)
    psnr_, ssim_, psnrall_ = [], [], []
    for kf, (x_k, ps, ss, pa) in enumerate(allr):
        x_[..., kf * nmask:(kf + 1) * nmask] = x_k
        psnr_.extend(ps)
        ssim_.extend(ss)
        psnrall_.append(pa)
    return x_, time.time() - t0, psnr_, ssim_, psnrall_
