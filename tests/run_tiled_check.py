#!/usr/bin/env python
"""Multi-GPU check (run under torchrun on a multi-GPU box, one rank per GPU):
the row-tiled solve over all ranks equals the single-GPU solve of the same scene,
for both halo transports.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/run_tiled_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sci-algorithms_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from scipnp import Solver, synth
    from scipnp.tiled import TiledSolver
    H, W, C, iters = 64 * world + 8, 320, 24, 7
    meas, mask, _ = synth.make_cacti(H, W, C, 1, cfg=31)
    y = meas[:, :, 0] / np.float32(255.)
    with Solver(1, H, W, C, method="gap", tv_weight=0.3, tv_iter_max=5) as so:
        so.load(y[None], mask)
        so.run(iters)
        ref = so.get_x()[0]
    ok = True
    for transport in ("nccl", "p2p"):
        for k in (1, 2, 3):
            ts = TiledSolver(H, W, C, rank, world, tv_weight=0.3, tv_iter_max=5, exchange_every=k,
                             transport=transport)
            ts.load(torch.from_numpy(y[ts.row_lo:ts.row_hi]).cuda(), torch.from_numpy(mask[ts.row_lo:ts.row_hi]).cuda())
            ts.run(3)
            ts.run(iters - 3)             # two runs: halos must be fresh at a run boundary
            got = ts.owned().cpu().numpy()
            err = float(np.abs(got - ref[ts.lo:ts.hi]).max())
            t = torch.tensor([err], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print("transport=%s(%s) k=%d  max|tiled - single| = %.3g" % (transport, ts.transport, k, float(t[0])))
            ok = ok and float(t[0]) <= 1e-6 and ts.transport == transport
            ts.close()
            dist.barrier()
    if rank == 0:
        print("TILED CHECK", "PASS" if ok else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
