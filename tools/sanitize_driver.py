#!/usr/bin/env python
"""Tiny runs of the fused kernels for compute-sanitizer (tools/sanitize.sh): a few outer iterations of the
warp-specialised kernel (GAP accelerated / plain, standalone TV) and of the stream kernel (ADMM, CASSI shape) on
scenes small enough for racecheck, compared with the exact path so that a run that "passes" also computed something."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sci-algorithms_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import scipnp  # noqa: E402
from scipnp import synth  # noqa: E402
from scipnp.engine import Solver  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
# (method, B, H, W, C, accelerated): one scene; a batch of three whose groups are packed across the measurements; plain GAP
# at C = 24; ADMM with the multiplier staged by TMA (C = 8) and read from global memory (C = 24)
cases = [("gap", 1, 72, 128, 8, True), ("gap", 3, 40, 128, 8, True), ("gap", 1, 40, 64, 24, False),
         ("admm", 1, 48, 64, 8, True), ("admm", 1, 40, 64, 24, True)]
for method, B, H, W, C, acc in cases:
    if which not in ("all", method):
        continue
    meas, mask, _ = synth.make_cacti(H, W, C, B, cfg=5)
    y = np.ascontiguousarray(np.moveaxis(meas, 2, 0)) / np.float32(255.)
    out = []
    for fused in (True, False):
        with Solver(B, H, W, C, method=method, accelerate=acc, tv_weight=0.3, tv_iter_max=5, fused=fused) as s:
            s.load(y, mask)
            s.run(2)
            out.append(s.get_x())
    print("%s %dx%dx%dx%d acc=%s: max|fused - exact| = %.3g" % (method, B, H, W, C, acc, float(np.abs(out[0] - out[1]).max())))
if which in ("all", "tv"):
    f = torch.rand((48, 64, 8), device="cuda")
    a = scipnp.denoise_tv_chambolle(f, 0.3, n_iter_max=5, multichannel=True)
    print("tv 48x64x8: mean shift %.3g" % float((a.mean() - f.mean()).abs()))
torch.cuda.synchronize()
print("launches:", scipnp._lib.lib.scipnp_launch_count())
