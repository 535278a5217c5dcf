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
cases = [("gap", 72, 128, 8, True), ("gap", 40, 64, 24, False), ("admm", 48, 64, 8, True)]
for method, H, W, C, acc in cases:
    if which not in ("all", method):
        continue
    meas, mask, _ = synth.make_cacti(H, W, C, 1, cfg=5)
    y = meas[:, :, 0] / np.float32(255.)
    out = []
    for fused in (True, False):
        with Solver(1, H, W, C, method=method, accelerate=acc, tv_weight=0.3, tv_iter_max=5, fused=fused) as s:
            s.load(y[None], mask)
            s.run(2)
            out.append(s.get_x()[0])
    print("%s %dx%dx%d acc=%s: max|fused - exact| = %.3g" % (method, H, W, C, acc, float(np.abs(out[0] - out[1]).max())))
if which in ("all", "tv"):
    f = torch.rand((48, 64, 8), device="cuda")
    a = scipnp.denoise_tv_chambolle(f, 0.3, n_iter_max=5, multichannel=True)
    print("tv 48x64x8: mean shift %.3g" % float((a.mean() - f.mean()).abs()))
torch.cuda.synchronize()
print("launches:", scipnp._lib.lib.scipnp_launch_count())
