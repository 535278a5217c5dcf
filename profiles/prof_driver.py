#!/usr/bin/env python
"""Small driver for ncu captures: a few outer GAP-TV iterations on the config-5
scene (3840x2160x24) so the profiler sees the kernels of the timed path only.
Usage: python profiles/prof_driver.py [iters] [H] [W] [C] [B] [gap|admm]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sci-algorithms_b200"))
import torch  # noqa: E402
from scipnp.engine import Solver  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
W = int(sys.argv[3]) if len(sys.argv) > 3 else 3840
C = int(sys.argv[4]) if len(sys.argv) > 4 else 24
B = int(sys.argv[5]) if len(sys.argv) > 5 else 1
method = sys.argv[6] if len(sys.argv) > 6 else "gap"
fused = int(os.environ.get("SCIPNP_FUSED", "1"))
g = torch.Generator(device="cuda").manual_seed(1)
Phi = (torch.rand((B, H, W, C), device="cuda", generator=g) <= 0.5).float()
yy = torch.arange(H, device="cuda", dtype=torch.float32)[None, :, None, None]
xx = torch.arange(W, device="cuda", dtype=torch.float32)[None, None, :, None]
tt = torch.arange(C, device="cuda", dtype=torch.float32)[None, None, None, :]
bb = torch.arange(B, device="cuda", dtype=torch.float32)[:, None, None, None]
# a smooth moving scene with an edge (noise alone makes skimage's stopping rule fire and the run roll back)
orig = 0.45 + 0.25 * torch.sin((xx + 3. * tt + 5. * bb) / 37.) * torch.cos(yy / 29.) + 0.2 * (((xx + 2. * tt) // 64 + yy // 48) % 2)
y = (Phi * orig).sum(3)
tv_eps = float(os.environ.get("TV_EPS", "2e-4"))      # TV_EPS=0: timing experiments whose results are not meaningful
s = Solver(B, H, W, C, method=method, tv_weight=0.3, tv_iter_max=5, fused=bool(fused), tv_eps=tv_eps, phi_batched=B > 1)
Pin = Phi if B > 1 else Phi[0].contiguous()
s.load(y, Pin, borrow_phi=bool(int(os.environ.get('BORROW', '0'))))
s.run(iters)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
s.run(iters)
e1.record()
torch.cuda.synchronize()
print("%dx%dx%dx%d %s path=%s refined=%d  %.4f ms / outer iteration" % (B, H, W, C, method, "fused" if s.uses_fused else "exact", s.refined_iters, e0.elapsed_time(e1) / iters))
