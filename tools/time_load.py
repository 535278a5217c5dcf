"""Time of Solver.load on the config-5 scene (device-resident inputs, borrowed masks): y copy + one-pass init."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sci-algorithms_b200"))
import torch
from scipnp.engine import Solver
H, W, C = 2160, 3840, 24
Phi = (torch.rand((H, W, C), device="cuda") <= 0.5).float()
y = torch.rand((1, H, W), device="cuda")
with Solver(1, H, W, C, method="gap", tv_weight=0.3, tv_iter_max=5) as s:
    for borrow in (True, False):
        s.load(y, Phi, borrow_phi=borrow); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): s.load(y, Phi, borrow_phi=borrow)
        b.record(); torch.cuda.synchronize()
        print("load (borrow=%s): %.3f ms" % (borrow, a.elapsed_time(b) / 5))
