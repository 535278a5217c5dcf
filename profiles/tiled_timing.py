"""torchrun --nproc-per-node N profiles/tiled_timing.py : per-iteration time (max over ranks) of the
tiled sweep scipnp_solver_run_tiled on the config-5 scene for several refresh periods k (KS=1,2,4) and
without any refresh; no per-reconstruction set-up in the timed region.  Source of the N=2 / N=8 sweep
figures in DESIGN.md section 5."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # profiles/ -> repo root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sci-algorithms_b200"))
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
from scipnp.tiled import TiledSolver
H, W, C = 2160, 3840, 24
def measure(k, iters=40, exchange=True):
    s = TiledSolver(H, W, C, rank, world, tv_weight=0.3, tv_iter_max=5, exchange_every=k, transport="p2p")
    rows = s.row_hi - s.row_lo
    g = torch.Generator(device="cuda").manual_seed(1 + rank)
    Phi = (torch.rand((rows, W, C), device="cuda", generator=g) <= 0.5).float()
    y = (Phi * torch.rand((rows, W, C), device="cuda", generator=g)).sum(2)
    s.load(y, Phi)
    del Phi
    def sweep():
        if exchange: s._sweep(iters)
        else: s.solver.step_async(iters)
    sweep(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sweep(); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("N=%d k=%d rows/rank=%d exchange=%s: %.4f ms/iteration (max over ranks)" % (world, k, rows, exchange, t.item()), flush=True)
    s.close(); del s; torch.cuda.empty_cache()
for k in [int(v) for v in os.environ.get('KS', '1,2,4').split(',')]:
    measure(k)
measure(2, exchange=False)
dist.destroy_process_group()
