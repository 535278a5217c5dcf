"""Row-tiled GAP-TV over several GPUs (one process per GPU, ``torch.distributed``).

A single scene [H, W, C] is split into contiguous row blocks, one per rank.
Each rank keeps ``halo = k*(tv_iter_max-1)`` extra rows above and below its
block and runs the ordinary single-GPU solver on its local rows.  One outer
iteration moves information by at most ``tv_iter_max-1`` rows (the TV stencil;
the projection is pointwise), so after ``k`` local iterations exactly the halo
rows are stale while the owned rows are still identical to the single-GPU
result.  Every ``k`` iterations the ranks then refresh their halo rows of ``x``
and ``y1`` from the neighbours that own them (ring-neighbour send/recv, no
global collective).  Chambolle's boundary rules apply at the true image edges
only; a tile seam is never treated as an edge for owned rows.

The reference has no counterpart (it is single-process); this is SURVEY.md
section 8(e) "single UHD scene".
"""
import numpy as np
import torch
import torch.distributed as dist

__all__ = ["partition_rows", "exchange_halos", "TiledSolver", "tiled_reference_run", "gap_denoise_tiled",
           "stop_rule_hits"]


def stop_rule_hits(energy, eps):
    """skimage's stopping test of ``denoise_tv_chambolle`` on logged energies ``[..., n_dual]`` (one row per outer
    iteration and channel slice): true where ``|E_{i-1} - E_i| < eps * E_0`` for some dual iteration ``i >= 1`` -- a
    stop at the last executed iteration changes nothing, so the last column is only compared as ``E_i``."""
    e = energy
    return ((e[..., :-1] - e[..., 1:]).abs() < eps * e[..., 0:1]).any(dim=-1)


def partition_rows(H, world, rank, halo):
    """Rows owned by ``rank`` ([lo, hi)) and rows held locally ([row_lo, row_hi))."""
    base, rem = divmod(H, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi, max(0, lo - halo), min(H, hi + halo)


def _plan(H, world, halo):
    return [partition_rows(H, world, r, halo) for r in range(world)]


def exchange_halos(fields, H, halo, rank, world, group=None):
    """Refresh the halo rows of every tensor in ``fields`` (axis 0 = local rows of this rank).

    Every rank derives all partitions from (H, world, halo), so no metadata travels.  Halos
    may span more than one neighbour when blocks are shorter than ``halo``: the loop below
    sends each remote rank exactly the intersection of its halo with this rank's owned rows."""
    if world == 1:
        return
    plan = _plan(H, world, halo)
    lo, hi, row_lo, row_hi = plan[rank]
    ops, bufs = [], []
    # gloo moves host memory: stage device rows through the host there (tests); NCCL sends
    # device rows directly over NVLink
    via_host = dist.get_backend(group) == "gloo" and any(f.is_cuda for f in fields)
    for other in range(world):
        if other == rank:
            continue
        olo, ohi, orow_lo, orow_hi = plan[other]
        # rows I own that `other` holds as halo
        s0, s1 = max(lo, orow_lo), min(hi, orow_hi)
        # rows `other` owns that I hold as halo
        r0, r1 = max(olo, row_lo), min(ohi, row_hi)
        for f in fields:
            if s1 > s0:
                t = f[s0 - row_lo:s1 - row_lo].contiguous()
                if via_host:
                    t = t.cpu()
                bufs.append(t)
                ops.append(dist.P2POp(dist.isend, t, other, group))
            if r1 > r0:
                t = torch.empty_like(f[r0 - row_lo:r1 - row_lo], device="cpu" if via_host else f.device)
                bufs.append((t, f, r0 - row_lo, r1 - row_lo))
                ops.append(dist.P2POp(dist.irecv, t, other, group))
    if not ops:
        return
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for b in bufs:
        if isinstance(b, tuple):
            t, f, a, z = b
            f[a:z].copy_(t)


class _DevView:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3, "strides": None}


def _wrap(ptr, shape, device, typestr="<f4"):
    return torch.as_tensor(_DevView(ptr, shape, typestr), device=device)


class TiledSolver:
    """One rank of the row-tiled GAP-TV / ADMM-TV solver (TV prior)."""

    def __init__(self, H, W, C, rank, world, tv_weight=0.1, tv_iter_max=5, _lambda=1.0,
                 accelerate=True, exchange_every=1, group=None, fused=True, transport="nccl",
                 tv_eps=2.e-4, method="gap", gamma=0.01):
        """transport: "nccl" (send/recv pairs through torch.distributed), "p2p" (CUDA-IPC
        mapped neighbour buffers, halo rows pulled over NVLink by the library, device-side
        flags) or "auto" (p2p when every rank can set it up, else nccl)."""
        from .engine import Solver
        self.method = str(method).lower()
        if self.method not in ("gap", "admm"):
            raise ValueError("method must be 'gap' or 'admm'")
        self.H, self.W, self.C = H, W, C
        self.rank, self.world, self.group = rank, world, group
        self.k = max(1, int(exchange_every))
        self.halo = self.k * (int(tv_iter_max) - 1)
        self.lo, self.hi, self.row_lo, self.row_hi = partition_rows(H, world, rank, self.halo)
        self.local_rows = self.row_hi - self.row_lo
        self.accelerate = accelerate
        self.solver = Solver(1, self.local_rows, W, C, method=self.method, accelerate=accelerate,
                             _lambda=_lambda, gamma=gamma, tv_weight=tv_weight, tv_iter_max=tv_iter_max,
                             tv_eps=tv_eps, fused=fused)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.exchanges = 0
        self.transport = "nccl"
        self.push = False          # halo push inside the fused kernel (p2p transport, exchange_every = 1)
        self.tv_eps = float(tv_eps)
        self.R = int(tv_iter_max) - 1
        if world > 1:
            self._setup_energy_reduce()
        if transport in ("p2p", "auto") and world > 1:
            self._setup_p2p(required=(transport == "p2p"))

    def _setup_energy_reduce(self):
        """skimage's stopping rule sums energies over the whole image: on the exact path the library hands the
        per-slice partial sums over this rank's owned rows to this callback, which all-reduces them in place
        (stream-ordered with NCCL; through the host with gloo)."""
        import ctypes as ct
        from ._lib import lib, check
        check(lib.scipnp_solver_tiling(self.solver._h, self.lo, self.hi, self.row_lo, self.row_hi))

        def reduce(dev_ptr, n, stream, user):
            try:
                t = _wrap(dev_ptr, (n,), self.device, "<f8")
                dist.all_reduce(t, group=self.group)
                return 0
            except Exception:       # noqa: BLE001  (must not propagate through the C frame)
                return 1
        self._reduce_cb = ct.CFUNCTYPE(ct.c_int, ct.c_void_p, ct.c_int, ct.c_void_p, ct.c_void_p)(reduce)
        check(lib.scipnp_solver_set_energy_reduce(self.solver._h, ct.cast(self._reduce_cb, ct.c_void_p), None, self.H))

    def _setup_p2p(self, required):
        """Map the neighbours' solver buffers (CUDA IPC) so that halo rows are pulled straight
        over NVLink by the library, with device-side flags instead of NCCL calls.  Needs every
        halo to lie inside the direct neighbour's owned rows."""
        import ctypes as ct
        from ._lib import lib, check
        plan = _plan(self.H, self.world, self.halo)
        ok = all(p[1] - p[0] >= self.halo for p in plan)
        err = None
        blob = b""
        try:
            if ok:
                check(lib.scipnp_solver_tiling(self.solver._h, self.lo, self.hi, self.row_lo, self.row_hi))
                n = lib.scipnp_solver_ipc_blob_bytes()
                buf = ct.create_string_buffer(n)
                check(lib.scipnp_solver_ipc_export(self.solver._h, buf))
                blob = buf.raw
        except Exception as e:      # noqa: BLE001
            err = e
        blobs = [None] * self.world
        dist.all_gather_object(blobs, (blob, self.row_lo, err is None and ok), group=self.group)
        good = all(b[2] for b in blobs)
        if good:
            try:
                for side, other in ((0, self.rank - 1), (1, self.rank + 1)):
                    if 0 <= other < self.world:
                        check(lib.scipnp_solver_ipc_attach(self.solver._h, side, blobs[other][0], blobs[other][1]))
            except Exception as e:      # noqa: BLE001
                err, good = e, False
        flag = torch.tensor([1 if good else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()):
            self.transport = "p2p"
        elif required:
            raise RuntimeError("peer-to-peer halo transport unavailable: %s" % (err or "halo wider than a block"))
        if self.transport == "p2p" and self.k == 1 and self.solver.uses_fused:
            # one exchange per iteration: let the fused kernel push its seam rows itself (no exchange kernel)
            rows = [p[3] - p[2] for p in plan]
            up = rows[self.rank - 1] if self.rank > 0 else 0
            dn = rows[self.rank + 1] if self.rank + 1 < self.world else 0
            ok = 1 if lib.scipnp_solver_enable_push(self.solver._h, up, dn) == 0 else 0
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            self.push = bool(int(flag.item()))
            if ok and not self.push:
                raise RuntimeError("halo push could be set up on some ranks only")

    # -- data ----------------------------------------------------------------------------
    def load(self, y_local, Phi_local, x0_local=None, X_orig_local=None, borrow_phi=False):
        """Inputs restricted to rows [row_lo, row_hi) (device tensors or host arrays); ``borrow_phi`` as in
        ``Solver.load``."""
        self.solver.load(y_local[None], Phi_local,
                         x0=None if x0_local is None else x0_local[None],
                         X_orig=None if X_orig_local is None else X_orig_local[None], borrow_phi=borrow_phi)

    def _fields(self):
        """The carried arrays of the iteration as tensors over the solver's buffers: x and y1 (accelerated GAP), or
        theta and the multiplier b (ADMM)."""
        xp, y1p = self.solver.state_ptrs()
        f = [_wrap(xp, (self.local_rows, self.W, self.C), self.device)]
        if self.method == "admm":
            f.append(_wrap(y1p, (self.local_rows, self.W, self.C), self.device))
        elif self.accelerate:
            f.append(_wrap(y1p, (self.local_rows, self.W), self.device))
        return f

    def _result_field(self):
        """What the reference returns: x for GAP, the projection output x (not theta) for ADMM
        (pnp_sci_algo.py:840,864)."""
        if self.method != "admm":
            return self._fields()[0]
        return _wrap(self.solver.admm_state_ptrs()[2], (self.local_rows, self.W, self.C), self.device)

    def _sweep(self, iters):
        if self.transport == "p2p":
            from ._lib import lib, check
            from .engine import stream_ptr
            check(lib.scipnp_solver_run_tiled(self.solver._h, int(iters), self.k, stream_ptr()))
            self.exchanges += (iters + self.k - 1) // self.k
            return
        done = 0
        while done < iters:
            n = min(self.k, iters - done)
            self.solver.step_async(n)
            done += n
            # always refresh, so that a following run() starts from exact halos
            exchange_halos(self._fields(), self.H, self.halo, self.rank, self.world, self.group)
            self.exchanges += 1

    def run(self, iters):
        """`iters` outer iterations, everything enqueued asynchronously; one host sync at the
        end to agree (over all ranks) on whether the TV early stop fired anywhere, in which
        case every rank rolls back and redoes the run on the exact path."""
        fused = self.solver.uses_fused
        if fused:
            self.solver.begin()
        self._sweep(iters)
        if self.transport == "p2p" and not (self.push and fused):
            self._check_sync()
        if not fused:
            return
        if self.push:
            fired = self._stop_rule_whole_scene(iters)
        else:
            flag = torch.tensor([1 if self.solver.fired() else 0], dtype=torch.int32, device=self.device)
            if self.world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            fired = bool(int(flag.item()))
        if fired:
            self.solver.rollback()
            self.solver.set_path(False)
            try:
                self._sweep(iters)
            finally:
                self.solver.set_path(True)
            self.solver.add_refined(iters)

    def _stop_rule_whole_scene(self, iters):
        """skimage's eps test (pnp_sci_algo.py:650 -> denoise_tv_chambolle) on the energies of the WHOLE scene: the
        fused kernel logs the energies of every iteration over this rank's owned rows; one all-reduce per run sums
        them over the ranks, and every rank takes the same decision.  The neighbours' time-out flag rides along, so
        a run costs one collective and one host synchronisation."""
        import ctypes as ct
        from ._lib import lib, check
        dev, n, per, tflag = ct.c_void_p(), ct.c_int(0), ct.c_int(0), ct.c_void_p()
        check(lib.scipnp_solver_energy_log(self.solver._h, ct.byref(dev), ct.byref(n), ct.byref(per)))
        check(lib.scipnp_solver_sync_flag(self.solver._h, ct.byref(tflag)))
        have = n.value > 0 and self.R > 1
        parts = []
        if have:
            parts.append(_wrap(dev.value, (n.value * per.value,), self.device, "<f8"))
        tail = torch.zeros(2, dtype=torch.float64, device=self.device)
        if tflag.value:
            tail[1] = _wrap(tflag.value, (1,), self.device, "<i4")[0]
        parts.append(tail)
        buf = torch.cat(parts)
        if n.value < iters:              # log capacity exceeded: the per-tile side check covers the rest
            buf[-2] = 1.0 if self.solver.fired() else 0.0
        if self.world > 1:
            dist.all_reduce(buf, group=self.group)
        res = buf[-2:]
        if have:
            e = buf[:-2].reshape(n.value, per.value // self.R, self.R)
            hit = stop_rule_hits(e, self.tv_eps).any().to(torch.float64).reshape(1)
            res = torch.cat([res, hit])
        res = res.tolist()               # the one host synchronisation of the run
        if res[1] > 0:
            raise RuntimeError("rank %d: a neighbour never signalled its halo rows (timeout)" % self.rank)
        return res[0] > 0 or (have and res[2] > 0)

    def _check_sync(self):
        import ctypes as ct
        from ._lib import lib, check
        from .engine import stream_ptr
        t = ct.c_int(0)
        check(lib.scipnp_solver_sync_error(self.solver._h, ct.byref(t), stream_ptr()))
        if t.value:
            raise RuntimeError("rank %d: a neighbour never signalled its halo rows (timeout)" % self.rank)

    def owned(self, out=None):
        """Owned rows of the current estimate as a device tensor [hi-lo, W, C]."""
        x = self._result_field()
        res = x[self.lo - self.row_lo:self.hi - self.row_lo]
        if out is not None:
            out.copy_(res)
            return out
        return res.clone()

    def result(self):
        """Owned rows of the current estimate (a copy; the solver's buffers ping-pong)."""
        return self.owned()

    @property
    def uses_fused(self):
        return self.solver.uses_fused

    @property
    def refined_iters(self):
        return self.solver.refined_iters

    def close(self):
        self.solver.close()

    # -- host-buffer path: a stream of reconstructions, copies under the kernels -------------------------
    def run_host_stream(self, jobs, iters):
        """Reconstruct a sequence of scenes whose row blocks live in pinned host memory.  ``jobs`` is a list of
        ``(y_host, Phi_host, out_host)`` (this rank's rows [row_lo, row_hi) of y and Phi, its owned rows of the
        result), all pinned.  The H2D copy of job i+1 and the D2H copy of job i-1 run on a copy stream under the
        iterations of job i (two device staging sets, two result buffers); nothing is allocated per job.
        Returns when every result has landed in its ``out_host``."""
        dev = self.device
        main = torch.cuda.current_stream()
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = [(torch.empty((self.local_rows, self.W), dtype=torch.float32, device=dev),
                            torch.empty((self.local_rows, self.W, self.C), dtype=torch.float32, device=dev))
                           for _ in range(2)]
            self._res = [torch.empty((self.hi - self.lo, self.W, self.C), dtype=torch.float32, device=dev)
                         for _ in range(2)]
        cs = self._copy_stream
        up = [None, None]          # event: staging set i holds its job
        taken = [None, None]       # event: the solver has copied staging set i into its own buffers
        solved = [None, None]      # event: result buffer i holds a finished job
        drained = [None, None]     # event: result buffer i has left for the host

        def upload(i):
            k = i % 2
            with torch.cuda.stream(cs):
                if taken[k] is not None:
                    cs.wait_event(taken[k])
                self._stage[k][0].copy_(jobs[i][0], non_blocking=True)
                self._stage[k][1].copy_(jobs[i][1], non_blocking=True)
                up[k] = cs.record_event()

        if jobs:
            upload(0)
        for i in range(len(jobs)):
            k = i % 2
            if i + 1 < len(jobs):
                upload(i + 1)
            main.wait_event(up[k])
            self.load(self._stage[k][0], self._stage[k][1])
            taken[k] = main.record_event()
            self.run(iters)
            if drained[k] is not None:
                main.wait_event(drained[k])
            self.owned(out=self._res[k])
            solved[k] = main.record_event()
            with torch.cuda.stream(cs):
                cs.wait_event(solved[k])
                jobs[i][2].copy_(self._res[k], non_blocking=True)
                drained[k] = cs.record_event()
        cs.synchronize()
        main.synchronize()

    # -- end-to-end measurement used by bench.py -------------------------------------------
    def e2e_measure(self, y_local, Phi_local, iters, steps):
        """Host-buffer path at N GPUs: every step copies this rank's rows of y and Phi up from pinned host memory
        and its owned rows of the result down; the copies of neighbouring steps overlap the iterations
        (``run_host_stream``)."""
        import time
        yh = y_local.cpu().pin_memory()
        Ph = Phi_local.cpu().pin_memory()
        outs = [torch.empty((self.hi - self.lo, self.W, self.C), dtype=torch.float32).pin_memory() for _ in range(2)]
        steps = max(1, steps)
        self.run_host_stream([(yh, Ph, outs[0])], iters)                     # warm-up (allocates the staging sets)
        if self.world > 1:
            dist.barrier(self.group)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        self.run_host_stream([(yh, Ph, outs[i % 2]) for i in range(steps)], iters)
        if self.world > 1:
            dist.barrier(self.group)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX, group=self.group)
        h2d = torch.tensor([yh.numel() * 4 + Ph.numel() * 4], dtype=torch.float64, device="cuda")
        d2h = torch.tensor([outs[0].numel() * 4], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(h2d, group=self.group)
            dist.all_reduce(d2h, group=self.group)
        assert float(outs[(steps - 1) % 2].abs().sum()) > 0
        return {"value": steps * iters / float(dt[0]), "unit": "it/s",
                "h2d_bytes_per_step": int(h2d[0]), "d2h_bytes_per_step": int(d2h[0]), "steps": steps,
                "api": "scipnp.tiled.TiledSolver.run_host_stream (pinned host buffers per rank, copies on a second "
                       "stream under the iterations of the neighbouring steps)"}


def tiled_reference_run(step_fn, fields, H, halo, k, iters, rank, world, group=None):
    """Backend-agnostic driver of the tiling protocol, used by the CPU (gloo) tests: ``step_fn``
    advances the local state by one outer iteration in place; halos are refreshed every k."""
    done = 0
    while done < iters:
        n = min(k, iters - done)
        for _ in range(n):
            step_fn()
        done += n
        exchange_halos(fields(), H, halo, rank, world, group)


def gap_denoise_tiled(y, Phi_sum=None, A=None, At=None, _lambda=1, accelerate=True, denoiser='tv', iter_max=50,
                      noise_estimate=False, sigma=None, tv_weight=0.1, tv_iter_max=5, multichannel=True, x0=None,
                      X_orig=None, model=None, show_iqa=True, tvm='tv_chambolle', Phi=None, group=None,
                      exchange_every=1, transport="auto"):
    """``gap_denoise`` (pnp_sci_algo.py:536-706) for ONE scene over all ranks of ``group``: the one-call entry of the
    row-tiled mode.  Every rank calls it with the same whole-scene arguments (host arrays, the reference's
    signature plus ``Phi=``); the function slices this rank's rows, runs the tiled solver and gathers the owned
    rows, so every rank returns the complete ``(x, psnr_, ssim_, psnr_all)``.  ``psnr_all`` (the per-iteration
    track) is not kept in tiled mode and comes back empty; ``Phi_sum`` is recomputed on the device."""
    from .pnp_sci_algo import _check_tv, _recover_phi, _host, _total_iters
    from .engine import f32c
    from .iqa import frames_iqa
    _check_tv(denoiser, tvm, multichannel)
    Phi = _recover_phi(A, At, y, Phi)
    yh = f32c(_host(y))
    H, W, Cc = Phi.shape
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    ts = TiledSolver(H, W, Cc, rank, world, tv_weight=tv_weight, tv_iter_max=tv_iter_max, _lambda=_lambda,
                     accelerate=accelerate, exchange_every=exchange_every, group=group, transport=transport)
    try:
        a, b = ts.row_lo, ts.row_hi
        dev = ts.device
        ts.load(torch.from_numpy(yh[a:b]).to(dev), torch.from_numpy(Phi[a:b]).to(dev),
                None if x0 is None else torch.from_numpy(f32c(_host(x0))[a:b]).to(dev))
        ts.run(_total_iters(sigma, iter_max))
        mine = ts.owned().cpu().numpy()
        lo = ts.lo
    finally:
        ts.close()
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (lo, mine), group=group)
        x = np.concatenate([p[1] for p in sorted(parts, key=lambda t: t[0])], axis=0)
    else:
        x = mine
    Xo = None if X_orig is None else f32c(_host(X_orig))
    ps, ss = frames_iqa(Xo, x)
    return x, ps, ss, []
