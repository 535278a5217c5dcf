"""Host-side plumbing over the C ABI: device buffers, streams, the solver handle.

PyTorch is used only for device memory, pinned host memory and streams.  Every
function here ends in a call into libscipnp.so; nothing computes on the CPU.
"""
import ctypes as ct

import numpy as np
import torch

from ._lib import lib, check, Params, require_device, ScipnpError

__all__ = ["Solver", "HostPipeline", "make_params", "to_device", "stream_ptr", "is_torch", "dptr", "f32c"]

METHOD_GAP, METHOD_ADMM = 0, 1


def is_torch(a):
    return isinstance(a, torch.Tensor)


def f32c(a):
    """float32, C-contiguous NumPy view/copy of an array-like (the engine
    computes in single precision; float64 inputs of the pnp_sci_test_* drivers
    are narrowed here, once)."""
    return np.ascontiguousarray(a, dtype=np.float32)


def to_device(a, device=None):
    """array-like / tensor -> contiguous float32 CUDA tensor."""
    require_device()
    if is_torch(a):
        t = a
        if not t.is_cuda:
            t = t.to(device or "cuda", non_blocking=True)
        return t.contiguous().to(torch.float32)
    return torch.from_numpy(f32c(a)).to(device or "cuda")


def stream_ptr():
    return ct.c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t):
    """Raw pointer of a tensor / ndarray (or None)."""
    if t is None:
        return None
    if is_torch(t):
        return ct.c_void_p(t.data_ptr())
    return ct.c_void_p(t.ctypes.data)


def make_params(B, H, W, C, method="gap", accelerate=True, _lambda=1.0, gamma=0.01, tv_weight=0.1,
                tv_iter_max=5, tv_eps=2.e-4, phi_batched=False, fused=True, clip=False):
    """``scipnp_params`` for the given problem (argument names as in pnp_sci_algo.py:536-539)."""
    p = Params()
    p.method = METHOD_ADMM if str(method).lower() == "admm" else METHOD_GAP
    p.accelerate = 1 if accelerate else 0
    p.lambda_ = float(_lambda)
    p.gamma = float(gamma)
    p.tv_weight = float(tv_weight)
    p.tv_eps = float(tv_eps)
    p.tv_iter_max = int(tv_iter_max)
    p.fused = 1 if fused else 0
    p.B, p.H, p.W, p.C = int(B), int(H), int(W), int(C)
    p.phi_batched = 1 if phi_batched else 0
    p.clip01 = 1 if clip else 0
    return p


class HostPipeline:
    """Stream of reconstructions from host buffers (``scipnp_pipeline_*``): the frame loop of
    ``admmdenoise_cacti`` (pnp_sci_algo.py:498-529) with the copies of one reconstruction running
    under the kernels of the next.  ``submit`` returns a ticket at once, ``wait`` returns when the
    output array of that ticket is complete.  Inputs/outputs are host arrays: NumPy (pageable, the
    copies then serialise) or pinned CPU tensors (``torch.Tensor.pin_memory``)."""

    def __init__(self, B, H, W, C, depth=2, **kw):
        require_device()
        self.shape = (int(B), int(H), int(W), int(C))
        self.params = make_params(B, H, W, C, **kw)
        self.depth = int(depth)
        h = ct.c_void_p()
        check(lib.scipnp_pipeline_create(ct.byref(self.params), self.depth, ct.byref(h)))
        self._h = h
        self._keep = {}

    def close(self):
        if getattr(self, "_h", None):
            lib.scipnp_pipeline_destroy(self._h)
            self._h = None
            self._keep = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @staticmethod
    def _host(a, shape, name):
        if a is None:
            return None
        if is_torch(a):
            if a.is_cuda or a.dtype != torch.float32 or not a.is_contiguous():
                raise ValueError("%s must be a contiguous float32 host tensor" % name)
        else:
            a = f32c(a)
        if tuple(a.shape) != tuple(shape):
            raise ValueError("%s has shape %s, expected %s" % (name, tuple(a.shape), tuple(shape)))
        return a

    def submit(self, y, Phi, iters, out, x0=None, X_orig=None):
        """Enqueue one reconstruction; ``out`` ([B,H,W,C] host array) is filled when ``wait``
        returns for the ticket this call hands back."""
        B, H, W, Cc = self.shape
        pb = self.params.phi_batched
        y = self._host(y, (B, H, W), "y")
        Phi = self._host(Phi, (B, H, W, Cc) if pb else (H, W, Cc), "Phi")
        x0 = self._host(x0, self.shape, "x0")
        X_orig = self._host(X_orig, self.shape, "X_orig")
        if is_torch(out):
            if out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != self.shape:
                raise ValueError("out must be a contiguous float32 host tensor of shape %s" % (self.shape,))
        elif out.dtype != np.float32 or not out.flags.c_contiguous or tuple(out.shape) != self.shape:
            raise ValueError("out must be a C-contiguous float32 array of shape %s" % (self.shape,))
        t = ct.c_int(-1)
        check(lib.scipnp_pipeline_submit(self._h, dptr(y), dptr(Phi), dptr(x0), dptr(X_orig), int(iters),
                                         dptr(out), ct.byref(t)))
        self._keep[t.value] = (y, Phi, x0, X_orig, out, int(iters), X_orig is not None)
        return t.value

    def wait(self, ticket):
        """Block until the ticket's output is complete; returns (out, psnr_all or None)."""
        n = ct.c_int(0)
        if ticket not in self._keep:             # unknown / collected: let the library say so
            check(lib.scipnp_pipeline_wait(self._h, int(ticket), None, 0, ct.byref(n)))
        y, Phi, x0, X_orig, out, iters, has_orig = self._keep.pop(ticket)
        if has_orig:
            cap = iters * self.shape[0]
            buf = (ct.c_double * max(cap, 1))()
            check(lib.scipnp_pipeline_wait(self._h, int(ticket), buf, cap, ct.byref(n)))
            psnr = np.array(buf[:n.value], dtype=np.float64).reshape(-1, self.shape[0])
            return out, psnr
        check(lib.scipnp_pipeline_wait(self._h, int(ticket), None, 0, ct.byref(n)))
        return out, None

    @property
    def refined_iters(self):
        n = ct.c_int(0)
        check(lib.scipnp_pipeline_refined_iters(self._h, ct.byref(n)))
        return n.value


class Solver:
    """Persistent GAP-TV / ADMM-TV solver (``scipnp_solver_*``).

    All arrays are [B,H,W,C] / [B,H,W] float32; ``Phi`` is [H,W,C] (shared) or
    [B,H,W,C] (``phi_batched``).  Inputs may be NumPy arrays (copied H2D by the
    library) or CUDA tensors (copied D2D).
    """

    def __init__(self, B, H, W, C, method="gap", accelerate=True, _lambda=1.0, gamma=0.01,
                 tv_weight=0.1, tv_iter_max=5, tv_eps=2.e-4, phi_batched=False, fused=True, clip=False):
        require_device()
        self.shape = (int(B), int(H), int(W), int(C))
        self.method = METHOD_ADMM if str(method).lower() == "admm" else METHOD_GAP
        p = Params()
        p.method = self.method
        p.accelerate = 1 if accelerate else 0
        p.lambda_ = float(_lambda)
        p.gamma = float(gamma)
        p.tv_weight = float(tv_weight)
        p.tv_eps = float(tv_eps)
        p.tv_iter_max = int(tv_iter_max)
        p.fused = 1 if fused else 0
        p.B, p.H, p.W, p.C = self.shape
        p.phi_batched = 1 if phi_batched else 0
        p.clip01 = 1 if clip else 0
        self.params = p
        h = ct.c_void_p()
        check(lib.scipnp_solver_create(ct.byref(p), ct.byref(h)))
        self._h = h
        self._keep = []

    # -- lifecycle ----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            lib.scipnp_solver_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- data -----------------------------------------------------------------
    def _chk(self, a, shape, name):
        if a is None:
            return None
        if not is_torch(a):
            a = f32c(a)
        elif a.dtype != torch.float32 or not a.is_contiguous():
            a = a.contiguous().to(torch.float32)
        if tuple(a.shape) != tuple(shape):
            raise ValueError("%s has shape %s, expected %s" % (name, tuple(a.shape), tuple(shape)))
        self._keep.append(a)
        return a

    def load(self, y, Phi, Phi_sum=None, x0=None, X_orig=None, borrow_phi=False):
        """Inputs of one reconstruction (host arrays or device tensors).  ``borrow_phi=True`` with a device tensor:
        the mask stack is read in place by every iteration instead of being copied into the handle (it must not
        change until the results were read; the handle keeps a reference)."""
        B, H, W, Cc = self.shape
        pb = self.params.phi_batched
        self._keep = []
        y = self._chk(y, (B, H, W), "y")
        Phi = self._chk(Phi, (B, H, W, Cc) if pb else (H, W, Cc), "Phi")
        Phi_sum = self._chk(Phi_sum, (B, H, W) if pb else (H, W), "Phi_sum")
        x0 = self._chk(x0, (B, H, W, Cc), "x0")
        X_orig = self._chk(X_orig, (B, H, W, Cc), "X_orig")
        borrow = bool(borrow_phi) and is_torch(Phi) and Phi.is_cuda
        fn = lib.scipnp_solver_load_borrow_phi if borrow else lib.scipnp_solver_load
        check(fn(self._h, dptr(y), dptr(Phi), dptr(Phi_sum), dptr(x0), dptr(X_orig), stream_ptr()))
        self._phi_ref = Phi if borrow else None
        if any(not is_torch(a) for a in self._keep):
            torch.cuda.current_stream().synchronize()   # pageable host sources
        self._keep = []

    def load_cassi(self, y, mask2d, step, x0=None, X_orig=None):
        """Single-disperser CASSI (B = 1): ``mask2d`` [H, W-(C-1)*step] is the coded aperture, the
        solver's W the sheared canvas; the fused iterations read the aperture at per-band offsets."""
        B, H, W, Cc = self.shape
        self._keep = []
        y = self._chk(y, (B, H, W), "y")
        mask2d = self._chk(mask2d, (H, W - (Cc - 1) * int(step)), "mask2d")
        x0 = self._chk(x0, (B, H, W, Cc), "x0")
        X_orig = self._chk(X_orig, (B, H, W, Cc), "X_orig")
        check(lib.scipnp_solver_load_cassi(self._h, dptr(y), dptr(mask2d), int(step), dptr(x0),
                                           dptr(X_orig), stream_ptr()))
        if any(not is_torch(a) for a in self._keep):
            torch.cuda.current_stream().synchronize()
        self._keep = []

    def run(self, iters):
        check(lib.scipnp_solver_run(self._h, int(iters), stream_ptr()))

    # pieces of run() for callers that interleave their own work (see include/scipnp.h)
    def begin(self):
        check(lib.scipnp_solver_begin(self._h, stream_ptr()))

    def step_async(self, iters):
        check(lib.scipnp_solver_step_async(self._h, int(iters), stream_ptr()))

    def fired(self):
        f = ct.c_int(0)
        check(lib.scipnp_solver_fired(self._h, ct.byref(f), stream_ptr()))
        return bool(f.value)

    def rollback(self):
        check(lib.scipnp_solver_rollback(self._h, stream_ptr()))

    def set_path(self, fused):
        check(lib.scipnp_solver_set_path(self._h, 1 if fused else 0))

    def set_tv(self, tv_weight, gamma=0.0):
        """TV weight / ADMM regulariser of the iterations that follow (ADMM_TV_rec shrinks both per iteration)."""
        check(lib.scipnp_solver_set_tv(self._h, float(tv_weight), float(gamma)))

    def add_refined(self, iters):
        check(lib.scipnp_solver_add_refined(self._h, int(iters)))

    def get_x(self, out=None):
        """Current estimate as NumPy (default) or into a given tensor/array."""
        if out is None:
            out = np.empty(self.shape, dtype=np.float32)
        check(lib.scipnp_solver_get_x(self._h, dptr(out), stream_ptr()))
        return out

    def _per_iter(self, fn):
        n = ct.c_int(0)
        check(fn(self._h, None, 0, ct.byref(n), stream_ptr()))
        B = self.shape[0]
        if n.value == 0:
            return np.zeros((0, B))
        buf = (ct.c_double * n.value)()
        check(fn(self._h, buf, n.value, ct.byref(n), stream_ptr()))
        return np.array(buf[:n.value], dtype=np.float64).reshape(-1, B)

    def psnr_all(self):
        """utils.psnr per (iteration, batch element): array [iters, B]."""
        return self._per_iter(lib.scipnp_solver_psnr)

    def sqerr_all(self):
        """sum (x - X_orig)^2 per (iteration, batch element): array [iters, B]."""
        return self._per_iter(lib.scipnp_solver_sqerr)

    @property
    def refined_iters(self):
        n = ct.c_int(0)
        check(lib.scipnp_solver_refined_iters(self._h, ct.byref(n)))
        return n.value

    @property
    def uses_fused(self):
        return bool(lib.scipnp_solver_uses_fused(self._h))

    @property
    def launches(self):
        return int(lib.scipnp_solver_launch_count(self._h))

    def admm_state_ptrs(self):
        """(theta, b, x) device pointers of an ADMM handle."""
        t, b, x = ct.c_void_p(), ct.c_void_p(), ct.c_void_p()
        check(lib.scipnp_solver_admm_state(self._h, ct.byref(t), ct.byref(b), ct.byref(x)))
        return t.value, b.value, x.value

    def state_ptrs(self):
        a, b = ct.c_void_p(), ct.c_void_p()
        check(lib.scipnp_solver_state(self._h, ct.byref(a), ct.byref(b)))
        return a.value, b.value
