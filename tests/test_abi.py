"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/scipnp.h declares, and refuses to compute without a GPU (no CPU
fallback).  No kernels are launched here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scipnp.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(scipnp_[a-z_0-9A-Z]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from scipnp import _lib
    decl = _declared()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(_lib.lib, name), "libscipnp.so lacks " + name
    # and the binding covers the header
    assert set(decl) == set(_lib.EXPORTS)


def test_exports_are_plain_c_symbols():
    from scipnp import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    syms = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for name in _declared():
        assert name in syms


def test_header_cites_reference_lines():
    src = open(HEADER).read()
    for ref in ("utils.py:10-15", "utils.py:17-26", "pnp_sci_algo.py:640-645",
                "pnp_sci_algo.py:808-809", "pnp_sci_algo.py:491-492", "utils.py:28-36"):
        assert ref in src


def test_params_struct_layout():
    from scipnp._lib import Params
    # int,int,float,float,double,double,int*8  (see scipnp_params in the header)
    assert C.sizeof(Params) == 4 * 4 + 2 * 8 + 8 * 4
    assert Params.tv_weight.offset == 16 and Params.B.offset == 40


def test_version_and_error_string():
    from scipnp._lib import lib
    assert lib.scipnp_version() >= 100
    assert isinstance(lib.scipnp_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    import scipnp
    from scipnp._lib import lib, Params
    assert lib.scipnp_device_count() == 0
    p = Params()
    p.B, p.H, p.W, p.C = 1, 8, 8, 4
    p.tv_weight, p.tv_iter_max = 0.1, 5
    h = C.c_void_p()
    assert lib.scipnp_solver_create(C.byref(p), C.byref(h)) == -2      # SCIPNP_ECUDA
    assert b"no CPU fallback" in lib.scipnp_last_error()
    y = np.zeros((8, 8), np.float32)
    m = np.ones((8, 8, 4), np.float32)
    with pytest.raises(scipnp.ScipnpError):
        scipnp.gap_denoise(y, m.sum(2), Phi=m, iter_max=1)
    with pytest.raises(scipnp.ScipnpError):
        scipnp.A_(m, m)
    # the host pipeline and the joint module's hand-off loops fail the same way
    pl = C.c_void_p()
    assert lib.scipnp_pipeline_create(C.byref(p), 2, C.byref(pl)) == -2
    with pytest.raises(scipnp.ScipnpError):
        scipnp.HostPipeline(1, 8, 8, 4)
    from scipnp import joint_pnp_sci_algo as J
    with pytest.raises(scipnp.ScipnpError):
        J.gap_multistep_denoise(y, m.sum(2), Phi=m, iter_max=1, sigma=0.1, second_denoiser=lambda x, s, mdl: x)


def test_argument_validation_happens_before_any_device_work():
    import scipnp
    y = np.zeros((8, 8), np.float32)
    m = np.ones((8, 8, 4), np.float32)
    with pytest.raises(ValueError):
        scipnp.gap_denoise(y, m.sum(2), Phi=m, denoiser='ffdnet', iter_max=1)
    with pytest.raises(ValueError):
        scipnp.admm_denoise(y, m.sum(2), Phi=m, denoiser='bm3d', iter_max=1)
    with pytest.raises(ValueError):
        scipnp.admmdenoise_cacti(y[..., None], m, projmeth='ista', denoiser='tv')
    from scipnp import joint_pnp_sci_algo as J
    with pytest.raises(ValueError):
        J.gap_multistep_denoise(y, m.sum(2), Phi=m, denoiser='tv', iter_max=1, second_denoiser=lambda x, s, mdl: x)
    with pytest.raises(ValueError):
        J.admm_multistep_denoise(y, m.sum(2), Phi=m, tvm='tv_bregman', iter_max=1, second_denoiser=lambda x, s, mdl: x)
    with pytest.raises(NotImplementedError):      # the learned denoisers are the caller's
        J.gap_multistep_denoise(y, m.sum(2), Phi=m, iter_max=1)
    with pytest.raises(ValueError):      # opaque operators that are not mask operators
        scipnp.gap_denoise(y, m.sum(2), A=lambda x: x.sum(2) * 2 + 1, At=lambda v: np.ones((8, 8, 4), np.float32),
                           iter_max=1)
