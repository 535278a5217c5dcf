"""ctypes binding of libscipnp.so (the C ABI declared in include/scipnp.h).

The library is built in-tree by ``sci-algorithms_b200/build.sh`` (or
``__graft_entry__.build()``).  There is no Python or CPU fallback: if the shared
object is missing the import fails, and every compute entry returns an error
without a CUDA device.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCIPNP_LIB") or os.path.join(_HERE, "libscipnp.so")   # SCIPNP_LIB: experiment builds
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "include", "scipnp.h"))


class ScipnpError(RuntimeError):
    pass


class Params(C.Structure):
    """Mirror of ``scipnp_params`` (include/scipnp.h)."""
    _fields_ = [
        ("method", C.c_int), ("accelerate", C.c_int),
        ("lambda_", C.c_float), ("gamma", C.c_float),
        ("tv_weight", C.c_double), ("tv_eps", C.c_double),
        ("tv_iter_max", C.c_int), ("fused", C.c_int),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int),
        ("phi_batched", C.c_int), ("clip01", C.c_int),
    ]


if not os.path.isfile(LIB_PATH):
    raise ImportError(
        "libscipnp.so not found at %s -- build it with sci-algorithms_b200/build.sh "
        "(scipnp has no CPU fallback)" % LIB_PATH)

lib = C.CDLL(LIB_PATH)

_fp = C.c_void_p      # float* / device or host pointer
_vp = C.c_void_p
_i = C.c_int

_SIGS = {
    "scipnp_version": (C.c_int, []),
    "scipnp_last_error": (C.c_char_p, []),
    "scipnp_device_count": (C.c_int, []),
    "scipnp_launch_count": (C.c_longlong, []),
    "scipnp_A": (C.c_int, [_fp, _fp, _fp, _i, _i, _i, _i, _i, _vp]),
    "scipnp_At": (C.c_int, [_fp, _fp, _fp, _i, _i, _i, _i, _i, _vp]),
    "scipnp_phi_sum": (C.c_int, [_fp, _fp, _i, _i, _i, _i, _vp]),
    "scipnp_gap_project": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_float, _i,
                                     _i, _i, _i, _i, _i, _vp]),
    "scipnp_admm_project": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_float, C.c_float,
                                      _i, _i, _i, _i, _i, _vp]),
    "scipnp_admm_dual_update": (C.c_int, [_fp, _fp, _fp, C.c_size_t, _vp]),
    "scipnp_tv_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i]),
    "scipnp_tv_chambolle": (C.c_int, [_fp, _fp, C.c_double, C.c_double, _i, _i, _i, _i, _i,
                                      _vp, C.c_size_t, _vp, _vp, _i, _vp]),
    "scipnp_tv_fused_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i, _i]),
    "scipnp_tv_fused_supported": (C.c_int, [_i, _i, _i, _i, _i]),
    "scipnp_tv_chambolle_fused": (C.c_int, [_fp, _fp, C.c_double, C.c_double, _i, _i, _i, _i, _i,
                                            _vp, C.c_size_t, _vp, _vp]),
    "scipnp_tv_matlab_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i]),
    "scipnp_tv_matlab": (C.c_int, [_fp, _fp, _i, C.c_float, _i, _i, _i, _i, _i, _vp, C.c_size_t, _vp]),
    "scipnp_tv_atv_clip_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i]),
    "scipnp_tv_atv_clip": (C.c_int, [_fp, _fp, C.c_float, _i, _i, _i, _i, _i, _vp, C.c_size_t, _vp]),
    "scipnp_sq_err": (C.c_int, [_fp, _fp, C.c_size_t, _vp, _vp]),
    "scipnp_frames_iqa": (C.c_int, [_fp, _fp, _i, _i, _i, _vp, _vp, _vp]),
    "scipnp_gap_tv_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i, _i]),
    "scipnp_gap_tv_fused": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_float, _i,
                                      C.c_double, C.c_double, _i, _i, _i, _i, _i, _i,
                                      _vp, C.c_size_t, _vp, _vp]),
    "scipnp_set_fused_variant": (C.c_int, [_i]),
    "scipnp_bayer_split": (C.c_int, [_fp, _fp, _i, _i, _i, _vp]),
    "scipnp_bayer_merge": (C.c_int, [_fp, _fp, _i, _i, _i, _vp]),
    "scipnp_cassi_shift_mask": (C.c_int, [_fp, _fp, _i, _i, _i, _i, _vp]),
    "scipnp_solver_create": (C.c_int, [C.POINTER(Params), C.POINTER(_vp)]),
    "scipnp_solver_destroy": (C.c_int, [_vp]),
    "scipnp_solver_load": (C.c_int, [_vp, _fp, _fp, _fp, _fp, _fp, _vp]),
    "scipnp_solver_load_borrow_phi": (C.c_int, [_vp, _fp, _fp, _fp, _fp, _fp, _vp]),
    "scipnp_solver_load_cassi": (C.c_int, [_vp, _fp, _fp, _i, _fp, _fp, _vp]),
    "scipnp_solver_run": (C.c_int, [_vp, _i, _vp]),
    "scipnp_solver_begin": (C.c_int, [_vp, _vp]),
    "scipnp_solver_step_async": (C.c_int, [_vp, _i, _vp]),
    "scipnp_solver_fired": (C.c_int, [_vp, C.POINTER(_i), _vp]),
    "scipnp_solver_rollback": (C.c_int, [_vp, _vp]),
    "scipnp_solver_set_path": (C.c_int, [_vp, _i]),
    "scipnp_solver_set_tv": (C.c_int, [_vp, C.c_double, C.c_double]),
    "scipnp_solver_add_refined": (C.c_int, [_vp, _i]),
    "scipnp_solver_ipc_blob_bytes": (C.c_int, []),
    "scipnp_solver_tiling": (C.c_int, [_vp, _i, _i, _i, _i]),
    "scipnp_solver_ipc_export": (C.c_int, [_vp, C.c_char_p]),
    "scipnp_solver_ipc_attach": (C.c_int, [_vp, _i, C.c_char_p, _i]),
    "scipnp_solver_set_energy_reduce": (C.c_int, [_vp, _vp, _vp, C.c_longlong]),
    "scipnp_solver_enable_push": (C.c_int, [_vp, _i, _i]),
    "scipnp_solver_uses_push": (C.c_int, [_vp]),
    "scipnp_solver_energy_log": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i)]),
    "scipnp_solver_exchange": (C.c_int, [_vp, _vp]),
    "scipnp_solver_run_tiled": (C.c_int, [_vp, _i, _i, _vp]),
    "scipnp_solver_sync_error": (C.c_int, [_vp, C.POINTER(_i), _vp]),
    "scipnp_solver_sync_flag": (C.c_int, [_vp, C.POINTER(_vp)]),
    "scipnp_solver_get_x": (C.c_int, [_vp, _fp, _vp]),
    "scipnp_solver_psnr": (C.c_int, [_vp, C.POINTER(C.c_double), _i, C.POINTER(_i), _vp]),
    "scipnp_solver_sqerr": (C.c_int, [_vp, C.POINTER(C.c_double), _i, C.POINTER(_i), _vp]),
    "scipnp_solver_refined_iters": (C.c_int, [_vp, C.POINTER(_i)]),
    "scipnp_solver_state": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "scipnp_solver_admm_state": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "scipnp_solver_launch_count": (C.c_longlong, [_vp]),
    "scipnp_solver_uses_fused": (C.c_int, [_vp]),
    "scipnp_host_release": (C.c_int, []),
    "scipnp_gap_denoise_host": (C.c_int, [_fp, _fp, _fp, _fp, C.POINTER(Params), _i, _fp,
                                          C.POINTER(C.c_double), C.POINTER(_i)]),
    "scipnp_admm_denoise_host": (C.c_int, [_fp, _fp, _fp, _fp, C.POINTER(Params), _i, _fp,
                                           C.POINTER(C.c_double), C.POINTER(_i)]),
    "scipnp_pipeline_create": (C.c_int, [C.POINTER(Params), _i, C.POINTER(_vp)]),
    "scipnp_pipeline_destroy": (C.c_int, [_vp]),
    "scipnp_pipeline_submit": (C.c_int, [_vp, _fp, _fp, _fp, _fp, _i, _fp, C.POINTER(_i)]),
    "scipnp_pipeline_wait": (C.c_int, [_vp, _i, C.POINTER(C.c_double), _i, C.POINTER(_i)]),
    "scipnp_pipeline_refined_iters": (C.c_int, [_vp, C.POINTER(_i)]),
}

for _name, (_res, _args) in _SIGS.items():
    _f = getattr(lib, _name)          # AttributeError here = header/library drift
    _f.restype = _res
    _f.argtypes = _args

EXPORTS = tuple(_SIGS)


def check(rc):
    if rc != 0:
        msg = lib.scipnp_last_error()
        raise ScipnpError("libscipnp error %d: %s" % (rc, msg.decode() if msg else "?"))


def require_device():
    if lib.scipnp_device_count() < 1:
        raise ScipnpError("no CUDA device visible: scipnp has no CPU fallback")
