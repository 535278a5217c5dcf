"""Oracle (test infrastructure): import the reference's OWN ``utils.py`` and
``pnp_sci_algo.py`` unmodified from ``/root/reference/PnP_SCI/python``.

Works only where the reference tree is mounted (the build container), never on
the GPU box.  Used by ``tests/golden/make_golden.py`` to generate the committed
fixtures and by the CPU tests that pin ``oracle/pnp_sci.py`` to the reference.

The reference imports packages that are absent here (matplotlib, scikit-image,
its CNN denoiser packages, colour-demosaicing).  They are replaced by empty
stub modules; the one third-party routine on the hot path,
``skimage.restoration.denoise_tv_chambolle``, and the two IQA helpers are
supplied by the oracle's restatements (``oracle/tv_chambolle.py``,
``oracle/iqa.py``).  Nothing is written to the read-only reference tree.
"""
import importlib
import os
import sys
import types

REFERENCE_PY = "/root/reference/PnP_SCI/python"

_STUBS = [
    "matplotlib", "matplotlib.pyplot",
    "skimage", "skimage.restoration", "skimage.measure", "skimage.metrics",
    "packages", "packages.ffdnet", "packages.ffdnet.test_ffdnet_ipol",
    "packages.fastdvdnet", "packages.fastdvdnet.test_fastdvdnet",
    "packages.colour_demosaicing", "packages.colour_demosaicing.bayer",
]


def available():
    return os.path.isfile(os.path.join(REFERENCE_PY, "pnp_sci_algo.py"))


def load(joint=False):
    """Return ``(utils_module, pnp_sci_algo_module)`` of the reference; with ``joint=True`` the third
    element is its ``joint_pnp_sci_algo`` module."""
    if not available():
        raise RuntimeError("reference tree not mounted at " + REFERENCE_PY)
    from . import tv_chambolle, iqa

    sys.dont_write_bytecode = True
    saved = {k: sys.modules.get(k) for k in _STUBS + ["utils", "pnp_sci_algo", "joint_pnp_sci_algo"]}
    try:
        for name in _STUBS:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        sk = sys.modules["skimage"]
        sk.__version__ = "0.17.2"
        r = sys.modules["skimage.restoration"]
        r.denoise_tv_chambolle = tv_chambolle.denoise_tv_chambolle
        for n in ("denoise_bilateral", "denoise_wavelet", "estimate_sigma", "denoise_tv_bregman"):
            setattr(r, n, _absent(n))
        ms = sys.modules["skimage.measure"]
        ms.compare_psnr = iqa.compare_psnr
        ms.compare_ssim = iqa.compare_ssim
        f = sys.modules["packages.ffdnet.test_ffdnet_ipol"]
        f.ffdnet_vdenoiser = _absent("ffdnet_vdenoiser")
        f.ffdnet_rgb_denoise = _absent("ffdnet_rgb_denoise")
        sys.modules["packages.fastdvdnet.test_fastdvdnet"].fastdvdnet_denoiser = \
            _absent("fastdvdnet_denoiser")
        sys.modules["packages.colour_demosaicing.bayer"] \
            .demosaicing_CFA_Bayer_Menon2007 = _absent("demosaicing")
        sys.path.insert(0, REFERENCE_PY)
        try:
            for n in ("utils", "pnp_sci_algo", "joint_pnp_sci_algo"):
                sys.modules.pop(n, None)
            ref_utils = importlib.import_module("utils")
            ref_algo = importlib.import_module("pnp_sci_algo")
            ref_joint = importlib.import_module("joint_pnp_sci_algo") if joint else None
        finally:
            sys.path.remove(REFERENCE_PY)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    if joint:
        return ref_utils, ref_algo, ref_joint
    return ref_utils, ref_algo


def _absent(name):
    def _f(*a, **k):
        raise RuntimeError("%s is outside the hot path and not available" % name)
    return _f
