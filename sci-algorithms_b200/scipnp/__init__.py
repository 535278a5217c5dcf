"""scipnp -- B200-native GAP/ADMM-TV reconstruction engine for snapshot
compressive imaging: the iterative hot path of the reference's
``PnP_SCI/python`` behind the reference's own call surface.

    from scipnp import pnp_sci_algo, utils       # same names as the reference
    pnp_sci_algo.admmdenoise_cacti(meas, mask, A, At, projmeth='gap', ...)

All computation happens in hand-written sm_100a CUDA kernels reached through
the C ABI of ``libscipnp.so`` (``include/scipnp.h``).  There is no CPU
fallback: importing this package without the built library raises ImportError.
"""
from . import _lib                                   # fails loudly if the .so is missing
from ._lib import ScipnpError, LIB_PATH
from .engine import Solver, HostPipeline
from .utils import A_, At_, psnr, phi_sum
from .pnp_sci_algo import (gap_denoise, admm_denoise, admmdenoise_cacti, gap_denoise_bayer,
                           gap_denoise_cassi, denoise_tv_chambolle)

__version__ = "0.1.0"
__all__ = ["Solver", "HostPipeline", "A_", "At_", "psnr", "phi_sum", "gap_denoise", "admm_denoise",
           "admmdenoise_cacti", "gap_denoise_bayer", "gap_denoise_cassi",
           "denoise_tv_chambolle", "ScipnpError", "LIB_PATH"]
