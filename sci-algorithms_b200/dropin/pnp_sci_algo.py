"""Module-name shim: put this directory first on ``sys.path`` and the reference's
drivers (``from pnp_sci_algo import admmdenoise_cacti``, pnp_sci_demo_kobe.py:22)
pick up the B200 engine without being edited."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipnp.pnp_sci_algo import *            # noqa: F401,F403,E402
from scipnp.pnp_sci_algo import (gap_denoise, admm_denoise, admmdenoise_cacti,   # noqa: F401,E402
                                 gap_denoise_bayer, gap_denoise_cassi, denoise_tv_chambolle)
