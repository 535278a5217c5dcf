"""Module-name shim for ``from utils import (A_, At_)`` (pnp_sci_demo_kobe.py:24) and
``from utils import (A_, At_, show_n_save_res)`` (pnp_sci_test_orig.py:22)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipnp.utils import (A_, At_, psnr, phi_sum, show_n_save_res, save_rgb_img, cli_run,     # noqa: F401,E402
                          rescale)
