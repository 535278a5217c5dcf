"""Module-name shim for ``from utils import (A_, At_)`` (pnp_sci_demo_kobe.py:24)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipnp.utils import A_, At_, psnr, phi_sum     # noqa: F401,E402
