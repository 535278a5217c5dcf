"""Module-name shim for ``from joint_pnp_sci_algo import ...`` (pnp_sci_test_* drivers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipnp.joint_pnp_sci_algo import admm_denoise, gap_denoise, A_, At_, psnr    # noqa: F401,E402
