"""Module-name shim for ``from joint_pnp_sci_algo import ...`` (pnp_sci_test_* drivers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipnp.joint_pnp_sci_algo import (joint_admmdenoise_cacti, admm_denoise, gap_denoise, gap_multistep_denoise,     # noqa: F401,E402
                                       gap_joint_denoise, admm_multistep_denoise, admm_joint_denoise,
                                       A_, At_, psnr)
