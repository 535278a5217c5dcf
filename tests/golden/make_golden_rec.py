#!/usr/bin/env python
"""Golden vectors of the stand-alone TV loops ``GAP_TV_rec`` / ``ADMM_TV_rec`` (pnp_sci_algo.py:866-907), produced by
the REFERENCE's own functions (imported unmodified through ``oracle/reference_loader.py``; the TV step inside is the
oracle's restatement of scikit-image's ``denoise_tv_chambolle``, as for the other loop fixtures).

    python tests/golden/make_golden_rec.py      # build container only: needs /root/reference
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sci-algorithms_b200"))

from oracle import reference_loader, pnp_sci as opnp  # noqa: E402
from scipnp import synth  # noqa: E402


def main():
    ref_utils, ref_algo = reference_loader.load()[:2]
    f32 = np.float32
    H, W, C = 36, 44, 8
    meas, mask, orig = synth.make_cacti(H, W, C, 1, cfg=31)
    y = meas[:, :, 0] / f32(255.)
    Xo = orig[:, :, :C] / f32(255.)
    ms = opnp.phi_sum(mask)
    with contextlib.redirect_stdout(io.StringIO()):
        g = ref_algo.GAP_TV_rec(y, mask, ref_utils.A_, ref_utils.At_, ms, 6, 1.0, 0.3, H, W, C, Xo)
        a = ref_algo.ADMM_TV_rec(y, mask, ref_utils.A_, ref_utils.At_, ms, 6, 1.0, 0.3, H, W, C, 0.01, Xo)
    path = os.path.join(HERE, "tv_rec.npz")
    np.savez_compressed(path, y=y, mask=mask, X_orig=Xo, Phi_sum=ms, gap=g, admm=a, maxiter=6, step_size=1.0,
                        weight=0.3, eta=0.01)
    print("tv_rec %.1f KB  gap %s admm %s" % (os.path.getsize(path) / 1024., g.dtype, a.dtype))


if __name__ == "__main__":
    main()
