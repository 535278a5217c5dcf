"""The reference's drivers through the module-name shims of ``sci-algorithms_b200/dropin``.

A driver of the reference does ``from pnp_sci_algo import admmdenoise_cacti``, ``from utils import (A_, At_)``
(pnp_sci_demo_kobe.py:22-24) or, for the pnp_sci_test_* family, additionally
``from joint_pnp_sci_algo import joint_admmdenoise_cacti`` and ``from utils import (A_, At_, show_n_save_res)``
(pnp_sci_test_orig.py:20-22).  With ``dropin/`` first on ``PYTHONPATH`` those statements must resolve to this
engine.  The CPU test runs the import blocks in a fresh interpreter; the GPU test runs the kobe demo's flow
(pnp_sci_demo_kobe.py:52-122: load a v5 .mat, A/At lambdas over the mask, GAP-TV then ADMM-TV through
admmdenoise_cacti, save with show_n_save_res) the same way and compares with the oracle.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "sci-algorithms_b200", "dropin")

IMPORT_BLOCKS = '''
from pnp_sci_algo import admmdenoise_cacti
from utils import (A_, At_)
from pnp_sci_algo import admmdenoise_cacti
from joint_pnp_sci_algo import joint_admmdenoise_cacti
from utils import (A_, At_, show_n_save_res)
from pnp_sci_algo import (gap_denoise, admm_denoise, gap_denoise_bayer)
from joint_pnp_sci_algo import (gap_joint_denoise, admm_joint_denoise, gap_multistep_denoise, admm_multistep_denoise)
from utils import (psnr, save_rgb_img, cli_run, rescale)
import scipnp, pnp_sci_algo, utils, joint_pnp_sci_algo
assert pnp_sci_algo.admmdenoise_cacti is scipnp.pnp_sci_algo.admmdenoise_cacti
assert utils.A_ is scipnp.utils.A_
assert joint_pnp_sci_algo.joint_admmdenoise_cacti is scipnp.joint_pnp_sci_algo.joint_admmdenoise_cacti
print("imports ok")
'''

DEMO_FLOW = '''
import sys
import numpy as np
import scipy.io as sio
from statistics import mean
from pnp_sci_algo import admmdenoise_cacti
from utils import (A_, At_, show_n_save_res)

matfile, resultsdir = sys.argv[1], sys.argv[2]
file = sio.loadmat(matfile)
meas = np.float32(file['meas']); mask = np.float32(file['mask']); orig = np.float32(file['orig'])
iframe, nframe, MAXB = 0, 2, 255.
A = lambda x: A_(x, mask)
At = lambda y: At_(y, mask)
common = dict(v0=None, orig=orig, iframe=iframe, nframe=nframe, MAXB=MAXB, maskdirection='plain', _lambda=1,
              denoiser='tv', iter_max=12, tv_weight=0.3, tv_iter_max=5)
vgaptv, tgaptv, psnr_gaptv, ssim_gaptv, psnrall_gaptv = admmdenoise_cacti(meas, mask, A, At, projmeth='gap',
                                                                            accelerate=True, **common)
print('GAP-TV PSNR {:2.2f} dB, SSIM {:.4f}, running time {:.1f} seconds.'.format(mean(psnr_gaptv), mean(ssim_gaptv), tgaptv))
vadmmtv, tadmmtv, psnr_admmtv, ssim_admmtv, psnrall_admmtv = admmdenoise_cacti(meas, mask, A, At, projmeth='admm',
                                                                                 gamma=0.01, **common)
show_n_save_res(vgaptv, tgaptv, psnr_gaptv, ssim_gaptv, psnrall_gaptv, orig, mask.shape[2], resultsdir, 'gaptv',
                iframe=iframe, nframe=nframe, MAXB=MAXB, show_res_flag=0, save_res_flag=1, tv_weight=0.3)
show_n_save_res(vadmmtv, tadmmtv, psnr_admmtv, ssim_admmtv, psnrall_admmtv, orig, mask.shape[2], resultsdir, 'admmtv',
                iframe=iframe, nframe=nframe, MAXB=MAXB, show_res_flag=0, save_res_flag=1)
'''


def _run(code, *argv):
    env = dict(os.environ)
    env["PYTHONPATH"] = DROPIN + os.pathsep + env.get("PYTHONPATH", "")
    return subprocess.run([sys.executable, "-c", code, *argv], env=env, cwd="/tmp", capture_output=True, text=True,
                          timeout=600)


def test_reference_driver_import_blocks_resolve_to_the_engine():
    r = _run(IMPORT_BLOCKS)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "imports ok" in r.stdout


def test_show_n_save_res_writes_the_reference_layout(tmp_path):
    import scipy.io as sio
    sys.path.insert(0, os.path.join(ROOT, "sci-algorithms_b200"))
    from scipnp.utils import show_n_save_res, rescale
    v = np.random.default_rng(0).random((8, 8, 8)).astype(np.float32)
    show_n_save_res(v, 1.5, [30.] * 8, [0.9] * 8, [[1., 2.]], v * 255, 4, str(tmp_path), 'run', iframe=1, nframe=2,
                    MAXB=255., show_res_flag=0, save_res_flag=1, note='x')
    rec = sio.loadmat(str(tmp_path / 'savedmat' / 'run_kmeas1_2.mat'))
    assert rec['vdenoise'].shape == (8, 8, 8) and float(rec['psnr_mean'].item()) == 30. and rec['note'][0] == 'x'
    assert int(rec['Cr'].item()) == 4 and int(rec['iframe'].item()) == 1
    show_n_save_res(v, 1.5, [], [], [], None, 4, str(tmp_path), 'blind', show_res_flag=0)
    assert 'orig' not in sio.loadmat(str(tmp_path / 'savedmat' / 'blind_kmeas0_0.mat'))
    r = rescale(np.array([2., 4., 6.]))
    assert r.min() == 0. and r.max() == 1.


@pytest.mark.gpu
def test_kobe_demo_flow_through_the_dropin(tmp_path):
    import scipy.io as sio
    sys.path.insert(0, os.path.join(ROOT, "sci-algorithms_b200"))
    from scipnp import synth
    from oracle import pnp_sci as O
    meas, mask, orig = synth.make_cacti(64, 80, 8, 2, cfg=61)
    mat = str(tmp_path / "toy_cacti.mat")
    sio.savemat(mat, {'meas': meas, 'mask': mask, 'orig': orig})
    r = _run(DEMO_FLOW, mat, str(tmp_path))
    assert r.returncode == 0, r.stderr[-3000:]
    assert "GAP-TV PSNR" in r.stdout
    A = lambda x: O.A_(x, mask)
    At = lambda y: O.At_(y, mask)
    common = dict(v0=None, orig=orig, iframe=0, nframe=2, MAXB=255., maskdirection='plain', _lambda=1,
                  denoiser='tv', iter_max=12, tv_weight=0.3, tv_iter_max=5)
    for name, kw in (('gaptv', dict(projmeth='gap', accelerate=True)), ('admmtv', dict(projmeth='admm', gamma=0.01))):
        xo, _, pso, sso, pao = O.admmdenoise_cacti(meas, mask, A, At, **kw, **common)
        rec = sio.loadmat(str(tmp_path / 'savedmat' / (name + '_kmeas0_1.mat')))
        assert float(np.abs(rec['vdenoise'] - xo).max()) <= 1e-4
        assert float(np.abs(rec['psnr_denoise'].ravel() - np.array(pso)).max()) <= 0.01
        assert float(np.abs(rec['psnrall_denoise'] - np.array(pao)).max()) <= 0.01
        assert float(np.abs(rec['ssim_denoise'].ravel() - np.array(sso)).max()) <= 1e-4
