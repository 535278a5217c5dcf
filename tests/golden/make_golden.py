#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the REFERENCE's own
``utils.py`` / ``pnp_sci_algo.py`` (imported unmodified from /root/reference via
``oracle/reference_loader.py``) on small seeded inputs.

Run in the build container only (the reference tree is not on the GPU box):

    python tests/golden/make_golden.py

Every ``*.npz`` stores the inputs and the reference's outputs.  The TV denoiser
inside the reference loops is the oracle's restatement of scikit-image's
``denoise_tv_chambolle`` (third-party, absent -- see ``oracle/__init__.py``), so
``tv_*.npz`` are regression vectors of that restatement, not reference outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sci-algorithms_b200"))

from oracle import reference_loader, tv_chambolle, pnp_sci as opnp  # noqa: E402
from scipnp import synth  # noqa: E402


def second_standin(x, nsig, model=None):
    """Stand-in for the learned denoiser of the TV+CNN period (same arithmetic in
    tests/test_gpu_parity.py with torch): an affine shrink that depends on the noise level."""
    a = np.float32(1.0 - 0.1 * float(nsig))
    return np.clip(x * a + np.float32(0.01), 0, 1).astype(np.float32)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("%-28s %7.1f KB" % (name, os.path.getsize(path) / 1024.))


def main():
    ref_utils, ref_algo, ref_joint = reference_loader.load(joint=True)
    f32 = np.float32

    # -- operators (R1, R2, R3, R10) -----------------------------------------
    rng = np.random.default_rng(7)
    for tag, (H, W, C) in {"ops_8": (24, 20, 8), "ops_5": (17, 23, 5)}.items():
        x = rng.random((H, W, C), dtype=f32)
        Phi = (rng.random((H, W, C)) <= 0.5).astype(f32)
        Phi[0, 0, :] = 0                       # a pixel no mask ever opens
        y = rng.random((H, W), dtype=f32)
        x2 = x + f32(0.05) * rng.standard_normal((H, W, C)).astype(f32)
        ms = np.sum(Phi, axis=2)
        ms[ms == 0] = 1
        save(tag, x=x, Phi=Phi, y=y, x2=x2,
             A=ref_utils.A_(x, Phi), At=ref_utils.At_(y, Phi), Phi_sum=ms,
             psnr=np.float64(ref_utils.psnr(x, x2)),
             psnr_same=np.float64(ref_utils.psnr(x, x)))

    # -- solver loops (R4, R5, R7) -------------------------------------------
    def cacti(H, W, C, F, cfg):
        meas, mask, orig = synth.make_cacti(H, W, C, F, cfg=cfg)
        return meas, mask, orig

    meas, mask, orig = cacti(40, 48, 8, 2, cfg=11)
    A = lambda x: ref_utils.A_(x, mask)
    At = lambda y: ref_utils.At_(y, mask)
    ms = opnp.phi_sum(mask)
    y = meas[:, :, 0] / f32(255.)
    Xo = orig[:, :, :8] / f32(255.)

    x, ps, ss, pa = ref_algo.gap_denoise(y, ms, A, At, _lambda=1, accelerate=True,
                                         denoiser='tv', iter_max=12, tv_weight=0.3,
                                         tv_iter_max=5, X_orig=Xo)
    save("gap_acc", y=y, mask=mask, X_orig=Xo, x=x, psnr=np.array(ps),
         ssim=np.array(ss), psnr_all=np.array(pa), iter_max=12, tv_weight=0.3,
         tv_iter_max=5, _lambda=1.0, accelerate=True)

    x, ps, ss, pa = ref_algo.gap_denoise(y, ms, A, At, _lambda=0.75, accelerate=False,
                                         denoiser='tv', iter_max=[3, 4], sigma=[0.2, 0.1],
                                         tv_weight=0.1, tv_iter_max=3, X_orig=Xo)
    save("gap_plain", y=y, mask=mask, X_orig=Xo, x=x, psnr=np.array(ps),
         ssim=np.array(ss), psnr_all=np.array(pa), iter_max=np.array([3, 4]),
         sigma=np.array([0.2, 0.1]), tv_weight=0.1, tv_iter_max=3, _lambda=0.75,
         accelerate=False)

    x, ps, ss, pa = ref_algo.admm_denoise(y, ms, A, At, _lambda=1, gamma=0.01,
                                          denoiser='tv', iter_max=12, tv_weight=0.3,
                                          tv_iter_max=5, X_orig=Xo)
    save("admm", y=y, mask=mask, X_orig=Xo, x=x, psnr=np.array(ps),
         ssim=np.array(ss), psnr_all=np.array(pa), iter_max=12, tv_weight=0.3,
         tv_iter_max=5, _lambda=1.0, gamma=0.01)

    # joint module's ADMM (clip of theta, gamma = 0 default): SURVEY 8f-1
    x, ps, ss, pa = ref_joint.admm_denoise(y, ms, A, At, _lambda=1, gamma=0.0, denoiser='tv',
                                           iter_max=12, tv_weight=0.3, tv_iter_max=5, X_orig=Xo)
    save("joint_admm", y=y, mask=mask, X_orig=Xo, x=x, psnr=np.array(ps), ssim=np.array(ss),
         psnr_all=np.array(pa), iter_max=12, tv_weight=0.3, tv_iter_max=5, _lambda=1.0, gamma=0.0)

    # TV + learned-denoiser period and the two-period driver of the joint module.
    # The reference's FFDNet is absent; a fixed elementwise stand-in is injected in its place so
    # that the reference's own loop (projection, TV, hand-off, PSNR) produces the vectors.
    ref_joint.ffdnet_vdenoiser = second_standin
    x, ps, ss, pa = ref_joint.gap_multistep_denoise(y, ms, A, At, _lambda=1, accelerate=True,
                                                    denoiser='tv+ffdnet', iter_max=[3, 3],
                                                    sigma=[0.2, 0.1], tv_weight=0.3, tv_iter_max=5,
                                                    X_orig=Xo)
    save("joint_multistep", y=y, mask=mask, X_orig=Xo, x=x, psnr=np.array(ps), ssim=np.array(ss),
         psnr_all=np.array(pa), iter_max=np.array([3, 3]), sigma=np.array([0.2, 0.1]), tv_weight=0.3,
         tv_iter_max=5)
    x, ps, ss, pa = ref_joint.gap_joint_denoise(y, ms, A, At, X_orig=Xo, denoiser='tv+ffdnet',
                                                iter_max1=4, iter_max2=[2, 2], sigma1=None,
                                                sigma2=[0.2, 0.1], _lambda=1, accelerate=True,
                                                tv_weight=0.3, tv_iter_max=5)
    save("joint_two_period", y=y, mask=mask, X_orig=Xo, x=x, psnr=np.array(ps), ssim=np.array(ss),
         psnr_all=np.array(pa), iter_max1=4, iter_max2=np.array([2, 2]), sigma2=np.array([0.2, 0.1]),
         tv_weight=0.3, tv_iter_max=5)

    x, ps, ss, pa = ref_joint.admm_multistep_denoise(y, ms, A, At, _lambda=1, gamma=0.01,
                                                     denoiser='tv+ffdnet', iter_max=[3, 3],
                                                     sigma=[0.2, 0.1], tv_weight=0.3, tv_iter_max=5,
                                                     X_orig=Xo)
    save("joint_admm_multistep", y=y, mask=mask, X_orig=Xo, x=x, psnr=np.array(ps), ssim=np.array(ss),
         psnr_all=np.array(pa), iter_max=np.array([3, 3]), sigma=np.array([0.2, 0.1]), tv_weight=0.3,
         tv_iter_max=5, gamma=0.01)
    x, ps, ss, pa = ref_joint.admm_joint_denoise(y, ms, A, At, X_orig=Xo, denoiser='tv+ffdnet',
                                                 iter_max1=4, iter_max2=[2, 2], sigma1=None,
                                                 sigma2=[0.2, 0.1], _lambda=1, gamma=0.01,
                                                 tv_weight=0.3, tv_iter_max=5)
    save("joint_admm_two_period", y=y, mask=mask, X_orig=Xo, x=x, psnr=np.array(ps), ssim=np.array(ss),
         psnr_all=np.array(pa), iter_max1=4, iter_max2=np.array([2, 2]), sigma2=np.array([0.2, 0.1]),
         tv_weight=0.3, tv_iter_max=5, gamma=0.01)

    # warm start + ragged channel count (C=5, odd sizes)
    meas5, mask5, orig5 = cacti(33, 29, 5, 1, cfg=12)
    A5 = lambda x: ref_utils.A_(x, mask5)
    At5 = lambda y: ref_utils.At_(y, mask5)
    y5 = meas5[:, :, 0] / f32(255.)
    x0 = np.clip(orig5 / f32(255.) + f32(0.1), 0, 1).astype(f32)
    x, ps, ss, pa = ref_algo.gap_denoise(y5, opnp.phi_sum(mask5), A5, At5, iter_max=6,
                                         tv_weight=0.2, tv_iter_max=4, x0=x0,
                                         X_orig=orig5 / f32(255.))
    save("gap_c5_warm", y=y5, mask=mask5, x0=x0, X_orig=orig5 / f32(255.), x=x,
         psnr_all=np.array(pa), iter_max=6, tv_weight=0.2, tv_iter_max=4)

    for md in ("plain", "updown"):
        for pm in ("gap", "admm"):
            kw = dict(_lambda=1, denoiser='tv', iter_max=5, tv_weight=0.3, tv_iter_max=5)
            if pm == "gap":
                kw["accelerate"] = True
            else:
                kw["gamma"] = 0.01
            x_, t_, ps, ss, pa = ref_algo.admmdenoise_cacti(
                meas, mask, A, At, projmeth=pm, v0=None, orig=orig, iframe=0,
                nframe=2, MAXB=255., maskdirection=md, **kw)
            save("cacti_%s_%s" % (pm, md), meas=meas, mask=mask, orig=orig, x=x_,
                 psnr=np.array(ps), ssim=np.array(ss), psnr_all=np.array(pa),
                 iter_max=5, tv_weight=0.3, tv_iter_max=5, MAXB=255.)

    # -- Bayer (R8) -----------------------------------------------------------
    yb, Pb, ob = synth.make_bayer(32, 40, 4, cfg=13)
    x, ps, ss, pa = ref_algo.gap_denoise_bayer(yb, Pb, _lambda=1, accelerate=True,
                                               denoiser='tv', iter_max=8, tv_weight=0.1,
                                               tv_iter_max=5, X_orig=ob)
    save("bayer", y_bayer=yb, Phi_bayer=Pb, X_orig=ob, x=x, psnr=np.array(ps),
         ssim=np.array(ss), psnr_all=np.array(pa), iter_max=8, tv_weight=0.1,
         tv_iter_max=5)

    # -- CASSI (R9): reference operators on the explicit shifted stack ----------
    yc, m2, cube = synth.make_cassi(24, 20, 6, step=2, cfg=14)
    Phic = opnp.cassi_shift_mask(m2, 6, 2)
    Ac = lambda x: ref_utils.A_(x, Phic)
    Atc = lambda y: ref_utils.At_(y, Phic)
    x, ps, ss, pa = ref_algo.gap_denoise(yc, opnp.phi_sum(Phic), Ac, Atc, iter_max=8,
                                         tv_weight=0.1, tv_iter_max=5, X_orig=cube)
    save("cassi", y=yc, mask2d=m2, step=2, nband=6, Phi=Phic, X_orig=cube, x=x,
         psnr_all=np.array(pa), iter_max=8, tv_weight=0.1, tv_iter_max=5)

    # -- TV (R6): regression vectors of the restatement --------------------------
    rng = np.random.default_rng(21)
    img = (synth.moving_scene(37, 45, 3)[:, :, :] / 255.).astype(f32)
    img += f32(0.08) * rng.standard_normal(img.shape).astype(f32)
    for T, w in ((1, 0.1), (2, 0.1), (5, 0.3), (200, 0.1)):
        en = []
        out = tv_chambolle.denoise_tv_chambolle(img, w, n_iter_max=T, multichannel=True,
                                                energy_out=en)
        nexec = np.array([len(e) for e in en])
        emax = max(len(e) for e in en)
        E = np.full((len(en), emax), np.nan)
        for c, e in enumerate(en):
            E[c, :len(e)] = e
        save("tv_T%d" % T, image=img, weight=w, n_iter_max=T, out=out, energy=E,
             n_exec=nexec)


if __name__ == "__main__":
    main()
