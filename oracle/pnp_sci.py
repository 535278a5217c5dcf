"""Oracle (test infrastructure): NumPy restatement of the GAP/ADMM-TV path of
``/root/reference/PnP_SCI/python``.

Pinned against the reference's own ``utils.py`` / ``pnp_sci_algo.py`` (imported
unmodified by ``oracle/reference_loader.py`` in the build container; fixtures in
``tests/golden/``).  Only the TV branches are restated -- wavelet / FFDNet /
FastDVDnet priors are outside the hot path and raise like the reference does
for an unknown denoiser.

Each function cites the reference lines it follows.  The arithmetic order of
every array statement is the reference's, so float32 results agree bit for bit
with the reference run under the same NumPy.
"""
import math
import time

import numpy as np

from .tv_chambolle import denoise_tv_chambolle
from .iqa import compare_psnr, compare_ssim

__all__ = ["A_", "At_", "psnr", "phi_sum", "gap_denoise", "admm_denoise", "joint_admm_denoise",
           "admmdenoise_cacti", "gap_denoise_bayer", "cassi_shift_mask",
           "cassi_shift_cube", "GAP_TV_rec", "ADMM_TV_rec", "admm_denoise_bayer"]


# -- R1 / R2 / R10 / R3 -------------------------------------------------------

def A_(x, Phi):
    """Forward model y = sum_c x[:,:,c]*Phi[:,:,c]  (utils.py:10-15)."""
    return np.sum(x * Phi, axis=2)


def At_(y, Phi):
    """Adjoint x[:,:,c] = y*Phi[:,:,c]  (utils.py:17-26)."""
    return np.multiply(np.repeat(y[:, :, np.newaxis], Phi.shape[2], axis=2), Phi)


def psnr(ref, img):
    """PSNR on [0,1] data, 100 when identical  (utils.py:28-36)."""
    mse = np.mean((ref - img) ** 2)
    if mse == 0:
        return 100
    return 20 * math.log10(1. / math.sqrt(mse))


def phi_sum(mask):
    """sum_c Phi with zeros replaced by one  (pnp_sci_algo.py:491-492)."""
    s = np.sum(mask, axis=tuple(range(2, mask.ndim)))
    s[s == 0] = 1
    return s


def _as_schedule(sigma, iter_max):
    # pnp_sci_algo.py:628-631 / 796-799
    if not isinstance(sigma, list):
        sigma = [sigma]
    if not isinstance(iter_max, list):
        iter_max = [iter_max] * len(sigma)
    return sigma, iter_max


def _frame_iqa(X_orig, x):
    # pnp_sci_algo.py:699-705 / 857-863
    ps, ss = [], []
    if X_orig is not None:
        for c in range(x.shape[-1]):
            ps.append(compare_psnr(X_orig[..., c], x[..., c], data_range=1.))
            ss.append(compare_ssim(X_orig[..., c], x[..., c], data_range=1.))
    return ps, ss


# -- R4 -----------------------------------------------------------------------

def gap_denoise(y, Phi_sum, A, At, _lambda=1, accelerate=True, denoiser='tv',
                iter_max=50, noise_estimate=False, sigma=None, tv_weight=0.1,
                tv_iter_max=5, multichannel=True, x0=None, X_orig=None,
                model=None, show_iqa=True, tvm='tv_chambolle'):
    """GAP with a TV prior  (pnp_sci_algo.py:536-706; loop :638-650,:682)."""
    if denoiser.lower() != 'tv' or tvm != 'tv_chambolle':
        raise ValueError('Unsupported denoiser {}!'.format(denoiser))
    if x0 is None:
        x0 = At(y)                                    # :625-627
    sigma, iter_max = _as_schedule(sigma, iter_max)
    y1 = np.zeros_like(y)                             # :633
    x = x0
    psnr_all = []
    for idx, _ in enumerate(sigma):
        for _it in range(iter_max[idx]):
            yb = A(x)                                 # :640
            if accelerate:
                y1 = y1 + (y - yb)                    # :642
                x = x + _lambda * (At((y1 - yb) / Phi_sum))   # :643
            else:
                x = x + _lambda * (At((y - yb) / Phi_sum))    # :645
            x = denoise_tv_chambolle(x, tv_weight, n_iter_max=tv_iter_max,
                                     multichannel=multichannel)  # :650
            if show_iqa and X_orig is not None:
                psnr_all.append(psnr(X_orig, x))      # :682
    ps, ss = _frame_iqa(X_orig, x)
    return x, ps, ss, psnr_all


# -- the stand-alone TV loops at the end of the module ---------------------------

def GAP_TV_rec(y, Phi, A, At, Phi_sum, maxiter, step_size, weight, row, col, ColT, X_ori):
    """Accelerated GAP with 30 Chambolle iterations per step (pnp_sci_algo.py:866-882).  ``A`` / ``At`` take the
    mask as their second argument here (``A_`` / ``At_`` of utils.py); ``y1`` starts as float64 zeros (:867), so
    the whole loop runs in float64 like the reference's.  The progress print (:877-881) is not restated."""
    y1 = np.zeros((row, col))                                                   # :867
    f = At(y, Phi)                                                              # :869
    for ni in range(maxiter):
        fb = A(f, Phi)                                                          # :871
        y1 = y1 + (y - fb)                                                      # :872
        f = f + np.multiply(step_size, At(np.divide(y1 - fb, Phi_sum), Phi))    # :873
        f = denoise_tv_chambolle(f, weight, n_iter_max=30, multichannel=True)   # :874
    return f


def ADMM_TV_rec(y, Phi, A, At, Phi_sum, maxiter, step_size, weight, row, col, ColT, eta, X_ori):
    """ADMM with 30 Chambolle iterations per step and geometrically decaying TV weight (x0.999) and
    regulariser eta (x0.998) (pnp_sci_algo.py:884-907).  Returns ``v`` (the projection output)."""
    theta = At(y, Phi)                                                          # :887
    v = theta
    b = np.zeros((row, col, ColT))                                              # :889
    for ni in range(maxiter):
        yb = A(theta + b, Phi)                                                  # :891
        v = (theta + b) + np.multiply(step_size, At(np.divide(y - yb, Phi_sum + eta), Phi))   # :893
        theta = denoise_tv_chambolle(v - b, weight, n_iter_max=30, multichannel=True)        # :895
        b = b - (v - theta)                                                     # :897
        weight = 0.999 * weight                                                 # :898
        eta = 0.998 * eta                                                       # :899
    return v


# -- R5 -----------------------------------------------------------------------

def admm_denoise(y, Phi_sum, A, At, _lambda=1, gamma=0.01, denoiser='tv',
                 iter_max=50, noise_estimate=False, sigma=None, tv_weight=0.1,
                 tv_iter_max=5, multichannel=True, x0=None, model=None,
                 X_orig=None, show_iqa=True):
    """ADMM with a TV prior  (pnp_sci_algo.py:708-864; loop :805-812,:836-840).

    Returns ``x`` (the projection output, not ``theta``) and reports the PSNR
    of ``x``, as the reference does.
    """
    if denoiser.lower() != 'tv':
        raise ValueError('Unsupported denoiser {}!'.format(denoiser))
    if x0 is None:
        x0 = At(y)                                    # :793-794
    sigma, iter_max = _as_schedule(sigma, iter_max)
    x = x0
    theta = x0
    b = np.zeros_like(x0)                             # :802
    psnr_all = []
    for idx, _ in enumerate(sigma):
        for _it in range(iter_max[idx]):
            yb = A(theta + b)                         # :808
            x = (theta + b) + _lambda * (At((y - yb) / (Phi_sum + gamma)))  # :809
            theta = denoise_tv_chambolle(x - b, tv_weight,
                                         n_iter_max=tv_iter_max,
                                         multichannel=multichannel)  # :812
            b = b - (x - theta)                       # :836
            if show_iqa and X_orig is not None:
                psnr_all.append(psnr(X_orig, x))      # :840
    ps, ss = _frame_iqa(X_orig, x)
    return x, ps, ss, psnr_all


def joint_admm_denoise(y, Phi_sum, A, At, _lambda=1, gamma=0.0, accelerate=None, denoiser='tv',
                       iter_max=50, noise_estimate=False, sigma=None, tv_weight=0.1, tv_iter_max=5,
                       multichannel=True, x0=None, model=None, X_orig=None, show_iqa=True,
                       tvm='tv_chambolle'):
    """ADMM-TV of the joint module (joint_pnp_sci_algo.py:502-665): as ``admm_denoise`` with
    ``theta = np.clip(theta, 0, 1)`` after the denoiser (:633) and ``gamma`` defaulting to 0.
    ``tvm`` 'ITV3D_FGP' and 'ITV2D_cham' also call denoise_tv_chambolle there (:608-611)."""
    if denoiser.lower() != 'tv' or tvm not in ('tv_chambolle', 'ITV3D_FGP', 'ITV2D_cham'):
        raise ValueError('Unsupported denoiser {}!'.format(denoiser))
    if x0 is None:
        x0 = At(y)
    sigma, iter_max = _as_schedule(sigma, iter_max)
    x = x0
    theta = x0
    b = np.zeros_like(x0)
    psnr_all = []
    for idx, _ in enumerate(sigma):
        for _it in range(iter_max[idx]):
            yb = A(theta + b)                                             # :598
            x = (theta + b) + _lambda * (At((y - yb) / (Phi_sum + gamma)))  # :599
            theta = denoise_tv_chambolle(x - b, tv_weight, n_iter_max=tv_iter_max,
                                         multichannel=multichannel)        # :607
            theta = np.clip(theta, 0, 1)                                   # :633
            b = b - (x - theta)                                            # :635
            if show_iqa and X_orig is not None:
                psnr_all.append(psnr(X_orig, x))
    ps, ss = _frame_iqa(X_orig, x)
    return x, ps, ss, psnr_all


def gap_multistep_denoise(y, Phi_sum, A, At, second, _lambda=1, accelerate=True,
                          denoiser='tv+ffdnet', iter_max=50, noise_estimate=False, sigma=None,
                          tv_weight=0.1, tv_iter_max=5, multichannel=True, x0=None, X_orig=None,
                          model=None, show_iqa=True, tvm='tv_chambolle'):
    """TV + second-denoiser period of the joint module (joint_pnp_sci_algo.py:309-500): every
    iteration is the GAP projection (:412-417), the Chambolle TV step (:428-429; 'ITV3D_FGP' and
    'ITV2D_cham' name functions that do not exist in that file, so only 'tv_chambolle' runs) and
    then a learned denoiser ``x = second(x, nsig, model)`` (:441, :466 -- FFDNet / FastDVDnet in
    the reference, absent here and injected by the caller).  PSNR is taken after both (:473)."""
    if denoiser.lower() not in ('tv+ffdnet', 'tv+fastdvdnet'):
        raise ValueError('Unsupported denoiser {}!'.format(denoiser))
    if tvm != 'tv_chambolle':
        raise ValueError('Unsupported TV denoiser {}!'.format(tvm))
    if x0 is None:
        x0 = At(y)
    sigma, iter_max = _as_schedule(sigma, iter_max)
    y1 = np.zeros_like(y)
    x = x0
    psnr_all = []
    for idx, nsig in enumerate(sigma):
        for _it in range(iter_max[idx]):
            yb = A(x)
            if accelerate:
                y1 = y1 + (y - yb)
                x = x + _lambda * (At((y1 - yb) / Phi_sum))
            else:
                x = x + _lambda * (At((y - yb) / Phi_sum))
            x = denoise_tv_chambolle(x, tv_weight, n_iter_max=tv_iter_max, multichannel=multichannel)
            x = second(x, nsig, model)
            if show_iqa and X_orig is not None:
                psnr_all.append(psnr(X_orig, x))
    ps, ss = _frame_iqa(X_orig, x)
    return x, ps, ss, psnr_all


def gap_joint_denoise(y, Phi_sum, A, At, second, x0=None, X_orig=None, denoiser='tv+ffdnet',
                      iter_max1=50, iter_max2=50, sigma1=None, sigma2=None, **args):
    """Two periods (joint_pnp_sci_algo.py:100-116): GAP-TV, then the TV + second-denoiser loop
    started from its result; returns what the second period returns."""
    x, _, _, _ = gap_denoise(y, Phi_sum, A, At, x0=x0, X_orig=X_orig, denoiser='tv',
                             iter_max=iter_max1, sigma=sigma1, **args)
    return gap_multistep_denoise(y, Phi_sum, A, At, second, x0=x, X_orig=X_orig, denoiser=denoiser,
                                 iter_max=iter_max2, sigma=sigma2, **args)


def admm_multistep_denoise(y, Phi_sum, A, At, second, _lambda=1, gamma=0.0, accelerate=None,
                           denoiser='tv+ffdnet', iter_max=50, noise_estimate=False, sigma=None,
                           tv_weight=0.1, tv_iter_max=5, multichannel=True, x0=None, model=None,
                           X_orig=None, show_iqa=True, tvm='tv_chambolle'):
    """ADMM twin of ``gap_multistep_denoise`` (joint_pnp_sci_algo.py:118-306): projection (:213-214),
    ``theta = TV(x - b)`` (:230; 'ITV3D_FGP' and 'ITV2D_cham' call the same function here),
    ``theta = second(theta, nsig, model)`` (:242, :263), clip to [0, 1] (:268), multiplier (:270);
    PSNR of ``x`` (:273) and ``x`` is returned."""
    if denoiser.lower() not in ('tv+ffdnet', 'tv+fastdvdnet'):
        raise ValueError('Unsupported denoiser {}!'.format(denoiser))
    if tvm not in ('tv_chambolle', 'ITV3D_FGP', 'ITV2D_cham'):
        raise ValueError('Unsupported TV denoiser {}!'.format(tvm))
    if x0 is None:
        x0 = At(y)
    sigma, iter_max = _as_schedule(sigma, iter_max)
    x = x0
    theta = x0
    b = np.zeros_like(x0)
    psnr_all = []
    for idx, nsig in enumerate(sigma):
        for _it in range(iter_max[idx]):
            yb = A(theta + b)
            x = (theta + b) + _lambda * (At((y - yb) / (Phi_sum + gamma)))
            theta = denoise_tv_chambolle(x - b, tv_weight, n_iter_max=tv_iter_max, multichannel=multichannel)
            theta = second(theta, nsig, model)
            theta = np.clip(theta, 0, 1)
            b = b - (x - theta)
            if show_iqa and X_orig is not None:
                psnr_all.append(psnr(X_orig, x))
    ps, ss = _frame_iqa(X_orig, x)
    return x, ps, ss, psnr_all


def admm_joint_denoise(y, Phi_sum, A, At, second, x0=None, X_orig=None, denoiser='tv+ffdnet',
                       iter_max1=50, iter_max2=50, sigma1=None, sigma2=None, **args):
    """Two periods (joint_pnp_sci_algo.py:81-98): the joint module's ADMM-TV, then
    ``admm_multistep_denoise`` started from its result."""
    x, _, _, _ = joint_admm_denoise(y, Phi_sum, A, At, x0=x0, X_orig=X_orig, denoiser='tv',
                                    iter_max=iter_max1, sigma=sigma1, **args)
    return admm_multistep_denoise(y, Phi_sum, A, At, second, x0=x, X_orig=X_orig, denoiser=denoiser,
                                  iter_max=iter_max2, sigma=sigma2, **args)


# -- R7 -----------------------------------------------------------------------

def admmdenoise_cacti(meas, mask, A, At, projmeth='admm', v0=None, orig=None,
                      iframe=0, nframe=1, MAXB=1., maskdirection='plain',
                      **args):
    """Coded-frame loop around the solvers  (pnp_sci_algo.py:479-534).

    ``orig=None`` raises UnboundLocalError in the reference (``orig_k`` is
    never bound, :501-502,:513); here it is passed on as ``X_orig=None`` the
    way ``joint_pnp_sci_algo.py:38-41`` does.
    """
    nmask = mask.shape[-1]
    mask_sum = phi_sum(mask)
    x_ = np.zeros((*mask.shape[:-1], nmask * nframe), dtype=np.float32)
    psnr_, ssim_, psnrall_ = [], [], []
    t0 = time.time()
    t_ = 0.
    md = maskdirection.lower()
    for kf in range(nframe):
        orig_k = None
        if orig is not None:
            orig_k = orig[..., (kf + iframe) * nmask:(kf + iframe + 1) * nmask] / MAXB
        meas_k = meas[..., kf + iframe] / MAXB
        flip = (md == 'updown' and (kf + iframe) % 2 == 1) or \
               (md == 'downup' and (kf + iframe) % 2 == 0)
        v0_k = None
        if v0 is not None:
            v0_k = v0[:, :, kf * nmask:(kf + 1) * nmask]
            if flip:
                v0_k = v0_k[..., ::-1]
        pm = projmeth.lower()
        if pm == 'admm':
            x_k, p_k, s_k, pa_k = admm_denoise(meas_k, mask_sum, A, At,
                                               x0=v0_k, X_orig=orig_k, **args)
        elif pm == 'gap':
            x_k, p_k, s_k, pa_k = gap_denoise(meas_k, mask_sum, A, At,
                                              x0=v0_k, X_orig=orig_k, **args)
        else:
            raise ValueError('Unsupported projection method %s' % projmeth.upper())
        if flip:
            x_k = x_k[..., ::-1]
            p_k, s_k, pa_k = p_k[::-1], s_k[::-1], pa_k[::-1]
        t_ = time.time() - t0
        x_[..., kf * nmask:(kf + 1) * nmask] = x_k
        psnr_.extend(p_k)
        ssim_.extend(s_k)
        psnrall_.append(pa_k)
    return x_, t_, psnr_, ssim_, psnrall_


# -- R8 -----------------------------------------------------------------------

_BAYER = ((0, 0), (0, 1), (1, 0), (1, 1))     # pnp_sci_algo.py:99


def gap_denoise_bayer(y_bayer, Phi_bayer, _lambda=1, accelerate=True,
                      denoiser='tv', iter_max=50, noise_estimate=True,
                      sigma=None, tv_weight=0.1, tv_iter_max=5,
                      multichannel=True, x0_bayer=None, X_orig=None,
                      model=None, show_iqa=True):
    """Bayer GAP-TV  (pnp_sci_algo.py:20-265): four sub-lattice projections
    (:150-156) and one TV call over the [H/2, W/2, 4*Cr] stack (:163-166)."""
    if denoiser.lower() != 'tv':
        raise ValueError('Unsupported denoiser {}!'.format(denoiser))
    sigma, iter_max = _as_schedule(sigma, iter_max)
    nrow, ncol, nmask = Phi_bayer.shape
    h2, w2 = nrow // 2, ncol // 2
    f32 = np.float32
    yall = np.zeros([h2, w2, 4], dtype=f32)
    Phiall = np.zeros([h2, w2, nmask, 4], dtype=f32)
    Phi_sumall = np.zeros([h2, w2, 4], dtype=f32)
    xall = np.zeros([h2, w2, nmask, 4], dtype=f32)
    for ib, (r0, c0) in enumerate(_BAYER):
        yall[..., ib] = y_bayer[r0::2, c0::2]
        Phiall[..., ib] = Phi_bayer[r0::2, c0::2]
        Phi_sumall[..., ib] = phi_sum(Phiall[..., ib])
        if x0_bayer is None:
            xall[..., ib] = At_(yall[..., ib], Phiall[..., ib])
        else:
            xall[..., ib] = x0_bayer[r0::2, c0::2]
    y1all = np.zeros_like(yall)
    x_bayer = np.zeros_like(Phi_bayer)
    psnr_all = []
    for idx, _ in enumerate(sigma):
        for _it in range(iter_max[idx]):
            for ib in range(4):
                yb = A_(xall[..., ib], Phiall[..., ib])
                if accelerate:
                    y1all[..., ib] += (yall[..., ib] - yb)
                    xall[..., ib] += _lambda * (At_(
                        (y1all[..., ib] - yb) / Phi_sumall[..., ib], Phiall[..., ib]))
                else:
                    xall[..., ib] += _lambda * (At_(
                        (yall[..., ib] - yb) / Phi_sumall[..., ib], Phiall[..., ib]))
            v = xall.reshape([h2, w2, nmask * 4])
            v = denoise_tv_chambolle(v, tv_weight, n_iter_max=tv_iter_max,
                                     multichannel=multichannel)
            xall = v.reshape([h2, w2, nmask, 4])
            if show_iqa and X_orig is not None:
                for ib, (r0, c0) in enumerate(_BAYER):
                    x_bayer[r0::2, c0::2] = xall[..., ib]
                psnr_all.append(compare_psnr(X_orig, x_bayer, data_range=1.))
    for ib, (r0, c0) in enumerate(_BAYER):
        x_bayer[r0::2, c0::2] = xall[..., ib]
    ps, ss = [], []
    if X_orig is not None:
        for c in range(nmask):
            ps.append(compare_psnr(X_orig[:, :, c], x_bayer[:, :, c], data_range=1.))
            ss.append(compare_ssim(X_orig[:, :, c], x_bayer[:, :, c], data_range=1.))
    return x_bayer, ps, ss, psnr_all


# -- R9 (spec only in the reference: DeSCI/test_desci_cassi.m:53-75) ----------

def admm_denoise_bayer(y_bayer, Phi_bayer, _lambda=1, gamma=0.01, denoiser='tv', iter_max=50,
                       noise_estimate=True, sigma=None, tv_weight=0.1, tv_iter_max=5, multichannel=True,
                       x0_bayer=None, X_orig=None, model=None, show_iqa=True):
    """Bayer ADMM-TV (pnp_sci_algo.py:268-475).  PARITY UNPINNED: the reference's body raises NameError at :399
    (``ball`` is never bound; :389 binds ``b`` and the sub-lattice loop then rebinds it), so it cannot be run.
    Restated with the one evident repair -- ``ball = np.zeros_like(x0all)`` -- and otherwise statement for
    statement: set-up :347-385, projection per sub-lattice :397-399, one TV call over the stacked channels
    :402-406, multiplier update :427, PSNR of the merged mosaic :429-433, return :475."""
    if denoiser.lower() != 'tv':
        raise ValueError('Unsupported denoiser {}!'.format(denoiser))
    bayer = [[0, 0], [0, 1], [1, 0], [1, 1]]                              # :347
    sigma, iter_max = _as_schedule(sigma, iter_max)
    (nrow, ncol, nmask) = Phi_bayer.shape
    yall = np.zeros([nrow // 2, ncol // 2, 4], dtype=np.float32)
    Phiall = np.zeros([nrow // 2, ncol // 2, nmask, 4], dtype=np.float32)
    Phi_sumall = np.zeros([nrow // 2, ncol // 2, 4], dtype=np.float32)
    x0all = np.zeros([nrow // 2, ncol // 2, nmask, 4], dtype=np.float32)
    for ib, bb in enumerate(bayer):
        yall[..., ib] = y_bayer[bb[0]::2, bb[1]::2]
        Phiall[..., ib] = Phi_bayer[bb[0]::2, bb[1]::2]
        Phib_sum = np.sum(Phiall[..., ib], axis=2)
        Phib_sum[Phib_sum == 0] = 1
        Phi_sumall[..., ib] = Phib_sum
        if x0_bayer is None:
            x0all[..., ib] = At_(yall[..., ib], Phiall[..., ib])
        else:
            x0all[..., ib] = x0_bayer[bb[0]::2, bb[1]::2]
    xall = x0all.copy()                 # the reference aliases xall, thetaall and x0all (:386-387); the first
    thetaall = x0all                    # projection reads theta of every sub-lattice before it writes x of it
    x_bayer = np.zeros_like(Phi_bayer)
    ball = np.zeros_like(x0all)         # the repair
    psnr_all = []
    for idx, _ in enumerate(sigma):
        for _it in range(iter_max[idx]):
            for ib in range(4):
                yb = A_(thetaall[..., ib] + ball[..., ib], Phiall[..., ib])
                xall[..., ib] = thetaall[..., ib] + ball[..., ib] + _lambda * (
                    At_((yall[..., ib] - yb) / (Phi_sumall[..., ib] + gamma), Phiall[..., ib]))
            v = (xall - ball).reshape([nrow // 2, ncol // 2, nmask * 4])
            v = denoise_tv_chambolle(v, tv_weight, n_iter_max=tv_iter_max, multichannel=multichannel)
            thetaall = v.reshape([nrow // 2, ncol // 2, nmask, 4])
            ball = ball - (xall - thetaall)
            if show_iqa and X_orig is not None:
                for ib, bb in enumerate(bayer):
                    x_bayer[bb[0]::2, bb[1]::2] = xall[..., ib]
                psnr_all.append(psnr(X_orig, x_bayer))
    for ib, bb in enumerate(bayer):
        x_bayer[bb[0]::2, bb[1]::2] = xall[..., ib]
    return x_bayer, psnr_all


def cassi_shift_mask(mask2d, nband, step):
    """Shifted mask stack of a single-disperser CASSI system:
    ``Phi[h, w + step*k, k] = M[h, w]`` on a canvas [H, W+(nband-1)*step, nband].
    The reference's CASSI data arrive pre-shifted (``toy31_cassi.mat``); its
    operators are the same elementwise ``A_xy``/``At_xy_nonorm``."""
    H, W = mask2d.shape
    Phi = np.zeros((H, W + (nband - 1) * step, nband), dtype=mask2d.dtype)
    for k in range(nband):
        Phi[:, step * k:step * k + W, k] = mask2d
    return Phi


def cassi_shift_cube(cube, step):
    """Place band k of ``cube[H, W, nband]`` at column offset ``step*k``."""
    H, W, nband = cube.shape
    out = np.zeros((H, W + (nband - 1) * step, nband), dtype=cube.dtype)
    for k in range(nband):
        out[:, step * k:step * k + W, k] = cube[:, :, k]
    return out
