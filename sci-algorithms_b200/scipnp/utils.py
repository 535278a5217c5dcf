"""Drop-in for the hot-path part of the reference's ``PnP_SCI/python/utils.py``
(``A_``, ``At_``, ``psnr``; utils.py:10-36), computed on the GPU through
libscipnp.so.  NumPy in -> NumPy out; CUDA tensors in -> CUDA tensors out.
The result helpers of the reference module (``show_n_save_res`` utils.py:41-125,
``save_rgb_img`` :128-146, ``cli_run`` :150-181, ``rescale`` :183-184) are host-side
bookkeeping; they are provided so that the reference's drivers import unchanged
(``from utils import (A_, At_, show_n_save_res)``, pnp_sci_test_orig.py:22).  Figures
need matplotlib and image files need OpenCV -- both optional, imported on use.
"""
import ctypes as C
import math
import os
from statistics import mean

import numpy as np
import torch

from ._lib import lib, check
from .engine import to_device, stream_ptr, is_torch, dptr

__all__ = ["A_", "At_", "psnr", "phi_sum", "show_n_save_res", "save_rgb_img", "cli_run", "rescale"]


def _ret(t, like):
    return t if is_torch(like) else t.cpu().numpy()


def A_(x, Phi):
    """Forward model ``y = sum_c x[:,:,c]*Phi[:,:,c]`` (utils.py:10-15)."""
    xd, pd = to_device(x), to_device(Phi)
    if xd.shape != pd.shape or xd.dim() != 3:
        raise ValueError("A_ expects x and Phi of equal shape [H, W, C]")
    H, W, Cc = xd.shape
    y = torch.empty((H, W), dtype=torch.float32, device=xd.device)
    check(lib.scipnp_A(dptr(xd), dptr(pd), dptr(y), 1, H, W, Cc, 0, stream_ptr()))
    return _ret(y, x)


def At_(y, Phi):
    """Adjoint ``x[:,:,c] = y*Phi[:,:,c]`` (utils.py:17-26)."""
    yd, pd = to_device(y), to_device(Phi)
    if pd.dim() != 3 or tuple(yd.shape) != tuple(pd.shape[:2]):
        raise ValueError("At_ expects y [H, W] and Phi [H, W, C]")
    H, W, Cc = pd.shape
    x = torch.empty((H, W, Cc), dtype=torch.float32, device=pd.device)
    check(lib.scipnp_At(dptr(yd), dptr(pd), dptr(x), 1, H, W, Cc, 0, stream_ptr()))
    return _ret(x, y)


def phi_sum(Phi):
    """``sum_c Phi`` with zeros replaced by one (pnp_sci_algo.py:491-492)."""
    pd = to_device(Phi)
    H, W, Cc = pd.shape
    s = torch.empty((H, W), dtype=torch.float32, device=pd.device)
    check(lib.scipnp_phi_sum(dptr(pd), dptr(s), 1, H, W, Cc, stream_ptr()))
    return _ret(s, Phi)


def psnr(ref, img):
    """PSNR on [0,1] data, 100 when identical (utils.py:28-36)."""
    a, b = to_device(ref), to_device(img)
    if a.shape != b.shape:
        raise ValueError("psnr expects arrays of equal shape")
    acc = torch.zeros(1, dtype=torch.float64, device=a.device)
    check(lib.scipnp_sq_err(dptr(a), dptr(b), a.numel(), dptr(acc), stream_ptr()))
    mse = float(acc.item()) / a.numel()
    if mse == 0:
        return 100
    return 20 * math.log10(1. / math.sqrt(mse))


# -- result helpers of the drivers (host side) ---------------------------------------------------------

def _frame_grid(plt, frames, titles, path, cols):
    """One figure with a grey-level panel per frame, saved to ``path``."""
    n = frames.shape[2]
    fig = plt.figure(figsize=(12, 6.5))
    for k in range(n):
        ax = fig.add_subplot(max(n // cols, 1), cols, k + 1)
        ax.imshow(frames[:, :, k], cmap='gray', vmin=0, vmax=1)
        ax.axis('off')
        ax.set_title(titles[k], fontsize=12)
    fig.subplots_adjust(wspace=0.02, hspace=0.02, bottom=0, top=1, left=0, right=1)
    fig.savefig(path)
    plt.close(fig)


def show_n_save_res(vdenoise, tdenoise, psnr_denoise, ssim_denoise, psnrall_denoise, orig, Cr, resultsdir,
                    save_name, iframe=0, nframe=1, MAXB=255, show_res_flag=1, save_res_flag=1, **kwargs):
    """Figures under ``resultsdir/savedfig/`` and a MATLAB file under ``resultsdir/savedmat/`` with the
    reference's file names and variable names (utils.py:41-125)."""
    if show_res_flag:
        import matplotlib
        matplotlib.use(os.environ.get("MPLBACKEND", "Agg"))
        import matplotlib.pyplot as plt
        figdir = resultsdir + '/savedfig/'
        os.makedirs(figdir, exist_ok=True)
        cols = max(Cr // 2, 1)
        for kf in range(nframe):
            k0 = (kf + iframe) * Cr
            stem = '{}{}_kmeas{:d}'.format(figdir, save_name, kf + iframe)
            if orig is not None:
                _frame_grid(plt, orig[:, :, k0:k0 + Cr] / MAXB,
                            ['Ground truth: Frame #{0:d}'.format(k0 + nt + 1) for nt in range(Cr)],
                            stem + '_orig.png', cols)
                titles = ['Frame #{0:d} ({1:2.2f} dB)'.format(k0 + nt + 1, psnr_denoise[nt]) for nt in range(Cr)]
            else:
                titles = ['Frame #{0:d})'.format(k0 + nt + 1) for nt in range(Cr)]
            _frame_grid(plt, vdenoise[:, :, kf * Cr:(kf + 1) * Cr], titles, stem + '_vdenoise.png', cols)
            if orig is not None:
                fig = plt.figure()
                plt.plot(psnrall_denoise[kf], 'r')
                fig.savefig(stem + '_psnr_all.png')
                plt.close(fig)
        if orig is not None:
            fig = plt.figure()
            plt.plot(psnr_denoise)
            fig.savefig('{}{}_psnr_framewise.png'.format(figdir, save_name))
            plt.close(fig)
    if save_res_flag:
        import scipy.io as sio
        matdir = resultsdir + '/savedmat/'
        os.makedirs(matdir, exist_ok=True)
        path = '{}{}_kmeas{:d}_{:d}.mat'.format(matdir, save_name, iframe, iframe + nframe - 1)
        print('Results saved to: {}\n'.format(path))
        rec = {'vdenoise': vdenoise, 'tdenoise': tdenoise, 'iframe': iframe, 'nframe': nframe, 'Cr': Cr}
        if orig is not None:
            rec.update(orig=orig, psnr_denoise=psnr_denoise, ssim_denoise=ssim_denoise,
                       psnrall_denoise=psnrall_denoise, psnr_mean=mean(psnr_denoise))
        rec.update(kwargs)
        sio.savemat(path, rec)


def rescale(data):
    """Affine map of ``data`` onto [0, 1] (utils.py:183-184)."""
    lo, hi = np.min(data), np.max(data)
    return (data - lo) / (hi - lo)


def save_rgb_img(img, save_dir, prefix='img', save_format='.jpg', rescale_ch=False):
    """``img`` [H, W, 3, N] in [0, 1] -> N image files (utils.py:128-146)."""
    import cv2
    os.makedirs(save_dir, exist_ok=True)
    if rescale_ch:
        for ch in range(3):
            img[:, :, ch, :] = rescale(img[:, :, ch, :])
    for k in range(img.shape[-1]):
        bgr = np.uint8(img[:, :, :, k] * 255)[..., ::-1]
        cv2.imwrite(os.path.join(save_dir, prefix + '%04d' % k + save_format), bgr)
    print('images saved to: ', save_dir)


def cli_run(script_name, orig_name, scale, Cr, mask_name, test_algo_flag, root_dir='.', result_path='/results/tmp',
            iframe=0, nframe=1, MAXB=255, show_res_flag=0, save_res_flag=0, log_result_flag=0,
            gaussian_noise_level=0, poisson_noise=0, gamma=0, tv_weight=None, iter_max1=0, sigma1=0,
            iter_max2=[0], sigma2=[0]):
    """Run a test script with the reference's command-line flags (utils.py:150-181)."""
    flags = [('orig_name', orig_name), ('scale', scale), ('Cr', Cr), ('mask_name', mask_name),
             ('test_algo_flag', test_algo_flag), ('root_dir', root_dir), ('result_path', result_path),
             ('iframe', iframe), ('nframe', nframe), ('MAXB', MAXB), ('show_res_flag', show_res_flag),
             ('save_res_flag', save_res_flag), ('log_result_flag', log_result_flag),
             ('gaussian_noise_level', gaussian_noise_level), ('poisson_noise', poisson_noise), ('gamma', gamma),
             ('tv_weight', '{:.2f}'.format(tv_weight)), ('iter_max1', iter_max1), ('sigma1', sigma1),
             ('iter_max2', ' '.join(str(v) for v in iter_max2)),
             ('sigma2', ' '.join('{:.4f}'.format(v) for v in sigma2))]
    command_str = 'python {} '.format(script_name) + ' '.join('--{} {}'.format(k, v) for k, v in flags)
    print(command_str)
    os.system(command_str)
