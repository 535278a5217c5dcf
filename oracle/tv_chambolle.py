"""Oracle (test infrastructure): NumPy restatement of the Chambolle TV denoiser.

Restates ``skimage.restoration.denoise_tv_chambolle`` as shipped in
scikit-image 0.17.2 (``skimage/restoration/_denoise.py``), the third-party
routine the reference calls at ``PnP_SCI/python/pnp_sci_algo.py:164,409,650,
812,874,895``.  scikit-image is absent from /root/reference and from this
image, so this file is the specification of R6 and is
**parity unpinned** by the reference itself (see ``oracle/__init__.py``).

Algorithm (A. Chambolle, "An algorithm for total variation minimization and
applications", JMIV 20, 2004), 2-D, isotropic, step tau = 1/(2*ndim) = 1/4:

    p^0 = 0
    for i = 0 .. T-1:
        out_i  = f                      (i == 0)
               = f + D(p^i)             (i  > 0)
        g      = forward differences of out_i (zero on the last row / column)
        p^{i+1}= (p^i - tau*g) / (1 + (tau/w)*|g|)
        E_i    = (sum D(p^i)^2 + w*sum|g|) / size
        stop if i > 0 and |E_{i-1} - E_i| < eps*E_0
    return out_i of the last executed iteration

    D(p)[r,c] = -(p0[r,c] + p1[r,c]) + p0[r-1,c] + p1[r,c-1]   (terms with a
    negative index are absent)

so ``n_iter_max = T`` applies at most T-1 dual updates to the returned image.

The arithmetic order of the array statements follows the published source so
that float32 rounding agrees; the scalar energy ``E`` is carried as a Python
float (NumPy 1.x value-based promotion, the stack the reference ran on).
"""
import numpy as np

__all__ = ["denoise_tv_chambolle", "tv_chambolle_2d"]


def tv_chambolle_2d(image, weight=0.1, eps=2.e-4, n_iter_max=200,
                    energy_out=None):
    """One 2-D ROF problem; ``image`` float array [H, W].

    ``energy_out`` (optional list) receives E_i for every executed iteration.
    """
    f = image
    H, W = f.shape
    dt = f.dtype
    tau = 1.0 / 4.0
    p0 = np.zeros((H, W), dtype=dt)   # dual component along axis 0 (rows)
    p1 = np.zeros((H, W), dtype=dt)   # dual component along axis 1 (cols)
    g0 = np.zeros((H, W), dtype=dt)
    g1 = np.zeros((H, W), dtype=dt)
    d = np.zeros((H, W), dtype=dt)
    out = f
    e_init = e_prev = 0.0
    i = 0
    while i < n_iter_max:
        if i > 0:
            d = -(p0 + p1)
            d[1:, :] += p0[:-1, :]
            d[:, 1:] += p1[:, :-1]
            out = f + d
        else:
            out = f
        energy = float((d ** 2).sum())
        g0[:-1, :] = out[1:, :] - out[:-1, :]
        g1[:, :-1] = out[:, 1:] - out[:, :-1]
        norm = np.sqrt(g0 ** 2 + g1 ** 2)
        energy += weight * float(norm.sum())
        norm *= dt.type(tau / weight)
        norm += dt.type(1.0)
        p0 -= dt.type(tau) * g0
        p1 -= dt.type(tau) * g1
        p0 /= norm
        p1 /= norm
        energy /= float(f.size)
        if energy_out is not None:
            energy_out.append(energy)
        if i == 0:
            e_init = energy
            e_prev = energy
        elif abs(e_prev - energy) < eps * e_init:
            break
        else:
            e_prev = energy
        i += 1
    return out


def denoise_tv_chambolle(image, weight=0.1, eps=2.e-4, n_iter_max=200,
                         multichannel=False, energy_out=None):
    """Same surface as the skimage (<0.19) function the reference imports.

    ``multichannel=True``: every slice ``image[..., c]`` is an independent 2-D
    problem.  ``multichannel=False`` on a 2-D array: one problem.  (The n-D
    coupled case, a 3-D array with ``multichannel=False``, is not on the
    reference's path and is rejected.)
    """
    image = np.asarray(image)
    if image.dtype.kind != 'f':
        raise TypeError("oracle covers float input only (the reference path)")
    if multichannel:
        if image.ndim != 3:
            raise ValueError("multichannel oracle expects [H, W, C]")
        out = np.zeros_like(image)
        for c in range(image.shape[-1]):
            e_c = [] if energy_out is not None else None
            out[..., c] = tv_chambolle_2d(image[..., c], weight, eps,
                                          n_iter_max, e_c)
            if energy_out is not None:
                energy_out.append(e_c)
        return out
    if image.ndim != 2:
        raise ValueError("single-channel oracle expects [H, W]")
    return tv_chambolle_2d(image, weight, eps, n_iter_max, energy_out)
