"""Synthetic SCI scenes, masks and measurements (host side, NumPy).

The reference ships no data (every ``.mat`` is a Git-LFS pointer), so all
inputs are synthetic.  Recipes follow the reference's own generators:
Bernoulli(0.5) binary masks (``[dataset]/#code/binary_mask.m:63-68``) and
``meas = sum_c mask*orig`` (``[dataset]/#code/gen_data.m:134-152``).  Seeds and
shapes are fixed by SURVEY.md section 8d.

Not on the hot path: this only feeds tests and ``bench.py``.
"""
import numpy as np

__all__ = ["binary_mask", "moving_scene", "cacti_measure", "make_cacti",
           "make_bayer", "make_cassi", "CONFIGS"]

# name -> (H, W, Cr, nframe, seed-offset); sizes from BASELINE.json configs
CONFIGS = {
    "c1_kobe_like": dict(H=256, W=256, C=8, F=4, cfg=1),
    "c2_gray6": dict(H=256, W=256, C=8, F=28, cfg=2),
    "c3_bayer": dict(H=512, W=512, C=24, F=1, cfg=3),
    "c4_cassi": dict(H=256, W=256, C=28, F=1, cfg=4, step=2),
    "c5_uhd": dict(H=2160, W=3840, C=24, F=1, cfg=5),
}


def binary_mask(H, W, C, seed):
    rng = np.random.default_rng(seed)
    return (rng.random((H, W, C), dtype=np.float32) <= 0.5).astype(np.float32)


def moving_scene(H, W, nframes, seed=None, texture=0.0, t0=0):
    """Smooth scene in [0,255]: a low-frequency sinusoid drifting 3 px/frame
    plus a bright disc moving 4 px/frame; optional Gaussian texture."""
    yy = np.arange(H, dtype=np.float32)[:, None]
    xx = np.arange(W, dtype=np.float32)[None, :]
    out = np.empty((H, W, nframes), dtype=np.float32)
    rad = 0.12 * min(H, W)
    for t in range(nframes):
        tt = t + t0
        bg = 110. + 60. * np.sin(2 * np.pi * (xx + 3. * tt) / (0.45 * W)) \
            * np.cos(2 * np.pi * yy / (0.6 * H))
        cx = (0.25 * W + 4. * tt) % W
        cy = (0.35 * H + 2. * tt) % H
        disc = ((xx - cx) ** 2 + (yy - cy) ** 2) <= rad * rad
        out[:, :, t] = np.where(disc, 230., bg)
    if texture > 0:
        rng = np.random.default_rng(seed)
        out += rng.normal(0., texture, size=out.shape).astype(np.float32)
    np.clip(out, 0., 255., out=out)
    return out


def cacti_measure(orig, mask):
    """meas[:,:,k] = sum_c mask[:,:,c]*orig[:,:,k*C+c]."""
    H, W, C = mask.shape
    F = orig.shape[2] // C
    meas = np.empty((H, W, F), dtype=np.float32)
    for k in range(F):
        meas[:, :, k] = np.sum(mask * orig[:, :, k * C:(k + 1) * C], axis=2)
    return meas


def make_cacti(H, W, C, F=1, cfg=1, texture=2.0):
    """Returns (meas[H,W,F], mask[H,W,C], orig[H,W,C*F]) in [0,255] units."""
    mask = binary_mask(H, W, C, 1000 + cfg)
    orig = moving_scene(H, W, C * F, seed=2000 + cfg, texture=texture)
    meas = cacti_measure(orig, mask)
    return meas, mask, orig


def make_bayer(H, W, C, cfg=3, texture=2.0):
    """RGGB mosaic of an RGB synthetic scene; returns (y_bayer[H,W],
    Phi_bayer[H,W,C], orig_bayer[H,W,C]) with orig in [0,1]."""
    mask = binary_mask(H, W, C, 1000 + cfg)
    gains = (1.0, 0.8, 0.8, 0.6)       # R, G, G, B
    base = moving_scene(H, W, C, seed=2000 + cfg, texture=texture) / 255.
    orig = np.empty_like(base)
    for ib, (r0, c0) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        orig[r0::2, c0::2] = base[r0::2, c0::2] * gains[ib]
    y = np.sum(mask * orig, axis=2).astype(np.float32)
    return y, mask, orig.astype(np.float32)


def make_cassi(H, W, nband, step=2, cfg=4, texture=1.0):
    """Single-disperser CASSI: returns (y[H,Wc], mask2d[H,W], cube_shift[H,Wc,nband])
    with Wc = W+(nband-1)*step and the cube in [0,1]."""
    rng = np.random.default_rng(1000 + cfg)
    mask2d = (rng.random((H, W), dtype=np.float32) <= 0.5).astype(np.float32)
    scene = moving_scene(H, W, 1, seed=2000 + cfg, texture=texture)[:, :, 0] / 255.
    lam = np.linspace(0., 1., nband, dtype=np.float32)
    spectra = 0.35 + 0.65 * np.exp(-((lam - 0.55) ** 2) / 0.08)
    cube = scene[:, :, None] * spectra[None, None, :]
    Wc = W + (nband - 1) * step
    shifted = np.zeros((H, Wc, nband), dtype=np.float32)
    y = np.zeros((H, Wc), dtype=np.float32)
    for k in range(nband):
        shifted[:, step * k:step * k + W, k] = cube[:, :, k]
        y[:, step * k:step * k + W] += mask2d * cube[:, :, k]
    return y, mask2d, shifted
