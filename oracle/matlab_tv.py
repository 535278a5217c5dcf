"""Oracle (test infrastructure -- only tests/, smoke() and bench.py's cpu legs may import this):
NumPy restatement of the MATLAB twin's default TV denoiser, SURVEY.md section 8f-2.

    TV_denoising            PnP_SCI/matlab/algorithms/tvdenoisers/TV_denoising.m:1-44  (helpers :51-68)
    TV_denoising_clip_LB    .../TV_denoising_clip_LB.m:1-46 (3-D branch :25-36; helpers :50-85)
    tvdenoise_cham_ATV2D    .../tvdenoise_cham_ATV2D.m:45-91   (3-D branch :72-87)
    tvdenoise_cham_ITV2D    .../tvdenoise_cham_ITV2D.m:42-94   (3-D branch :73-90)
    tvdenoise_cham_ITV3D    .../tvdenoise_cham_ITV3D.m:40-94   (3-D branch :72-90)
    fgp_denoise_ATV2D / ITV2D / ITV3D   .../fgp_denoise_*.m:49-168 (loop :73-121, Lforward_3d / Ltrans_3d :127-168)
The TV family of gapdenoise.m:86-108 (`tvm` = ATV_ClipA, ATV_ClipB, ATV_cham, ATV_FGP, ITV2D_cham, ITV2D_FGP,
ITV3D_cham, ITV3D_FGP), all on [H, W, F] stacks as gapdenoise.m calls them.

Parity unpinned: there is no MATLAB or Octave in this image, so this file was never run against
the .m source; it follows it line by line (cited below) and is checked by properties
(tests/test_oracle_tv.py): mean preservation, constants are fixed points, lambda -> 0 is the
identity, the adjoint pairs (dh, dht) / (dv, dvt), frames are independent.
Arithmetic follows the input dtype (float32 in -> float32 throughout, as MATLAB does for single).
"""
import math

import numpy as np

__all__ = ["TV_denoising", "dh", "dv", "dht", "dvt", "TV_denoising_clip_LB", "tvdenoise_cham_ATV2D",
           "tvdenoise_cham_ITV2D", "tvdenoise_cham_ITV3D", "fgp_denoise_ATV2D", "fgp_denoise_ITV2D",
           "fgp_denoise_ITV3D"]

ALPHA = 5                                            # TV_denoising.m:8


def dv(x):                                           # :51-52   diff(x)
    return x[1:] - x[:-1]


def dh(x):                                           # :55-56   diff(x,1,2)
    return x[:, 1:] - x[:, :-1]


def dvt(z):                                          # :59-60 / :67-68 (3-D)   [-z(1,:); -diff(z); z(end,:)]
    return np.concatenate([-z[:1], -(z[1:] - z[:-1]), z[-1:]], axis=0)


def dht(z):                                          # :63-64 / :71-72 (3-D)   [-z(:,1) -diff(z,1,2) z(:,end)]
    return np.concatenate([-z[:, :1], -(z[:, 1:] - z[:, :-1]), z[:, -1:]], axis=1)


def _clip(x, lam):                                   # :83-84   sign(x).*min(abs(x),lambda)
    return np.sign(x) * np.minimum(np.abs(x), lam)


def TV_denoising(y0, lam, iters=100):
    """2-D image [H, W] or stack of frames [H, W, F] (TV per frame, :20-30)."""
    y0 = np.asarray(y0)
    if y0.ndim not in (2, 3) or y0.shape[0] < 2 or y0.shape[1] < 2:
        raise ValueError("TV_denoising restated for [H, W] and [H, W, F] inputs with H, W >= 2")
    ft = y0.dtype if y0.dtype in (np.float32, np.float64) else np.float64
    y0 = y0.astype(ft, copy=False)
    c = ft.type(1.0 / ALPHA)
    half = ft.type(lam / 2.0)
    zh = np.zeros((y0.shape[0], y0.shape[1] - 1) + y0.shape[2:], ft)       # :14-15 / :22-23
    zv = np.zeros((y0.shape[0] - 1, y0.shape[1]) + y0.shape[2:], ft)
    x0 = y0
    for _ in range(int(iters)):
        x0h = y0 - dht(zh)                                                  # :17 / :25
        x0v = y0 - dvt(zv)                                                  # :18 / :26
        x0 = (x0h + x0v) / ft.type(2)                                       # :19 / :27
        zh = _clip(zh + c * dh(x0), half)                                   # :20 / :28
        zv = _clip(zv + c * dv(x0), half)                                   # :21 / :29
    return x0


# -- the rest of the family (gapdenoise.m:86-108) ------------------------------------------------------------------

def _ft(a):
    a = np.asarray(a)
    ft = a.dtype if a.dtype in (np.float32, np.float64) else np.dtype(np.float64)
    if a.ndim != 3 or a.shape[0] < 2 or a.shape[1] < 2:
        raise ValueError("restated for [H, W, F] stacks with H, W >= 2 (the 3-D branch gapdenoise.m uses)")
    return a.astype(ft, copy=False), ft


def TV_denoising_clip_LB(y0, lam, iters=20):
    """'ATV_ClipB' (gapdenoise.m:95-96): the 3-D branch :25-36 -- no averaging of the two half-steps and the clip
    level is lambda, not lambda/2."""
    y0, ft = _ft(y0)
    c = ft.type(1.0 / ALPHA)                                               # alpha = 5, :8
    lam = ft.type(lam)
    zh = np.zeros((y0.shape[0], y0.shape[1] - 1, y0.shape[2]), ft)         # :27-28
    zv = np.zeros((y0.shape[0] - 1, y0.shape[1], y0.shape[2]), ft)
    x0 = y0
    for _ in range(int(iters)):
        x0 = y0 - dht(zh) - dvt(zv)                                        # :33
        zh = _clip(zh + c * dh(x0), lam)                                   # :34
        zv = _clip(zv + c * dv(x0), lam)                                   # :35
    return x0


def _cham(f, lam, iters, dt, kind):
    """Common body of the three Getreuer-style Chambolle variants (3-D branch).  Index vectors :54-57:
    id/ir clamp at the last row/column (forward difference, zero there), iu/il = [1, 1:N-1] repeat the FIRST
    row/column, so the backward difference of the divergence vanishes at the first row/column."""
    f, ft = _ft(f)
    lam, dt, one = ft.type(lam), ft.type(dt), ft.type(1)
    p1 = np.zeros_like(f)
    p2 = np.zeros_like(f)
    divp = np.zeros_like(f)
    for _ in range(int(iters)):
        z = divp - f * lam                                                 # z = divp - f*lambda
        z1 = np.concatenate([z[:, 1:], z[:, -1:]], axis=1) - z             # z(:,ir,:) - z
        z2 = np.concatenate([z[1:], z[-1:]], axis=0) - z                   # z(id,:,:) - z
        if kind == "itv2d":
            denom = one + dt * np.sqrt(z1 ** 2 + z2 ** 2)                  # ITV2D :83
            p1 = (p1 + dt * z1) / denom
            p2 = (p2 + dt * z2) / denom
        elif kind == "itv3d":
            denom = one + dt * np.sqrt(np.sum(z1 ** 2 + z2 ** 2, axis=2))  # ITV3D :82-83 (sum over the frames)
            denom = denom[:, :, None]
            p1 = (p1 + dt * z1) / denom
            p2 = (p2 + dt * z2) / denom
        else:
            t1 = p1 + dt * z1                                              # ATV2D :83-84
            t2 = p2 + dt * z2
            p1 = t1 / np.maximum(one, np.abs(t1))
            p2 = t2 / np.maximum(one, np.abs(t2))
        divp = p1 - np.concatenate([p1[:, :1], p1[:, :-1]], axis=1) + p2 - np.concatenate([p2[:1], p2[:-1]], axis=0)
    return f - divp / lam                                                  # u = f - divp/lambda


def tvdenoise_cham_ATV2D(f, lam, iters):
    return _cham(f, lam, iters, 1.0 / 8, "atv2d")                          # dt = 1/8, :49


def tvdenoise_cham_ITV2D(f, lam, iters):
    return _cham(f, lam, iters, 1.0 / 8, "itv2d")                          # dt = 1/8, :52


def tvdenoise_cham_ITV3D(f, lam, iters):
    return _cham(f, lam, iters, 1.0 / 4, "itv3d")                          # dt = 1/4, :50


def _lforward(P1, P2):                                                     # Lforward_3d, fgp_denoise_*.m:127-150
    m, n, B = P2.shape[0], P1.shape[1], P1.shape[2]
    X = np.zeros((m, n, B), P1.dtype)
    X[:m - 1] = P1
    X[:, :n - 1] = X[:, :n - 1] + P2
    X[1:] = X[1:] - P1
    X[:, 1:] = X[:, 1:] - P2
    return X


def _fgp(Xobs, lam, maxiter, kind):
    """Fast gradient projection (Beck & Teboulle) as in fgp_denoise_*.m:56-121.  `count` is never incremented
    (:117-119 are commented out), so the loop runs exactly MAXITER times; X_den is the D of the LAST iteration,
    computed before that iteration's dual update."""
    Xobs, ft = _ft(Xobs)
    m, n, B = Xobs.shape
    lam = ft.type(lam)
    c = ft.type(1.0) / (ft.type(8) * lam)                                  # 1/(8*lambda)
    one = ft.type(1)
    P1 = np.zeros((m - 1, n, B), ft); P2 = np.zeros((m, n - 1, B), ft)
    R1 = np.zeros_like(P1); R2 = np.zeros_like(P2)
    tkp1 = 1.0
    D = np.zeros_like(Xobs)
    for _ in range(int(maxiter)):
        P1o, P2o = P1, P2
        tk = tkp1
        D = Xobs - lam * _lforward(R1, R2)                                 # :85
        Q1 = D[:m - 1] - D[1:]                                             # Ltrans_3d :164-165
        Q2 = D[:, :n - 1] - D[:, 1:]
        P1 = R1 + c * Q1                                                   # :90-91
        P2 = R2 + c * Q2
        if kind == "atv2d":
            P1 = P1 / np.maximum(np.abs(P1), one)                          # ATV2D :95-96
            P2 = P2 / np.maximum(np.abs(P2), one)
        else:
            A = np.concatenate([P1, np.zeros((1, n, B), ft)], axis=0) ** 2 + \
                np.concatenate([P2, np.zeros((m, 1, B), ft)], axis=1) ** 2  # :95
            if kind == "itv3d":
                A = np.sqrt(np.maximum(np.sum(A, axis=2), one))[:, :, None]    # ITV3D :96-97
                A = np.broadcast_to(A, (m, n, B))
            else:
                A = np.sqrt(np.maximum(A, one))                            # ITV2D :96
            P1 = P1 / A[:m - 1]
            P2 = P2 / A[:, :n - 1]
        tkp1 = (1 + math.sqrt(1 + 4 * tk ** 2)) / 2                        # :106 (double scalars)
        w = ft.type((tk - 1) / tkp1)
        R1 = P1 + w * (P1 - P1o)                                           # :108-109
        R2 = P2 + w * (P2 - P2o)
    return D


def fgp_denoise_ATV2D(Xobs, lam, maxiter):
    return _fgp(Xobs, lam, maxiter, "atv2d")


def fgp_denoise_ITV2D(Xobs, lam, maxiter):
    return _fgp(Xobs, lam, maxiter, "itv2d")


def fgp_denoise_ITV3D(Xobs, lam, maxiter):
    return _fgp(Xobs, lam, maxiter, "itv3d")
