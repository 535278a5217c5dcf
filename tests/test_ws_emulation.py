"""Executable description of the warp-specialised kernel's consumer pipeline (csrc/gap_tv_ws.cuh: ws_step, general
path), in NumPy: a lane owns two horizontally adjacent pixels, a warp 64 pixels of which 4 per side are halo, the
right neighbour of a lane's second pixel and the left neighbour of its first pixel come by "shuffles" (lane 31 / lane 0
receive their own value: those lanes are halo), stage i of step t advances dual iteration i on row t-i-1, the last
stage takes f(t-R) from the f ring, rows above a segment are masked (warm-up), rows of the image edge get Chambolle's
boundary rules.  Segments come from the kernel's work split (tests/test_ws_split_cpu.py).

Against the oracle's Chambolle TV (no early stop) the emulation must agree to rounding for every shape, strip count
and dual-iteration count: the design check of the ownership rules, the masks and the warm-up depth.  CPU only, float64.
"""
import numpy as np
import pytest

from oracle.tv_chambolle import tv_chambolle_2d
from test_ws_split_cpu import seg_iter

GW, HALO, TAU = 64, 4, 0.25


def ws_segment(f, out, weight, R, grp, own, r0, r1):
    """One consumer warp (32 lanes x 2 pixels) walks rows [r0, r1) of pixel group `grp` of one channel."""
    H, W = f.shape
    lane = np.arange(32)
    pxa = grp * own - HALO + 2 * lane                       # pixel A of the lane; B = A + 1
    pair_in = ((pxa >= 0) & (pxa < W)).astype(float)        # W is even and the base is even: pairs are atomic
    right_in = pair_in * (pxa + 2 < W)
    own_lane = (pair_in > 0) & (lane >= HALO // 2) & (lane < HALO // 2 + own // 2)
    c = TAU / weight
    z = lambda: np.zeros((2, 32))
    o_prev = [z() for _ in range(R)]
    g1b = [np.zeros(32) for _ in range(R)]
    P0 = [z() for _ in range(R)]
    P1 = [z() for _ in range(R)]
    fd = [z() for _ in range(R - 1)]
    rs, t_end = max(r0 - R, 0), r1 + R
    fring = {}

    def frow(t):                                            # what the projection warps leave in the f tile
        v = np.zeros((2, 32))
        if t < H:
            for q in range(2):
                px = pxa + q
                ok = (px >= 0) & (px < W)
                v[q, ok] = f[t, px[ok]]
        return v

    for t in range(rs, t_end):
        f_new = frow(t)
        fring[t] = f_new
        f_old = fring[t - R] if t - R >= rs else z()
        o_new = f_new.copy()
        pi0, pi1 = z(), z()
        for i in range(R):
            row_new, u = t - i, t - i - 1
            m = pair_in * float(rs <= u < H)
            md = float(row_new < H)
            o_rb = np.r_[o_new[0, 1:], o_new[0, -1]]        # shfl_down: the next lane's A (lane 31: itself)
            g1 = np.stack([o_prev[i][1] - o_prev[i][0], g1b[i]])
            g0 = (o_new - o_prev[i]) * md
            r = m / (1.0 + c * np.sqrt(g0 ** 2 + g1 ** 2))
            if i == 0:
                pn0, pn1 = g0 * (r * -TAU), g1 * (r * -TAU)
            else:
                pn0, pn1 = (g0 * -TAU + pi0) * r, (g1 * -TAU + pi1) * r
            p1l_a = np.r_[pn1[1, 0], pn1[1, :-1]]           # shfl_up: the previous lane's B (lane 0: itself)
            d = np.stack([(P0[i][0] - pn0[0]) + (p1l_a - pn1[0]), (P0[i][1] - pn0[1]) + (pn1[0] - pn1[1])])
            g1b[i] = (o_rb - o_new[1]) * right_in
            o_next = (f_old if i == R - 1 else fd[i]) + d
            pi0, pi1 = P0[i], P1[i]
            P0[i], P1[i] = pn0, pn1
            o_prev[i] = o_new
            o_new = o_next
        fd = [f_new] + fd[:-1]
        orow = t - R
        if r0 <= orow < r1:
            for q in range(2):
                px = pxa + q
                out[orow, px[own_lane]] = o_new[q, own_lane]


def ws_tv(f, weight, T, grid, own=56, cost=8):
    H, W = f.shape
    R = T - 1
    ngroups = -(-W // own)
    out = np.full_like(f, np.nan)
    edge = 1 if ngroups > 2 else 0
    for segs in seg_iter(1, 0, H, ngroups, cost, edge, grid, False):
        for (_, grp, r0, r1) in segs:
            ws_segment(f, out, weight, R, grp, own, r0, r1)
    return out


@pytest.mark.parametrize("T", [3, 4, 5])
@pytest.mark.parametrize("shape,grid,own", [((37, 60), 5, 56), ((64, 130), 9, 56), ((9, 200), 148, 56), ((50, 116), 3, 52),
                                            ((23, 8), 2, 56)])
def test_ws_pipeline_equals_chambolle(T, shape, grid, own):
    rng = np.random.default_rng(T * 100 + shape[0])
    f = rng.random(shape)
    got = ws_tv(f, 0.3, T, grid, own=own)
    want = tv_chambolle_2d(f, 0.3, eps=0.0, n_iter_max=T)
    assert not np.isnan(got).any()                          # every pixel of every row was produced by its owner
    assert np.abs(got - want).max() < 1e-12


def test_warm_up_depth_is_exactly_R():
    """A segment that starts in the middle of the image is exact with R warm-up rows above it and not with R-1."""
    rng = np.random.default_rng(9)
    f = rng.random((40, 60))
    want = tv_chambolle_2d(f, 0.3, eps=0.0, n_iter_max=5)
    out = np.full_like(f, np.nan)
    ws_segment(f, out, 0.3, 4, 0, 56, 20, 30)
    assert np.abs(out[20:30, :56] - want[20:30, :56]).max() < 1e-12
