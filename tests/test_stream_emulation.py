"""Executable description of the fused kernel's algorithm (csrc/gap_tv_stream.cuh), in NumPy:

* row streaming with a one-row lag per dual iteration (step rho: stage i works on row rho-i, the
  dual variable of row rho-i-1 advances, out_R(rho-R) leaves the pipeline),
* 32-lane pixel groups that own 32-2R pixels, neighbours by "shuffles" (out from lane+1 with the
  last image column clamped, p1 from lane-1 with lane 0 receiving itself),
* row segments with R warm-up rows above and R drain rows below, masks at true image edges only,
* the balanced work split with strip boundaries charged kSegCost units.

Against the oracle's Chambolle TV (no early stop) the emulation must agree to rounding: this is the
design check the CUDA kernel was written from, kept as a regression of the masks, the ownership
rules and the work split.  CPU only, float64.
"""
import numpy as np
import pytest

from oracle.tv_chambolle import tv_chambolle_2d

LANES, SEG_COST, TAU = 32, 12, 0.25


def work_split(ngroups, H, grid):
    """(strip, r0, r1) row segments of every CTA; mirrors the kernel's unit mapping (NG = 1)."""
    Hv = H + SEG_COST
    total = ngroups * Hv
    per_cta = -(-total // grid)
    out = []
    for cta in range(grid):
        unit, unit_end = cta * per_cta, min(total, (cta + 1) * per_cta)
        segs = []
        while unit < unit_end:
            strip, v0 = divmod(unit, Hv)
            v1 = min(Hv, v0 + (unit_end - unit))
            unit += v1 - v0
            r0, r1 = max(v0 - SEG_COST, 0), v1 - SEG_COST
            if r1 > r0:
                segs.append((strip, r0, r1))
        out.append(segs)
    return out


def stream_segment(f, out, weight, R, grp, r0, r1):
    """One pixel group (32 lanes) walks the rows of one segment: gap_tv_stream_kernel's general path."""
    H, W = f.shape
    own_n = LANES - 2 * R
    lane = np.arange(LANES)
    px = grp * own_n - R + lane
    px_in = (px >= 0) & (px < W)
    own = px_in & (lane >= R) & (lane < LANES - R)
    src_right = np.where(px_in & (px < W - 1) & (lane < LANES - 1), lane + 1, lane)
    pxc = np.clip(px, 0, W - 1)
    c = TAU / weight
    z = lambda: np.zeros(LANES)
    o_prev = [z() for _ in range(R)]
    g1_prev = [z() for _ in range(R)]
    P0 = [z() for _ in range(R + 1)]
    P1 = [z() for _ in range(R + 1)]
    fd = [z() for _ in range(R)]
    rs, rend = max(0, r0 - R), r1 + R
    for rho in range(rs, rend):
        f_new = np.where(px_in, f[rho, pxc], 0.0) if rho < H else z()
        o_new = f_new
        pend0 = pend1 = None
        for i in range(R):
            row_new = rho - i
            u = row_new - 1
            m = (1.0 if (rs <= u < H) else 0.0) * px_in
            md = 1.0 if row_new < H else 0.0
            pi0, pi1 = P0[i], P1[i]
            if i > 0:
                P0[i], P1[i] = pend0, pend1
            o_right = o_new[src_right]
            g0 = (o_new - o_prev[i]) * md
            g1 = g1_prev[i]
            nrm = np.sqrt(g0 * g0 + g1 * g1)
            r = m / (1.0 + c * nrm)
            pn0 = (pi0 - TAU * g0) * r
            pn1 = (pi1 - TAU * g1) * r
            p1l = np.concatenate([pn1[:1], pn1[:-1]])            # __shfl_up_sync(.., 1): lane 0 keeps its own
            d = (P0[i + 1] - pn0) + (p1l - pn1)
            o_next = fd[i] + d
            g1_prev[i] = o_right - o_new
            o_prev[i] = o_new
            pend0, pend1 = pn0, pn1
            o_new = o_next
        P0[R], P1[R] = pend0, pend1
        fd = [f_new] + fd[:-1]
        orow = rho - R
        if r0 <= orow < r1:
            out[orow, px[own]] = o_new[own]


def stream_tv(f, weight, T, grid):
    R = T - 1
    H, W = f.shape
    own_n = LANES - 2 * R
    ngroups = -(-W // own_n)
    out = np.full_like(f, np.nan)
    for segs in work_split(ngroups, H, grid):
        for strip, r0, r1 in segs:
            stream_segment(f, out, weight, R, strip, r0, r1)
    return out


@pytest.mark.parametrize("T", [3, 4, 5])
@pytest.mark.parametrize("shape,grid", [((37, 45), 5), ((64, 24), 3), ((9, 70), 7), ((50, 33), 1), ((23, 100), 16)])
def test_streaming_pipeline_equals_chambolle(T, shape, grid):
    rng = np.random.default_rng(T * 100 + shape[0])
    f = rng.random(shape)
    want = tv_chambolle_2d(f, weight=0.3, eps=0.0, n_iter_max=T)
    got = stream_tv(f, 0.3, T, grid)
    assert not np.isnan(got).any()                         # every pixel has exactly one owner
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-13)


def test_work_split_covers_every_row_once_and_is_balanced():
    for ngroups, H, grid in ((160, 2160, 296), (160, 286, 296), (11, 256, 296), (3, 40, 5), (1, 7, 4)):
        seen = np.zeros((ngroups, H), int)
        loads = []
        for segs in work_split(ngroups, H, grid):
            cost = 0
            for strip, r0, r1 in segs:
                seen[strip, r0:r1] += 1
                cost += (r1 - r0) + SEG_COST                 # rows plus the charge of a segment start
            loads.append(cost)
        assert (seen == 1).all()
        per_cta = -(-(ngroups * (H + SEG_COST)) // grid)
        # a CTA's charged cost never exceeds its share by more than the one start that is not charged
        # (its own), and every CTA except the tail of the split is within two charges of that share
        assert max(loads) <= per_cta + SEG_COST
        full = [l for l in loads if l][:-2]
        assert all(l >= per_cta - 2 * SEG_COST for l in full)


# -- the whole outer iteration as the kernel forms it ---------------------------------------------

def _tv_stack(x, weight, T):
    return np.stack([tv_chambolle_2d(x[:, :, c], weight=weight, eps=0.0, n_iter_max=T)
                     for c in range(x.shape[2])], axis=2)


def test_gap_iteration_as_fused():
    """Phase A + stage 0 + pipeline: s = (y1_new - yb)/Phi_sum per pixel, f = x + lambda*s*Phi per
    channel, TV per channel (pnp_sci_algo.py:640-650)."""
    from oracle import pnp_sci as O
    rng = np.random.default_rng(2)
    H, W, C, T, lam = 21, 40, 4, 5, 0.75
    Phi = (rng.random((H, W, C)) <= 0.5).astype(np.float64)
    x = rng.random((H, W, C))
    y = rng.random((H, W)) * C / 2
    y1 = rng.random((H, W)) * 0.1
    ps = O.phi_sum(Phi)
    # reference statements
    yb = O.A_(x, Phi)
    y1_ref = y1 + (y - yb)
    want = _tv_stack(x + lam * O.At_((y1_ref - yb) / ps, Phi), 0.3, T)
    # kernel form: one scale per pixel from the full dot product, then channel by channel
    s = ((y1 + (y - (x * Phi).sum(2))) - (x * Phi).sum(2)) / ps
    got = np.stack([stream_tv(x[:, :, c] + (lam * s) * Phi[:, :, c], 0.3, T, grid=3) for c in range(C)], axis=2)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)


def test_admm_multiplier_comes_out_of_the_f_delay_line():
    """ADMM (pnp_sci_algo.py:808-836): with f = x - b = theta + lambda*s*Phi the TV input, the
    multiplier update b - (x - theta_new) is theta_new - f, so the fused kernel needs no copy of b
    beyond the projection."""
    from oracle import pnp_sci as O
    rng = np.random.default_rng(4)
    H, W, C, T, lam, gamma = 17, 33, 3, 4, 1.0, 0.01
    Phi = (rng.random((H, W, C)) <= 0.5).astype(np.float64)
    theta = rng.random((H, W, C))
    b = 0.1 * rng.standard_normal((H, W, C))
    y = rng.random((H, W)) * C / 2
    ps = O.phi_sum(Phi)
    yb = O.A_(theta + b, Phi)
    x = (theta + b) + lam * O.At_((y - yb) / (ps + gamma), Phi)
    theta_ref = _tv_stack(x - b, 0.3, T)
    b_ref = b - (x - theta_ref)
    s = (y - ((theta + b) * Phi).sum(2)) / (ps + gamma)
    f = theta + (lam * s)[:, :, None] * Phi
    theta_new = np.stack([stream_tv(f[:, :, c], 0.3, T, grid=2) for c in range(C)], axis=2)
    np.testing.assert_allclose(theta_new, theta_ref, rtol=0, atol=1e-12)
    np.testing.assert_allclose(theta_new - f, b_ref, rtol=0, atol=1e-12)
    np.testing.assert_allclose(f + b, x, rtol=0, atol=1e-12)          # the x that admm_denoise returns
