"""CPU check of the work split of the warp-specialised kernel (csrc/gap_tv_ws.cuh: WsSegIter, ws_slot, ws_groups).

The kernel deals "ticks" to the CTAs: a row of an interior strip weighs 16 ticks, a row of the first / last strip of a
scene 16 + edge, every strip is charged `cost` rows in front of its rows, a row belongs to the CTA that holds its first
tick, and with several measurements whose groups do not fill the strips the groups of all measurements are laid end to
end (packing).  This file restates that arithmetic in Python, statement by statement, and checks what the kernel
relies on: every (measurement, group, output row) is produced exactly once, for full scenes, row windows of tiles,
ragged sizes and packed batches; and the shares are balanced.
"""
import itertools

import pytest

TICK = 16      # kWsTick


def ws_groups(Q):                      # gap_tv_ws.cuh: ws_groups
    return 8 // Q if Q <= 4 else max(1, 12 // Q)


def seg_iter(B, out_lo, out_hi, nstrips, cost, edge, grid, pack):
    """WsSegIter for every CTA: list over CTAs of (b, strip, r0, r1)."""
    Hw = out_hi - out_lo
    wi, we, chg = TICK, TICK + edge, cost * TICK
    Li, Le = chg + Hw * wi, chg + Hw * we
    Tb = Le if nstrips == 1 else 2 * Le + (nstrips - 2) * Li
    total = (1 if pack else B) * Tb
    q, rem = divmod(total, grid)
    out = []
    for c in range(grid):
        unit = c * q + min(c, rem)
        end = unit + q + (1 if c < rem else 0)
        segs = []
        while unit < end:
            b = unit // Tb
            v = unit - b * Tb
            if v < Le:
                s, ln, w = 0, Le, we
            else:
                v -= Le
                s = 1 + v // Li
                if s >= nstrips - 1:
                    s, ln, w = nstrips - 1, Le, we
                    v -= (nstrips - 2) * Li
                else:
                    v -= (s - 1) * Li
                    ln, w = Li, wi
            v0 = v
            left = end - unit
            v1 = ln if ln - v0 < left else v0 + left
            unit += v1 - v0
            a0, a1 = max(v0 - chg, 0), max(v1 - chg, 0)
            r0, r1 = out_lo + (a0 + w - 1) // w, out_lo + (a1 + w - 1) // w
            if r1 > r0:
                segs.append((b, s, r0, r1))
        out.append(segs)
    return out


def ws_slot(pack, B, ngroups, b, strip, gi, ngrp):     # gap_tv_ws.cuh: ws_slot
    if not pack:
        grp = strip * ngrp + gi
        return b, grp, grp < ngroups
    G = strip * ngrp + gi
    bb = G // ngroups
    if bb >= B:
        return 0, ngroups, False
    return bb, G - bb * ngroups, True


def plan(B, H, W, C, own=56, cost=8, edge=1, grid=148, out_lo=0, out_hi=None, allow_pack=True):
    """What launch_fused_ws sets up (gap_tv_ws.cu) and what the kernel's roles derive from it."""
    Q = C // 2
    ngrp = ws_groups(Q)
    out_hi = H if out_hi is None else out_hi
    ngroups = -(-W // own)
    nstrips = -(-ngroups // ngrp)
    pack = allow_pack and B > 1 and ngroups % ngrp != 0
    if pack:
        nstrips = -(-(B * ngroups) // ngrp)
        edge = 0
    elif nstrips <= 2:
        edge = 0
    ctas = seg_iter(B, out_lo, out_hi, nstrips, cost, edge, grid, pack)
    produced = {}
    for c, segs in enumerate(ctas):
        for (b, s, r0, r1) in segs:
            for gi in range(ngrp):
                bb, grp, live = ws_slot(pack, B, ngroups, b, s, gi, ngrp)
                if not live:
                    continue
                for r in range(r0, r1):
                    produced[(bb, grp, r)] = produced.get((bb, grp, r), 0) + 1
    return ctas, produced, ngroups


@pytest.mark.parametrize("B,H,W,C", [(1, 2160, 3840, 24), (1, 278, 3840, 24), (28, 256, 256, 8), (4, 256, 256, 24),
                                     (1, 256, 256, 8), (3, 40, 128, 8), (2, 33, 100, 4), (5, 7, 60, 12), (1, 5, 64, 16)])
def test_every_row_of_every_group_is_produced_once(B, H, W, C):
    for grid, allow_pack in itertools.product((148, 37, 3), (True, False)):
        ctas, produced, ngroups = plan(B, H, W, C, grid=grid, allow_pack=allow_pack)
        assert len(produced) == B * ngroups * H
        assert set(produced.values()) == {1}


@pytest.mark.parametrize("lo,hi", [(4, 274), (0, 270), (4, 273)])
def test_row_window_of_a_tile(lo, hi):
    """Row-tiled mode: only the owned rows [out_lo, out_hi) are produced (the halo rows belong to the neighbours)."""
    ctas, produced, ngroups = plan(1, hi + 4, 3840, 24, out_lo=lo, out_hi=hi)
    rows = {r for (_, _, r) in produced}
    assert rows == set(range(lo, hi)) and set(produced.values()) == {1}
    assert len(produced) == ngroups * (hi - lo)


def test_shares_are_balanced_and_edge_strips_get_fewer_rows():
    ctas, _, _ = plan(1, 2160, 3840, 24)
    rows = [sum(r1 - r0 for (_, _, r0, r1) in segs) for segs in ctas]
    assert len(ctas) == 148 and min(rows) > 0
    # 69 strips: the CTAs of the first and the last strip carry pixel masks and get about 6 % fewer rows
    interior = [rows[c] for c, segs in enumerate(ctas) if all(0 < s < 68 for (_, s, _, _) in segs)]
    edge = [rows[c] for c, segs in enumerate(ctas) if all(s in (0, 68) for (_, s, _, _) in segs)]
    assert max(interior) - min(interior) <= 9          # one segment charge at most
    assert edge and max(edge) < min(interior)
    # at most three row segments per CTA on the UHD scene
    assert max(len(segs) for segs in ctas) <= 3


def test_packing_fills_the_group_slots():
    """28 measurements x 5 groups, two groups per CTA: 70 strips instead of 84."""
    assert ws_groups(4) == 2
    packed, prod_p, _ = plan(28, 256, 256, 8)
    plain, prod_u, _ = plan(28, 256, 256, 8, allow_pack=False)
    assert prod_p.keys() == prod_u.keys()
    rows_p = max(sum(r1 - r0 for (_, _, r0, r1) in segs) for segs in packed)
    rows_u = max(sum(r1 - r0 for (_, _, r0, r1) in segs) for segs in plain)
    assert rows_p < rows_u
