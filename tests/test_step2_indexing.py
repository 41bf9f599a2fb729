"""CPU model of the index arithmetic of the two-step kernels (chemsim_b200/csrc/step2_impl.cuh): no GPU, no
compute — it pins the properties the kernels rely on, with the tile constants parsed from step_decl.cuh so that a
change of the tile shape that breaks one of them fails here first.

  * phase A on adjacent cell pairs: the incremental (row, pair) walk of a thread block visits every cell of the
    (TY+2) x (TX+2) ext region exactly once, and for the six populations that stream along x the pair's first
    element sits on an even element index in HBM (8-byte aligned 64-bit load) and on an even shared-memory column
    (64-bit store);
  * phase B: every shifted shared-memory read of a lane starts on a multiple of V columns (16-byte aligned);
  * the L2 prefetch of a later tile never leaves the own rows / columns of the lattice, whatever the distance.
"""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECL = open(os.path.join(ROOT, "chemsim_b200", "csrc", "step_decl.cuh")).read()
GHOST = 2
# d2q9.cuh: ex_q = cy_q (State::stream moves population q by dx = +c_iy)
EX_OF = [0, 0, 1, 0, -1, 1, 1, -1, -1]


def default_of(macro):
    m = re.search(r"#ifndef %s\s*\n#define %s\s+(\S+)" % (macro, macro), DECL)
    assert m, macro
    return int(m.group(1))


def tile(elem_bytes):
    v = 16 // elem_bytes
    ty = default_of("CHEMSIM_STEP2_TY")
    tx = 32 * v
    ex, ey = tx + 2, ty + 2
    sp = ((ex + v - 1 + v - 1) // v) * v
    return dict(V=v, TY=ty, TX=tx, NT=32 * ty, EX=ex, EY=ey, SP=sp)


def shift_of(q, v):
    return ((EX_OF[q] - 1) % v + v) % v


def test_defaults_are_the_measured_ones():
    assert default_of("CHEMSIM_STEP2_TY") == 8
    assert default_of("CHEMSIM_STEP2_HPAIR") == 1
    assert default_of("CHEMSIM_PACKED_STEP2") == 1
    assert default_of("CHEMSIM_PACKED_VEC") == 0            # single-step kernels stay scalar (64-register cap)


def test_adjacent_pair_walk_covers_the_ext_region_once_and_stays_aligned():
    t = tile(4)
    assert t["EX"] % 2 == 0 and t["SP"] % 2 == 0 and (t["EY"] * t["SP"]) % 2 == 0
    px_per_row = t["EX"] // 2
    dyp, dxp = divmod(t["NT"], px_per_row)
    seen = {}
    for tid in range(t["NT"]):
        ey0, px0 = divmod(tid, px_per_row)
        idx = tid
        while idx < t["EY"] * px_per_row:
            assert (ey0, px0) == divmod(idx, px_per_row)          # the carry update equals the division
            for cell in ((ey0, 2 * px0), (ey0, 2 * px0 + 1)):
                assert cell not in seen
                seen[cell] = tid
            ey0 += dyp
            px0 += dxp
            if px0 >= px_per_row:
                px0 -= px_per_row
                ey0 += 1
            idx += t["NT"]
    assert len(seen) == t["EY"] * t["EX"]
    assert all(0 <= ey < t["EY"] and 0 <= ex < t["EX"] for ey, ex in seen)
    # alignment: interior tiles start at tx0 = k*TX (k >= 1); pitch and plane are multiples of 32 elements
    for tx0 in (t["TX"], 5 * t["TX"]):
        for px in range(px_per_row):
            for q in range(9):
                elem = (tx0 - 1) + 2 * px - EX_OF[q]               # column of the pair's first source element
                col = 2 * px + shift_of(q, t["V"])                 # shared-memory column of its first result
                if EX_OF[q] != 0:
                    assert elem % 2 == 0 and col % 2 == 0          # one LDG.64, one STS.64
                else:
                    assert elem % 2 == 1 and col % 2 == 1          # two scalar accesses each
                assert col + 1 < t["SP"]


@pytest.mark.parametrize("elem_bytes", [4, 8])
def test_phase_b_reads_are_vector_aligned(elem_bytes):
    t = tile(elem_bytes)
    for q in range(9):
        for lane in range(32):
            col = lane * t["V"] + (1 - EX_OF[q] + shift_of(q, t["V"]))
            assert col % t["V"] == 0 and col + t["V"] <= t["SP"]


@pytest.mark.parametrize("elem_bytes", [4, 8])
@pytest.mark.parametrize("w,h", [(4096, 4096), (1024, 40), (128, 128), (640, 17), (260, 64)])
@pytest.mark.parametrize("ahead", [1, 37, 148, 592, 5000])
def test_prefetch_targets_stay_inside_the_lattice(elem_bytes, w, h, ahead):
    """step2_kernel: block (bx, trow) prefetches tile (bx + ahead % ntx [carry], trow + ahead / ntx); rows
    pty0-1 .. pty0+TY and columns ptx0 .. ptx0+TX-1 must be own rows / columns (no ghost row, no pitch padding)."""
    t = tile(elem_bytes)
    if w % t["V"]:
        pytest.skip("ragged widths take the scalar kernel")
    ntx = (w + t["TX"] - 1) // t["TX"]
    nty = (h + t["TY"] - 1) // t["TY"]
    rows, cols = divmod(ahead, ntx)
    lines = t["TX"] * elem_bytes // 128
    assert t["EY"] * lines <= t["NT"]                              # one thread per (row, line)
    hit = 0
    for trow in range(nty):
        for bx in range(ntx):
            pcol, prow = bx + cols, trow + rows
            if pcol >= ntx:
                pcol -= ntx
                prow += 1
            pty0, ptx0 = prow * t["TY"], pcol * t["TX"]
            if not pty0 + t["TY"] <= h:                            # kernel: beyond the launch's rows
                continue
            if pty0 < 1 or pty0 + t["TY"] + 1 > h or ptx0 < 0 or ptx0 + t["TX"] > w:   # step2_prefetch_tile's guard
                continue
            hit += 1
            first_row, last_row = pty0 - 1, pty0 - 1 + t["EY"] - 1
            assert 0 <= first_row and last_row <= h - 1
            assert 0 <= ptx0 and ptx0 + lines * (128 // elem_bytes) <= w
            assert 0 <= first_row + GHOST and last_row + GHOST < h + 2 * GHOST
    if (w, h) == (4096, 4096) and ahead <= 592:
        assert hit > 0.9 * ntx * nty                               # and at the headline size it does prefetch
