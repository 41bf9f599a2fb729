"""Generates tests/golden/*.npz — committed known-answer vectors for the D2Q9 path.

The reference ships no golden data and cannot be run here (Rust +
ArrayFire, neither present), so these vectors are produced by the *literal*
array-at-a-time numpy/scipy restatement (oracle/lbm_numpy.py: nine
`convolve2d(f_i, stencil_i^T)` calls, `np.where` for af::replace, elementwise ops
in reference order).  They pin the C oracle and the CUDA path against an
implementation that shares no code with either.  Parity with the real reference
remains UNPINNED (oracle/lbm_oracle.h).

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from chemsim_b200 import scenarios  # noqa: E402
from oracle import lbm_numpy as N  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run(rho, vx, vy, solid, dtype, collision, periodic, steps, dx=1.0, dt=1.0):
    k = N.Consts(dtype, dx, dt)
    f = N.compute_equilibrium(rho.astype(dtype), vx.astype(dtype), vy.astype(dtype), k)
    snaps = {}
    for s in range(1, max(steps) + 1):
        f = N.step(f, solid.astype(bool), k, collision, periodic)
        if s in steps:
            snaps[s] = f.copy()
    return snaps


def main():
    cases = {}
    # (1) main.rs scenario, scaled to 48x48 with radius 6, literal zero-fill edges, BGK tau=15
    for dtype in (np.float32, np.float64):
        tag = np.dtype(dtype).name
        rho, vx, vy, solid = scenarios.main_rs(48, 48, dtype, walls=True, radius=6.0)
        for s, f in run(rho, vx, vy, solid, dtype, ("bgk", 15.0), False, (1, 2, 10)).items():
            cases[f"mainrs48_zerofill_bgk15_{tag}_n{s}"] = f
        # main.rs's ACTIVE configuration: Regularized<KBC> (src/main.rs:198-199), literal zero-fill edges
        for s, f in run(rho, vx, vy, solid, dtype, ("regularized",), False, (1, 2, 10)).items():
            cases[f"mainrs48_zerofill_regularized_{tag}_n{s}"] = f
        rho, vx, vy, solid = scenarios.main_rs(48, 48, dtype, walls=False, radius=6.0)
        for s, f in run(rho, vx, vy, solid, dtype, ("bgk", 15.0), True, (1, 25)).items():
            cases[f"mainrs48_periodic_bgk15_{tag}_n{s}"] = f
        # (2) ragged, non-square random state with random solids (seeded), tau = 0.8
        rho, vx, vy, solid = scenarios.random_state(40, 24, dtype, seed=7)
        for periodic in (False, True):
            e = "periodic" if periodic else "zerofill"
            for s, f in run(rho, vx, vy, solid, dtype, ("bgk", 0.8), periodic, (1, 5)).items():
                cases[f"random40x24_{e}_bgk08_{tag}_n{s}"] = f
        # (3) the other operators of src/lbm.rs on the same ragged state, 3 steps
        for name, col in (("trt", ("trt", 0.8, 1.1)), ("regularized", ("regularized",)), ("kbc", ("kbc", 0.1))):
            for s, f in run(rho, vx, vy, solid, dtype, col, True, (3,)).items():
                cases[f"random40x24_periodic_{name}_{tag}_n{s}"] = f
    np.savez_compressed(os.path.join(OUT, "d2q9_golden.npz"), **cases)
    print(f"wrote {len(cases)} arrays")


if __name__ == "__main__":
    main()
