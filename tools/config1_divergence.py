#!/usr/bin/env python
"""BASELINE.json config 1 run to N = 1000 on the literal main.rs setup (256^2, BGK tau = 15,
zero-fill edges, walls + cylinder): the reference scheme leaks mass through its zero-fill
edges and diverges at N ~ 210-250.  This reports — it does not gate — that
the GPU path and the CPU oracle blow up the same way: per checkpoint the total mass of both,
whether they are bit-identical, the first step at which is_unstable() fires and the first
step with a non-finite population.  Writes a markdown table to stdout."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from chemsim_b200 import lbm, scenarios  # noqa: E402
from oracle import lbm_oracle as O  # noqa: E402


def main():
    for dtype in (np.float32, np.float64):
        w = h = 256
        rho, vx, vy, solid = scenarios.main_rs(w, h, dtype)
        m = lambda a: lbm.Matrix.new(a.reshape(-1), (w, h), dtype=dtype)
        disc = lbm.Discretization(1.0, 1.0)
        pops = lbm.compute_equilibrium(m(rho), (m(vx), m(vy)), lbm.D2Q9.directions(), disc)
        state = lbm.State.initial(lbm.D2Q9.new(pops), solid, lbm.BGK(15.0), disc)
        f = O.compute_equilibrium(rho, vx, vy)
        col = O.collision(O.BGK, tau=15.0)
        u = np.uint32 if dtype == np.float32 else np.uint64
        first_unstable = {"gpu": None, "cpu": None}
        first_nonfinite = {"gpu": None, "cpu": None}
        print(f"\n### {np.dtype(dtype).name}\n")
        print("| N | mass (GPU) | mass (oracle) | bit-identical | finite entries equal | max |f| (GPU) |")
        print("|---|---|---|---|---|---|")
        with np.errstate(all="ignore"):
            for n in range(1, 1001):
                state.step(1)
                f = O.step_ref(f, solid, 1, col, O.EDGE_ZEROFILL)
                if n % 10 == 0 or n in (1, 2):
                    g = state.populations_array()
                    for name, arr, unstable in (("gpu", g, state.is_unstable()), ("cpu", f, O.is_unstable(f))):
                        if unstable and first_unstable[name] is None:
                            first_unstable[name] = n
                        if not np.isfinite(arr).all() and first_nonfinite[name] is None:
                            first_nonfinite[name] = n
                    if n in (1, 2, 10, 50, 100, 200, 250, 300, 500, 1000):
                        both = np.isfinite(g) & np.isfinite(f)
                        same_mask = bool((np.isfinite(g) == np.isfinite(f)).all())
                        print(f"| {n} | {state.total_mass():.10g} | {O.total_mass(f):.10g} | "
                              f"{bool((g.view(u) == f.view(u)).all())} | "
                              f"{same_mask and bool((g[both] == f[both]).all())} | {np.nanmax(np.abs(g)):.3g} |")
        print(f"\nfirst is_unstable(): GPU step {first_unstable['gpu']}, oracle step {first_unstable['cpu']} "
              f"(checked every 10 steps); first non-finite population: GPU step {first_nonfinite['gpu']}, "
              f"oracle step {first_nonfinite['cpu']}")
        state.close()


if __name__ == "__main__":
    main()
