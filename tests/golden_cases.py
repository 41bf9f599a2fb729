"""Shared description of the committed golden vectors (tests/golden/d2q9_golden.npz): for a
case name, the inputs, the edge mode and the collision operator — as the oracle's struct and
as the host mirror's object.  Used by the CPU oracle tests and the GPU parity tests, so the
name parsing is exercised on CPU."""
import os

import numpy as np

from chemsim_b200 import lbm, scenarios
from oracle import lbm_oracle as O

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "d2q9_golden.npz"))


def parse(name):
    """-> dict(dtype, steps, inputs=(rho, vx, vy, solid), edge, oracle_collision, mirror_collision)"""
    parts = name.split("_")
    dtype = np.float32 if "float32" in parts else np.float64
    steps = int(parts[-1][1:])
    scenario, edge_name, col_name = parts[0], parts[1], parts[2]
    if scenario == "mainrs48":
        inputs = scenarios.main_rs(48, 48, dtype, walls=(edge_name == "zerofill"), radius=6.0)
    elif scenario == "random40x24":
        inputs = scenarios.random_state(40, 24, dtype, seed=7)
    else:
        raise KeyError(name)
    edge = {"zerofill": O.EDGE_ZEROFILL, "periodic": O.EDGE_PERIODIC}[edge_name]
    oracle_col = {"bgk15": O.collision(O.BGK, tau=15.0), "bgk08": O.collision(O.BGK, tau=0.8),
                  "trt": O.collision(O.TRT, tau_plus=0.8, tau_minus=1.1),
                  "regularized": O.collision(O.REGULARIZED), "kbc": O.collision(O.KBC, viscosity=0.1)}[col_name]
    mirror_col = {"bgk15": lbm.BGK(15.0), "bgk08": lbm.BGK(0.8), "trt": lbm.TRT(tau_minus=1.1, tau_plus=0.8),
                  "regularized": lbm.Regularized.new(lbm.KBC.new(10.0)), "kbc": lbm.KBC.new(0.1)}[col_name]
    return dict(dtype=dtype, steps=steps, inputs=inputs, edge=edge, oracle_collision=oracle_col,
                mirror_collision=mirror_col)


def names():
    return list(GOLDEN.files)
