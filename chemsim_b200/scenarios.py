"""Initial conditions for the D2Q9 path (host-side numpy; caller-level code).

`main_rs` restates what the reference's binary builds in
/root/reference/src/main.rs:180-328 (`initial_state`): rho == 1, (vx, vy) ==
(0.02, 0), geometry = disc of radius 25 at (w/2, h/2) plus the four border
lines.  The other builders are the synthetic configurations of BASELINE.json
made concrete in SURVEY.md §8(d).  All fields are (h, w) arrays, element (y, x)
at [y, x] == host index y*w+x, the order Matrix::new takes (src/matrix.rs:24-30).
"""
from __future__ import annotations

import numpy as np


def main_rs(w: int, h: int, dtype=np.float32, walls: bool = True, radius: float = 25.0):
    """main.rs `initial_state((w, h))`.  walls=False gives the stable periodic twin
    (outer border lines removed) of SURVEY.md §8(d) config 1."""
    rho = np.full((h, w), 1.0, dtype=dtype)                 # main.rs:223
    vx = np.full((h, w), 0.02, dtype=dtype)                 # main.rs:208-211
    vy = np.zeros((h, w), dtype=dtype)                      # main.rs:212
    x = np.arange(w, dtype=np.float64)[None, :]
    y = np.arange(h, dtype=np.float64)[:, None]
    r = np.sqrt((x - w / 2.0) ** 2 + (y - h / 2.0) ** 2)    # main.rs:283-286 (f64)
    solid = r < radius                                      # main.rs:287-289
    if walls:                                               # main.rs:290-293
        solid[:, 0] = True
        solid[0, :] = True
        solid[:, w - 1] = True
        solid[h - 1, :] = True
    return rho, vx, vy, solid.astype(np.uint8)


def smooth_periodic(w: int, h: int, dtype=np.float32):
    """rho = 1 + 0.01 sin(2pi x/W) cos(2pi y/H),
    u = 0.05 (sin(2pi y/H), sin(2pi x/W)); no solids; run with tau = 0.8."""
    x = (np.arange(w, dtype=np.float64) / w)[None, :]
    y = (np.arange(h, dtype=np.float64) / h)[:, None]
    two_pi = 2.0 * np.pi
    rho = 1.0 + 0.01 * np.sin(two_pi * x) * np.cos(two_pi * y)
    vx = 0.05 * np.sin(two_pi * y) * np.ones_like(x)
    vy = 0.05 * np.sin(two_pi * x) * np.ones_like(y)
    solid = np.zeros((h, w), dtype=np.uint8)
    return rho.astype(dtype), vx.astype(dtype), vy.astype(dtype), solid


def smooth_periodic_rows(w: int, h_global: int, y0: int, y1: int, dtype=np.float32):
    """Rows [y0, y1) of smooth_periodic(w, h_global) without building the whole
    field (a rank's y-slab of a large global lattice).  Same values bit for bit; evaluated in
    cache-sized row blocks because the 32768-wide workloads build tens of GiB with it."""
    x = (np.arange(w, dtype=np.float64) / w)[None, :]
    y = (np.arange(y0, y1, dtype=np.float64) / h_global)[:, None]
    two_pi = 2.0 * np.pi
    sx = 0.01 * np.sin(two_pi * x)                       # (1, w)   rho = 1.0 + (0.01 sin) * cos
    cy = np.cos(two_pi * y)                              # (h, 1)
    rows = y1 - y0
    rho = np.empty((rows, w), dtype=dtype)
    block = max(1, (1 << 20) // max(w, 1))
    tmp = np.empty((min(block, rows), w), dtype=np.float64)
    for r in range(0, rows, block):
        n = min(block, rows - r)
        np.multiply(sx, cy[r:r + n], out=tmp[:n])
        np.add(1.0, tmp[:n], out=tmp[:n])
        rho[r:r + n] = tmp[:n]
    # (0.05 sin(2 pi y)) * 1.0 and (0.05 sin(2 pi x)) * 1.0: products by one are exact
    vx = np.ascontiguousarray(np.broadcast_to((0.05 * np.sin(two_pi * y)).astype(dtype), (rows, w)))
    vy = np.ascontiguousarray(np.broadcast_to((0.05 * np.sin(two_pi * x)).astype(dtype), (rows, w)))
    solid = np.zeros((rows, w), dtype=np.uint8)
    return rho, vx, vy, solid


def channel_cylinder(w: int = 8192, h: int = 2048, dtype=np.float32, radius: float = 64.0,
                     cx: float = 1024.0, cy: float | None = None, u0: float = 0.05):
    """channel walls on rows y=0 and y=h-1, a cylinder
    of `radius` at (cx, cy), uniform flow u0 along the channel (+x in memory,
    which is the reference's "vy" component: population 2 moves by dx=+1)."""
    if cy is None:
        cy = h / 2.0
    rho = np.ones((h, w), dtype=dtype)
    vx = np.zeros((h, w), dtype=dtype)
    vy = np.full((h, w), u0, dtype=dtype)
    x = np.arange(w, dtype=np.float64)[None, :]
    y = np.arange(h, dtype=np.float64)[:, None]
    solid = np.sqrt((x - cx) ** 2 + (y - cy) ** 2) < radius
    solid[0, :] = True
    solid[h - 1, :] = True
    return rho, vx, vy, solid.astype(np.uint8)


def random_state(w: int, h: int, dtype=np.float32, seed: int = 0, solid_fraction: float = 0.1):
    """Seeded ragged test input: positive random rho, small random u, random solids."""
    rng = np.random.default_rng(seed)
    rho = (0.8 + 0.4 * rng.random((h, w))).astype(dtype)
    vx = (0.1 * (rng.random((h, w)) - 0.5)).astype(dtype)
    vy = (0.1 * (rng.random((h, w)) - 0.5)).astype(dtype)
    solid = (rng.random((h, w)) < solid_fraction).astype(np.uint8)
    return rho, vx, vy, solid
