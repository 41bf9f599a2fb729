"""oracle/lbm_numpy.py — array-at-a-time numpy/scipy restatement of lbm.rs.

TEST INFRASTRUCTURE ONLY (parity unpinned, see oracle/lbm_oracle.h).  This is the
*second*, independent restatement: where the C oracle hard-codes the stream
shift table, this file performs the reference's literal calls —
`convolve2(f_i, stencil_i^T)` with the 3x3 one-hot stencils of
/root/reference/src/lbm.rs:233-269 via `scipy.signal.convolve2d(mode="same",
fillvalue=0)` (a true, flipped convolution with zero padding, which is what
ArrayFire 3.6.1 documents for ConvMode::DEFAULT / ConvDomain::SPATIAL), and
`af::replace(a, cond, b)` as `np.where(cond, a, b)`.  tests/test_oracle.py checks
that the C oracle and this file agree bit for bit; the shift table in
lbm_oracle.c is thereby derived, not assumed.

It is slow (pure numpy, 9 convolutions per step) and used for small cases only.

Layout: a `Matrix` built from a host slice s of shape (w, h) holds s[y*w+x] at
ArrayFire position (dim0=y, dim1=x) (src/matrix.rs:24-30); here that is simply
the numpy array `s.reshape(h, w)`, indexed [y, x].
"""
from __future__ import annotations

import numpy as np
from scipy.signal import convolve2d, correlate2d

# src/lbm.rs:209-219 (numerators over 36), :221-231
W_NUM = [16, 4, 4, 4, 4, 1, 1, 1, 1]
CX = [0, 1, 0, -1, 0, 1, -1, -1, 1]
CY = [0, 0, 1, 0, -1, 1, 1, -1, -1]
OPP = [0, 3, 4, 1, 2, 7, 8, 5, 6]  # src/lbm.rs:298-309

# src/lbm.rs:233-269, row-major 3x3 exactly as written in make_m(&[...])
STENCILS = [
    [0, 0, 0, 0, 1, 0, 0, 0, 0],
    [0, 0, 0, 1, 0, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 0, 1, 0],
    [0, 0, 0, 0, 0, 1, 0, 0, 0],
    [0, 1, 0, 0, 0, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 1, 0, 0],
    [0, 0, 0, 0, 0, 0, 0, 0, 1],
    [0, 0, 1, 0, 0, 0, 0, 0, 0],
    [1, 0, 0, 0, 0, 0, 0, 0, 0],
]


def stencil_matrix(i: int, dtype) -> np.ndarray:
    """Matrix::new(&temp, (3,3)) (src/lbm.rs:203-207): element (y,x) = vals[y*3+x]."""
    return np.array(STENCILS[i], dtype=dtype).reshape(3, 3)


def derived_shift(i: int) -> tuple[int, int]:
    """(dy, dx) by which State::stream moves population i, found by convolving a
    single 1 with stencil_i^T — the derivation behind ORACLE_EY/ORACLE_EX."""
    probe = np.zeros((5, 5))
    probe[2, 2] = 1.0
    out = convolve2d(probe, stencil_matrix(i, np.float64).T, mode="same", fillvalue=0)
    (y,), (x,) = np.nonzero(out)
    return int(y) - 2, int(x) - 2


class Consts:
    """Host scalars computed in `dtype` as src/lbm.rs:54-56, :64-66, :84 do."""

    def __init__(self, dtype, dx, dt):
        t = np.dtype(dtype).type
        self.t = t
        self.dx, self.dt = t(dx), t(dt)
        self.cs = self.dx / (np.sqrt(t(3.0)) * self.dt)
        self.cs2 = self.cs * self.cs
        self.cs4 = self.cs2 * self.cs2
        self.k1 = t(1.0) / self.cs2
        self.k2 = t(1.0) / (t(2.0) * self.cs4)
        self.k3 = t(-1.0) / (t(2.0) * self.cs2)
        self.w = [t(n) / t(36.0) for n in W_NUM]
        self.cx = [t(c) for c in CX]
        self.cy = [t(c) for c in CY]


def density(f):
    """src/lbm.rs:117-121"""
    rho = np.zeros_like(f[0])
    for i in range(9):
        rho = rho + f[i]
    return rho


def momentum_density(f, k: Consts):
    """src/lbm.rs:123-131"""
    mx = np.zeros_like(f[0])
    my = np.zeros_like(f[0])
    for i in range(9):
        mx = mx + f[i] * k.cx[i]
        my = my + f[i] * k.cy[i]
    return mx, my


def velocity(f, k: Consts):
    """src/lbm.rs:133-138, src/matrix.rs:133-136"""
    inv = np.ones_like(f[0]) / density(f)
    mx, my = momentum_density(f, k)
    return inv * mx, inv * my


def speed(f, k: Consts):
    """src/lbm.rs:151-154"""
    vx, vy = velocity(f, k)
    return np.sqrt(vx * vx + vy * vy)


def compute_equilibrium(rho, vx, vy, k: Consts):
    """src/lbm.rs:43-71"""
    v2 = vx * vx + vy * vy
    out = []
    for i in range(9):
        vc = vx * k.cx[i] + vy * k.cy[i]
        vc2 = vc * vc
        s = np.ones_like(rho) + vc * k.k1 + vc2 * k.k2 + v2 * k.k3
        out.append((rho * k.w[i]) * s)
    return np.stack(out)


def equilibrium(f, k: Consts):
    """src/lbm.rs:156-160"""
    vx, vy = velocity(f, k)
    return compute_equilibrium(density(f), vx, vy, k)


def stream(f, periodic=False, mirrored=False):
    """src/lbm.rs:716-729.  periodic=True is the extension (boundary='wrap').
    mirrored=True is the OTHER reading of af::convolve2 — the kernel applied unflipped (a
    correlation) — which moves every population the opposite way; the product's switch for it is
    chemsim_lbm_set_stream_convention (a reference-generated golden vector decides, see
    tests/test_reference_golden.py)."""
    conv = correlate2d if mirrored else convolve2d
    out = []
    for i in range(9):
        st = stencil_matrix(i, f.dtype).T  # pair.0.stencil.transpose()
        if periodic:
            o = conv(f[i], st, mode="same", boundary="wrap")
        else:
            o = conv(f[i], st, mode="same", boundary="fill", fillvalue=0)
        out.append(o.astype(f.dtype))
    return np.stack(out)


def bounce_back(f, solid):
    """src/lbm.rs:741-751: replace(sw_i, geometry, f_i) keeps sw_i where solid."""
    return np.stack([np.where(solid, f[OPP[i]], f[i]) for i in range(9)])


def collide_bgk(f, feq, k: Consts, tau):
    """src/lbm.rs:349-364"""
    factor = -k.dt / k.t(tau)
    return f + (f - feq) * factor


def collide_trt(f, feq, k: Consts, tau_plus, tau_minus):
    """src/lbm.rs:401-444 (+ the swap_equilibrium quirk, :311-322)."""
    f_sw = np.stack([f[OPP[i]] for i in range(9)])
    feq_sw = np.stack([feq[0]] + [f[OPP[i]] for i in range(1, 9)])
    f_p, f_m = f + f_sw, f - f_sw
    e_p, e_m = feq + feq_sw, feq - feq_sw
    om_m = k.t(1.0) / k.t(tau_minus)
    om_p = k.t(1.0) / k.t(tau_plus)
    omega = ((f_p - e_p) * om_p + (f_m - e_m) * om_m) * (-k.dt * k.t(0.5))
    return f + omega


def collide_regularized(f, feq, k: Consts):
    """src/lbm.rs:606-661"""
    fneq = f - feq
    z = np.zeros_like(f[0])
    sxx, sxy, syx, syy = z, z, z, z
    for i in range(9):
        sxx = sxx + fneq[i] * (k.cx[i] * k.cx[i])
        sxy = sxy + fneq[i] * (k.cx[i] * k.cy[i])
        syx = syx + fneq[i] * (k.cy[i] * k.cx[i])
        syy = syy + fneq[i] * (k.cy[i] * k.cy[i])
    out = []
    for i in range(9):
        qxx = k.cx[i] * k.cx[i] - k.cs2
        qxy = k.cx[i] * k.cy[i]
        qyx = k.cy[i] * k.cx[i]
        qyy = k.cy[i] * k.cy[i] - k.cs2
        sf = k.w[i] / (k.t(2.0) * k.cs4)
        reg = feq[i]
        reg = reg + sxx * (qxx * sf)
        reg = reg + sxy * (qxy * sf)
        reg = reg + syx * (qyx * sf)
        reg = reg + syy * (qyy * sf)
        out.append(reg)
    return np.stack(out)


def collide_kbc(f, feq, k: Consts, visc):
    """src/lbm.rs:468-585"""
    t = k.t
    dx = k.dx
    rho = density(f)
    u, v = velocity(f, k)
    uv, u2, v2 = u * v, u * u, v * v
    temp = np.zeros_like(rho)
    for i in range(9):
        temp = temp + f[i] * (dx * dx)
    pi_t = temp - uv
    n_t = v2 - u2
    s0 = ((uv * t(8.0)) * pi_t + n_t * n_t) * rho * t(0.5)
    s13 = ((((u * dx - n_t) + t(1.0)) * n_t) - (v * (dx * t(4.0)) + uv * t(8.0)) * pi_t) * rho * t(0.25)
    s24 = ((((v * (-dx) - n_t) + t(-1.0)) * n_t) - (u * (dx * t(4.0)) + uv * t(8.0)) * pi_t) * rho * t(0.25)
    s58 = ((((uv * t(8.0) + u * (t(4.0) * dx)) + (t(2.0) * dx * dx)) * pi_t) + (n_t - (v - u) * dx) * n_t) * rho * t(0.125)
    ds = [s0, s13, s24, s13, s24, s58, s58, s58, s58]
    dh = [f[i] - feq[i] - ds[i] for i in range(9)]
    cs = k.cs
    beta = t(1.0) / ((t(2.0) * t(visc) / (cs * cs)) + t(1.0))
    num = np.zeros_like(rho)
    den = np.zeros_like(rho)
    with np.errstate(all="ignore"):
        for i in range(9):
            num = num + (ds[i] * dh[i]) / feq[i]
            den = den + (dh[i] * dh[i]) / feq[i]
        gamma = (((num / den) * (t(2.0) - t(1.0) / beta)) + (t(-1.0) / beta)) * t(-1.0)
        out = [f[i] + (ds[i] * (t(2.0) * -beta) + (dh[i] * gamma) * (-beta)) for i in range(9)]
    return np.stack(out)


def step(f, solid, k: Consts, collision=("bgk", 15.0), periodic=False, mirrored=False):
    """State::step src/lbm.rs:694-714: stream -> bounce_back -> collide."""
    f = stream(f, periodic, mirrored)
    f = bounce_back(f, solid)
    feq = equilibrium(f, k)
    kind = collision[0]
    if kind == "bgk":
        return collide_bgk(f, feq, k, collision[1])
    if kind == "trt":
        return collide_trt(f, feq, k, collision[1], collision[2])
    if kind == "regularized":
        return collide_regularized(f, feq, k)
    if kind == "kbc":
        return collide_kbc(f, feq, k, collision[1])
    raise ValueError(kind)
