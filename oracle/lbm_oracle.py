"""oracle/lbm_oracle.py — ctypes loader for the C oracle (oracle/lbm_oracle.c).

TEST INFRASTRUCTURE ONLY (parity unpinned, see oracle/lbm_oracle.h).  Importable
from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never from
chemsim_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblbm_oracle.so")

EDGE_ZEROFILL, EDGE_PERIODIC = 0, 1
BGK, TRT, REGULARIZED, KBC = 0, 1, 2, 3


class Collision(C.Structure):
    _fields_ = [("kind", C.c_int), ("tau", C.c_double), ("tau_plus", C.c_double),
                ("tau_minus", C.c_double), ("viscosity", C.c_double)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, n) for n in ("lbm_oracle.c", "lbm_oracle_impl.h", "lbm_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.lbm_oracle_total_mass_f32.restype = C.c_double
        _lib.lbm_oracle_total_mass_f64.restype = C.c_double
    return _lib


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", C.c_float
    if dtype == np.float64:
        return "f64", C.c_double
    raise TypeError(dtype)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _fn(name, dtype):
    return getattr(lib(), f"{name}_{_sfx(dtype)[0]}")


def max_threads() -> int:
    return int(lib().lbm_oracle_max_threads())


def use_all_cores() -> int:
    """Use every core this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().lbm_oracle_set_threads(n)
    return max_threads()


def constants(dtype, dx=1.0, dt=1.0):
    s, ct = _sfx(dtype)
    out = np.zeros(5, dtype=dtype)
    _fn("lbm_oracle_constants", dtype)(ct(dx), ct(dt), _p(out))
    return dict(zip(("cs2", "cs4", "k1", "k2", "k3"), out))


def compute_equilibrium(rho, vx, vy, dx=1.0, dt=1.0):
    """compute_equilibrium, /root/reference/src/lbm.rs:43-71.  Fields are (h, w)."""
    dtype = rho.dtype
    _, ct = _sfx(dtype)
    rho, vx, vy = (np.ascontiguousarray(a, dtype=dtype) for a in (rho, vx, vy))
    out = np.empty((9,) + rho.shape, dtype=dtype)
    _fn("lbm_oracle_equilibrium", dtype)(_p(rho), _p(vx), _p(vy), C.c_size_t(rho.size), ct(dx), ct(dt), _p(out))
    return out


def collision(kind=BGK, tau=0.0, tau_plus=0.0, tau_minus=0.0, viscosity=0.0):
    return Collision(kind, tau, tau_plus, tau_minus, viscosity)


def step_ref(f, solid, nsteps, col: Collision, edge=EDGE_ZEROFILL, dx=1.0, dt=1.0):
    """State::step x nsteps, reference-structured three passes (src/lbm.rs:694-714).
    f: (9, h, w), updated copy returned."""
    dtype = f.dtype
    _, ct = _sfx(dtype)
    f = np.array(f, dtype=dtype, order="C", copy=True)
    _, h, w = f.shape
    sp = None
    if solid is not None:
        solid = np.ascontiguousarray(solid, dtype=np.uint8)
        assert solid.shape == (h, w)
        sp = _p(solid)
    _fn("lbm_oracle_step_ref", dtype)(_p(f), sp, w, h, edge, ct(dx), ct(dt), C.byref(col), nsteps)
    return f


def step_fused(f, solid, nsteps, tau, edge=EDGE_ZEROFILL, dx=1.0, dt=1.0):
    """Fused BGK step (OpenMP over rows), bit-identical to step_ref with BGK."""
    dtype = f.dtype
    _, ct = _sfx(dtype)
    a = np.array(f, dtype=dtype, order="C", copy=True)
    b = np.empty_like(a)
    _, h, w = a.shape
    sp = None
    if solid is not None:
        solid = np.ascontiguousarray(solid, dtype=np.uint8)
        sp = _p(solid)
    fn = _fn("lbm_oracle_step_fused", dtype)
    for _ in range(nsteps):
        fn(_p(a), _p(b), sp, w, h, edge, 0, ct(dx), ct(dt), ct(tau))
        a, b = b, a
    return a


class FusedStepper:
    """Persistent A/B buffers for timing the fused step (bench.py's CPU legs): nothing is
    allocated, copied or first-touched inside `step`.  Construction runs two untimed steps so
    that BOTH buffers of the pair that is kept have been first-touched by the OpenMP threads
    that will work on them (the caller's array was touched by one thread only)."""

    def __init__(self, f, solid, tau, edge=EDGE_ZEROFILL, dx=1.0, dt=1.0):
        dtype = f.dtype
        _, self._ct = _sfx(dtype)
        f = np.ascontiguousarray(f)
        _, self.h, self.w = f.shape
        self._solid = None if solid is None else np.ascontiguousarray(solid, dtype=np.uint8)
        self._fn = _fn("lbm_oracle_step_fused", dtype)
        self._par = (edge, 0, self._ct(dx), self._ct(dt), self._ct(tau))
        b = np.empty_like(f)
        a = np.empty_like(f)
        self._call(f, b)
        self._call(b, a)
        self.a, self.b = a, b          # a = state after the two construction steps
        self.steps_done = 2

    def _call(self, src, dst):
        sp = None if self._solid is None else _p(self._solid)
        self._fn(_p(src), _p(dst), sp, self.w, self.h, self._par[0], self._par[1], *self._par[2:])

    def step(self, nsteps):
        for _ in range(nsteps):
            self._call(self.a, self.b)
            self.a, self.b = self.b, self.a
        self.steps_done += nsteps
        return self.a


def step_fused_slab(src_with_ghosts, dst_with_ghosts, solid, edge, tau, dx=1.0, dt=1.0):
    """One fused step on a y-slab whose planes carry one ghost row above and below
    (shape (9, h+2, w)); x edges follow `edge`, y neighbours come from the ghosts."""
    dtype = src_with_ghosts.dtype
    _, ct = _sfx(dtype)
    _, hp, w = src_with_ghosts.shape
    sp = None
    if solid is not None:
        solid = np.ascontiguousarray(solid, dtype=np.uint8)
        sp = _p(solid)
    _fn("lbm_oracle_step_fused", dtype)(_p(src_with_ghosts), _p(dst_with_ghosts), sp, w, hp - 2, edge, 1,
                                        ct(dx), ct(dt), ct(tau))


def _field(name, f, *extra):
    dtype = f.dtype
    f = np.ascontiguousarray(f)
    n = f[0].size
    out = np.empty(f.shape[1:], dtype=dtype)
    _fn(name, dtype)(_p(f), C.c_size_t(n), *extra, _p(out))
    return out


def density(f):
    return _field("lbm_oracle_density", f)


def speed(f):
    return _field("lbm_oracle_speed", f)


def pressure(f, dx=1.0, dt=1.0):
    _, ct = _sfx(f.dtype)
    return _field("lbm_oracle_pressure", f, ct(dx), ct(dt))


def _pair(name, f):
    dtype = f.dtype
    f = np.ascontiguousarray(f)
    a = np.empty(f.shape[1:], dtype=dtype)
    b = np.empty(f.shape[1:], dtype=dtype)
    _fn(name, dtype)(_p(f), C.c_size_t(f[0].size), _p(a), _p(b))
    return a, b


def momentum_density(f):
    return _pair("lbm_oracle_momentum", f)


def velocity(f):
    return _pair("lbm_oracle_velocity", f)


def lattice_equilibrium(f, dx=1.0, dt=1.0):
    dtype = f.dtype
    _, ct = _sfx(dtype)
    f = np.ascontiguousarray(f)
    out = np.empty_like(f)
    _fn("lbm_oracle_lattice_equilibrium", dtype)(_p(f), C.c_size_t(f[0].size), ct(dx), ct(dt), _p(out))
    return out


def is_unstable(f, dx=1.0, dt=1.0) -> bool:
    _, ct = _sfx(f.dtype)
    f = np.ascontiguousarray(f)
    return bool(_fn("lbm_oracle_is_unstable", f.dtype)(_p(f), C.c_size_t(f[0].size), ct(dx), ct(dt)))


def total_mass(f) -> float:
    f = np.ascontiguousarray(f)
    _, h, w = f.shape
    return float(_fn("lbm_oracle_total_mass", f.dtype)(_p(f), w, h))
