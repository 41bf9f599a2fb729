"""CPU check of the PRODUCT's per-cell arithmetic: chemsim_b200/csrc/d2q9.cuh (bounce_back,
collide<COL> in their strength-reduced form) and consts.hpp are compiled with g++ — every CUDA
round-to-nearest intrinsic mapped to the plain IEEE operation, no contraction — and compared
bit for bit with the literal oracle.  The CUDA kernels execute the same expression tree with the
same individually rounded operations, so this pins the exactness of the reduced forms (and the
host scalars) on every machine, before the GPU parity tests repeat it on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from chemsim_b200 import scenarios
from oracle import lbm_oracle as O

import golden_cases

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_arith")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KIND = {O.BGK: 1, O.TRT: 2, O.REGULARIZED: 3, O.KBC: 4}      # d2q9.cuh: enum Collision


@pytest.fixture(scope="module")
def host():
    src = os.path.join(HERE, "host_step.cpp")
    lib = os.path.join(HERE, "libhost_step.so")
    deps = [src] + [os.path.join(ROOT, "chemsim_b200", "csrc", n) for n in ("d2q9.cuh", "consts.hpp")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                        "-I/usr/local/cuda/include", "-o", lib, src], check=True)
    return C.CDLL(lib)


def host_step(lib, f, solid, nsteps, col, edge, dx=1.0, dt=1.0, pairs=False):
    a = np.array(f, order="C", copy=True)
    b = np.empty_like(a)
    _, h, w = a.shape
    fn = lib.host_step_f32 if a.dtype == np.float32 else lib.host_step_f64
    if pairs:                                     # collide2<COL, true>: the F32x2 instantiation of the operators
        assert a.dtype == np.float32
        fn = lib.host_step_f32_pairs
    sp = None
    if solid is not None:
        solid = np.ascontiguousarray(solid, dtype=np.uint8)
        sp = solid.ctypes.data_as(C.c_void_p)
    for _ in range(nsteps):
        rc = fn(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), sp, w, h, int(edge == O.EDGE_PERIODIC),
                C.c_double(dx), C.c_double(dt), KIND[col.kind], C.c_double(col.tau), C.c_double(col.tau_plus),
                C.c_double(col.tau_minus), C.c_double(col.viscosity))
        assert rc == 0
        a, b = b, a
    return a


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32 if a.dtype == np.float32 else np.uint64)


COLLISIONS = {
    "bgk": O.collision(O.BGK, tau=0.8),
    "bgk15": O.collision(O.BGK, tau=15.0),
    "trt": O.collision(O.TRT, tau_plus=0.8, tau_minus=1.1),
    "regularized": O.collision(O.REGULARIZED),
    "kbc": O.collision(O.KBC, viscosity=0.1),
}


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("edge", [O.EDGE_ZEROFILL, O.EDGE_PERIODIC])
@pytest.mark.parametrize("name", sorted(COLLISIONS))
def test_product_arithmetic_is_bit_identical_to_the_literal_oracle(host, name, edge, dtype):
    col = COLLISIONS[name]
    for (w, h, seed) in ((40, 24, 1), (37, 9, 2), (5, 3, 3), (96, 50, 4)):
        rho, vx, vy, solid = scenarios.random_state(w, h, dtype, seed=seed)
        f0 = O.compute_equilibrium(rho, vx, vy)
        # perturb away from equilibrium so that f - feq is not tiny
        f0 = (f0 * (1.0 + 0.2 * (np.random.default_rng(seed).random(f0.shape) - 0.5))).astype(dtype)
        ref = O.step_ref(f0, solid, 4, col, edge)
        got = host_step(host, f0, solid, 4, col, edge)
        assert np.isfinite(ref).all()
        np.testing.assert_array_equal(bits(got), bits(ref), err_msg=f"{name} {w}x{h}")


@pytest.mark.parametrize("edge", [O.EDGE_ZEROFILL, O.EDGE_PERIODIC])
@pytest.mark.parametrize("name", sorted(COLLISIONS))
def test_two_cell_f32_instantiation_is_bit_identical_too(host, name, edge):
    """The f32 step kernels collide two cells at once on F32x2 values (packed additions, d2q9.cuh):
    the same templates instantiated for that type must evaluate the very same expression tree —
    every overload resolving to the f32 operation, every constant broadcast in f32."""
    col = COLLISIONS[name]
    for (w, h, seed) in ((40, 24, 1), (37, 9, 2), (5, 3, 3)):
        rho, vx, vy, solid = scenarios.random_state(w, h, np.float32, seed=seed)
        f0 = O.compute_equilibrium(rho, vx, vy)
        f0 = (f0 * (1.0 + 0.2 * (np.random.default_rng(seed).random(f0.shape) - 0.5))).astype(np.float32)
        ref = O.step_ref(f0, solid, 4, col, edge)
        got = host_step(host, f0, solid, 4, col, edge, pairs=True)
        np.testing.assert_array_equal(bits(got), bits(ref), err_msg=f"{name} {w}x{h}")
    # signed zeros and exact zeros (fluid at rest, zero-fill edges)
    rho = np.ones((10, 24), np.float32)
    z = np.zeros((10, 24), np.float32)
    solid = np.zeros((10, 24), np.uint8)
    solid[4:6, 7:9] = 1
    f0 = O.compute_equilibrium(rho, z, -z)
    ref = O.step_ref(f0, solid, 5, col, edge)
    got = host_step(host, f0, solid, 5, col, edge, pairs=True)
    ok = ~np.isnan(ref)
    assert (np.isnan(ref) == np.isnan(got)).all()
    np.testing.assert_array_equal(bits(got)[ok], bits(ref)[ok], err_msg=name)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_zero_velocity_and_exact_zero_populations(host, dtype):
    """Signed zeros: a fluid at rest (mx = my = +-0) and zero-fill edges that inject exact zeros."""
    w, h = 24, 10
    rho = np.ones((h, w), dtype)
    z = np.zeros((h, w), dtype)
    solid = np.zeros((h, w), np.uint8)
    solid[4:6, 7:9] = 1
    f0 = O.compute_equilibrium(rho, z, -z)        # vy = -0.0
    for name, col in COLLISIONS.items():
        for edge in (O.EDGE_ZEROFILL, O.EDGE_PERIODIC):
            ref = O.step_ref(f0, solid, 5, col, edge)
            got = host_step(host, f0, solid, 5, col, edge)
            if name == "kbc":                     # 0/0 in gamma at rest: NaN on both sides, same cells
                assert (np.isnan(ref) == np.isnan(got)).all()
                ok = ~np.isnan(ref)
                np.testing.assert_array_equal(bits(got)[ok], bits(ref)[ok], err_msg=name)
            else:
                np.testing.assert_array_equal(bits(got), bits(ref), err_msg=f"{name} edge={edge}")


def test_other_discretization(host):
    dtype = np.float64
    rho, vx, vy, solid = scenarios.random_state(33, 12, dtype, seed=9)
    f0 = O.compute_equilibrium(rho, vx, vy, 0.5, 0.25)
    for name in ("bgk", "trt", "regularized", "kbc"):
        col = COLLISIONS[name]
        ref = O.step_ref(f0, solid, 3, col, O.EDGE_PERIODIC, 0.5, 0.25)
        got = host_step(host, f0, solid, 3, col, O.EDGE_PERIODIC, 0.5, 0.25)
        np.testing.assert_array_equal(bits(got), bits(ref), err_msg=name)


@pytest.mark.parametrize("name", golden_cases.names())
def test_product_arithmetic_reproduces_the_golden_vectors(host, name):
    case = golden_cases.parse(name)
    rho, vx, vy, solid = case["inputs"]
    f0 = O.compute_equilibrium(rho, vx, vy)
    got = host_step(host, f0, solid, case["steps"], case["oracle_collision"], case["edge"])
    np.testing.assert_array_equal(bits(got), bits(golden_cases.GOLDEN[name]), err_msg=name)


def test_config1_blow_up_follows_the_oracle_while_finite(host):
    """main.rs's literal setup diverges; the reduced forms must follow the
    literal tree bit for bit for as long as every value is finite (here: 300 steps, |f| ~ 1e3)."""
    dtype = np.float32
    rho, vx, vy, solid = scenarios.main_rs(64, 64, dtype, radius=6.0)
    f0 = O.compute_equilibrium(rho, vx, vy)
    col = O.collision(O.BGK, tau=15.0)
    ref = O.step_ref(f0, solid, 300, col, O.EDGE_ZEROFILL)
    got = host_step(host, f0, solid, 300, col, O.EDGE_ZEROFILL)
    assert np.isfinite(ref).all()
    np.testing.assert_array_equal(bits(got), bits(ref))
