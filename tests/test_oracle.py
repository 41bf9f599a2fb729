"""CPU tests (no GPU): the C oracle against the literal numpy/scipy restatement,
the committed golden vectors and the derived known-answer numbers of SURVEY.md
§6.2.  The reference itself has no tests; parity is unpinned."""
import os

import numpy as np
import pytest

from chemsim_b200 import scenarios
from oracle import lbm_numpy as N
from oracle import lbm_oracle as O

import golden_cases  # noqa: E402  (tests/ is on sys.path under pytest's rootdir conftest)

GOLDEN = golden_cases.GOLDEN
DTYPES = [np.float32, np.float64]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32 if a.dtype == np.float32 else np.uint64)


def assert_bit_equal(a, b):
    assert a.dtype == b.dtype and a.shape == b.shape
    np.testing.assert_array_equal(bits(a), bits(b))


def test_shift_table_is_derived_from_the_literal_convolution():
    # ORACLE_EY/EX in lbm_oracle.c == what convolve2d(f, stencil^T) does
    ey = [0, -1, 0, 1, 0, -1, 1, 1, -1]
    ex = [0, 0, 1, 0, -1, 1, 1, -1, -1]
    for i in range(9):
        assert N.derived_shift(i) == (ey[i], ex[i])
        # (dy, dx) = (-c_x, +c_y): an isometry of the lattice
        assert (ey[i], ex[i]) == (-N.CX[i], N.CY[i])


def test_host_constants_match_survey():
    k = O.constants(np.float32)
    assert k["cs2"] == np.float32(0.333333313) and k["cs4"] == np.float32(0.111111097)
    assert k["k1"] == np.float32(3.00000024) and k["k2"] == np.float32(4.50000048)
    assert k["k3"] == np.float32(-1.50000012)
    k = O.constants(np.float64)
    assert k["k1"] == 2.9999999999999991 and k["k2"] == 4.4999999999999973 and k["k3"] == -1.4999999999999996


@pytest.mark.parametrize("dtype", DTYPES)
def test_initial_equilibrium_analytic(dtype):
    # rho=1, u=(0.02,0): f0=4/9*0.9994, f1=1/9*1.0612, f3=1/9*0.9412, ...
    rho, vx, vy, _ = scenarios.main_rs(8, 8, dtype)
    f = O.compute_equilibrium(rho, vx, vy)
    expect = [4 / 9 * 0.9994, 1 / 9 * 1.0612, 1 / 9 * 0.9994, 1 / 9 * 0.9412, 1 / 9 * 0.9994,
              1 / 36 * 1.0612, 1 / 36 * 0.9412, 1 / 36 * 0.9412, 1 / 36 * 1.0612]
    tol = 1e-6 if dtype == np.float32 else 1e-14
    for i in range(9):
        np.testing.assert_allclose(f[i], expect[i], rtol=tol)
    np.testing.assert_allclose(O.density(f), 1.0, rtol=tol)


def test_known_answers_main_rs_256_f64():
    """SURVEY.md §6.2 spot values (derived by the survey's own throwaway restatement)."""
    rho, vx, vy, solid = scenarios.main_rs(256, 256, np.float64)
    f = O.compute_equilibrium(rho, vx, vy)
    col = O.collision(O.BGK, tau=15.0)
    mass = {1: 65365.3421777778, 2: 65360.4220250771, 10: 65324.3377294086, 50: 65205.8375620200}
    f1 = {1: 0.11791111111111112, 2: 0.11791111111111112, 10: 0.11727809081775965, 50: 0.1161610097575969}
    f5 = {1: 0.02947777777777778, 2: 0.026979438204001879}
    rho00 = {1: 0.70261111111111096, 2: 0.67475868394235217, 50: 0.34080113191318356}
    prev = 0
    for n in (1, 2, 10, 50):
        f = O.step_ref(f, solid, n - prev, col)
        prev = n
        assert abs(O.total_mass(f) - mass[n]) < 1e-9
        assert abs(f[1][128][100] - f1[n]) <= 2e-16
        if n in f5:
            assert abs(f[5][1][1] - f5[n]) <= 1e-17
        if n in rho00:
            assert abs(O.density(f)[0][0] - rho00[n]) <= 3e-16
    assert not O.is_unstable(f)


def test_known_answers_periodic_twin_f64():
    rho, vx, vy, solid = scenarios.main_rs(256, 256, np.float64, walls=False)
    f = O.compute_equilibrium(rho, vx, vy)
    f = O.step_fused(f, solid, 100, 15.0, edge=O.EDGE_PERIODIC)
    assert abs(f[1][128][100] - 0.11569394721835483) <= 2e-16
    assert abs(O.density(f)[100][128] - 0.97846645088456441) <= 2e-16
    f = O.step_fused(f, solid, 900, 15.0, edge=O.EDGE_PERIODIC)
    assert abs(f[1][128][100] - 0.112608525951684) <= 1e-15
    assert abs(O.total_mass(f) - 65536.0) < 1e-6  # mass conserved to rounding


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("colname", ["bgk", "trt", "regularized", "kbc"])
def test_c_oracle_equals_numpy_restatement_bitwise(dtype, periodic, colname):
    rho, vx, vy, solid = scenarios.random_state(21, 13, dtype, seed=3)
    k = N.Consts(dtype, 1.0, 1.0)
    f_np = N.compute_equilibrium(rho, vx, vy, k)
    f_c = O.compute_equilibrium(rho, vx, vy)
    assert_bit_equal(f_np, f_c)
    ncol = {"bgk": ("bgk", 0.9), "trt": ("trt", 0.8, 1.1), "regularized": ("regularized",), "kbc": ("kbc", 0.1)}[colname]
    ccol = {"bgk": O.collision(O.BGK, tau=0.9), "trt": O.collision(O.TRT, tau_plus=0.8, tau_minus=1.1),
            "regularized": O.collision(O.REGULARIZED), "kbc": O.collision(O.KBC, viscosity=0.1)}[colname]
    edge = O.EDGE_PERIODIC if periodic else O.EDGE_ZEROFILL
    for _ in range(4):
        f_np = N.step(f_np, solid.astype(bool), k, ncol, periodic)
        f_c = O.step_ref(f_c, solid, 1, ccol, edge)
        assert_bit_equal(f_np, f_c)
    assert np.isfinite(f_c).all()


@pytest.mark.parametrize("dtype", DTYPES)
def test_readouts_equal_numpy_bitwise(dtype):
    rho, vx, vy, solid = scenarios.random_state(17, 9, dtype, seed=5)
    k = N.Consts(dtype, 1.0, 1.0)
    f = O.step_ref(O.compute_equilibrium(rho, vx, vy), solid, 3, O.collision(O.BGK, tau=0.7), O.EDGE_PERIODIC)
    assert_bit_equal(O.density(f), N.density(f))
    mx, my = O.momentum_density(f)
    nmx, nmy = N.momentum_density(f, k)
    assert_bit_equal(mx, nmx) and assert_bit_equal(my, nmy)
    ux, uy = O.velocity(f)
    nux, nuy = N.velocity(f, k)
    assert_bit_equal(ux, nux) and assert_bit_equal(uy, nuy)
    assert_bit_equal(O.speed(f), N.speed(f, k))
    assert_bit_equal(O.pressure(f), N.density(f) * k.cs2)
    assert_bit_equal(O.lattice_equilibrium(f), N.equilibrium(f, k))
    assert abs(O.total_mass(f) - float(np.sum(f.astype(np.float64)))) < 1e-9


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("edge", [O.EDGE_ZEROFILL, O.EDGE_PERIODIC])
def test_fused_equals_three_pass_bitwise(dtype, edge):
    rho, vx, vy, solid = scenarios.random_state(37, 19, dtype, seed=11)
    f = O.compute_equilibrium(rho, vx, vy)
    a = O.step_ref(f, solid, 6, O.collision(O.BGK, tau=0.8), edge)
    b = O.step_fused(f, solid, 6, 0.8, edge)
    assert_bit_equal(a, b)


@pytest.mark.parametrize("name", golden_cases.names())
def test_c_oracle_reproduces_golden_vectors(name):
    case = golden_cases.parse(name)
    rho, vx, vy, solid = case["inputs"]
    f = O.step_ref(O.compute_equilibrium(rho, vx, vy), solid, case["steps"], case["oracle_collision"], case["edge"])
    assert_bit_equal(f, GOLDEN[name])
    assert case["mirror_collision"] is not None        # the host mirror's operator object builds on CPU


@pytest.mark.parametrize("dtype", DTYPES)
def test_uniform_periodic_state_is_a_fixed_point(dtype):
    # SURVEY.md §8(c) self-consistency: periodic, no solids, u = const
    rho, vx, vy, _ = scenarios.main_rs(16, 12, dtype, walls=False, radius=0.0)
    f0 = O.compute_equilibrium(rho, vx, vy)
    f = O.step_fused(f0, None, 5, 0.8, O.EDGE_PERIODIC)
    np.testing.assert_allclose(f, f0, rtol=1e-6 if dtype == np.float32 else 1e-14)


def test_lattice_isometry_transpose():
    # reflecting the inputs across the diagonal permutes the outputs:
    # (y,x)->(x,y) maps shift (ey,ex)->(ex,ey), i.e. c=(cx,cy) -> (-cy,-cx)
    dtype = np.float64
    rho, vx, vy, solid = scenarios.random_state(14, 14, dtype, seed=2)
    f = O.step_fused(O.compute_equilibrium(rho, vx, vy), solid, 7, 0.8, O.EDGE_PERIODIC)
    # reflected problem: u' = (-vy^T, -vx^T)
    g = O.step_fused(O.compute_equilibrium(rho.T.copy(), -vy.T.copy(), -vx.T.copy()), solid.T.copy(), 7, 0.8,
                     O.EDGE_PERIODIC)
    perm = [[(N.CX[j], N.CY[j]) for j in range(9)].index((-N.CY[i], -N.CX[i])) for i in range(9)]
    for i in range(9):
        np.testing.assert_allclose(g[perm[i]], f[i].T, rtol=1e-12, atol=1e-15)


def test_mass_conserved_with_solids_periodic():
    rho, vx, vy, solid = scenarios.random_state(32, 20, np.float64, seed=9)
    f0 = O.compute_equilibrium(rho, vx, vy)
    f = O.step_fused(f0, solid, 50, 0.8, O.EDGE_PERIODIC)
    assert abs(O.total_mass(f) - O.total_mass(f0)) < 1e-9 * O.total_mass(f0)


def test_f32_mass_drift_is_the_weight_rounding():
    """DESIGN.md §1 finding: in f32 the reference's weights 16/36, 4/36, 1/36 sum to 1 + 7.45e-9,
    so BGK creates mass at (sum(w) - 1)/tau per step — the drift bench.py reports is the
    reference's own arithmetic, not a conservation bug.  In f64 the sum is exact to 1e-16."""
    w32 = [np.float32(n) / np.float32(36.0) for n in (16, 4, 4, 4, 4, 1, 1, 1, 1)]
    eps = float(np.sum(np.array(w32, dtype=np.float64))) - 1.0
    assert abs(eps - 7.45e-9) < 0.05e-9
    tau, steps = 0.8, 400
    rho, vx, vy, _ = scenarios.smooth_periodic(128, 96, np.float32)
    f0 = O.compute_equilibrium(rho, vx, vy)
    f = O.step_fused(f0, None, steps, tau, O.EDGE_PERIODIC)
    drift = O.total_mass(f) / O.total_mass(f0) - 1.0
    predicted = eps / tau * steps
    assert 0.7 * predicted < drift < 1.3 * predicted, (drift, predicted)
    rho, vx, vy, _ = scenarios.smooth_periodic(128, 96, np.float64)
    g0 = O.compute_equilibrium(rho, vx, vy)
    g = O.step_fused(g0, None, steps, tau, O.EDGE_PERIODIC)
    assert abs(O.total_mass(g) / O.total_mass(g0) - 1.0) < 1e-13
