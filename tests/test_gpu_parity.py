"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle
on the same inputs.  Tolerances are BASELINE.json's: max relative error <= 1e-5
(f32) / <= 1e-12 (f64); on top of that the kernels follow the reference's operation
order without FMA contraction, so bit-equality with the oracle is asserted too.
"""
import os

import numpy as np
import pytest

from chemsim_b200 import lbm, scenarios
from oracle import lbm_oracle as O

pytestmark = pytest.mark.gpu

import golden_cases  # noqa: E402

GOLDEN = golden_cases.GOLDEN
DTYPES = [np.float32, np.float64]
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


def max_rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)))


def assert_parity(a, b, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, what
    err = max_rel_err(a, b)
    assert err <= TOL[a.dtype], f"{what}: max rel err {err:g}"
    u = np.uint32 if a.dtype == np.float32 else np.uint64
    np.testing.assert_array_equal(np.ascontiguousarray(a).view(u), np.ascontiguousarray(b).view(u),
                                  err_msg=f"{what}: not bit-identical")


def make_state(rho, vx, vy, solid, tau, edge, dtype):
    h, w = rho.shape
    m = lambda a: lbm.Matrix.new(a.reshape(-1), (w, h), dtype=dtype)
    disc = lbm.Discretization(1.0, 1.0)
    pops = lbm.compute_equilibrium(m(rho), (m(vx), m(vy)), lbm.D2Q9.directions(), disc)
    return lbm.State.initial(lbm.D2Q9.new(pops), solid, lbm.BGK(tau), disc, edge=edge)


def check_all_fields(state, f_ref, what):
    f = state.populations_array()
    assert_parity(f, f_ref, what + " f")
    assert_parity(state.density().array, O.density(f_ref), what + " rho")
    ux, uy = state.velocity()
    rux, ruy = O.velocity(f_ref)
    assert_parity(ux.array, rux, what + " ux")
    assert_parity(uy.array, ruy, what + " uy")
    mass = state.total_mass()
    ref_mass = O.total_mass(f_ref)
    assert abs(mass - ref_mass) <= 1e-12 * abs(ref_mass), what + " mass"


@pytest.mark.parametrize("dtype", DTYPES)
def test_config1_main_rs_256_literal(dtype):
    """BASELINE.json config 1 as SURVEY.md §8(d) makes it concrete: main.rs initial_state
    at 256^2, BGK tau=15, zero-fill edges; gate N = 1, 2, 10, 50 incl. the mass series."""
    rho, vx, vy, solid = scenarios.main_rs(256, 256, dtype)
    state = make_state(rho, vx, vy, solid, 15.0, lbm.EDGE_ZEROFILL, dtype)
    f_ref = O.compute_equilibrium(rho, vx, vy)
    assert_parity(state.populations_array(), f_ref, "initial equilibrium")
    col = O.collision(O.BGK, tau=15.0)
    prev = 0
    for n in (1, 2, 10, 50):
        state.step(n - prev)
        f_ref = O.step_ref(f_ref, solid, n - prev, col, O.EDGE_ZEROFILL)
        prev = n
        check_all_fields(state, f_ref, f"N={n}")
    assert abs(state.time - 50.0) < 1e-6
    if dtype == np.float64:   # SURVEY.md §6.2 derived known answers
        assert abs(state.total_mass() - 65205.8375620200) < 1e-8
        assert abs(state.population(1).array[128, 100] - 0.1161610097575969) < 1e-15


@pytest.mark.parametrize("dtype", DTYPES)
def test_config1_stable_twin_periodic_1000_steps(dtype):
    rho, vx, vy, solid = scenarios.main_rs(256, 256, dtype, walls=False)
    state = make_state(rho, vx, vy, solid, 15.0, lbm.EDGE_PERIODIC, dtype)
    m0 = state.total_mass()
    state.step(1000)
    f_ref = O.step_fused(O.compute_equilibrium(rho, vx, vy), solid, 1000, 15.0, O.EDGE_PERIODIC)
    check_all_fields(state, f_ref, "N=1000")
    drift = abs(state.total_mass() - m0) / m0
    assert drift <= (1e-6 if dtype == np.float32 else 1e-12)
    assert not state.is_unstable()


@pytest.mark.parametrize("name", golden_cases.names())
def test_golden_vectors(name):
    case = golden_cases.parse(name)
    dtype = case["dtype"]
    rho, vx, vy, solid = case["inputs"]
    h, w = rho.shape
    m = lambda a: lbm.Matrix.new(a.reshape(-1), (w, h), dtype=dtype)
    disc = lbm.Discretization(1.0, 1.0)
    pops = lbm.compute_equilibrium(m(rho), (m(vx), m(vy)), lbm.D2Q9.directions(), disc)
    state = lbm.State.initial(lbm.D2Q9.new(pops), solid, case["mirror_collision"], disc, edge=case["edge"])
    state.step(case["steps"])
    assert_parity(state.populations_array(), GOLDEN[name], name)


RAGGED = [(1, 1), (2, 1), (1, 5), (3, 2), (4, 4), (5, 3), (8, 1), (31, 7), (32, 3), (33, 2), (36, 5), (124, 3),
          (127, 2), (128, 4), (129, 3), (132, 2), (260, 9), (1028, 3), (2052, 2)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("edge", [lbm.EDGE_ZEROFILL, lbm.EDGE_PERIODIC])
def test_ragged_sizes_with_random_solids(dtype, edge):
    """Edge cases: tiny, odd and non-multiple-of-vector widths (scalar kernel), widths
    that end mid-warp and mid-block (vector kernel), single rows and columns."""
    for (w, h) in RAGGED:
        rho, vx, vy, solid = scenarios.random_state(w, h, dtype, seed=w * 131 + h)
        state = make_state(rho, vx, vy, solid, 0.8, edge, dtype)
        state.step(3)
        f_ref = O.step_fused(O.compute_equilibrium(rho, vx, vy), solid, 3, 0.8, edge)
        assert_parity(state.populations_array(), f_ref, f"{w}x{h} edge={edge}")
        state.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_all_readouts(dtype):
    rho, vx, vy, solid = scenarios.random_state(132, 37, dtype, seed=21)
    state = make_state(rho, vx, vy, solid, 0.9, lbm.EDGE_PERIODIC, dtype)
    state.step(4)
    f = state.populations_array()
    f_ref = O.step_fused(O.compute_equilibrium(rho, vx, vy), solid, 4, 0.9, O.EDGE_PERIODIC)
    assert_parity(f, f_ref, "f")
    assert_parity(state.density().array, O.density(f_ref), "density")
    assert_parity(state.pressure().array, O.pressure(f_ref), "pressure")
    assert_parity(state.speed().array, O.speed(f_ref), "speed")
    for got, ref, name in zip(state.velocity(), O.velocity(f_ref), "xy"):
        assert_parity(got.array, ref, "velocity " + name)
    for got, ref, name in zip(state.momentum_density(), O.momentum_density(f_ref), "xy"):
        assert_parity(got.array, ref, "momentum " + name)
    feq_ref = O.lattice_equilibrium(f_ref)
    feq = np.stack([m.array for _, m in state.equilibrium()])
    assert_parity(feq, feq_ref, "equilibrium")
    fneq = np.stack([m.array for _, m in state.non_equilibrium()])
    assert_parity(fneq, f_ref - feq_ref, "non_equilibrium")
    assert state.is_unstable() == O.is_unstable(f_ref)
    assert abs(state.total_mass() - O.total_mass(f_ref)) <= 1e-12 * O.total_mass(f_ref)
    np.testing.assert_array_equal(state.geometry, solid.astype(bool))
    assert state.size() == (132, 37)


def test_is_unstable_detects_negative_equilibrium():
    dtype = np.float32
    rho, vx, vy, solid = scenarios.main_rs(64, 64, dtype, walls=False, radius=0.0)
    vx[10, 10] = 2.0      # |u| >> cs  ->  f_eq,0 = 4/9 rho (1 - 1.5 u^2) < 0
    state = make_state(rho, vx, vy, solid, 1.0, lbm.EDGE_PERIODIC, dtype)
    f_ref = O.compute_equilibrium(rho, vx, vy)
    assert O.is_unstable(f_ref) and state.is_unstable()


@pytest.mark.parametrize("dtype", DTYPES)
def test_set_get_population_round_trip_and_explicit_populations(dtype):
    rng = np.random.default_rng(5)
    w, h = 100, 17
    f0 = (0.05 + rng.random((9, h, w))).astype(dtype)
    pops = [lbm.Matrix.new(f0[q].reshape(-1), (w, h), dtype=dtype) for q in range(9)]
    solid = rng.random((h, w)) < 0.2
    state = lbm.State.initial(lbm.D2Q9.new(pops), solid, lbm.BGK(1.3), lbm.Discretization(), edge=lbm.EDGE_ZEROFILL)
    assert_parity(state.populations_array(), f0, "round trip")
    state.step(2)
    f_ref = O.step_ref(f0, solid.astype(np.uint8), 2, O.collision(O.BGK, tau=1.3), O.EDGE_ZEROFILL)
    assert_parity(state.populations_array(), f_ref, "explicit populations")


def test_geometry_can_be_rewritten_between_steps():
    """main.rs:77-89 rewrites state.geometry while the simulation runs."""
    dtype = np.float32
    rho, vx, vy, solid = scenarios.main_rs(128, 128, dtype, walls=False, radius=0.0)
    assert not solid.any()
    state = make_state(rho, vx, vy, solid, 0.8, lbm.EDGE_PERIODIC, dtype)
    f_ref = O.compute_equilibrium(rho, vx, vy)
    state.step(3)
    f_ref = O.step_fused(f_ref, solid, 3, 0.8, O.EDGE_PERIODIC)
    solid2 = solid.copy()
    solid2[60:69, 40:49] = 1          # the 9x9 block the mouse handler paints
    state.geometry = solid2
    state.step(5)
    f_ref = O.step_fused(f_ref, solid2, 5, 0.8, O.EDGE_PERIODIC)
    assert_parity(state.populations_array(), f_ref, "after geometry edit")
    state.geometry = solid            # and removed again
    state.step(2)
    f_ref = O.step_fused(f_ref, solid, 2, 0.8, O.EDGE_PERIODIC)
    assert_parity(state.populations_array(), f_ref, "after geometry removal")


def test_discretization_other_than_unity():
    dtype = np.float64
    rho, vx, vy, solid = scenarios.random_state(64, 16, dtype, seed=4)
    h, w = rho.shape
    disc = lbm.Discretization(0.5, 0.25)
    m = lambda a: lbm.Matrix.new(a.reshape(-1), (w, h), dtype=dtype)
    pops = lbm.compute_equilibrium(m(rho), (m(vx), m(vy)), lbm.D2Q9.directions(), disc)
    state = lbm.State.initial(lbm.D2Q9.new(pops), solid, lbm.BGK(0.4), disc, edge=lbm.EDGE_PERIODIC)
    state.step(3)
    f_ref = O.step_fused(O.compute_equilibrium(rho, vx, vy, 0.5, 0.25), solid, 3, 0.4, O.EDGE_PERIODIC, 0.5, 0.25)
    assert_parity(state.populations_array(), f_ref, "dx=0.5 dt=0.25")
    assert abs(state.time - 0.75) < 1e-12


def test_errors_mirror_the_reference():
    state = lbm.State.create((32, 8), lbm.BGK(0.8))
    with pytest.raises(lbm.LbmError) as e:
        state.step()
    assert e.value.status == 5  # NOT_READY: no populations yet
    with pytest.raises(lbm.InvalidSliceSize):     # Matrix::new -> Err(InvalidSliceSize)
        state.init_equilibrium(np.ones(10, np.float32), np.ones(10, np.float32), np.ones(10, np.float32))
    with pytest.raises(lbm.InvalidSliceSize):
        state.geometry = np.zeros((3, 3), bool)


# ---- full-size, size-independent properties (BASELINE.json config 2 / 3 sizes) ----

@pytest.mark.parametrize("dtype", DTYPES)
def test_full_size_4096_against_oracle_and_mass(dtype):
    w = h = 4096
    rho, vx, vy, solid = scenarios.smooth_periodic(w, h, dtype)
    state = lbm.State.create((w, h), lbm.BGK(0.8), dtype=dtype, edge=lbm.EDGE_PERIODIC)
    state.init_equilibrium(rho, vx, vy)
    assert "scalar" not in state.step_kernel_name()          # a vector-width kernel (step2 or step_vec)
    m0 = state.total_mass()
    state.step(2)
    f_ref = O.step_fused(O.compute_equilibrium(rho, vx, vy), None, 2, 0.8, O.EDGE_PERIODIC)
    for q in range(9):
        assert_parity(state.population(q).array, f_ref[q], f"4096^2 q={q}")
    state.step(98)
    drift = abs(state.total_mass() - m0) / m0
    assert drift <= (1e-6 if dtype == np.float32 else 1e-12), drift
    assert not state.is_unstable()


def test_full_size_uniform_state_is_a_fixed_point():
    w = h = 4096
    dtype = np.float32
    rho = np.ones((h, w), dtype)
    vx = np.full((h, w), 0.03, dtype)
    vy = np.full((h, w), -0.02, dtype)
    state = lbm.State.create((w, h), lbm.BGK(0.8), dtype=dtype, edge=lbm.EDGE_PERIODIC)
    state.init_equilibrium(rho, vx, vy)
    f0 = state.population(5).array.copy()
    state.step(50)
    np.testing.assert_allclose(state.population(5).array, f0, rtol=2e-6)
    np.testing.assert_allclose(state.density().array, 1.0, rtol=2e-6)


def test_full_size_channel_8192x2048_transpose_isometry_and_oracle_band():
    """Config 3: channel walls + cylinder.  (a) a band of rows around the cylinder is
    compared with the oracle run on the whole lattice for 2 steps; (b) mass is conserved
    (periodic in x, solid walls bounce everything back)."""
    dtype = np.float32
    w, h = 8192, 2048
    rho, vx, vy, solid = scenarios.channel_cylinder(w, h, dtype)
    state = lbm.State.create((w, h), lbm.BGK(0.8), dtype=dtype, edge=lbm.EDGE_PERIODIC)
    state.init_equilibrium(rho, vx, vy)
    state.geometry = solid
    m0 = state.total_mass()
    state.step(2)
    f_ref = O.step_fused(O.compute_equilibrium(rho, vx, vy), solid, 2, 0.8, O.EDGE_PERIODIC)
    for q in range(9):
        assert_parity(state.population(q).array, f_ref[q], f"channel q={q}")
    state.step(200)
    assert abs(state.total_mass() - m0) / m0 <= 1e-6


# ---- chunked / asynchronous entry points ---------------------------------------------

def test_row_chunked_init_and_geometry_equal_whole_field_calls():
    dtype = np.float32
    rho, vx, vy, solid = scenarios.random_state(192, 50, dtype, seed=31)
    whole = make_state(rho, vx, vy, solid, 0.8, lbm.EDGE_PERIODIC, dtype)
    parts = lbm.State.create((192, 50), lbm.BGK(0.8), dtype=dtype, edge=lbm.EDGE_PERIODIC)
    for r0, r1 in ((0, 7), (7, 32), (32, 50)):
        parts.init_equilibrium_rows(r0, rho[r0:r1], vx[r0:r1], vy[r0:r1])
        parts.set_geometry_rows(r0, solid[r0:r1])
    whole.step(5)
    parts.step(5)
    assert_parity(parts.populations_array(), whole.populations_array(), "chunked upload")
    np.testing.assert_array_equal(parts.geometry, solid.astype(bool))
    with pytest.raises(lbm.LbmError):
        parts.set_geometry_rows(45, solid[:10])           # row range outside the slab
    with pytest.raises(lbm.InvalidSliceSize):
        parts.init_equilibrium_rows(0, rho[:3], vx[:2], vy[:3])


@pytest.mark.parametrize("dtype", DTYPES)
def test_async_geometry_and_density_snapshots(dtype):
    """The pipelined frame loop bench.py's e2e leg uses: asynchronous geometry upload,
    step, asynchronous density snapshot into two alternating pinned buffers."""
    import torch
    w, h = 384, 70
    rho, vx, vy, solid = scenarios.random_state(w, h, dtype, seed=41, solid_fraction=0.02)
    solid[:, 128:256] = 0                     # a solid-free band: exercises the segment-flag skip
    state = make_state(rho, vx, vy, np.zeros_like(solid), 0.8, lbm.EDGE_PERIODIC, dtype)
    masks = [solid, np.zeros_like(solid), solid[::-1].copy()]
    pinned_masks = [torch.from_numpy(m.copy()).pin_memory() for m in masks]
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    out = [torch.empty((h, w), dtype=tdt).pin_memory() for _ in range(2)]
    f_ref = O.compute_equilibrium(rho, vx, vy)
    for i in range(6):
        m = i % 3
        state.set_geometry_async(pinned_masks[m].data_ptr(), pinned_masks[m].numel())
        state.step(1)
        state.density_async(out[i & 1].data_ptr(), out[i & 1].numel())
        f_ref = O.step_fused(f_ref, masks[m], 1, 0.8, O.EDGE_PERIODIC)
        if i >= 1:
            pass
        state.synchronize()
        assert_parity(out[i & 1].numpy(), O.density(f_ref), f"async density frame {i}")
    # without intermediate synchronisation: two snapshots in flight
    for i in range(4):
        state.set_geometry_async(pinned_masks[0].data_ptr(), pinned_masks[0].numel())
        state.step(2)
        state.density_async(out[i & 1].data_ptr(), out[i & 1].numel())
        f_ref = O.step_fused(f_ref, masks[0], 2, 0.8, O.EDGE_PERIODIC)
    state.synchronize()
    assert_parity(out[1].numpy(), O.density(f_ref), "last snapshot")
    assert_parity(state.populations_array(), f_ref, "populations after async frames")


def test_solid_free_warps_skip_the_mask_but_results_are_identical():
    """Geometry with solids only in a corner: most warps take the flag-skip path."""
    dtype = np.float32
    w, h = 1024, 24
    rho, vx, vy, _ = scenarios.random_state(w, h, dtype, seed=51)
    solid = np.zeros((h, w), np.uint8)
    solid[3:6, 1000:1010] = 1
    solid[20, 63] = 1
    solid[21, 64] = 1
    state = make_state(rho, vx, vy, solid, 0.8, lbm.EDGE_ZEROFILL, dtype)
    state.step(6)
    f_ref = O.step_fused(O.compute_equilibrium(rho, vx, vy), solid, 6, 0.8, O.EDGE_ZEROFILL)
    assert_parity(state.populations_array(), f_ref, "sparse solids")


# ---- the other CollisionOperator impls of lbm.rs -------------------

def _operators(dtype):
    disc = lbm.Discretization(1.0, 1.0)
    trt = lbm.TRT.new(0.25, 0.1, disc, dtype)
    return {
        "trt": (trt, O.collision(O.TRT, tau_plus=trt.tau_plus, tau_minus=trt.tau_minus)),
        "regularized": (lbm.Regularized.new(lbm.KBC.new(10.0)), O.collision(O.REGULARIZED)),
        "kbc": (lbm.KBC.new(0.1), O.collision(O.KBC, viscosity=0.1)),
    }


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("edge", [lbm.EDGE_ZEROFILL, lbm.EDGE_PERIODIC])
@pytest.mark.parametrize("name", ["trt", "regularized", "kbc"])
def test_other_collision_operators_match_oracle(name, edge, dtype):
    op, ocol = _operators(dtype)[name]
    for (w, h) in ((40, 24), (37, 9), (256, 12)):       # vector kernel, scalar kernel, multi-warp rows
        rho, vx, vy, solid = scenarios.random_state(w, h, dtype, seed=w + h)
        m = lambda a: lbm.Matrix.new(a.reshape(-1), (w, h), dtype=dtype)
        disc = lbm.Discretization(1.0, 1.0)
        pops = lbm.compute_equilibrium(m(rho), (m(vx), m(vy)), lbm.D2Q9.directions(), disc)
        state = lbm.State.initial(lbm.D2Q9.new(pops), solid, op, disc, edge=edge)
        state.step(3)
        f_ref = O.step_ref(O.compute_equilibrium(rho, vx, vy), solid, 3, ocol, edge)
        assert np.isfinite(f_ref).all()
        assert_parity(state.populations_array(), f_ref, f"{name} {w}x{h}")
        state.close()


def test_main_rs_active_configuration_regularized_kbc_400():
    """What the reference's binary actually runs (src/main.rs:198-199, :343): 400x400,
    Regularized<KBC(viscosity 10)>, zero-fill edges, walls + cylinder."""
    dtype = np.float32
    w = h = 400
    rho, vx, vy, solid = scenarios.main_rs(w, h, dtype)
    m = lambda a: lbm.Matrix.new(a.reshape(-1), (w, h), dtype=dtype)
    disc = lbm.Discretization(1.0, 1.0)
    collision = lbm.Regularized.new(lbm.KBC.new(10.0))
    assert collision.kinematic_shear_viscosity(disc) == np.float32(10.0)
    pops = lbm.compute_equilibrium(m(rho), (m(vx), m(vy)), lbm.D2Q9.directions(), disc)
    state = lbm.State.initial(lbm.D2Q9.new(pops), solid, collision, disc)
    f_ref = O.compute_equilibrium(rho, vx, vy)
    prev = 0
    for n in (1, 2, 10, 40):
        state.step(n - prev)
        f_ref = O.step_ref(f_ref, solid, n - prev, O.collision(O.REGULARIZED), O.EDGE_ZEROFILL)
        prev = n
        check_all_fields(state, f_ref, f"regularized N={n}")


def test_trt_host_scalars():
    disc = lbm.Discretization(1.0, 1.0)
    trt = lbm.TRT.new(0.25, 10.0, disc)          # main.rs:189-190
    assert np.float32(trt.tau_plus) == np.float32(1.0) * (np.float32(10.0) / (np.float32(0.577350259) ** 2) + np.float32(0.5))
    assert abs(float(trt.lambda_(disc)) - 0.25) < 1e-6
    assert abs(float(trt.kinematic_shear_viscosity(disc)) - 10.0) < 1e-5


def test_full_size_translation_equivariance_4096():
    """Size-independent property at config-2 size: on a periodic lattice, shifting the input by
    (dy, dx) cells shifts the output by (dy, dx) — bit for bit.  Exercises the x wrap in the
    lane-edge loads, the y wrap and every warp/block boundary of the vector kernel."""
    dtype = np.float32
    w = h = 4096
    rho, vx, vy, _ = scenarios.smooth_periodic(w, h, dtype)
    rng = np.random.default_rng(3)
    rho = (rho + 0.001 * rng.random((h, w))).astype(dtype)       # break the analytic symmetry
    solid = np.zeros((h, w), np.uint8)
    solid[1000:1040, 2000:2300] = 1
    solid[0, 5] = 1
    dy, dx = 37, 4093

    def run(r, a, b, s):
        st = lbm.State.create((w, h), lbm.BGK(0.8), dtype=dtype, edge=lbm.EDGE_PERIODIC)
        st.init_equilibrium(r, a, b)
        st.geometry = s
        st.step(20)
        out = [st.population(q).array for q in (0, 2, 5, 7)]
        st.close()
        return out

    base = run(rho, vx, vy, solid)
    roll = lambda a: np.roll(np.roll(a, dy, axis=0), dx, axis=1)
    shifted = run(roll(rho), roll(vx), roll(vy), roll(solid))
    for a, b in zip(base, shifted):
        np.testing.assert_array_equal(roll(a).view(np.uint32), b.view(np.uint32))


def test_large_lattice_16384_needs_64bit_offsets():
    """Maximum-size style case: 16384^2 f32 (18 GiB of populations, plane offsets beyond 2^31
    elements), initialised in row chunks; a uniform periodic state must stay a fixed point and
    the f64 mass reduction must equal the analytic value."""
    dtype = np.float32
    w = h = 16384
    state = lbm.State.create((w, h), lbm.BGK(0.8), dtype=dtype, edge=lbm.EDGE_PERIODIC)
    chunk = 2048
    rho = np.ones((chunk, w), dtype)
    vx = np.full((chunk, w), 0.03, dtype)
    vy = np.full((chunk, w), -0.02, dtype)
    for r in range(0, h, chunk):
        state.init_equilibrium_rows(r, rho, vx, vy)
    f_cell = O.compute_equilibrium(rho[:1, :1], vx[:1, :1], vy[:1, :1])[:, 0, 0]      # the nine values of one cell
    mass0 = state.total_mass()
    assert abs(mass0 - float(np.sum(f_cell.astype(np.float64))) * w * h) <= 1e-9 * mass0
    state.step(10)
    for q in (0, 8):                                   # first and last plane (largest offsets)
        got = state.population(q).array
        np.testing.assert_allclose(got[::1021, ::509], f_cell[q], rtol=2e-6)
        assert got[h - 1, w - 1] == got[0, 0]
    assert abs(state.total_mass() - mass0) <= 2e-7 * mass0      # ~10 steps of the (sum w - 1)/tau drift
    state.close()


# ---- live geometry edit on the device, async snapshots of every field, checkpoints (f-3, f-4) ----

def brush_mask(w, h, pos):
    """The reference's mouse handler, literally (src/main.rs:71-91): EVERY cell is rewritten;
    solid iff |a - x| < 5 and |b - y| < 5 with x = floor(pos[1]), y = floor(pos[0]); a indexes
    rows (dim0), b columns (dim1) of the square geometry array."""
    x, y = int(np.floor(pos[1])), int(np.floor(pos[0]))
    a = np.arange(h)[:, None]
    b = np.arange(w)[None, :]
    return ((np.abs(a - x) < 5) & (np.abs(b - y) < 5)).astype(np.uint8)


@pytest.mark.parametrize("dtype", DTYPES)
def test_paint_brush_on_the_device_equals_the_host_rewrite(dtype):
    w = h = 192
    rho, vx, vy, solid = scenarios.main_rs(w, h, dtype, walls=True, radius=20.0)
    state = make_state(rho, vx, vy, solid, 0.8, lbm.EDGE_ZEROFILL, dtype)
    f_ref = O.compute_equilibrium(rho, vx, vy)
    state.step(2)
    f_ref = O.step_fused(f_ref, solid, 2, 0.8, O.EDGE_ZEROFILL)
    # interior, clipped at the left/top edge, clipped at the right/bottom edge, outside (ignored)
    for pos in ((100.7, 50.2), (1.0, 2.9), (190.5, 191.0), (500.0, 10.0)):
        state.paint_brush(pos)
        inside = 0 <= int(pos[1]) < h and 0 <= int(pos[0]) < w            # main.rs:76
        if inside:
            solid = brush_mask(w, h, pos)
        np.testing.assert_array_equal(state.geometry, solid.astype(bool))
        state.step(3)
        f_ref = O.step_fused(f_ref, solid, 3, 0.8, O.EDGE_ZEROFILL)
        assert_parity(state.populations_array(), f_ref, f"after brush at {pos}")
    # erase with a rectangle, add a wide one that crosses many flag segments
    state.paint_rect(0, 0, w, h, False)
    state.paint_rect(-7, 100, 150, 3, True)
    solid = np.zeros((h, w), np.uint8)
    solid[100:103, 0:143] = 1
    np.testing.assert_array_equal(state.geometry, solid.astype(bool))
    state.step(4)
    f_ref = O.step_fused(f_ref, solid, 4, 0.8, O.EDGE_ZEROFILL)
    assert_parity(state.populations_array(), f_ref, "after paint_rect")
    state.fill_geometry(False)
    state.step(2)
    f_ref = O.step_fused(f_ref, np.zeros_like(solid), 2, 0.8, O.EDGE_ZEROFILL)
    assert_parity(state.populations_array(), f_ref, "after fill_geometry(False)")
    assert "scalar" not in state.step_kernel_name()          # a vector-width kernel (step2 or step_vec)


@pytest.mark.parametrize("dtype", DTYPES)
def test_async_snapshots_of_every_field(dtype):
    """chemsim_lbm_get_async for all of main.rs's display modes (main.rs:157-174) and the rest of
    the readout surface: snapshots queued between steps, read after one synchronize."""
    import torch
    from chemsim_b200 import _ffi
    w, h = 260, 33
    rho, vx, vy, solid = scenarios.random_state(w, h, dtype, seed=61, solid_fraction=0.05)
    state = make_state(rho, vx, vy, solid, 0.9, lbm.EDGE_PERIODIC, dtype)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    buf = lambda: torch.empty((h, w), dtype=tdt).pin_memory()
    f_ref = O.compute_equilibrium(rho, vx, vy)
    step_ref = lambda f, n: O.step_fused(f, solid, n, 0.9, O.EDGE_PERIODIC)

    # (field, q, reference) for two snapshots in flight at a time, steps in between
    state.step(2); f_ref = step_ref(f_ref, 2)
    a0, a1 = buf(), buf()
    state.get_async(_ffi.FIELD_VELOCITY, a0.data_ptr(), a0.numel(), a1.data_ptr())
    want_v = O.velocity(f_ref)
    state.step(1); f_ref = step_ref(f_ref, 1)
    b0 = buf()
    state.get_async(_ffi.FIELD_SPEED, b0.data_ptr(), b0.numel())
    want_s = O.speed(f_ref)
    state.step(3); f_ref = step_ref(f_ref, 3)
    state.synchronize()
    assert_parity(a0.numpy(), want_v[0], "async velocity x")
    assert_parity(a1.numpy(), want_v[1], "async velocity y")
    assert_parity(b0.numpy(), want_s, "async speed")

    c0, c1, d0 = buf(), buf(), buf()
    state.get_async(_ffi.FIELD_MOMENTUM_DENSITY, c0.data_ptr(), c0.numel(), c1.data_ptr())
    want_m = O.momentum_density(f_ref)
    state.get_async(_ffi.FIELD_POPULATION, d0.data_ptr(), d0.numel(), q=7)
    want_p = f_ref[7].copy()
    state.step(2); f_ref = step_ref(f_ref, 2)             # the lattice buffer the snapshot came from is rewritten
    state.synchronize()
    assert_parity(c0.numpy(), want_m[0], "async momentum x")
    assert_parity(c1.numpy(), want_m[1], "async momentum y")
    assert_parity(d0.numpy(), want_p, "async population 7")

    e0, g0 = buf(), buf()
    state.get_async(_ffi.FIELD_PRESSURE, e0.data_ptr(), e0.numel())
    state.get_async(_ffi.FIELD_NON_EQUILIBRIUM, g0.data_ptr(), g0.numel(), q=3)
    state.synchronize()
    assert_parity(e0.numpy(), O.pressure(f_ref), "async pressure")
    assert_parity(g0.numpy(), f_ref[3] - O.lattice_equilibrium(f_ref)[3], "async non-equilibrium 3")
    h0 = buf()
    state.get_async(_ffi.FIELD_EQUILIBRIUM, h0.data_ptr(), h0.numel(), q=5)
    state.get_async(_ffi.FIELD_DENSITY, e0.data_ptr(), e0.numel())
    state.synchronize()
    assert_parity(h0.numpy(), O.lattice_equilibrium(f_ref)[5], "async equilibrium 5")
    assert_parity(e0.numpy(), O.density(f_ref), "async density")
    with pytest.raises(lbm.LbmError):
        state.get_async(_ffi.FIELD_VELOCITY, a0.data_ptr(), a0.numel())          # second component missing
    with pytest.raises(lbm.LbmError):
        state.get_async(_ffi.FIELD_DENSITY, a0.data_ptr(), a0.numel(), a1.data_ptr())
    with pytest.raises(lbm.LbmError):
        state.get_async(_ffi.FIELD_POPULATION, a0.data_ptr(), a0.numel(), q=9)
    assert_parity(state.populations_array(), f_ref, "populations after the async readouts")


@pytest.mark.parametrize("dtype", DTYPES)
def test_checkpoint_restore_round_trip(dtype):
    w, h = 132, 29
    rho, vx, vy, solid = scenarios.random_state(w, h, dtype, seed=71)
    state = make_state(rho, vx, vy, solid, 0.8, lbm.EDGE_PERIODIC, dtype)
    state.step(5)
    blob = state.checkpoint()
    hdr = blob[:8].tobytes()
    header_bytes = int(blob[8:12].view(np.uint32)[0])
    assert hdr == b"CSLBMCK1" and header_bytes == 112
    assert blob.size == header_bytes + 9 * w * h * np.dtype(dtype).itemsize + w * h
    assert int(blob[48:52].view(np.uint32)[0]) == 5                        # step_index
    f5 = state.populations_array()
    state.step(7)
    f12 = state.populations_array()
    t12 = state.time
    # restore into the same handle: back to step 5, then the same 7 steps again
    state.restore(blob)
    assert abs(state.time - 5.0) < 1e-6
    assert_parity(state.populations_array(), f5, "restored populations")
    np.testing.assert_array_equal(state.geometry, solid.astype(bool))
    state.step(7)
    assert_parity(state.populations_array(), f12, "replayed steps")
    assert state.time == t12
    # ... and into a fresh handle with a different geometry/populations
    other = lbm.State.create((w, h), lbm.BGK(0.8), dtype=dtype, edge=lbm.EDGE_PERIODIC)
    other.restore(blob.tobytes())
    other.step(7)
    assert_parity(other.populations_array(), f12, "fresh handle")
    f_ref = O.step_fused(O.compute_equilibrium(rho, vx, vy), solid, 12, 0.8, O.EDGE_PERIODIC)
    assert_parity(f12, f_ref, "oracle")
    # a checkpoint of another lattice is refused
    wrong = lbm.State.create((w, h + 1), lbm.BGK(0.8), dtype=dtype)
    with pytest.raises(lbm.LbmError):
        wrong.restore(blob)
    with pytest.raises(lbm.LbmError):
        other.restore(blob[:200].copy())
    bad = blob.copy(); bad[0] = 0
    with pytest.raises(lbm.LbmError):
        other.restore(bad)


def test_regularized_reports_the_underlying_viscosity_with_the_state_discretization():
    """Regularized::kinematic_shear_viscosity(disc) = underlying.kinematic_shear_viscosity(disc)
    (src/lbm.rs:663-665) — with the State's dx/dt, not the unit discretization."""
    import ctypes as C
    from chemsim_b200 import _ffi
    disc = lbm.Discretization(0.5, 0.25)
    op = lbm.Regularized.new(lbm.BGK(0.9))
    state = lbm.State.create((16, 4), op, disc, dtype=np.float32)
    out = C.c_double()
    _ffi.check(_ffi.load().chemsim_lbm_kinematic_shear_viscosity(state._h, C.byref(out)), state._h)
    assert np.float32(out.value) == lbm.BGK(0.9).kinematic_shear_viscosity(disc, np.float32)
    assert _ffi.load().chemsim_lbm_kinematic_bulk_viscosity(state._h, None) == _ffi.ERR_INVALID_ARGUMENT


# ---- the other reading of af::convolve2 (chemsim_lbm_set_stream_convention) --------------------------

@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("edge", [lbm.EDGE_ZEROFILL, lbm.EDGE_PERIODIC])
def test_mirrored_stream_convention_matches_the_correlation_restatement(dtype, edge):
    """Convention 1 = the stencil applied unflipped: compared with oracle/lbm_numpy.py's literal
    scipy.signal.correlate2d calls (an implementation that shares nothing with the switch, which
    reverses the fields at the C-ABI boundary), for populations, every readout, the geometry, the
    painted rectangle and the rendered image."""
    from oracle import lbm_numpy as N
    w, h = 44, 19
    rho, vx, vy, solid = scenarios.random_state(w, h, dtype, seed=81)
    k = N.Consts(dtype, 1.0, 1.0)
    periodic = edge == lbm.EDGE_PERIODIC
    state = lbm.State.create((w, h), lbm.BGK(0.8), dtype=dtype, edge=edge)
    state.set_stream_convention(True)
    state.init_equilibrium(rho, vx, vy)
    state.geometry = solid
    np.testing.assert_array_equal(state.geometry, solid.astype(bool))
    f = N.compute_equilibrium(rho, vx, vy, k)
    assert_parity(state.populations_array(), f, "initial state")
    state.step(5)                                        # two double steps + one single
    for _ in range(5):
        f = N.step(f, solid.astype(bool), k, ("bgk", 0.8), periodic, mirrored=True)
    assert_parity(state.populations_array(), f.astype(dtype), "mirrored stream, 5 steps")
    assert_parity(state.density().array, N.density(f).astype(dtype), "density")
    for got, ref, name in zip(state.velocity(), N.velocity(f, k), "xy"):
        assert_parity(got.array, ref.astype(dtype), "velocity " + name)
    # ... and it differs from convention 0 (otherwise the test proves nothing)
    plain = lbm.State.create((w, h), lbm.BGK(0.8), dtype=dtype, edge=edge)
    plain.init_equilibrium(rho, vx, vy)
    plain.geometry = solid
    plain.step(5)
    assert not np.array_equal(plain.populations_array(), state.populations_array())
    # the same run equals the point reflection of convention 0 on reflected inputs
    rev = lambda a: np.ascontiguousarray(a[..., ::-1, ::-1])
    refl = lbm.State.create((w, h), lbm.BGK(0.8), dtype=dtype, edge=edge)
    refl.init_equilibrium(rev(rho), rev(vx), rev(vy))
    refl.geometry = rev(solid)
    refl.step(5)
    assert_parity(state.populations_array(), rev(refl.populations_array()), "reflection identity")
    # explicit populations, painted rectangle, image
    state.paint_rect(3, 2, 9, 4, True)
    solid2 = solid.copy(); solid2[2:6, 3:12] = 1
    np.testing.assert_array_equal(state.geometry, solid2.astype(bool))
    for q in range(9):
        state.set_population(q, f[q].astype(dtype))
    assert_parity(state.populations_array(), f.astype(dtype), "set/get population")
    img, img_refl = state.render(1), None
    refl.geometry = rev(solid2)
    for q in range(9):
        refl.set_population(q, rev(f[q].astype(dtype)))
    img_refl = refl.render(1)
    assert int(np.abs(img.astype(np.int16) - img_refl[::-1, ::-1].astype(np.int16)).max()) <= 1
    with pytest.raises(lbm.LbmError):
        state.set_stream_convention(False)               # not after the upload
