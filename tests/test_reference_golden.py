"""Parity against vectors dumped by the REAL reference (rust/tools/dump_golden.rs, run by a
maintainer on a machine with nightly Rust + ArrayFire 3.6.1; impossible in this image).

The directory tests/golden/reference/ (or $CHEMSIM_REFERENCE_GOLDEN) holds, per case, the inputs
the reference used and the nine populations + readouts after N steps.  When it is present the
oracle (CPU test) and the CUDA path (`-m gpu`) are held to it bit for bit, and "parity unpinned"
(oracle/lbm_oracle.h) becomes "pinned".  When it is absent the tests SKIP and say
so: nothing here may be mistaken for a pinned result.

If the populations come out point-reflected (element (y, x) <-> (H-1-y, W-1-x)), ArrayFire's
convolve2 centres/flips the other way round than SURVEY.md §8 a-2 derived: the tests then name
`chemsim_lbm_set_stream_convention(h, 1)` as the switch instead of just failing."""
import glob
import os
import re

import numpy as np
import pytest

from chemsim_b200 import lbm
from oracle import lbm_oracle as O

REF_DIR = os.environ.get("CHEMSIM_REFERENCE_GOLDEN") or os.path.join(os.path.dirname(__file__), "golden", "reference")
SKIP = ("no reference-generated vectors under %s: parity with the real reference stays UNPINNED "
        "(generate them with rust/tools/dump_golden.rs)" % REF_DIR)


def cases():
    out = []
    for path in sorted(glob.glob(os.path.join(REF_DIR, "*_n*.npy"))):
        m = re.match(r"(.+)_n(\d+)\.npy$", os.path.basename(path))
        if m:
            out.append((m.group(1), int(m.group(2))))
    return out


def collision_of(case):
    op = case.split("_")[2]
    disc = lbm.Discretization(1.0, 1.0)
    trt = lbm.TRT.new(0.25, 0.1, disc, np.float32)
    return {"bgk15": (O.collision(O.BGK, tau=15.0), lbm.BGK(15.0)),
            "bgk08": (O.collision(O.BGK, tau=0.8), lbm.BGK(0.8)),
            "trt": (O.collision(O.TRT, tau_plus=trt.tau_plus, tau_minus=trt.tau_minus), trt),
            "regularized": (O.collision(O.REGULARIZED), lbm.Regularized.new(lbm.KBC.new(10.0))),
            "kbc": (O.collision(O.KBC, viscosity=0.1), lbm.KBC.new(0.1))}[op]


def load(case, steps):
    get = lambda suffix: np.load(os.path.join(REF_DIR, f"{case}_{suffix}.npy"))
    return (get("rho"), get("vx"), get("vy"), get("solid")), get(f"n{steps}")


def explain(got, ref, what):
    u = np.uint32
    if (got.view(u) == ref.view(u)).all():
        return
    mirrored = got[:, ::-1, ::-1]
    if (np.ascontiguousarray(mirrored).view(u) == ref.view(u)).all():
        pytest.fail(f"{what}: the reference's lattice is the POINT REFLECTION of ours — af::convolve2 uses the other "
                    "flip/centre convention; select it with chemsim_lbm_set_stream_convention(h, 1) "
                    "(lbm.State.set_stream_convention(True)) and flip ORACLE_EX/EY")
    err = np.max(np.abs(got.astype(np.float64) - ref) / np.maximum(np.abs(ref), 1e-30))
    pytest.fail(f"{what}: differs from the reference, max rel err {err:g}")


@pytest.mark.skipif(not cases(), reason=SKIP)
@pytest.mark.parametrize("case,steps", cases() or [("none", 0)])
def test_oracle_equals_the_real_reference(case, steps):
    (rho, vx, vy, solid), ref = load(case, steps)
    f0 = O.compute_equilibrium(rho, vx, vy)
    got = O.step_ref(f0, solid, steps, collision_of(case)[0], O.EDGE_ZEROFILL)
    explain(got, ref, f"oracle {case} n={steps}")
    for name, fn in (("density", O.density), ("speed", O.speed)):
        path = os.path.join(REF_DIR, f"{case}_n{steps}_{name}.npy")
        if os.path.exists(path):
            np.testing.assert_array_equal(fn(got).view(np.uint32), np.load(path).view(np.uint32), err_msg=name)
    stats = os.path.join(REF_DIR, f"{case}_n{steps}_density_stats.npy")
    if os.path.exists(stats):                       # af::stdev_all = population sigma (src/render.rs:42)
        mean, std = np.load(stats)
        rho_n = O.density(got).astype(np.float64)
        assert abs(rho_n.mean() - mean) <= 1e-6 * abs(mean)
        assert abs(rho_n.std(ddof=0) - std) <= 1e-4 * abs(std), "stdev_all is not the population standard deviation"


@pytest.mark.gpu
@pytest.mark.skipif(not cases(), reason=SKIP)
@pytest.mark.parametrize("case,steps", cases() or [("none", 0)])
def test_cuda_path_equals_the_real_reference(case, steps):
    (rho, vx, vy, solid), ref = load(case, steps)
    h, w = rho.shape
    m = lambda a: lbm.Matrix.new(a.reshape(-1), (w, h), dtype=np.float32)
    disc = lbm.Discretization(1.0, 1.0)
    pops = lbm.compute_equilibrium(m(rho), (m(vx), m(vy)), lbm.D2Q9.directions(), disc)
    state = lbm.State.initial(lbm.D2Q9.new(pops), solid, collision_of(case)[1], disc)
    state.step(steps)
    explain(state.populations_array(), ref, f"CUDA {case} n={steps}")


def test_reference_vector_ingestion_is_wired(tmp_path, monkeypatch):
    """The ingestion path itself, exercised with vectors written in dump_golden.rs's format by the
    oracle (so this test pins nothing about the reference — it keeps the one-command recipe alive)."""
    from chemsim_b200 import scenarios
    rho, vx, vy, solid = scenarios.main_rs(16, 16, np.float32, radius=3.0)
    f = O.step_ref(O.compute_equilibrium(rho, vx, vy), solid, 2, O.collision(O.BGK, tau=15.0), O.EDGE_ZEROFILL)
    case = "mainrs16_zerofill_bgk15_float32"
    for name, arr in (("rho", rho), ("vx", vx), ("vy", vy), ("solid", solid), ("n2", f), ("n2_density", O.density(f)),
                      ("n2_density_stats", np.array([O.density(f).astype(np.float64).mean(),
                                                     O.density(f).astype(np.float64).std()]))):
        np.save(tmp_path / f"{case}_{name}.npy", arr)
    import test_reference_golden as T
    monkeypatch.setattr(T, "REF_DIR", str(tmp_path))
    assert T.cases() == [(case, 2)]
    T.test_oracle_equals_the_real_reference.__wrapped__(case, 2) if hasattr(T.test_oracle_equals_the_real_reference, "__wrapped__") \
        else T.test_oracle_equals_the_real_reference(case, 2)
    mirrored = np.ascontiguousarray(f[:, ::-1, ::-1])
    with pytest.raises(pytest.fail.Exception, match="POINT REFLECTION"):
        T.explain(mirrored, f, "demo")
