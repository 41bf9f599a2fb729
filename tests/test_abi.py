"""CPU tests of the boundary: the C-ABI library loads and exports every symbol
include/chemsim_lbm.h declares, the Python mirror's host logic behaves like the
reference's, and nothing computes without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from chemsim_b200 import _ffi, build, lbm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _ffi.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "chemsim_lbm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(chemsim_lbm_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    syms = declared_symbols()
    assert len(syms) >= 30
    assert sorted(_ffi.PROTOTYPES) == syms


def test_library_exports_every_declared_symbol(lib):
    raw = C.CDLL(_ffi.LIB_PATH)
    for name in declared_symbols():
        assert getattr(raw, name) is not None
    assert lib.chemsim_lbm_abi_version() == _ffi.ABI_VERSION == 2


def test_no_oracle_in_the_product_path():
    # the product must not import / link the oracle
    pkg = os.path.join(ROOT, "chemsim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "lbm_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_invalid_arguments_are_reported_not_crashed(lib):
    h = C.c_void_p()
    assert lib.chemsim_lbm_create(0, 16, 0, 0, -1, C.byref(h)) == _ffi.ERR_INVALID_ARGUMENT
    assert b"positive" in lib.chemsim_lbm_last_error(None)
    assert lib.chemsim_lbm_create(16, 16, 7, 0, -1, C.byref(h)) == _ffi.ERR_INVALID_ARGUMENT
    assert lib.chemsim_lbm_create(16, 16, 0, 5, -1, C.byref(h)) == _ffi.ERR_INVALID_ARGUMENT
    assert lib.chemsim_lbm_create_slab(16, 16, 0, 0, -1, 2, 2, None, C.byref(h)) == _ffi.ERR_INVALID_ARGUMENT
    assert lib.chemsim_lbm_step(None, 1) == _ffi.ERR_INVALID_ARGUMENT
    assert lib.chemsim_lbm_destroy(None) == _ffi.OK


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.chemsim_lbm_create(16, 16, 0, 0, -1, C.byref(h)) == _ffi.ERR_CUDA
    with pytest.raises(lbm.LbmError):
        lbm.State.create((16, 16), lbm.BGK(0.8))


def test_matrix_new_checks_slice_size():
    m = lbm.Matrix.new(np.arange(12, dtype=np.float32), (4, 3))
    assert m.get_shape() == (4, 3) and m.get_width() == 4 and m.get_height() == 3
    assert m.array[2, 1] == 2 * 4 + 1                      # element (y,x) = slice[y*w+x]
    np.testing.assert_array_equal(m.get_underlying(), np.arange(12, dtype=np.float32))
    with pytest.raises(lbm.InvalidSliceSize):
        lbm.Matrix.new(np.zeros(11, dtype=np.float32), (4, 3))
    assert lbm.Matrix.new_filled(1.0, (5, 2)).array.shape == (2, 5)


def test_directions_table_matches_reference():
    dirs = lbm.D2Q9.directions()
    assert [d.c_vector for d in dirs] == [(0, 0), (1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)]
    assert abs(sum(d.w_scalar for d in dirs) - 1.0) < 1e-15
    # stencil of direction 1 (src/lbm.rs:238-240): a single 1 at row 1, column 0
    assert dirs[1].stencil == (0, 0, 0, 1, 0, 0, 0, 0, 0)
    from oracle import lbm_numpy
    for i, d in enumerate(dirs):
        assert list(d.stencil) == lbm_numpy.STENCILS[i]


def test_host_scalars_of_the_mirror():
    disc = lbm.Discretization(1.0, 1.0)
    assert disc.isothermal_speed_of_sound(np.float32) == np.float32(0.577350259)
    bgk = lbm.BGK(15.0)
    assert bgk.kinematic_shear_viscosity(disc) == np.float32((1.0 / 3.0) * 14.5)
    assert abs(float(bgk.kinematic_bulk_viscosity(disc)) - 2 * 14.5 / 9) < 1e-5


def test_missing_extension_fails_loudly(tmp_path):
    """No silent fallback: if the CUDA library is not there, importing the binding's
    library raises instead of degrading to a CPU / PyTorch path."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['CHEMSIM_LBM_LIB'] = %r\n"
            "from chemsim_b200 import _ffi\n"
            "try:\n    _ffi.load()\nexcept ImportError as e:\n    print('LOUD', 'no CPU' in str(e))\n"
            % (ROOT, str(tmp_path / "nope.so")))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "LOUD True" in res.stdout, res.stdout + res.stderr


def test_rust_binding_declares_every_symbol_of_the_header():
    """rust/src/ffi.rs cannot be compiled here (no rustc), so at least hold it to the header textually:
    every entry point is declared, and nothing is declared that the header does not have."""
    text = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    declared = sorted(set(re.findall(r"pub fn (chemsim_lbm_[a-z0-9_]+)\s*\(", text)))
    assert declared == declared_symbols()


def test_rust_shim_keeps_the_reference_signatures():
    """The items of lbm.rs's public surface that main.rs touches, as text."""
    text = open(os.path.join(ROOT, "rust", "src", "lbm.rs")).read()
    for needle in ("pub type Scalar = f32;", "pub type Populations = Vec<(Direction, Population)>;",
                   "pub type Geometry = crate::af_compat::Array<bool>;",
                   "pub fn new(populations: &[Population]) -> Self",
                   "pub fn directions() -> [Direction; 9]", "pub struct BGK { pub tau: Scalar }",
                   "pub fn new(lambda: Scalar, ks_viscosity: Scalar, disc: &Discretization) -> Self",
                   "pub geometry: Geometry,", "pub time: Scalar,", "pub fn step(&mut self)",
                   "pub fn density(&self) -> Matrix", "pub fn velocity(&self) -> (Matrix, Matrix)",
                   "pub fn momentum_density(&self) -> (Matrix, Matrix)", "pub fn speed(&self) -> Matrix",
                   "pub fn is_unstable(&self) -> bool"):
        assert needle in text, needle
    patch = open(os.path.join(ROOT, "rust", "patches", "main_rs.patch")).read()
    assert patch.count("\n@@") == 2 and "use chemsim::af_compat as af;" in patch


def test_packed_f32_additions_are_never_contracted(lib):
    """The f32 two-step kernels collide two cells at once with packed additions (SASS FADD2, d2q9.cuh
    F32x2).  ptxas contracts a packed multiplication feeding a packed addition into FFMA2 even under
    --fmad=false, which would change the rounding — so the library must contain packed ADDITIONS only:
    no FFMA2, no FMUL2, and FADD2 in every f32 two-step kernel."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _ffi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    counts, fn = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = {"FADD2": 0, "FFMA2": 0, "FMUL2": 0}
            continue
        if fn:
            for op in ("FADD2", "FFMA2", "FMUL2"):
                if re.search(r"\b%s\b" % op, line):
                    counts[fn][op] += 1
    assert counts
    assert all(c["FFMA2"] == 0 and c["FMUL2"] == 0 for c in counts.values())
    step2_f32 = [c for name, c in counts.items() if re.search(r"step2_(slab_p2p_)?kernelIf", name)]
    assert len(step2_f32) >= 12                      # 3 operators x (periodic, mask) variants, plain and peer-memory
    assert all(c["FADD2"] > 100 for c in step2_f32)
    step2_f64 = [c for name, c in counts.items() if re.search(r"step2_(slab_p2p_)?kernelId", name)]
    assert step2_f64 and all(c["FADD2"] == 0 for c in step2_f64)
