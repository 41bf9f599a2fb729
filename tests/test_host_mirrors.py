"""CPU tests of the host-side mirrors and scenario builders (no GPU)."""
import os
import subprocess

import numpy as np

from chemsim_b200 import build, scenarios

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_mirror_compiles_warning_free(tmp_path):
    """chemsim_b200/cpp/lbm.hpp + the main.rs-shaped driver build with -Wall -Wextra -Werror
    against the C-ABI library (link check of every entry point the mirror uses)."""
    build.build()
    out = tmp_path / "harness"
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-pedantic",
           os.path.join(ROOT, "chemsim_b200", "cpp", "main_rs_harness.cpp"), "-o", str(out),
           "-L" + os.path.join(ROOT, "chemsim_b200"), "-lchemsim_lbm",
           "-Wl,-rpath," + os.path.join(ROOT, "chemsim_b200")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    # without a GPU the driver must fail loudly (status 3 = CUDA error), never fall back
    import torch
    if not torch.cuda.is_available():
        run = subprocess.run([str(out), "32", "32", "1"], capture_output=True, text=True)
        assert run.returncode == 1 and "LbmError 3" in run.stderr


def test_header_is_plain_c(tmp_path):
    """include/chemsim_lbm.h is consumable from C (the FFI boundary), not only C++."""
    src = tmp_path / "t.c"
    src.write_text('#include "chemsim_lbm.h"\nint main(void){ chemsim_lbm_t *h = 0; (void)h; '
                   'return chemsim_lbm_abi_version() == CHEMSIM_LBM_ABI_VERSION ? 0 : 1; }\n')
    build.build()
    exe = tmp_path / "t"
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
           str(src), "-o", str(exe), "-L" + os.path.join(ROOT, "chemsim_b200"), "-lchemsim_lbm",
           "-Wl,-rpath," + os.path.join(ROOT, "chemsim_b200")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert subprocess.run([str(exe)]).returncode == 0


def test_main_rs_scenario_geometry():
    """main.rs:269-312: disc of radius 25 (f64 compare) at (w/2, h/2) plus the four border lines."""
    rho, vx, vy, solid = scenarios.main_rs(256, 256, np.float32)
    assert rho.dtype == np.float32 and (rho == 1).all() and (vx == np.float32(0.02)).all() and (vy == 0).all()
    assert solid[0].all() and solid[-1].all() and solid[:, 0].all() and solid[:, -1].all()
    assert solid[128, 128] and solid[128, 152] and not solid[128, 153] and not solid[128, 128 + 25]
    inner = solid[1:-1, 1:-1].sum()
    assert abs(inner - np.pi * 25 ** 2) < 40           # lattice disc area
    _, _, _, twin = scenarios.main_rs(256, 256, np.float32, walls=False)
    assert twin.sum() == inner and not twin[0].any()


def test_smooth_periodic_rows_match_whole_field():
    full = scenarios.smooth_periodic(64, 48, np.float64)
    part = scenarios.smooth_periodic_rows(64, 48, 10, 30, np.float64)
    for a, b in zip(full[:3], part[:3]):
        np.testing.assert_array_equal(a[10:30], b)
    assert abs(full[0].mean() - 1.0) < 1e-12 and np.abs(full[1]).max() <= 0.05


def test_channel_scenario():
    rho, vx, vy, solid = scenarios.channel_cylinder(512, 128, np.float32, radius=8.0, cx=64.0)
    assert solid[0].all() and solid[-1].all() and solid[64, 64] and not solid[64, 80]
    assert (vy == np.float32(0.05)).all() and (vx == 0).all()


def test_c_example_builds_and_fails_loudly_without_gpu(tmp_path):
    build.build()
    exe = tmp_path / "c_api_demo"
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_api_demo.c"), "-o", str(exe), "-L" + os.path.join(ROOT, "chemsim_b200"),
           "-lchemsim_lbm", "-Wl,-rpath," + os.path.join(ROOT, "chemsim_b200"), "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    import torch
    run = subprocess.run([str(exe), "64", "64", "2"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert run.returncode == 0 and run.stdout.count("frame") == 2
    else:
        assert run.returncode == 1 and "status 3" in run.stderr      # CHEMSIM_LBM_ERR_CUDA, no fallback
