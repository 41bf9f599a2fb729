"""The C++ host mirror (chemsim_b200/cpp/lbm.hpp) driven like the reference's main.rs
drives lbm.rs, compared frame by frame with the oracle."""
import re
import subprocess

import numpy as np
import pytest

from chemsim_b200 import build, scenarios
from oracle import lbm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("operator", ["bgk", "regkbc"])
def test_main_rs_shaped_cpp_driver_matches_oracle(operator):
    exe = build.build_harness()
    w = h = 96
    frames, paint_frame = 12, 5
    res = subprocess.run([exe, str(w), str(h), str(frames), str(paint_frame), operator], capture_output=True, text=True,
                         timeout=120)
    assert res.returncode == 0, res.stderr
    assert "error-check ok" in res.stdout
    lines = [l for l in res.stdout.splitlines() if l.startswith("frame")]
    assert len(lines) == frames
    rho, vx, vy, solid = scenarios.main_rs(w, h, np.float32)
    f = O.compute_equilibrium(rho, vx, vy)
    col = O.collision(O.BGK, tau=15.0) if operator == "bgk" else O.collision(O.REGULARIZED)
    py, px = h // 2, w // 4
    for i, line in enumerate(lines):
        if i == paint_frame:          # main.rs:82-88: geometry becomes ONLY the 9x9 block
            yy, xx = np.mgrid[0:h, 0:w]
            solid = ((np.abs(xx - px) < 5) & (np.abs(yy - py) < 5)).astype(np.uint8)
        f = O.step_ref(f, solid, 2, col, O.EDGE_ZEROFILL)     # speed_factor = 2
        m = re.match(r"frame (\d+) time (\S+) mass (\S+) rho (\S+) speed (\S+) unstable (\d)", line)
        assert int(m.group(1)) == i
        assert float(m.group(2)) == 2.0 * (i + 1)
        assert abs(float(m.group(3)) - O.total_mass(f)) <= 1e-12 * O.total_mass(f)
        assert np.float32(m.group(4)) == O.density(f)[py, px]
        assert np.float32(m.group(5)) == O.speed(f)[py, px]
        assert int(m.group(6)) == int(O.is_unstable(f))
