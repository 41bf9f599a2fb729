"""Device-side render.rs against the numpy restatement.  The colour
pipeline uses exp/atan2 whose last-bit behaviour differs between libm and CUDA, so
pixels are compared with a tolerance of one 8-bit level (the reference's own GPU and
CPU ArrayFire backends would differ by as much)."""
import numpy as np
import pytest

from chemsim_b200 import lbm, scenarios
from oracle import lbm_oracle as O
from oracle import render_numpy as R

pytestmark = pytest.mark.gpu


def _state(dtype, w=160, h=96, steps=30):
    rho, vx, vy, solid = scenarios.main_rs(w, h, dtype, walls=True, radius=10.0)
    rho = rho + (0.01 * np.sin(np.arange(w) / 7.0)[None, :]).astype(dtype)
    m = lambda a: lbm.Matrix.new(a.reshape(-1), (w, h), dtype=dtype)
    disc = lbm.Discretization(1.0, 1.0)
    pops = lbm.compute_equilibrium(m(rho), (m(vx), m(vy)), lbm.D2Q9.directions(), disc)
    state = lbm.State.initial(lbm.D2Q9.new(pops), solid, lbm.BGK(15.0), disc)
    state.step(steps)
    f_ref = O.step_ref(O.compute_equilibrium(rho, vx, vy), solid, steps, O.collision(O.BGK, tau=15.0))
    return state, f_ref, solid


def close(img, ref, frac=0.999):
    d = np.abs(img.astype(np.int16) - ref.astype(np.int16))
    assert d.max() <= 2, d.max()
    assert (d <= 1).mean() >= frac


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_render_modes_match_numpy_restatement(dtype):
    state, f, solid = _state(dtype)
    ux, uy = O.velocity(f)
    mx, my = O.momentum_density(f)
    refs = {
        state.RENDER_DENSITY: R.render_scalar_field(O.density(f)),
        state.RENDER_SPEED: R.render_scalar_field(O.speed(f)),
        state.RENDER_VELOCITY: R.render_vector_field(ux, uy),
        state.RENDER_MOMENTUM: R.render_vector_field(mx, my),
    }
    for mode, ref in refs.items():
        img = state.render(mode, overlay_geometry=False)
        assert img.shape == (96, 160, 4) and (img[..., 3] == 255).all()
        close(img[..., :3], ref)
        over = state.render(mode, overlay_geometry=True)
        close(over[..., :3], R.render_geometry(ref, solid))
        assert (over[solid.astype(bool)][:, :3] == (0, 0, 255)).all()


def test_render_scalar_is_shades_of_red():
    state, f, solid = _state(np.float32, steps=5)
    img = state.render(state.RENDER_DENSITY, overlay_geometry=False)
    assert (img[..., 1] == 0).all() and (img[..., 2] == 0).all() and img[..., 0].std() > 0
