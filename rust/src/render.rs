//! `chemsim::render` on the device (reference: `src/render.rs`).  The reference's functions take
//! the macroscopic fields (`&Matrix`) and evaluate the colour mapping with ArrayFire
//! (`mean_all`, `stdev_all`, `join_many`, `hsv2rgb`, `slice`, :41-70, :113-160); here the whole of it —
//! field, z-score, logistic, HSV->RGB, geometry overlay — runs in `chemsim_lbm_render` and one
//! RGBA8 image crosses the bus.  The call sites in `LBMSim::render` (`src/main.rs:157-176`) change
//! from `render_scalar_field(&self.state.density(), buf)` etc. to `render_state(&self.state, mode, buf)`
//! (rust/patches/main_rs.patch).
use crate::lbm;

pub trait Drawable {                       // src/display.rs:27-31, as far as render.rs uses it
    fn dimensions(&self) -> (u32, u32);
    fn set_pixel(&mut self, pos: (u32, u32), rgb: (u8, u8, u8));
}

#[derive(Clone, Copy, Debug, PartialEq)]
pub enum Field { Density = 0, Speed = 1, Velocity = 2, MomentumDensity = 3 }   // main.rs DisplayMode, :36-41

/// render_scalar_field / render_vector_field + render_geometry, `src/render.rs:7-178`.
pub fn render_state<D: Drawable>(state: &lbm::State<lbm::D2Q9>, field: Field, buf: &mut D) {
    let (w, h) = buf.dimensions();
    assert_eq!((w as usize, h as usize), state.size());                          // render.rs:29
    let rgba = state.render_rgba(field as i32, true);
    for y in 0..h {
        for x in 0..w {
            let i = 4 * ((y * w + x) as usize);
            buf.set_pixel((x, y), (rgba[i], rgba[i + 1], rgba[i + 2]));
        }
    }
}
