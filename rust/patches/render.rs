//! Replacement for the reference's `src/render.rs` when `src/lbm.rs` is the device-backed shim
//! (rust/src/lbm.rs).  The reference's three functions evaluate the colour mapping of a macroscopic
//! field with ArrayFire (`mean_all`, `stdev_all`, `join_many`, `hsv2rgb`, `slice`; src/render.rs:23-178)
//! on `Matrix::get_array()`; the shim's `Matrix` is a host value without an ArrayFire array, and the
//! same arithmetic runs in `chemsim_lbm_render` on the device.  UNCOMPILED here.
use super::display::{Drawable, RGB, PixelPos};
use super::lbm;

#[derive(Clone, Copy, Debug, PartialEq)]
pub enum Field { Density = 0, Speed = 1, Velocity = 2, MomentumDensity = 3 }

/// render_scalar_field / render_vector_field followed by render_geometry (src/render.rs:7-178).
pub fn render_state<D: Drawable>(state: &lbm::State<lbm::D2Q9>, field: Field, buf: &mut D) {
    let (w, h) = buf.dimensions();
    assert_eq!((w as usize, h as usize), state.size());
    let rgba = state.render_rgba(field as i32, true);
    for x in 0 .. w {
        for y in 0 .. h {
            let i = 4 * ((y * w + x) as usize);
            buf.set_pixel(PixelPos(x, y), RGB(rgba[i], rgba[i + 1], rgba[i + 2]));
        }
    }
}
