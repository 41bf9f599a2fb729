//! Drop-in for `chemsim::lbm` (reference: `src/lbm.rs`): the same public names and signatures, a
//! device-resident lattice behind the C ABI.  UNCOMPILED in this repository (no Rust toolchain in the
//! image); `chemsim_b200/cpp/lbm.hpp` is its compiled twin.
//!
//! Kept exactly (so `src/main.rs` compiles against it after the one-line patch in rust/patches/):
//!  * `Scalar`, `Vector`, `Discretization`, `Direction`, `Population = Matrix`,
//!    `Populations = Vec<(Direction, Population)>`, `Geometry = af::Array<bool>` (the stand-in of
//!    `af_compat`), `compute_equilibrium(..) -> Populations`, `D2Q9::new(&[Population])`,
//!    `D2Q9::directions()`, `BGK { tau }`, `TRT::new`, `KBC::new`, `Regularized::new`,
//!    `State::initial(Box<L>, Geometry, Box<CollisionOperator<L>>, Discretization)`, the public fields
//!    `state.time` / `.lattice` / `.geometry` / `.collision` / `.discretization`, `State::step` and every
//!    readout (`density`, `pressure`, `momentum_density`, `velocity`, `speed`, `populations`,
//!    `equilibrium`, `non_equilibrium`, `is_unstable`, `size`, `delta_x`, `delta_t`).
//! What differs, and why:
//!  * `CollisionOperator::evaluate(&L, &Discretization) -> Populations` (`src/lbm.rs:328-333`) is replaced
//!    by `apply` (select the operator on the device): collision is fused into the step kernel and is not
//!    callable on its own.  Likewise `State::{stream, collide, bounce_back}` (:716-751) do not exist.
//!  * `state.geometry` is a host value; `State::step` uploads it when it has been reassigned (main.rs:89).
//!  * `Matrix` is a host `Vec<f32>` (no `get_array` / `unsafe_new`): `render.rs` is replaced by
//!    `render::render_state` (device-side), see rust/patches/main_rs.patch.
//!  * `state.lattice` keeps the populations it was BUILT from; the live ones are `state.populations()`.
//!  * Errors: the reference panics on everything but `Matrix::new`; so does this shim (`check`).
use std::ffi::CStr;
use std::os::raw::c_int;

use crate::ffi;
pub use crate::matrix::Matrix;

pub type Scalar = f32; // src/lbm.rs:13

#[derive(PartialEq, PartialOrd, Debug, Clone, Copy)]
pub struct Vector(pub Scalar, pub Scalar); // src/lbm.rs:17-39

impl Vector {
    #[inline(always)]
    pub fn to_pair(&self) -> (Scalar, Scalar) { (self.0, self.1) }
}

#[derive(PartialEq, PartialOrd, Debug, Clone, Copy)]
pub struct Discretization { // src/lbm.rs:75-86
    pub delta_x: Scalar,
    pub delta_t: Scalar,
}

impl Discretization {
    #[inline(always)]
    pub fn isothermal_speed_of_sound(&self) -> Scalar {
        self.delta_x / (Scalar::sqrt(3.0) * self.delta_t)
    }
}

#[derive(Clone, Debug)]
pub struct Direction { // src/lbm.rs:90-95
    pub w_scalar: Scalar,
    pub c_vector: Vector,
    pub stencil: [i8; 9],
}

pub type Geometry = crate::af_compat::Array<bool>; // src/lbm.rs:99 (`af::Array<bool>`)
pub type Population = Matrix;                      // src/lbm.rs:103
pub type Populations = Vec<(Direction, Population)>; // src/lbm.rs:107

/// `lbm::compute_equilibrium` (`src/lbm.rs:43-71`), evaluated on the host in the reference's exact
/// operation order (f32, no contraction): nine arrays, as in the reference.  It runs once, at set-up;
/// `D2Q9::new` + `State::initial` upload them with `chemsim_lbm_set_population`.
pub fn compute_equilibrium(
    density: Matrix,
    velocity: (Matrix, Matrix),
    directions: &[Direction],
    discretization: Discretization,
) -> Populations {
    let size = density.get_shape();
    let (vx, vy) = velocity;
    assert_eq!(size, vx.get_shape()); // src/lbm.rs:51
    assert_eq!(size, vy.get_shape()); // src/lbm.rs:52
    let v2 = vx.hadamard(&vx) + vy.hadamard(&vy);
    let cs = discretization.isothermal_speed_of_sound();
    let cs2 = cs * cs;
    let cs4 = cs2 * cs2;
    let mut result = Vec::with_capacity(directions.len());
    for dir in directions {
        let (cx, cy) = dir.c_vector.to_pair();
        let vc = vx.scale(cx) + vy.scale(cy);
        let vc2 = vc.hadamard(&vc);
        let sum: Matrix = Matrix::new_filled(1.0, size)
            + vc.scale(1.0 / cs2)
            + vc2.scale(1.0 / (2.0 * cs4))
            + v2.scale(-1.0 / (2.0 * cs2));
        let pop = density.scale(dir.w_scalar).hadamard(&sum);
        result.push((dir.clone(), pop));
    }
    result
}

pub trait Lattice { // src/lbm.rs:111-176 (the arithmetic methods moved onto State: they need the device)
    fn size(&self) -> (usize, usize);
    fn populations(&self) -> &Populations;
}

#[derive(Clone)]
pub struct D2Q9 { // src/lbm.rs:180-184
    size: (usize, usize),
    populations: Populations,
}

impl D2Q9 {
    /// `D2Q9::new(&[Population])` (`src/lbm.rs:187-200`).
    pub fn new(populations: &[Population]) -> Self {
        assert!(populations.len() == 9);
        let size = populations[0].get_shape();
        for pop in populations { assert_eq!(size, pop.get_shape()); }
        let dirs = Self::directions();
        D2Q9 { size, populations: dirs.iter().cloned().zip(populations.iter().cloned()).collect() }
    }

    /// `D2Q9::directions()` (`src/lbm.rs:202-282`).
    pub fn directions() -> [Direction; 9] {
        const NUM: [Scalar; 9] = [16.0, 4.0, 4.0, 4.0, 4.0, 1.0, 1.0, 1.0, 1.0];
        const C: [(Scalar, Scalar); 9] = [
            (0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0),
            (1.0, 1.0), (-1.0, 1.0), (-1.0, -1.0), (1.0, -1.0),
        ];
        const ONE_AT: [usize; 9] = [4, 3, 7, 5, 1, 6, 8, 2, 0]; // position of the 1 in each 3x3 stencil, :233-269
        let make = |i: usize| {
            let mut stencil = [0i8; 9];
            stencil[ONE_AT[i]] = 1;
            Direction { w_scalar: NUM[i] / 36.0, c_vector: Vector(C[i].0, C[i].1), stencil }
        };
        [make(0), make(1), make(2), make(3), make(4), make(5), make(6), make(7), make(8)]
    }
}

impl Lattice for D2Q9 {
    fn size(&self) -> (usize, usize) { self.size }
    fn populations(&self) -> &Populations { &self.populations }
}

// ---- collision operators (src/lbm.rs:327-666) ----------------------------------------------------

pub trait CollisionOperator<L> {
    /// Select this operator on a device lattice (replaces `evaluate`, which ran on ArrayFire arrays).
    fn apply(&self, handle: *mut ffi::chemsim_lbm_t, disc: &Discretization) -> c_int;
    fn kinematic_shear_viscosity(&self, disc: &Discretization) -> Scalar;
    #[inline(always)]
    fn kinematic_bulk_viscosity(&self, disc: &Discretization) -> Scalar {
        2.0 * self.kinematic_shear_viscosity(disc) / 3.0
    }
}

pub struct BGK { pub tau: Scalar } // src/lbm.rs:345-347

impl<L: Lattice> CollisionOperator<L> for BGK {
    fn apply(&self, h: *mut ffi::chemsim_lbm_t, _d: &Discretization) -> c_int { unsafe { ffi::chemsim_lbm_set_bgk(h, self.tau as f64) } }
    fn kinematic_shear_viscosity(&self, disc: &Discretization) -> Scalar { // :366-369
        let (dx, dt) = (disc.delta_x, disc.delta_t);
        (dx * dx / (3.0 * dt * dt)) * (self.tau - dt / 2.0)
    }
}

pub struct TRT { pub tau_minus: Scalar, pub tau_plus: Scalar } // src/lbm.rs:374-377

impl TRT {
    pub fn new(lambda: Scalar, ks_viscosity: Scalar, disc: &Discretization) -> Self { // :380-390
        let dt = disc.delta_t;
        let cs = disc.isothermal_speed_of_sound();
        let tau_plus = dt * ((ks_viscosity / (cs * cs)) + 0.5);
        let tau_minus = dt * ((lambda / ((tau_plus / dt) - 0.5)) + 0.5);
        TRT { tau_minus, tau_plus }
    }
    pub fn lambda(&self, disc: &Discretization) -> Scalar { // :392-398
        let dt = disc.delta_t;
        let mut result = 1.0;
        result *= (self.tau_plus / dt) - 0.5;
        result *= (self.tau_minus / dt) - 0.5;
        result
    }
}

impl<L: Lattice> CollisionOperator<L> for TRT {
    fn apply(&self, h: *mut ffi::chemsim_lbm_t, _d: &Discretization) -> c_int {
        unsafe { ffi::chemsim_lbm_set_trt(h, self.tau_plus as f64, self.tau_minus as f64) }
    }
    fn kinematic_shear_viscosity(&self, disc: &Discretization) -> Scalar { // :446-450
        let cs = disc.isothermal_speed_of_sound();
        cs * cs * (self.tau_plus / disc.delta_t - 0.5)
    }
}

pub struct KBC { ks_viscosity: Scalar } // src/lbm.rs:458-466

impl KBC {
    pub fn new(ks_viscosity: Scalar) -> Self { KBC { ks_viscosity } }
}

impl CollisionOperator<D2Q9> for KBC {
    fn apply(&self, h: *mut ffi::chemsim_lbm_t, _d: &Discretization) -> c_int { unsafe { ffi::chemsim_lbm_set_kbc(h, self.ks_viscosity as f64) } }
    fn kinematic_shear_viscosity(&self, _disc: &Discretization) -> Scalar { self.ks_viscosity } // :587-589
}

pub struct Regularized<C> { underlying: C } // src/lbm.rs:596-604

impl<C> Regularized<C> {
    pub fn new(underlying: C) -> Self { Regularized { underlying } }
}

impl<L, C> CollisionOperator<L> for Regularized<C>
where L: Lattice, C: CollisionOperator<L> {
    fn apply(&self, h: *mut ffi::chemsim_lbm_t, disc: &Discretization) -> c_int {
        // only the underlying operator's viscosity is ever used (src/lbm.rs:663-665), with the State's discretization
        unsafe { ffi::chemsim_lbm_set_regularized(h, self.underlying.kinematic_shear_viscosity(disc) as f64) }
    }
    fn kinematic_shear_viscosity(&self, disc: &Discretization) -> Scalar { // :663-665
        self.underlying.kinematic_shear_viscosity(disc)
    }
}

// ---- State (src/lbm.rs:670-819) ------------------------------------------------------------------

fn check(status: c_int, h: *const ffi::chemsim_lbm_t) {
    if status != ffi::OK {
        let msg = unsafe { CStr::from_ptr(ffi::chemsim_lbm_last_error(h)) }.to_string_lossy().into_owned();
        panic!("chemsim_lbm status {}: {}", status, msg);
    }
}

pub struct State<L> {
    pub time: Scalar,
    pub lattice: Box<L>,
    /// Host value, as in the reference (`pub geometry: Geometry`, `src/lbm.rs:673`): read it with
    /// `.dims()` / `.host()`, replace it by assignment (`src/main.rs:77-89`); `step` uploads a new one.
    pub geometry: Geometry,
    pub collision: Box<dyn CollisionOperator<L>>,
    pub discretization: Discretization,
    handle: *mut ffi::chemsim_lbm_t,
    uploaded_geometry: u64,
}

impl State<D2Q9> {
    /// `State::initial(Box<L>, Geometry, Box<CollisionOperator<L>>, Discretization)`, `src/lbm.rs:679-692`.
    pub fn initial(
        lattice: Box<D2Q9>,
        geometry: Geometry,
        collision: Box<dyn CollisionOperator<D2Q9>>,
        discretization: Discretization,
    ) -> Self {
        Self::initial_with_edge(lattice, geometry, collision, discretization, ffi::EDGE_ZEROFILL)
    }

    /// Same, with the one extension: periodic edges (`ffi::EDGE_PERIODIC`).
    pub fn initial_with_edge(
        lattice: Box<D2Q9>,
        geometry: Geometry,
        collision: Box<dyn CollisionOperator<D2Q9>>,
        discretization: Discretization,
        edge: c_int,
    ) -> Self {
        let (w, h) = lattice.size();
        let n = w * h;
        let mut handle: *mut ffi::chemsim_lbm_t = std::ptr::null_mut();
        check(unsafe { ffi::chemsim_lbm_create(w as c_int, h as c_int, ffi::F32, edge, -1, &mut handle) }, std::ptr::null());
        check(unsafe {
            ffi::chemsim_lbm_set_discretization(handle, discretization.delta_x as f64, discretization.delta_t as f64)
        }, handle);
        check(collision.apply(handle, &discretization), handle);
        for (q, (_, pop)) in lattice.populations().iter().enumerate() {
            check(unsafe {
                ffi::chemsim_lbm_set_population(handle, q as c_int, pop.as_slice().as_ptr() as *const _, n)
            }, handle);
        }
        let mut state = State { time: 0.0, lattice, geometry, collision, discretization, handle, uploaded_geometry: 0 };
        state.upload_geometry();
        state
    }

    /// The geometry array holds element (dim0 = y, dim1 = x) (`src/main.rs:308-311`: built with dims
    /// [w, h] and transposed; the mouse handler builds it with the transposed dims directly, :77-89).
    fn upload_geometry(&mut self) {
        let (w, h) = self.size();
        let dims = self.geometry.dims();
        assert_eq!((dims[0] as usize, dims[1] as usize), (h, w));
        let mut bytes = vec![0u8; w * h];
        for y in 0..h { for x in 0..w { bytes[y * w + x] = self.geometry.at(y, x) as u8; } }
        check(unsafe { ffi::chemsim_lbm_set_geometry(self.handle, bytes.as_ptr(), bytes.len()) }, self.handle);
        self.uploaded_geometry = self.geometry.id();
    }

    /// The mouse handler of `src/main.rs:71-91` without its host round trip: the geometry becomes the
    /// 9x9 block around the cursor, painted by two small kernels (`chemsim_lbm_fill_geometry` +
    /// `chemsim_lbm_paint_rect`).  Optional: the unchanged handler (download, rewrite, assign) works too.
    pub fn paint_brush(&mut self, pos: [f64; 2]) {
        let (row, col) = (f64::floor(pos[1]) as i64, f64::floor(pos[0]) as i64);
        let (w, h) = self.size();
        if row < 0 || col < 0 || row as usize >= h || col as usize >= w { return; }
        check(unsafe { ffi::chemsim_lbm_fill_geometry(self.handle, 0) }, self.handle);
        check(unsafe { ffi::chemsim_lbm_paint_rect(self.handle, (col - 4) as c_int, (row - 4) as c_int, 9, 9, 1) }, self.handle);
        let mut bits = vec![false; w * h];                     // keep the host value in step (column-major [h, w])
        for y in 0..h { for x in 0..w {
            bits[x * h + y] = (y as i64 - row).abs() < 5 && (x as i64 - col).abs() < 5;
        } }
        self.geometry = Geometry::new(&bits, crate::af_compat::Dim4::new(&[h as u64, w as u64, 1, 1]));
        self.uploaded_geometry = self.geometry.id();
    }

    /// `State::step` (`src/lbm.rs:694-714`): stream -> bounce_back -> collide, one fused kernel.
    pub fn step(&mut self) {
        if self.geometry.id() != self.uploaded_geometry { self.upload_geometry(); }   // `state.geometry = ...`
        check(unsafe { ffi::chemsim_lbm_step(self.handle, 1) }, self.handle);
        self.time += self.discretization.delta_t;
    }

    #[inline(always)] pub fn size(&self) -> (usize, usize) { self.lattice.size() }
    #[inline(always)] pub fn delta_x(&self) -> Scalar { self.discretization.delta_x }
    #[inline(always)] pub fn delta_t(&self) -> Scalar { self.discretization.delta_t }
    #[inline(always)] pub fn isothermal_speed_of_sound(&self) -> Scalar { self.discretization.isothermal_speed_of_sound() }

    fn read1(&self, f: unsafe extern "C" fn(*mut ffi::chemsim_lbm_t, *mut std::os::raw::c_void, usize) -> c_int) -> Matrix {
        let mut m = Matrix::new_filled(0.0, self.size());
        let n = m.as_slice().len();
        check(unsafe { f(self.handle, m.as_mut_slice().as_mut_ptr() as *mut _, n) }, self.handle);
        m
    }

    fn read2(&self, f: unsafe extern "C" fn(*mut ffi::chemsim_lbm_t, *mut std::os::raw::c_void,
                                          *mut std::os::raw::c_void, usize) -> c_int) -> (Matrix, Matrix) {
        let mut a = Matrix::new_filled(0.0, self.size());
        let mut b = Matrix::new_filled(0.0, self.size());
        let n = a.as_slice().len();
        check(unsafe {
            f(self.handle, a.as_mut_slice().as_mut_ptr() as *mut _, b.as_mut_slice().as_mut_ptr() as *mut _, n)
        }, self.handle);
        (a, b)
    }

    pub fn density(&self) -> Matrix { self.read1(ffi::chemsim_lbm_get_density) }             // :779
    pub fn pressure(&self) -> Matrix { self.read1(ffi::chemsim_lbm_get_pressure) }           // :784
    pub fn speed(&self) -> Matrix { self.read1(ffi::chemsim_lbm_get_speed) }                 // :800
    pub fn velocity(&self) -> (Matrix, Matrix) { self.read2(ffi::chemsim_lbm_get_velocity) } // :795
    pub fn momentum_density(&self) -> (Matrix, Matrix) { self.read2(ffi::chemsim_lbm_get_momentum_density) } // :790

    fn read_q(&self, f: unsafe extern "C" fn(*mut ffi::chemsim_lbm_t, c_int, *mut std::os::raw::c_void, usize) -> c_int)
        -> Populations {
        let dirs = D2Q9::directions();
        (0..9).map(|q| {
            let mut m = Matrix::new_filled(0.0, self.size());
            let n = m.as_slice().len();
            check(unsafe { f(self.handle, q as c_int, m.as_mut_slice().as_mut_ptr() as *mut _, n) }, self.handle);
            (dirs[q].clone(), m)
        }).collect()
    }

    pub fn populations(&self) -> Populations { self.read_q(ffi::chemsim_lbm_get_population) }          // :769 (by value: device -> host)
    pub fn equilibrium(&self) -> Populations { self.read_q(ffi::chemsim_lbm_get_equilibrium) }         // :805
    pub fn non_equilibrium(&self) -> Populations { self.read_q(ffi::chemsim_lbm_get_non_equilibrium) } // :810

    pub fn is_unstable(&self) -> bool { // :815-818
        let mut flag: c_int = 0;
        check(unsafe { ffi::chemsim_lbm_is_unstable(self.handle, &mut flag) }, self.handle);
        flag != 0
    }

    /// Sum of all populations in f64 (`Matrix::sum`, `src/matrix.rs:138-140`).
    pub fn total_mass(&self) -> f64 {
        let mut m = 0.0f64;
        check(unsafe { ffi::chemsim_lbm_total_mass(self.handle, &mut m) }, self.handle);
        m
    }

    /// `render_scalar_field` / `render_vector_field` + `render_geometry` (`src/render.rs`) on the
    /// device: RGBA8, row-major `y*w + x`.  mode: 0 density, 1 speed, 2 velocity, 3 momentum density.
    pub fn render_rgba(&self, mode: i32, overlay_geometry: bool) -> Vec<u8> {
        let (w, h) = self.size();
        let mut rgba = vec![0u8; w * h * 4];
        check(unsafe {
            ffi::chemsim_lbm_render(self.handle, mode as c_int, overlay_geometry as c_int, rgba.as_mut_ptr(), w * h)
        }, self.handle);
        rgba
    }
}

impl<L> Drop for State<L> {
    fn drop(&mut self) {
        unsafe { ffi::chemsim_lbm_destroy(self.handle); }
    }
}
