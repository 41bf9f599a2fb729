//! The handful of `arrayfire` items `src/main.rs` itself touches, as host-side stand-ins — so that
//! the ONE line `use arrayfire as af;` (`src/main.rs:16`) becomes `use chemsim::af_compat as af;`
//! and the rest of the file compiles unchanged (rust/patches/main_rs.patch):
//!
//!   main.rs:77-89    `self.state.geometry.dims()`, `dims.get()`, `.host(&mut vec)`,
//!                    `af::Array::new(&vec[..], dims)`
//!   main.rs:308-311  `af::Dim4::new(&[w, h, 1, 1])`, `af::transpose(&af::Array::new(&vec[..], dim4), false)`
//!   main.rs:331-341  `af::init()`, `af::set_backend(af::Backend::CUDA)`, `af::get_active_backend()`,
//!                    `af::get_available_backends()`, `af::device_info()`
//!
//! Semantics are ArrayFire's: arrays are COLUMN-major (dim0 fastest), `host` copies out in that
//! order, `transpose` swaps dim0 and dim1.  Only `bool` arrays are needed (`lbm::Geometry`).
use std::sync::atomic::{AtomicU64, Ordering};

static NEXT_ID: AtomicU64 = AtomicU64::new(1);

#[derive(Clone, Copy, Debug, PartialEq)]
pub struct Dim4([u64; 4]);

impl Dim4 {
    pub fn new(dims: &[u64; 4]) -> Self { Dim4(*dims) }
    pub fn get(&self) -> &[u64; 4] { &self.0 }
    pub fn elements(&self) -> u64 { self.0.iter().product() }
}

impl std::ops::Index<usize> for Dim4 {
    type Output = u64;
    fn index(&self, i: usize) -> &u64 { &self.0[i] }
}

/// `af::Array<bool>`: column-major host storage.  Every value gets a fresh id, so that
/// `lbm::State` notices `state.geometry = af::Array::new(..)` (a plain field assignment in the
/// reference, `src/main.rs:89`) and uploads the new mask before the next step.
#[derive(Clone, Debug)]
pub struct Array<T: Copy> {
    data: Vec<T>,
    dims: Dim4,
    id: u64,
}

impl<T: Copy> Array<T> {
    pub fn new(slice: &[T], dims: Dim4) -> Self {
        assert_eq!(slice.len() as u64, dims.elements());
        Array { data: slice.to_vec(), dims, id: NEXT_ID.fetch_add(1, Ordering::Relaxed) }
    }
    pub fn dims(&self) -> Dim4 { self.dims }
    pub fn host(&self, out: &mut [T]) { out.copy_from_slice(&self.data); }
    pub fn id(&self) -> u64 { self.id }
    /// element (i, j) = (dim0, dim1)
    pub fn at(&self, i: usize, j: usize) -> T { self.data[j * self.dims[0] as usize + i] }
}

pub fn transpose<T: Copy>(a: &Array<T>, _conjugate: bool) -> Array<T> {
    let (d0, d1) = (a.dims[0] as usize, a.dims[1] as usize);
    let mut data = Vec::with_capacity(d0 * d1);
    for j in 0..d0 {                      // new dim1 = old dim0
        for i in 0..d1 {                  // new dim0 = old dim1, fastest
            data.push(a.data[j + i * d0]);    // old element (dim0 = j, dim1 = i)
        }
    }
    Array { data, dims: Dim4([a.dims[1], a.dims[0], a.dims[2], a.dims[3]]), id: NEXT_ID.fetch_add(1, Ordering::Relaxed) }
}

#[derive(Clone, Copy, Debug, PartialEq)]
pub enum Backend { DEFAULT, CPU, CUDA, OPENCL }

pub fn init() {}
pub fn set_backend(_backend: Backend) {}
pub fn get_active_backend() -> Backend { Backend::CUDA }
pub fn get_available_backends() -> Vec<Backend> { vec![Backend::CUDA] }
pub fn device_info() -> (String, String, String, String) {
    ("chemsim_lbm (hand-written sm_100a CUDA)".to_string(), "CUDA".to_string(), String::new(), String::new())
}
