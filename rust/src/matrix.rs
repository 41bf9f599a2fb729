//! Host-side `matrix::Matrix` (reference: `src/matrix.rs`).  In the reference a `Matrix` is an
//! ArrayFire device array; here the device-resident data lives inside `lbm::State`, and `Matrix`
//! is the host value callers build initial fields from and receive readouts in.  Same layout at
//! the boundary: shape `(w, h)`, element `(y, x)` at `slice[y*w + x]` (`src/matrix.rs:24-30`).
use std::ops::{Add, Sub};

#[derive(Clone, Debug)]
pub struct Matrix {
    data: Vec<f32>,
    w: usize,
    h: usize,
}

#[derive(Debug, Clone, Copy)]
pub enum Error {
    /// The slice given to `Matrix::new` had the wrong size.
    InvalidSliceSize,
}

pub type Result<T> = std::result::Result<T, Error>;

impl Matrix {
    /// `Matrix::new(slice, (w, h))`, `src/matrix.rs:24-30`.
    pub fn new(slice: &[f32], dims: (usize, usize)) -> Result<Self> {
        let (w, h) = dims;
        if slice.len() != w * h {
            return Err(Error::InvalidSliceSize);
        }
        Ok(Matrix { data: slice.to_vec(), w, h })
    }

    /// `Matrix::new_filled(value, (w, h))`, `src/matrix.rs:40-44` (with the intended `(w, h)`
    /// meaning; the reference builds the ArrayFire dims untransposed and only works for w == h).
    pub fn new_filled(value: f32, dims: (usize, usize)) -> Self {
        let (w, h) = dims;
        Matrix { data: vec![value; w * h], w, h }
    }

    pub fn get_width(&self) -> usize { self.w }
    pub fn get_height(&self) -> usize { self.h }
    pub fn get_shape(&self) -> (usize, usize) { (self.w, self.h) }

    /// Row-major host copy, index `y*w + x` (`src/matrix.rs:120-126`).
    pub fn get_underlying(&self) -> Vec<f32> { self.data.clone() }

    pub fn as_slice(&self) -> &[f32] { &self.data }
    pub fn as_mut_slice(&mut self) -> &mut [f32] { &mut self.data }

    fn map<F: Fn(f32) -> f32>(&self, f: F) -> Self {
        Matrix { data: self.data.iter().map(|&v| f(v)).collect(), w: self.w, h: self.h }
    }

    fn zip<F: Fn(f32, f32) -> f32>(&self, rhs: &Self, f: F) -> Self {
        assert_eq!(self.get_shape(), rhs.get_shape());
        Matrix { data: self.data.iter().zip(&rhs.data).map(|(&a, &b)| f(a, b)).collect(), w: self.w, h: self.h }
    }

    // The element-wise helpers render.rs applies to readout fields (src/render.rs:29-72, :113-164).
    pub fn scale(&self, scalar: f32) -> Self { self.map(|v| v * scalar) }
    pub fn shift(&self, shifter: f32) -> Self { self.map(|v| v + shifter) }
    pub fn clamp(&self, min: f32, max: f32) -> Self { self.map(|v| v.max(min).min(max)) }
    pub fn logistic(&self) -> Self { self.map(|v| 1.0 / (1.0 + (-v).exp())) }
    pub fn sqrt(&self) -> Self { self.map(f32::sqrt) }
    pub fn abs(&self) -> Self { self.map(f32::abs) }
    pub fn recip(&self) -> Self { self.map(|v| 1.0 / v) }
    pub fn hadamard(&self, rhs: &Self) -> Self { self.zip(rhs, |a, b| a * b) }
    pub fn divide(&self, rhs: &Self) -> Self { self.zip(rhs, |a, b| a / b) }
    pub fn sum(&self) -> f64 { self.data.iter().map(|&v| v as f64).sum() }
    pub fn maximum_real(&self) -> f64 { self.data.iter().cloned().fold(f32::NEG_INFINITY, f32::max) as f64 }
}

impl<'a, 'b> Add<&'a Matrix> for &'b Matrix {
    type Output = Matrix;
    fn add(self, rhs: &'a Matrix) -> Matrix { self.zip(rhs, |a, b| a + b) }
}
impl Add<Matrix> for Matrix {
    type Output = Matrix;
    fn add(self, rhs: Matrix) -> Matrix { &self + &rhs }
}
impl<'a, 'b> Sub<&'a Matrix> for &'b Matrix {
    type Output = Matrix;
    fn sub(self, rhs: &'a Matrix) -> Matrix { self.zip(rhs, |a, b| a - b) }
}
impl Sub<Matrix> for Matrix {
    type Output = Matrix;
    fn sub(self, rhs: Matrix) -> Matrix { &self - &rhs }
}
