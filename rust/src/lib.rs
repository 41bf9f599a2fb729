//! Drop-in for `chemsim::lbm` + the parts of `chemsim::matrix` its callers use
//! (taktoa/chemsim `src/lbm.rs`, `src/matrix.rs`), backed by `libchemsim_lbm.so`.
//! UNCOMPILED in this repository (no Rust toolchain in the build image); the same
//! surface is compiled and tested as C++ (`chemsim_b200/cpp/lbm.hpp`) and Python
//! (`chemsim_b200/lbm.py`).
pub mod ffi;
pub mod lbm;
pub mod matrix;
