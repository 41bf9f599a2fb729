//! Drop-in for `chemsim::lbm` + the parts of `chemsim::matrix` / `chemsim::render` its callers use
//! (taktoa/chemsim `src/lbm.rs`, `src/matrix.rs`, `src/render.rs`), backed by `libchemsim_lbm.so`.
//! UNCOMPILED in this repository (no Rust toolchain in the build image); the same surface is
//! compiled and tested as C++ (`chemsim_b200/cpp/lbm.hpp`) and Python (`chemsim_b200/lbm.py`).
//! What a maintainer changes in the reference to use it: INTEGRATION.md and `rust/patches/`.
pub mod af_compat;
pub mod ffi;
pub mod lbm;
pub mod matrix;
pub mod render;
