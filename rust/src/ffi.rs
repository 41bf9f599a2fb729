//! `extern "C"` declarations of include/chemsim_lbm.h (ABI version 2).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

#[repr(C)]
pub struct chemsim_lbm_t {
    _private: [u8; 0],
}

pub const F32: c_int = 0;
pub const F64: c_int = 1;
pub const EDGE_ZEROFILL: c_int = 0;
pub const EDGE_PERIODIC: c_int = 1;

pub const OK: c_int = 0;
pub const ERR_INVALID_ARGUMENT: c_int = 1;
pub const ERR_INVALID_SLICE_SIZE: c_int = 2;
pub const ERR_CUDA: c_int = 3;
pub const ERR_NCCL: c_int = 4;
pub const ERR_NOT_READY: c_int = 5;
pub const ERR_UNSUPPORTED: c_int = 6;

pub const NCCL_ID_BYTES: usize = 128;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct chemsim_lbm_halo_msg {
    pub is_send: c_int,
    pub peer: c_int,
    pub q: c_int,
    pub row: c_int,
}

extern "C" {
    pub fn chemsim_lbm_abi_version() -> c_int;
    pub fn chemsim_lbm_last_error(h: *const chemsim_lbm_t) -> *const c_char;

    pub fn chemsim_lbm_create(width: c_int, height: c_int, dtype: c_int, edge: c_int, device: c_int,
                              out: *mut *mut chemsim_lbm_t) -> c_int;
    pub fn chemsim_lbm_create_slab(width: c_int, global_height: c_int, dtype: c_int, edge: c_int, device: c_int,
                                   rank: c_int, nranks: c_int, nccl_id: *const c_void,
                                   out: *mut *mut chemsim_lbm_t) -> c_int;
    pub fn chemsim_lbm_nccl_unique_id(out_id: *mut c_void) -> c_int;
    pub fn chemsim_lbm_enable_p2p_halo(h: *mut chemsim_lbm_t) -> c_int;
    pub fn chemsim_lbm_halo_mode(h: *const chemsim_lbm_t, mode: *mut c_int) -> c_int;
    pub fn chemsim_lbm_slab_rows(global_height: c_int, rank: c_int, nranks: c_int, row_offset: *mut c_int,
                                 rows: *mut c_int) -> c_int;
    pub fn chemsim_lbm_halo_plan(global_height: c_int, rank: c_int, nranks: c_int, edge: c_int,
                                 out: *mut chemsim_lbm_halo_msg, count: *mut c_int) -> c_int;
    pub fn chemsim_lbm_barrier(h: *mut chemsim_lbm_t) -> c_int;
    pub fn chemsim_lbm_set_p2p_timeout(h: *mut chemsim_lbm_t, seconds: c_double) -> c_int;
    pub fn chemsim_lbm_set_stream_convention(h: *mut chemsim_lbm_t, mirrored: c_int) -> c_int;
    pub fn chemsim_lbm_destroy(h: *mut chemsim_lbm_t) -> c_int;
    pub fn chemsim_lbm_shape(h: *const chemsim_lbm_t, width: *mut c_int, local_height: *mut c_int,
                             global_height: *mut c_int, row_offset: *mut c_int) -> c_int;

    pub fn chemsim_lbm_set_discretization(h: *mut chemsim_lbm_t, delta_x: c_double, delta_t: c_double) -> c_int;
    pub fn chemsim_lbm_set_bgk(h: *mut chemsim_lbm_t, tau: c_double) -> c_int;
    pub fn chemsim_lbm_set_trt(h: *mut chemsim_lbm_t, tau_plus: c_double, tau_minus: c_double) -> c_int;
    pub fn chemsim_lbm_set_regularized(h: *mut chemsim_lbm_t, underlying_viscosity: c_double) -> c_int;
    pub fn chemsim_lbm_set_kbc(h: *mut chemsim_lbm_t, ks_viscosity: c_double) -> c_int;
    pub fn chemsim_lbm_kinematic_shear_viscosity(h: *const chemsim_lbm_t, out: *mut c_double) -> c_int;
    pub fn chemsim_lbm_kinematic_bulk_viscosity(h: *const chemsim_lbm_t, out: *mut c_double) -> c_int;

    pub fn chemsim_lbm_init_equilibrium(h: *mut chemsim_lbm_t, rho: *const c_void, vx: *const c_void,
                                        vy: *const c_void, n: usize) -> c_int;
    pub fn chemsim_lbm_init_equilibrium_rows(h: *mut chemsim_lbm_t, row_begin: c_int, row_count: c_int,
                                             rho: *const c_void, vx: *const c_void, vy: *const c_void,
                                             n: usize) -> c_int;
    pub fn chemsim_lbm_set_population(h: *mut chemsim_lbm_t, q: c_int, src: *const c_void, n: usize) -> c_int;
    pub fn chemsim_lbm_set_geometry(h: *mut chemsim_lbm_t, solid: *const u8, n: usize) -> c_int;
    pub fn chemsim_lbm_set_geometry_rows(h: *mut chemsim_lbm_t, row_begin: c_int, row_count: c_int,
                                         solid: *const u8, n: usize) -> c_int;
    pub fn chemsim_lbm_set_geometry_async(h: *mut chemsim_lbm_t, solid: *const u8, n: usize) -> c_int;

    pub fn chemsim_lbm_fill_geometry(h: *mut chemsim_lbm_t, value: c_int) -> c_int;
    pub fn chemsim_lbm_paint_rect(h: *mut chemsim_lbm_t, x0: c_int, y0: c_int, width: c_int, height: c_int,
                                  value: c_int) -> c_int;

    pub fn chemsim_lbm_step(h: *mut chemsim_lbm_t, nsteps: c_int) -> c_int;
    pub fn chemsim_lbm_synchronize(h: *mut chemsim_lbm_t) -> c_int;
    pub fn chemsim_lbm_time(h: *const chemsim_lbm_t, out: *mut c_double) -> c_int;

    pub fn chemsim_lbm_get_density(h: *mut chemsim_lbm_t, dst: *mut c_void, n: usize) -> c_int;
    pub fn chemsim_lbm_get_density_async(h: *mut chemsim_lbm_t, dst: *mut c_void, n: usize) -> c_int;
    pub fn chemsim_lbm_get_async(h: *mut chemsim_lbm_t, field: c_int, q: c_int, dst0: *mut c_void, dst1: *mut c_void,
                                 n: usize) -> c_int;
    pub fn chemsim_lbm_get_pressure(h: *mut chemsim_lbm_t, dst: *mut c_void, n: usize) -> c_int;
    pub fn chemsim_lbm_get_speed(h: *mut chemsim_lbm_t, dst: *mut c_void, n: usize) -> c_int;
    pub fn chemsim_lbm_get_velocity(h: *mut chemsim_lbm_t, vx: *mut c_void, vy: *mut c_void, n: usize) -> c_int;
    pub fn chemsim_lbm_get_momentum_density(h: *mut chemsim_lbm_t, mx: *mut c_void, my: *mut c_void,
                                            n: usize) -> c_int;
    pub fn chemsim_lbm_get_population(h: *mut chemsim_lbm_t, q: c_int, dst: *mut c_void, n: usize) -> c_int;
    pub fn chemsim_lbm_get_equilibrium(h: *mut chemsim_lbm_t, q: c_int, dst: *mut c_void, n: usize) -> c_int;
    pub fn chemsim_lbm_get_non_equilibrium(h: *mut chemsim_lbm_t, q: c_int, dst: *mut c_void, n: usize) -> c_int;
    pub fn chemsim_lbm_get_geometry(h: *mut chemsim_lbm_t, dst: *mut u8, n: usize) -> c_int;
    pub fn chemsim_lbm_total_mass(h: *mut chemsim_lbm_t, out: *mut c_double) -> c_int;
    pub fn chemsim_lbm_total_mass_global(h: *mut chemsim_lbm_t, out: *mut c_double) -> c_int;
    pub fn chemsim_lbm_render(h: *mut chemsim_lbm_t, mode: c_int, overlay_geometry: c_int, rgba: *mut u8,
                              n_pixels: usize) -> c_int;
    pub fn chemsim_lbm_is_unstable(h: *mut chemsim_lbm_t, out: *mut c_int) -> c_int;

    pub fn chemsim_lbm_checkpoint_bytes(h: *const chemsim_lbm_t, out: *mut usize) -> c_int;
    pub fn chemsim_lbm_checkpoint(h: *mut chemsim_lbm_t, dst: *mut c_void, bytes: usize) -> c_int;
    pub fn chemsim_lbm_restore(h: *mut chemsim_lbm_t, src: *const c_void, bytes: usize) -> c_int;

    pub fn chemsim_lbm_cuda_stream(h: *const chemsim_lbm_t, stream: *mut *mut c_void) -> c_int;
    pub fn chemsim_lbm_kernel_launches(h: *const chemsim_lbm_t, out: *mut u64) -> c_int;
    pub fn chemsim_lbm_step_kernel_name(h: *const chemsim_lbm_t) -> *const c_char;
}
