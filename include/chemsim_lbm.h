/*
 * chemsim_lbm.h — C ABI of the B200-native D2Q9 collide+stream path.
 *
 * This is the drop-in boundary for the hot path of taktoa/chemsim's
 * src/lbm.rs: every entry point below is what an FFI binding of lbm.rs would
 * bind, and each cites the reference item (file:line, relative to the
 * reference tree) it replaces.  The reference has no FFI of its own — its
 * boundary is the set of lbm.rs / matrix.rs items that main.rs, render.rs and
 * display.rs touch — so the Rust shim that keeps those names
 * on top of this ABI is given in INTEGRATION.md.
 *
 * Conventions
 *  - Plain pointers and sizes only.  Host fields are row-major, element (y,x)
 *    at [y*w + x], exactly what Matrix::new takes and Matrix::get_underlying
 *    returns (src/matrix.rs:24-30, :120-126).  `n` arguments are element
 *    counts and must equal width*local_height, else
 *    CHEMSIM_LBM_ERR_INVALID_SLICE_SIZE (= matrix::Error::InvalidSliceSize,
 *    src/matrix.rs:15-19, :26).
 *  - Element type of every `void*` field is the lattice dtype given at create
 *    time (float for F32 — the reference's `Scalar`, src/lbm.rs:13 — or double).
 *  - Every function returns 0 on success or a chemsim_lbm_status; the message
 *    is available from chemsim_lbm_last_error().  Nothing unwinds or aborts
 *    across the boundary (the reference panics/aborts, Cargo.toml:128; the Rust
 *    shim maps non-zero statuses back to those panics / Results).
 *  - One caller thread per handle (the reference is single-threaded,
 *    src/display.rs:121-147).  All work is queued on the handle's CUDA stream;
 *    functions that return host data synchronise that stream, the others are
 *    asynchronous and ordered with later calls.
 *  - There is no CPU fallback: without a CUDA device every compute entry point
 *    fails with CHEMSIM_LBM_ERR_CUDA.
 *
 * Environment switches (read once per process; none of them changes a result —
 * they select between bit-identical execution strategies for A/B measurements):
 *    CHEMSIM_LBM_STEP2=0          every step on the single-step kernels (default: two steps per pass)
 *    CHEMSIM_LBM_PREFETCH=<n>     two-step kernels: L2 prefetch distance in tiles (0 = off; default a quarter
 *                                 of a wave of resident blocks; n < 0: percent of a wave)
 *    CHEMSIM_LBM_PDL=0            plain stream order instead of programmatic dependent launch
 *    CHEMSIM_LBM_P2P_TIMEOUT_S=<s> how long a face block waits for a neighbour GPU's step flag
 */
#ifndef CHEMSIM_LBM_H
#define CHEMSIM_LBM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHEMSIM_LBM_ABI_VERSION 2
#define CHEMSIM_LBM_Q 9
#define CHEMSIM_LBM_NCCL_ID_BYTES 128

typedef struct chemsim_lbm chemsim_lbm_t; /* opaque; owns all device memory */

typedef enum {
    CHEMSIM_LBM_F32 = 0, /* reference: `pub type Scalar = f32`, src/lbm.rs:13 */
    CHEMSIM_LBM_F64 = 1  /* extension: same operation order evaluated in f64  */
} chemsim_lbm_dtype;

typedef enum {
    CHEMSIM_LBM_EDGE_ZEROFILL = 0, /* reference: af::convolve2 zero padding, src/lbm.rs:722-724 */
    CHEMSIM_LBM_EDGE_PERIODIC = 1  /* extension: wrap-around (BASELINE.json configs 2,4,5)      */
} chemsim_lbm_edge;

typedef enum {
    CHEMSIM_LBM_OK = 0,
    CHEMSIM_LBM_ERR_INVALID_ARGUMENT = 1,
    CHEMSIM_LBM_ERR_INVALID_SLICE_SIZE = 2, /* matrix::Error::InvalidSliceSize */
    CHEMSIM_LBM_ERR_CUDA = 3,
    CHEMSIM_LBM_ERR_NCCL = 4,
    CHEMSIM_LBM_ERR_NOT_READY = 5, /* step before populations / collision were set */
    CHEMSIM_LBM_ERR_UNSUPPORTED = 6
} chemsim_lbm_status;

typedef enum {                            /* CollisionOperator impls, src/lbm.rs:327-666 */
    CHEMSIM_LBM_COLLISION_NONE = 0,
    CHEMSIM_LBM_COLLISION_BGK = 1,        /* src/lbm.rs:345-370 */
    CHEMSIM_LBM_COLLISION_TRT = 2,        /* src/lbm.rs:374-451 */
    CHEMSIM_LBM_COLLISION_REGULARIZED = 3,/* src/lbm.rs:596-666 */
    CHEMSIM_LBM_COLLISION_KBC = 4         /* src/lbm.rs:455-590 */
} chemsim_lbm_collision;

/* ---- library -------------------------------------------------------------- */

int chemsim_lbm_abi_version(void);
/* Message of the last failure on `h` (or, with h == NULL, of the last failed
 * create on this thread).  Never NULL. */
const char *chemsim_lbm_last_error(const chemsim_lbm_t *h);

/* ---- construction / ownership --------------------------------------------- */

/* One lattice on one GPU.  Replaces D2Q9::new + State::initial
 * (src/lbm.rs:187-200, :679-692): the handle is the device-resident State.
 * `device` < 0 keeps the current CUDA device. */
int chemsim_lbm_create(int width, int height, int dtype, int edge, int device, chemsim_lbm_t **out);

/* One y-slab of a lattice sharded over `nranks` processes (one per GPU): rank r
 * owns global rows [r*H/nranks, (r+1)*H/nranks).  Neighbouring slabs exchange
 * one row of three populations per face per step with ncclSend/ncclRecv on a
 * side stream, overlapped with the interior update.  `nccl_id` is the 128-byte
 * id from chemsim_lbm_nccl_unique_id() on rank 0, distributed by the caller
 * (e.g. torch.distributed / MPI broadcast).  New functionality: the reference is
 * single-device. */
int chemsim_lbm_create_slab(int width, int global_height, int dtype, int edge, int device, int rank,
                            int nranks, const void *nccl_id, chemsim_lbm_t **out);
int chemsim_lbm_nccl_unique_id(void *out_id /* CHEMSIM_LBM_NCCL_ID_BYTES */);

/* Switch a sharded lattice from the NCCL exchange to the fused peer-memory halo: the
 * face-row kernel stores the populations that cross a slab face straight into the
 * neighbouring GPU's ghost row through a cudaIpc-mapped pointer (NVLink) and publishes a
 * step counter the neighbour waits on — compute and exchange are one kernel, NCCL is only
 * used for the first exchange after an upload.  COLLECTIVE: every rank of the lattice must
 * call it at the same point.  If any rank cannot map its neighbours (no peer access, ragged
 * width) all ranks stay in NCCL mode; chemsim_lbm_halo_mode reports the outcome. */
typedef enum { CHEMSIM_LBM_HALO_NCCL = 0, CHEMSIM_LBM_HALO_P2P = 1 } chemsim_lbm_halo;
int chemsim_lbm_enable_p2p_halo(chemsim_lbm_t *h);
int chemsim_lbm_halo_mode(const chemsim_lbm_t *h, int *mode);

/* Host-only helpers that expose the sharding logic (no GPU needed; the CPU tests
 * drive a gloo emulation of the exchange with them).
 * slab_rows: the rows rank `rank` owns.  halo_plan: the messages one rank issues per
 * exchange, in issue order; `out` must hold CHEMSIM_LBM_HALO_PLAN_MAX entries.  A slab
 * keeps TWO ghost rows per side, because one pass may advance the lattice by two steps:
 * per face it sends its outermost row (all nine populations) and, of the next row, the
 * three populations that move towards the face.  If some slab has a single row
 * (global_height / nranks < 2) the plan is the one-row form: three populations per face. */
typedef enum {
    CHEMSIM_LBM_ROW_FIRST = 0,        /* local row 0            (send) */
    CHEMSIM_LBM_ROW_LAST = 1,         /* local row H-1          (send) */
    CHEMSIM_LBM_ROW_GHOST_ABOVE = 2,  /* ghost row -1           (recv) */
    CHEMSIM_LBM_ROW_GHOST_BELOW = 3,  /* ghost row H            (recv) */
    CHEMSIM_LBM_ROW_SECOND = 4,       /* local row 1            (send) */
    CHEMSIM_LBM_ROW_SECOND_LAST = 5,  /* local row H-2          (send) */
    CHEMSIM_LBM_ROW_GHOST_ABOVE2 = 6, /* ghost row -2           (recv) */
    CHEMSIM_LBM_ROW_GHOST_BELOW2 = 7  /* ghost row H+1          (recv) */
} chemsim_lbm_halo_row;
typedef struct {
    int is_send; /* 1 = ncclSend, 0 = ncclRecv */
    int peer;    /* rank of the neighbour */
    int q;       /* population index 0..8 */
    int row;     /* chemsim_lbm_halo_row */
} chemsim_lbm_halo_msg;
#define CHEMSIM_LBM_HALO_PLAN_MAX 48
int chemsim_lbm_slab_rows(int global_height, int rank, int nranks, int *row_offset, int *rows);
int chemsim_lbm_halo_plan(int global_height, int rank, int nranks, int edge, chemsim_lbm_halo_msg *out,
                          int *count);

/* Device-side barrier over the ranks of a sharded lattice: a one-element ncclAllReduce queued on
 * the handle's stream (asynchronous for the host).  Work queued after it starts on every rank at
 * the same time to within a collective's latency — what a multi-GPU timing needs before its start
 * event.  No-op for an unsharded lattice.  COLLECTIVE. */
int chemsim_lbm_barrier(chemsim_lbm_t *h);

/* How long a face block of the peer-memory halo waits for a neighbour's step flag before it gives
 * up (default 30 s, or CHEMSIM_LBM_P2P_TIMEOUT_S).  After a time-out the lattice is poisoned: the
 * rank stops publishing its steps, and chemsim_lbm_step, _synchronize and every readout return
 * CHEMSIM_LBM_ERR_CUDA until new populations are uploaded on every rank. */
int chemsim_lbm_set_p2p_timeout(chemsim_lbm_t *h, double seconds);

int chemsim_lbm_destroy(chemsim_lbm_t *h); /* Drop for State */

/* State::size (src/lbm.rs:753-756) and the slab this handle owns. */
int chemsim_lbm_shape(const chemsim_lbm_t *h, int *width, int *local_height, int *global_height,
                      int *row_offset);

/* ---- parameters ----------------------------------------------------------- */

/* Discretization{delta_x, delta_t} (src/lbm.rs:75-86).  Default 1, 1.  The host
 * scalars cs^2, 1/cs^2, 1/(2cs^4), -1/(2cs^2) are derived from these in the
 * lattice dtype exactly as src/lbm.rs:54-56, :64-66, :84 compute them. */
int chemsim_lbm_set_discretization(chemsim_lbm_t *h, double delta_x, double delta_t);

/* Which way State::stream moves the populations.  The reference streams with
 * af::convolve2(f_i, stencil_i^T) (src/lbm.rs:722-724).  0 (default): a true, flipped convolution centred
 * at floor(3/2) — ArrayFire 3.6's documented behaviour — which moves population i by
 * (dy, dx) = (-c_ix, +c_iy) in [y][x] memory terms.  1: the other reading (the stencil
 * applied unflipped), which moves every population the opposite way.  The two differ by a point reflection
 * of the lattice, (y, x) <-> (H-1-y, W-1-x), so convention 1 runs the same kernels and reverses every field
 * at this boundary (uploads, readouts, geometry, render, paint_rect coordinates).  Parity with the real
 * reference is unpinned (no Rust/ArrayFire in the build image): a reference-generated golden vector
 * (rust/tools/dump_golden.rs, tests/test_reference_golden.py) decides, and this switch makes the outcome a
 * flag instead of a rewrite.  Select it before the first upload; unsharded lattices only. */
int chemsim_lbm_set_stream_convention(chemsim_lbm_t *h, int mirrored);

/* collision = Box::new(BGK { tau })  (src/lbm.rs:345-347; factor = -dt/tau, :357) */
int chemsim_lbm_set_bgk(chemsim_lbm_t *h, double tau);

/* collision = Box::new(TRT { tau_plus, tau_minus })  (src/lbm.rs:374-377; evaluate :401-444,
 * including the swap_equilibrium quirk of :311-322).  TRT::new(lambda, viscosity, &disc)
 * (:380-390) is host arithmetic the caller's shim performs. */
int chemsim_lbm_set_trt(chemsim_lbm_t *h, double tau_plus, double tau_minus);

/* collision = Box::new(Regularized::new(underlying))  (src/lbm.rs:596-666).  evaluate never
 * calls the underlying operator; only its viscosity is reported (:663-665). */
int chemsim_lbm_set_regularized(chemsim_lbm_t *h, double underlying_viscosity);

/* collision = Box::new(KBC::new(ks_viscosity))  (src/lbm.rs:455-590; the DEBUG residual
 * print of :575-582 is not reproduced). */
int chemsim_lbm_set_kbc(chemsim_lbm_t *h, double ks_viscosity);

/* CollisionOperator::kinematic_shear_viscosity / _bulk_viscosity
 * (src/lbm.rs:335-340, :366-369), computed in the lattice dtype. */
int chemsim_lbm_kinematic_shear_viscosity(const chemsim_lbm_t *h, double *out);
int chemsim_lbm_kinematic_bulk_viscosity(const chemsim_lbm_t *h, double *out);

/* ---- state upload --------------------------------------------------------- */

/* populations = compute_equilibrium(rho, (vx, vy), D2Q9::directions(), disc)
 * followed by D2Q9::new (src/lbm.rs:43-71, main.rs:257-267), evaluated on the GPU. */
int chemsim_lbm_init_equilibrium(chemsim_lbm_t *h, const void *rho, const void *vx, const void *vy,
                                 size_t n);

/* The same for rows [row_begin, row_begin+row_count) of this handle's slab only
 * (n = width*row_count): lets a caller initialise a lattice that is larger than
 * the host memory it wants to spend, chunk by chunk. */
int chemsim_lbm_init_equilibrium_rows(chemsim_lbm_t *h, int row_begin, int row_count, const void *rho,
                                      const void *vx, const void *vy, size_t n);

/* D2Q9::new(&[Population; 9]) one array at a time (src/lbm.rs:187-200); q in 0..9. */
int chemsim_lbm_set_population(chemsim_lbm_t *h, int q, const void *src, size_t n);

/* state.geometry = ... (src/lbm.rs:673; main.rs:269-312 and the live edit at
 * main.rs:77-89): one byte per cell, non-zero = solid.  Callable between steps. */
int chemsim_lbm_set_geometry(chemsim_lbm_t *h, const uint8_t *solid, size_t n);
/* Rewrites rows [row_begin, row_begin+row_count) of the geometry only (the mouse
 * handler of main.rs:77-89 changes a 9x9 block but re-uploads everything). */
int chemsim_lbm_set_geometry_rows(chemsim_lbm_t *h, int row_begin, int row_count, const uint8_t *solid,
                                  size_t n);
/* Asynchronous form: `solid` must be page-locked and stay valid until the next
 * chemsim_lbm_synchronize(); the copy runs on its own stream and later steps wait
 * for it on the device, the host does not block. */
int chemsim_lbm_set_geometry_async(chemsim_lbm_t *h, const uint8_t *solid, size_t n);

/* Live geometry edits on the device.  The reference's mouse handler
 * (src/main.rs:71-91) downloads the mask, rewrites EVERY cell on the host — solid iff
 * |row - floor(pos[1])| < 5 and |col - floor(pos[0])| < 5 — and uploads it again; here that is
 * chemsim_lbm_fill_geometry(h, 0) followed by chemsim_lbm_paint_rect(h, px-4, py-4, 9, 9, 1): two
 * small kernels on the handle's stream, no host transfer.  Both are asynchronous.
 * paint_rect: cells [x0, x0+width) x [y0, y0+height) := value != 0; y is a GLOBAL row, the
 * rectangle is clipped to the lattice and to this handle's slab (on a sharded lattice every rank
 * makes the same call and paints its part). */
int chemsim_lbm_fill_geometry(chemsim_lbm_t *h, int value);
int chemsim_lbm_paint_rect(chemsim_lbm_t *h, int x0, int y0, int width, int height, int value);

/* ---- the hot path --------------------------------------------------------- */

/* State::step x nsteps (src/lbm.rs:694-714): stream -> bounce_back -> collide
 * fused into one pass per step; time += dt per step.  Asynchronous. */
int chemsim_lbm_step(chemsim_lbm_t *h, int nsteps);

int chemsim_lbm_synchronize(chemsim_lbm_t *h);

/* state.time (src/lbm.rs:671, :713), accumulated in the lattice dtype. */
int chemsim_lbm_time(const chemsim_lbm_t *h, double *out);

/* ---- macroscopic readout (device -> host on demand) ----------------------- */

int chemsim_lbm_get_density(chemsim_lbm_t *h, void *dst, size_t n);          /* State::density  :779 -> :117 */
/* Asynchronous State::density: snapshots the field on the device now and copies it
 * to page-locked `dst` on a separate stream while later steps run; `dst` is valid
 * after chemsim_lbm_synchronize().  At most two snapshots are in flight. */
int chemsim_lbm_get_density_async(chemsim_lbm_t *h, void *dst, size_t n);
/* The same for every readout main.rs's four display modes use (src/main.rs:157-174) and the rest
 * of the surface: field selects the getter, q the population for the per-direction fields, dst1
 * is the second component of VELOCITY / MOMENTUM_DENSITY and NULL otherwise.  This is what a
 * recorder (src/display.rs:157-185 steps and renders every frame) calls so that the transfer of
 * frame n overlaps the steps of frame n+1. */
typedef enum {
    CHEMSIM_LBM_FIELD_DENSITY = 0,          /* State::density          src/lbm.rs:779 */
    CHEMSIM_LBM_FIELD_PRESSURE = 1,         /* State::pressure         :784 */
    CHEMSIM_LBM_FIELD_SPEED = 2,            /* State::speed            :800 */
    CHEMSIM_LBM_FIELD_VELOCITY = 3,         /* State::velocity         :795 (dst0 = x, dst1 = y) */
    CHEMSIM_LBM_FIELD_MOMENTUM_DENSITY = 4, /* State::momentum_density :790 (dst0 = x, dst1 = y) */
    CHEMSIM_LBM_FIELD_POPULATION = 5,       /* State::populations      :769, direction q */
    CHEMSIM_LBM_FIELD_EQUILIBRIUM = 6,      /* State::equilibrium      :805, direction q */
    CHEMSIM_LBM_FIELD_NON_EQUILIBRIUM = 7   /* State::non_equilibrium  :810, direction q */
} chemsim_lbm_field;
int chemsim_lbm_get_async(chemsim_lbm_t *h, int field, int q, void *dst0, void *dst1, size_t n);
int chemsim_lbm_get_pressure(chemsim_lbm_t *h, void *dst, size_t n);         /* State::pressure :784          */
int chemsim_lbm_get_speed(chemsim_lbm_t *h, void *dst, size_t n);            /* State::speed    :800 -> :151 */
int chemsim_lbm_get_velocity(chemsim_lbm_t *h, void *vx, void *vy, size_t n);         /* :795 -> :133 */
int chemsim_lbm_get_momentum_density(chemsim_lbm_t *h, void *mx, void *my, size_t n); /* :790 -> :123 */
int chemsim_lbm_get_population(chemsim_lbm_t *h, int q, void *dst, size_t n);         /* State::populations :769 */
int chemsim_lbm_get_equilibrium(chemsim_lbm_t *h, int q, void *dst, size_t n);        /* State::equilibrium :805 */
int chemsim_lbm_get_non_equilibrium(chemsim_lbm_t *h, int q, void *dst, size_t n);    /* State::non_equilibrium :810 */
int chemsim_lbm_get_geometry(chemsim_lbm_t *h, uint8_t *dst, size_t n);               /* geometry.host(), main.rs:81 */

/* Sum of all nine populations over this handle's cells, accumulated in f64
 * (Matrix::sum -> af::sum_all, src/matrix.rs:138-140).  For a sharded lattice
 * chemsim_lbm_total_mass_global all-reduces the slab sums over NCCL. */
int chemsim_lbm_total_mass(chemsim_lbm_t *h, double *out);
int chemsim_lbm_total_mass_global(chemsim_lbm_t *h, double *out);

/* render.rs evaluated on the device: one RGBA8 image instead of 1-3 float
 * planes per frame.  mode 0/1 = render_scalar_field(&state.density() / &state.speed())
 * (src/render.rs:23-89), mode 2/3 = render_vector_field(&state.velocity() /
 * &state.momentum_density()) (:91-178): z-score with af::mean_all / af::stdev_all (population
 * standard deviation), logistic, HSV -> RGB, `(256*c).round().min(255).max(0) as u8`; with
 * overlay_geometry != 0 solid cells become RGB(0,0,255) as render_geometry does (:7-21).
 * rgba: n_pixels = width*local_height pixels of 4 bytes, row-major y*w+x, alpha 255
 * (display.rs:41-43).  On a sharded lattice the call is collective (the mean and the standard
 * deviation are all-reduced over the slabs) and each rank receives the image of its own rows. */
int chemsim_lbm_render(chemsim_lbm_t *h, int mode, int overlay_geometry, uint8_t *rgba, size_t n_pixels);

/* State::is_unstable (src/lbm.rs:815-818): min(f_eq,0) < 0 on this handle's cells. */
int chemsim_lbm_is_unstable(chemsim_lbm_t *h, int *out);

/* ---- checkpoint / restore ---------------------------------- */

/* A State as bytes: header, the nine populations of this handle's slab as dense row-major planes
 * in the lattice dtype (what State::populations + get_underlying would give, src/lbm.rs:769,
 * src/matrix.rs:120-126), then the geometry (one byte per cell).  The reference has no
 * serialisation; its recorder (src/display.rs:157-185, dead src/record.rs) only keeps rendered
 * frames.  restore() checks dtype / shape / slab, uploads populations and geometry and sets
 * state.time and the step counter; the collision operator and the discretization are recorded in
 * the header for the caller but NOT applied (the host State owns them).  On a sharded lattice
 * every rank checkpoints / restores its own slab; the first step after a restore re-exchanges the
 * halo, so restore is collective in the same sense as an upload. */
#define CHEMSIM_LBM_CHECKPOINT_MAGIC "CSLBMCK1"
typedef struct {
    char magic[8];
    uint32_t header_bytes, dtype, width, local_height, global_height, row_offset, rank, nranks, edge, collision;
    uint32_t step_index;
    float time_f32;           /* state.time as accumulated in f32 (the reference's Scalar) */
    double time_f64;          /* ... and in f64 (the value an F64 lattice reports) */
    double delta_x, delta_t, tau, tau_plus, tau_minus, viscosity;
} chemsim_lbm_checkpoint_header;
int chemsim_lbm_checkpoint_bytes(const chemsim_lbm_t *h, size_t *out);
int chemsim_lbm_checkpoint(chemsim_lbm_t *h, void *dst, size_t bytes);
int chemsim_lbm_restore(chemsim_lbm_t *h, const void *src, size_t bytes);

/* ---- interop / introspection ---------------------------------------------- */

/* The cudaStream_t all of the handle's work is queued on (for CUDA-event timing
 * by the caller). */
int chemsim_lbm_cuda_stream(const chemsim_lbm_t *h, void **stream);

/* Number of kernels this handle has launched since creation. */
int chemsim_lbm_kernel_launches(const chemsim_lbm_t *h, uint64_t *out);

/* Name of the fused step kernel variant the next chemsim_lbm_step will launch. */
const char *chemsim_lbm_step_kernel_name(const chemsim_lbm_t *h);

#ifdef __cplusplus
}
#endif
#endif /* CHEMSIM_LBM_H */
