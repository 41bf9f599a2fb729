/*
 * oracle/lbm_oracle.h — CPU oracle for chemsim's D2Q9 collide+stream hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it, and only as the checker or the reported CPU baseline.  Nothing
 * under chemsim_b200/ links, imports or calls it.
 *
 * What it is: a plain-C restatement of /root/reference/src/lbm.rs (State::step
 * = stream -> bounce_back -> collide, the BGK/TRT/Regularized/KBC operators,
 * compute_equilibrium and the macroscopic readouts) in the reference's own
 * array-at-a-time structure and operation order.  Each function cites the
 * reference lines it follows.
 *
 * PARITY UNPINNED.  The arithmetic of the reference lives in an un-vendored
 * third-party dependency: Rust crate `arrayfire` 3.6.0 (Cargo.lock:60-61) over
 * C++ ArrayFire 3.6.1 @ 25bb360659b091bbca711b463c0ad5f0cf818e9c
 * (nix/arrayfire/default.nix:38-46).  The reference has no tests, golden
 * vectors or fixtures, and neither Rust nor ArrayFire exists in this image, so
 * the reference cannot be run here.  The oracle is therefore anchored on
 *   (1) the reference's call sites (cited per function),
 *   (2) ArrayFire's documented semantics for convolve2 / replace / div,
 *   (3) an independent array-at-a-time numpy/scipy restatement that performs
 *       the literal `convolve2d(f, stencil^T, 'same', fillvalue=0)` calls
 *       (oracle/lbm_numpy.py) — the two must agree bit for bit, and
 *   (4) the derived known-answer numbers in SURVEY.md §6.2.
 *
 * Build: `make -C oracle` (gcc -O2 -ffp-contract=off -fopenmp).
 */
#ifndef LBM_ORACLE_H
#define LBM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_EDGE_ZEROFILL = 0, ORACLE_EDGE_PERIODIC = 1 };
enum {
    ORACLE_COLLISION_BGK = 0,          /* src/lbm.rs:345-370 */
    ORACLE_COLLISION_TRT = 1,          /* src/lbm.rs:374-451 */
    ORACLE_COLLISION_REGULARIZED = 2,  /* src/lbm.rs:596-666 */
    ORACLE_COLLISION_KBC = 3           /* src/lbm.rs:455-590 */
};

typedef struct {
    int    kind;
    double tau;        /* BGK */
    double tau_plus;   /* TRT */
    double tau_minus;  /* TRT */
    double viscosity;  /* KBC */
} lbm_oracle_collision_t;

/* D2Q9 tables (src/lbm.rs:221-231, :298-309; stream shift per SURVEY.md §8 a-2) */
extern const int ORACLE_CX[9], ORACLE_CY[9], ORACLE_EX[9], ORACLE_EY[9], ORACLE_OPP[9];

#define ORACLE_DECL(T, S)                                                                          \
    void   lbm_oracle_constants_##S(T dx, T dt, T *out5);                                          \
    void   lbm_oracle_density_##S(const T *f, size_t n, T *rho);                                   \
    void   lbm_oracle_momentum_##S(const T *f, size_t n, T *mx, T *my);                            \
    void   lbm_oracle_velocity_##S(const T *f, size_t n, T *vx, T *vy);                            \
    void   lbm_oracle_speed_##S(const T *f, size_t n, T *speed);                                   \
    void   lbm_oracle_pressure_##S(const T *f, size_t n, T dx, T dt, T *p);                        \
    void   lbm_oracle_equilibrium_##S(const T *rho, const T *vx, const T *vy, size_t n, T dx,      \
                                      T dt, T *feq);                                               \
    void   lbm_oracle_lattice_equilibrium_##S(const T *f, size_t n, T dx, T dt, T *feq);           \
    int    lbm_oracle_is_unstable_##S(const T *f, size_t n, T dx, T dt);                           \
    double lbm_oracle_total_mass_##S(const T *f, int w, int h);                                    \
    void   lbm_oracle_step_ref_##S(T *f, const uint8_t *solid, int w, int h, int edge, T dx, T dt, \
                                   const lbm_oracle_collision_t *col, int nsteps);                 \
    void   lbm_oracle_step_fused_##S(const T *src, T *dst, const uint8_t *solid, int w, int h,     \
                                     int edge, int ghost, T dx, T dt, T tau);

ORACLE_DECL(float, f32)
ORACLE_DECL(double, f64)

int lbm_oracle_max_threads(void);
void lbm_oracle_set_threads(int n);   /* launchers such as torchrun export OMP_NUM_THREADS=1 */

#ifdef __cplusplus
}
#endif
#endif
