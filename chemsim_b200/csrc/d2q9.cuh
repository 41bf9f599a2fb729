// d2q9.cuh — D2Q9 tables and the per-cell arithmetic shared by every kernel.
//
// The arithmetic restates /root/reference/src/lbm.rs in the reference's exact
// operation order.  Every floating-point operation goes through
// an explicit round-to-nearest intrinsic (__fadd_rn, __dmul_rn, ...), which the
// compiler never contracts into an FMA, so results are bit-identical to an IEEE
// evaluation of the reference's expression tree (ArrayFire's CPU backend).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace chemsim {

constexpr int Q = 9;

// Lattice velocities c_i (src/lbm.rs:221-231).
__host__ __device__ constexpr int cx_of(int i) { return i == 1 || i == 5 || i == 8 ? 1 : (i == 3 || i == 6 || i == 7 ? -1 : 0); }
__host__ __device__ constexpr int cy_of(int i) { return i == 2 || i == 5 || i == 6 ? 1 : (i == 4 || i == 7 || i == 8 ? -1 : 0); }
// State::stream (src/lbm.rs:716-729): convolve2 with stencil_i^T moves population
// i by (dy, dx) = (-c_ix, +c_iy) in [y][x] memory terms.
__host__ __device__ constexpr int ey_of(int i) { return -cx_of(i); }
__host__ __device__ constexpr int ex_of(int i) { return cy_of(i); }
// D2Q9::swap_populations (src/lbm.rs:298-309).
__host__ __device__ constexpr int opp_of(int i) { return i == 0 ? 0 : (i <= 4 ? ((i + 1) % 4) + 1 : ((i - 3) % 4) + 5); }
static_assert(opp_of(1) == 3 && opp_of(2) == 4 && opp_of(3) == 1 && opp_of(4) == 2, "opp");
static_assert(opp_of(5) == 7 && opp_of(6) == 8 && opp_of(7) == 5 && opp_of(8) == 6, "opp");

// Host scalars, computed on the host in T exactly as the reference does
// (src/lbm.rs:54-56, :64-66, :84, :209-219, :357).
enum Collision : int { COL_NONE = 0, COL_BGK = 1, COL_TRT = 2, COL_REGULARIZED = 3, COL_KBC = 4 };

template <typename T>
struct Consts {
    T w[Q];    // 16/36, 4/36 x4, 1/36 x4
    T cs2;     // cs*cs, cs = dx/(sqrt(3)*dt)
    T k1;      // 1/cs2
    T k2;      // 1/(2*cs4)
    T k3;      // -1/(2*cs2)
    T factor;  // BGK: -dt/tau                                   src/lbm.rs:357
    // TRT (src/lbm.rs:428-439)
    T omega_p, omega_m;   // 1/tau_plus, 1/tau_minus
    T half;               // -dt*0.5
    // Regularized (src/lbm.rs:638-656): q_tensor_ab[i] * (w_i / (2*cs4))
    T axx[Q], axy[Q], ayx[Q], ayy[Q];
    // KBC (src/lbm.rs:478-571)
    T dx, dx2;            // delta_x, dx*dx
    T dx_4, four_dx;      // dx*4.0, 4.0*dx
    T two_dx2;            // 2.0*dx*dx
    T neg_dx;             // -dx
    T neg_beta;           // -beta, beta = 1/(2*visc/(cs*cs) + 1)
    T two_neg_beta;       // 2.0 * -beta
    T gamma_scale;        // 2.0 - 1.0/beta
    T gamma_shift;        // -1.0/beta
};

// ---- rounded, never-contracted arithmetic -----------------------------------
__device__ __forceinline__ float  add(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ float  sub(float a, float b)   { return __fsub_rn(a, b); }
__device__ __forceinline__ float  mul(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ float  divi(float a, float b)  { return __fdiv_rn(a, b); }
__device__ __forceinline__ float  root(float a)           { return __fsqrt_rn(a); }
// correctly rounded reciprocal == IEEE 1/x (what af::div(1, x) yields), cheaper than a general division
__device__ __forceinline__ float  recip(float a)          { return __frcp_rn(a); }
__device__ __forceinline__ double recip(double a)         { return __drcp_rn(a); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double divi(double a, double b){ return __ddiv_rn(a, b); }
__device__ __forceinline__ double root(double a)          { return __dsqrt_rn(a); }

// Macroscopic moments of one cell, in the reference's order.
template <typename T>
struct Moments { T rho, mx, my, vx, vy; };

// Lattice::density (src/lbm.rs:117-121): ((0 + f0) + f1) + ... + f8
template <typename T>
__device__ __forceinline__ T density(const T (&g)[Q])
{
    T rho = T(0);
#pragma unroll
    for (int i = 0; i < Q; ++i) rho = add(rho, g[i]);
    return rho;
}

// Lattice::momentum_density (src/lbm.rs:123-131): md = md + f_i * c_i, every
// product formed (including *0 and *-1).
template <typename T>
__device__ __forceinline__ void momentum(const T (&g)[Q], T &mx, T &my)
{
    mx = T(0); my = T(0);
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        mx = add(mx, mul(g[i], T(cx_of(i))));
        my = add(my, mul(g[i], T(cy_of(i))));
    }
}

// Lattice::velocity (src/lbm.rs:133-138; Matrix::recip src/matrix.rs:133-136):
// r = 1/rho, v = r * m.
template <typename T>
__device__ __forceinline__ Moments<T> moments(const T (&g)[Q])
{
    Moments<T> m;
    m.rho = density(g);
    momentum(g, m.mx, m.my);
    const T r = recip(m.rho);                       // 1/rho, Matrix::recip
    m.vx = mul(r, m.mx);
    m.vy = mul(r, m.my);
    return m;
}

// compute_equilibrium for one direction (src/lbm.rs:58-68).
template <typename T>
__device__ __forceinline__ T equilibrium_i(int i, T rho, T vx, T vy, T v2, const Consts<T> &k)
{
    const T vc  = add(mul(vx, T(cx_of(i))), mul(vy, T(cy_of(i))));
    const T vc2 = mul(vc, vc);
    const T sum = add(add(add(T(1), mul(vc, k.k1)), mul(vc2, k.k2)), mul(v2, k.k3));
    return mul(mul(rho, k.w[i]), sum);
}

// State::collide with BGK (src/lbm.rs:731-739, :349-364): g <- g + (g - feq)*factor
template <typename T>
__device__ __forceinline__ void collide_bgk(T (&g)[Q], const Consts<T> &k)
{
    const Moments<T> m = moments(g);
    const T v2 = add(mul(m.vx, m.vx), mul(m.vy, m.vy));   // src/lbm.rs:53
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const T fe = equilibrium_i(i, m.rho, m.vx, m.vy, v2, k);
        g[i] = add(g[i], mul(sub(g[i], fe), k.factor));
    }
}

// ---- the other CollisionOperator impls of src/lbm.rs ---------------------------
// A tiny value wrapper whose operators are the rounded, never-contracted
// intrinsics above, so the expressions below read like the reference's.
template <typename T>
struct Num {
    T v;
    __device__ __forceinline__ Num(T x) : v(x) {}
    __device__ __forceinline__ Num operator+(Num o) const { return Num(add(v, o.v)); }
    __device__ __forceinline__ Num operator-(Num o) const { return Num(sub(v, o.v)); }
    __device__ __forceinline__ Num operator*(Num o) const { return Num(mul(v, o.v)); }
    __device__ __forceinline__ Num operator/(Num o) const { return Num(divi(v, o.v)); }
};

template <typename T>
__device__ __forceinline__ void equilibrium_all(const T (&g)[Q], const Consts<T> &k, Moments<T> &m, T (&fe)[Q])
{
    m = moments(g);
    const T v2 = add(mul(m.vx, m.vx), mul(m.vy, m.vy));
#pragma unroll
    for (int i = 0; i < Q; ++i) fe[i] = equilibrium_i(i, m.rho, m.vx, m.vy, v2, k);
}

// TRT::evaluate (src/lbm.rs:401-444).  swap_equilibrium (:311-322) overwrites slots
// 1..8 of the equilibrium with the opposite *population* — reproduced as written.
template <typename T>
__device__ __forceinline__ void collide_trt(T (&g)[Q], const Consts<T> &k)
{
    using N = Num<T>;
    Moments<T> m; T fe[Q];
    equilibrium_all(g, k, m, fe);
    T out[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const N fi(g[i]), fs(g[opp_of(i)]), ei(fe[i]), es(i == 0 ? fe[0] : g[opp_of(i)]);
        const N f_p = fi + fs, f_m = fi - fs, e_p = ei + es, e_m = ei - es;
        const N omega = ((f_p - e_p) * N(k.omega_p) + (f_m - e_m) * N(k.omega_m)) * N(k.half);
        out[i] = (fi + omega).v;
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) g[i] = out[i];
}

// Regularized::evaluate (src/lbm.rs:606-661); never calls the underlying operator.
template <typename T>
__device__ __forceinline__ void collide_regularized(T (&g)[Q], const Consts<T> &k)
{
    using N = Num<T>;
    Moments<T> m; T fe[Q];
    equilibrium_all(g, k, m, fe);
    N sxx(T(0)), sxy(T(0)), syx(T(0)), syy(T(0));
#pragma unroll
    for (int i = 0; i < Q; ++i) {                       // :625-632
        const N fneq = N(g[i]) - N(fe[i]);
        sxx = sxx + fneq * N(T(cx_of(i) * cx_of(i)));
        sxy = sxy + fneq * N(T(cx_of(i) * cy_of(i)));
        syx = syx + fneq * N(T(cy_of(i) * cx_of(i)));
        syy = syy + fneq * N(T(cy_of(i) * cy_of(i)));
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) {                       // :647-658
        N reg(fe[i]);
        reg = reg + sxx * N(k.axx[i]);
        reg = reg + sxy * N(k.axy[i]);
        reg = reg + syx * N(k.ayx[i]);
        reg = reg + syy * N(k.ayy[i]);
        g[i] = reg.v;
    }
}

// KBC::evaluate (src/lbm.rs:468-585), entropic stabiliser gamma*.
template <typename T>
__device__ __forceinline__ void collide_kbc(T (&g)[Q], const Consts<T> &k)
{
    using N = Num<T>;
    Moments<T> m; T fe[Q];
    equilibrium_all(g, k, m, fe);
    const N rho(m.rho), u(m.vx), v(m.vy), dx(k.dx);
    const N uv = u * v, u2 = u * u, v2 = v * v;                                        // :483-485
    N temp(T(0));
#pragma unroll
    for (int i = 0; i < Q; ++i) temp = temp + N(g[i]) * N(k.dx2);                       // :487-490
    const N pi_t = temp - uv, n_t = v2 - u2;                                           // :492-493
    const N uv8 = uv * N(T(8));
    const N s0  = ((uv8 * pi_t + n_t * n_t) * rho) * N(T(0.5));                        // :496-501
    const N s13 = (((((u * dx - n_t) + N(T(1))) * n_t) - (v * N(k.dx_4) + uv8) * pi_t) * rho) * N(T(0.25));        // :502-507
    const N s24 = (((((v * N(k.neg_dx) - n_t) + N(T(-1))) * n_t) - (u * N(k.dx_4) + uv8) * pi_t) * rho) * N(T(0.25)); // :508-513
    const N s58 = (((((uv8 + u * N(k.four_dx)) + N(k.two_dx2)) * pi_t) + (n_t - (v - u) * dx) * n_t) * rho) * N(T(0.125)); // :514-519
    const T s[Q] = {s0.v, s13.v, s24.v, s13.v, s24.v, s58.v, s58.v, s58.v, s58.v};    // :521-532
    T dh[Q];
    N num(T(0)), den(T(0));
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const N d = (N(g[i]) - N(fe[i])) - N(s[i]);                                    // :541
        dh[i] = d.v;
        num = num + (N(s[i]) * d) / N(fe[i]);                                          // :557
        den = den + (d * d) / N(fe[i]);                                                // :558
    }
    const N gamma = (((num / den) * N(k.gamma_scale)) + N(k.gamma_shift)) * N(T(-1)); // :560-564
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const N omega = N(s[i]) * N(k.two_neg_beta) + (N(dh[i]) * gamma) * N(k.neg_beta);   // :569-571
        g[i] = (N(g[i]) + omega).v;
    }
}

// State::collide (src/lbm.rs:731-739): dispatch on the operator; COL is a
// template parameter of the step kernels (the reference dispatches dynamically
// through Box<dyn CollisionOperator>, :674).
template <int COL, typename T>
__device__ __forceinline__ void collide(T (&g)[Q], const Consts<T> &k)
{
    if (COL == COL_BGK) collide_bgk(g, k);
    else if (COL == COL_TRT) collide_trt(g, k);
    else if (COL == COL_REGULARIZED) collide_regularized(g, k);
    else collide_kbc(g, k);
}

// State::bounce_back (src/lbm.rs:741-751): g_i <- solid ? g_opp(i) : g_i
template <typename T>
__device__ __forceinline__ void bounce_back(T (&g)[Q], bool solid)
{
    if (solid) {
        T t;
        t = g[1]; g[1] = g[3]; g[3] = t;
        t = g[2]; g[2] = g[4]; g[4] = t;
        t = g[5]; g[5] = g[7]; g[7] = t;
        t = g[6]; g[6] = g[8]; g[8] = t;
    }
}

}  // namespace chemsim
