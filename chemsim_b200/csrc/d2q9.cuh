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

// ---- two f32 cells per thread in one register pair (sm_100a packed FP32) ----------------------
// Blackwell issues add.rn.f32x2 (SASS FADD2) as ONE instruction for two IEEE round-to-nearest
// additions, lane by lane bit-identical to two scalar FADDs.  The f32 step kernels are bound by
// instruction issue, not by the FP32 pipe, so they run the collision of two cells at once on
// values of this type: every addition/subtraction is one packed instruction.  The
// MULTIPLICATIONS stay scalar on purpose: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2
// even with explicit .rn and --fmad=false (seen in SASS, CUDA 12.9), which would change the
// rounding; a scalar mul.rn.f32 feeding a packed add is never contracted.
// tests/test_abi.py::test_packed_f32_additions_are_never_contracted checks the built library's SASS
// (no FFMA2 / FMUL2 anywhere), tests/test_host_arith.py the F32x2 instantiation of the operators on the CPU.
// FADD2 issues at half rate (same FP32 pipe time as two FADDs): what it saves are issue slots —
// measured +1.7 % (BGK) to +3.2 % (Regularized) on the two-step kernels (profiles/r02_packed_prefetch_ab.md).
struct F32x2 {
    float lo, hi;
    F32x2() = default;
    __device__ __forceinline__ F32x2(float x) : lo(x), hi(x) {}
    __device__ __forceinline__ F32x2(float l, float h) : lo(l), hi(h) {}
};
#if defined(__CUDACC__)
#define CHEMSIM_F32X2_OP(NAME, PTX, HOSTOP)                                                           \
    __device__ __forceinline__ F32x2 NAME(F32x2 a, F32x2 b)                                           \
    {                                                                                                 \
        F32x2 r;                                                                                      \
        asm("{\n\t.reg .b64 pa, pb, pr;\n\tmov.b64 pa, {%2, %3};\n\tmov.b64 pb, {%4, %5};\n\t"       \
            PTX ".rn.f32x2 pr, pa, pb;\n\tmov.b64 {%0, %1}, pr;\n\t}"                                 \
            : "=f"(r.lo), "=f"(r.hi) : "f"(a.lo), "f"(a.hi), "f"(b.lo), "f"(b.hi));                   \
        return r;                                                                                     \
    }
#else   // tests/host_arith: the same two IEEE operations, lane by lane
#define CHEMSIM_F32X2_OP(NAME, PTX, HOSTOP)                                                           \
    inline F32x2 NAME(F32x2 a, F32x2 b) { return F32x2(HOSTOP(a.lo, b.lo), HOSTOP(a.hi, b.hi)); }
#endif
CHEMSIM_F32X2_OP(add, "add", __fadd_rn)
CHEMSIM_F32X2_OP(sub, "sub", __fsub_rn)
#undef CHEMSIM_F32X2_OP
// -DCHEMSIM_PACKED_MUL=1 (tools/variants.py pm; measured SLOWER, 125.5 vs 128.9 GLUPS: spills, constants forced
// into vector registers — not the default): packed multiplications after all, written as fma.rn.f32x2(a, b, -0) with the
// -0 read from constant memory, i.e. opaque to ptxas.  RN(a*b + (-0)) == RN(a*b) for every a, b
// (a zero product keeps its sign: (+0) + (-0) = +0, (-0) + (-0) = -0; inf*0 stays NaN), and an FMA
// cannot be contracted any further with the addition that consumes it.
#ifndef CHEMSIM_PACKED_MUL
#define CHEMSIM_PACKED_MUL 0
#endif
#if defined(__CUDACC__) && CHEMSIM_PACKED_MUL
static __constant__ float chemsim_neg_zero = -0.0f;
__device__ __forceinline__ F32x2 mul(F32x2 a, F32x2 b)
{
    F32x2 r;
    const float nz = chemsim_neg_zero;
    asm("{\n\t.reg .b64 pa, pb, pc, pr;\n\tmov.b64 pa, {%2, %3};\n\tmov.b64 pb, {%4, %5};\n\tmov.b64 pc, {%6, %6};\n\t"
        "fma.rn.f32x2 pr, pa, pb, pc;\n\tmov.b64 {%0, %1}, pr;\n\t}"
        : "=f"(r.lo), "=f"(r.hi) : "f"(a.lo), "f"(a.hi), "f"(b.lo), "f"(b.hi), "f"(nz));
    return r;
}
#else
__device__ __forceinline__ F32x2 mul(F32x2 a, F32x2 b)  { return F32x2(__fmul_rn(a.lo, b.lo), __fmul_rn(a.hi, b.hi)); }
#endif
__device__ __forceinline__ F32x2 divi(F32x2 a, F32x2 b) { return F32x2(__fdiv_rn(a.lo, b.lo), __fdiv_rn(a.hi, b.hi)); }
__device__ __forceinline__ F32x2 root(F32x2 a)          { return F32x2(__fsqrt_rn(a.lo), __fsqrt_rn(a.hi)); }
__device__ __forceinline__ F32x2 recip(F32x2 a)         { return F32x2(__frcp_rn(a.lo), __frcp_rn(a.hi)); }

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

// ---- the step kernels' form of the same arithmetic ----------------------------------
// The functions above evaluate the reference's expression tree literally (every product by
// 0 and +-1 formed); they serve the readouts and the initialisation, whose outputs ARE those
// intermediate values.  The collision operators below only have to produce the same nine
// post-collision populations, so they drop operations that provably cannot change them.
// Every identity used is exact in IEEE-754 round-to-nearest for FINITE inputs:
//   (1) x*(+1) == x;  x*(-1) == -x;  a + (-x) == a - x;  RN(-z) == -RN(z)
//   (2) x*0 == +-0, and acc + (+-0) == acc whenever acc is not -0.  An accumulator that starts
//       as (+0) + y is never -0 (x + y == -0 only if x == y == -0), so the zero terms of
//       momentum_density / the Pi_neq sums vanish bit for bit.
//   (3) 1 + (+-0) == 1: the sign of a zero velocity component never reaches the equilibrium
//       (vc enters it only through 1 + vc*k1 and vc*vc).
//   (4) equal operands give equal results: w_1..w_4 and w_5..w_8 are the same number
//       (4/36, 1/36, make_consts), vc_opp(i) == -vc_i, so a_i = vc_i*k1 and b_i = vc_i^2*k2 are
//       shared by a direction and its opposite: sum_i = ((1 + a) + b) + c, sum_opp = ((1 - a) + b) + c.
// Not preserved: which of inf/NaN appears once a population has overflowed (the literal tree
// turns inf*0 into NaN earlier), and, for Regularized only, the SIGN of an exactly-zero result
// when an equilibrium underflows to -0.  tests/ compare these kernels bit for bit with the
// literal oracle; tools/config1_divergence.py follows the blow-up of config 1.

// Lattice::density + momentum_density + velocity (src/lbm.rs:117-138) by identities (1), (2).
template <typename T>
__device__ __forceinline__ Moments<T> moments_reduced(const T (&g)[Q])
{
    Moments<T> m;
    m.rho = density(g);
    // c_x = (0, 1, 0,-1, 0, 1,-1,-1, 1):  ((0 + g1) - g3) + g5 - g6 - g7 + g8, in index order
    m.mx = add(sub(sub(add(sub(add(T(0), g[1]), g[3]), g[5]), g[6]), g[7]), g[8]);
    // c_y = (0, 0, 1, 0,-1, 1, 1,-1,-1):  ((0 + g2) - g4) + g5 + g6 - g7 - g8
    m.my = sub(sub(add(add(sub(add(T(0), g[2]), g[4]), g[5]), g[6]), g[7]), g[8]);
    const T r = recip(m.rho);
    m.vx = mul(r, m.mx);
    m.vy = mul(r, m.my);
    return m;
}

// compute_equilibrium for all nine directions (src/lbm.rs:58-68) by identities (1), (3), (4).
template <typename T, typename K>
__device__ __forceinline__ void equilibrium_pair(T vc, T c, T rw, const K &k, T &fe_i, T &fe_opp)
{
    const T a = mul(vc, k.k1);
    const T b = mul(mul(vc, vc), k.k2);
    fe_i   = mul(rw, add(add(add(T(1), a), b), c));
    fe_opp = mul(rw, add(add(sub(T(1), a), b), c));     // vc_opp = -vc: 1 + (-a) == 1 - a
}

template <typename T, typename K>
__device__ __forceinline__ void equilibrium_all(const T (&g)[Q], const K &k, Moments<T> &m, T (&fe)[Q])
{
    m = moments_reduced(g);
    const T v2 = add(mul(m.vx, m.vx), mul(m.vy, m.vy));   // src/lbm.rs:53
    const T c = mul(v2, k.k3);
    const T rw0 = mul(m.rho, k.w[0]), rws = mul(m.rho, k.w[1]), rwd = mul(m.rho, k.w[5]);
    fe[0] = mul(rw0, add(T(1), c));                       // vc_0 = +-0: ((1 + +-0) + 0) + c
    equilibrium_pair(m.vx, c, rws, k, fe[1], fe[3]);                 // c_1 = ( 1, 0) = -c_3
    equilibrium_pair(m.vy, c, rws, k, fe[2], fe[4]);                 // c_2 = ( 0, 1) = -c_4
    equilibrium_pair(add(m.vx, m.vy), c, rwd, k, fe[5], fe[7]);      // c_5 = ( 1, 1) = -c_7
    equilibrium_pair(sub(m.vy, m.vx), c, rwd, k, fe[6], fe[8]);      // c_6 = (-1, 1) = -c_8
}

// State::collide with BGK (src/lbm.rs:731-739, :349-364): g <- g + (g - feq)*factor
template <typename T, typename K>
__device__ __forceinline__ void collide_bgk(T (&g)[Q], const K &k)
{
    Moments<T> m; T fe[Q];
    equilibrium_all(g, k, m, fe);
#pragma unroll
    for (int i = 0; i < Q; ++i) g[i] = add(g[i], mul(sub(g[i], fe[i]), k.factor));
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

// TRT::evaluate (src/lbm.rs:401-444).  swap_equilibrium (:311-322) overwrites slots
// 1..8 of the equilibrium with the opposite *population* — reproduced as written.
template <typename T, typename K>
__device__ __forceinline__ void collide_trt(T (&g)[Q], const K &k)
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
// Pi_neq (:625-632) by identities (1), (2): c_x^2, c_y^2 are 0/1 and c_x*c_y is 0/+-1, and the
// yx sum repeats the xy sum operand for operand.  The second loop (:647-658) shares its products
// between directions of equal (c^2, w) (identity (4): axx_1 == axx_3, axx_5..8 equal, axy_6 ==
// -axy_5, ...) and skips the +-0 terms of the axis directions.
template <typename T, typename K>
__device__ __forceinline__ void collide_regularized(T (&g)[Q], const K &k)
{
    Moments<T> m; T fe[Q];
    equilibrium_all(g, k, m, fe);
    T n[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) n[i] = sub(g[i], fe[i]);
    const T sxx = add(add(add(add(add(add(T(0), n[1]), n[3]), n[5]), n[6]), n[7]), n[8]);
    const T syy = add(add(add(add(add(add(T(0), n[2]), n[4]), n[5]), n[6]), n[7]), n[8]);
    const T sxy = sub(add(sub(add(T(0), n[5]), n[6]), n[7]), n[8]);     // == syx
    const T xx0 = mul(sxx, k.axx[0]), xx1 = mul(sxx, k.axx[1]), xx2 = mul(sxx, k.axx[2]), xx5 = mul(sxx, k.axx[5]);
    const T yy0 = mul(syy, k.ayy[0]), yy1 = mul(syy, k.ayy[1]), yy2 = mul(syy, k.ayy[2]), yy5 = mul(syy, k.ayy[5]);
    const T xy5 = mul(sxy, k.axy[5]);                                   // axy_7 = axy_5, axy_6 = axy_8 = -axy_5
    g[0] = add(add(fe[0], xx0), yy0);
    g[1] = add(add(fe[1], xx1), yy1);
    g[2] = add(add(fe[2], xx2), yy2);
    g[3] = add(add(fe[3], xx1), yy1);
    g[4] = add(add(fe[4], xx2), yy2);
    g[5] = add(add(add(add(fe[5], xx5), xy5), xy5), yy5);
    g[6] = add(sub(sub(add(fe[6], xx5), xy5), xy5), yy5);
    g[7] = add(add(add(add(fe[7], xx5), xy5), xy5), yy5);
    g[8] = add(sub(sub(add(fe[8], xx5), xy5), xy5), yy5);
}

// KBC::evaluate (src/lbm.rs:468-585), entropic stabiliser gamma*.
template <typename T, typename K>
__device__ __forceinline__ void collide_kbc(T (&g)[Q], const K &k)
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
template <int COL, typename T, typename K>
__device__ __forceinline__ void collide(T (&g)[Q], const K &k)
{
    if (COL == COL_BGK) collide_bgk(g, k);
    else if (COL == COL_TRT) collide_trt(g, k);
    else if (COL == COL_REGULARIZED) collide_regularized(g, k);
    else collide_kbc(g, k);
}

// Two cells at once.  PACK (f32 only): the packed form above; otherwise one after the other.
template <int COL, bool PACK, typename T, typename K>
__device__ __forceinline__ void collide2(T (&c0)[Q], T (&c1)[Q], const K &k)
{
    if constexpr (PACK && sizeof(T) == 4) {
        F32x2 p[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) p[q] = F32x2(c0[q], c1[q]);
        collide<COL>(p, k);
#pragma unroll
        for (int q = 0; q < Q; ++q) { c0[q] = p[q].lo; c1[q] = p[q].hi; }
    } else {
        collide<COL>(c0, k);
        collide<COL>(c1, k);
    }
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
