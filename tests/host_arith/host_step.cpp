// tests/host_arith/host_step.cpp — TEST INFRASTRUCTURE.
// Compiles the product's per-cell arithmetic (chemsim_b200/csrc/d2q9.cuh: bounce_back and
// collide<COL>, and the host scalars of consts.hpp) with g++ for the CPU, each CUDA
// round-to-nearest intrinsic mapped to the plain IEEE operation (-ffp-contract=off), and
// wraps it in a cell loop with the same pull-stream indexing as the step kernels.  The CPU
// test suite compares it bit for bit with the literal oracle, so the exactness of the
// strength-reduced collision operators is checked without a GPU.  Not part of the product.
#include <cmath>
#include <cstddef>
#include <cstdint>

static inline float  __fadd_rn(float a, float b)   { return a + b; }
static inline float  __fsub_rn(float a, float b)   { return a - b; }
static inline float  __fmul_rn(float a, float b)   { return a * b; }
static inline float  __fdiv_rn(float a, float b)   { return a / b; }
static inline float  __fsqrt_rn(float a)           { return std::sqrt(a); }
static inline float  __frcp_rn(float a)            { return 1.0f / a; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a)          { return std::sqrt(a); }
static inline double __drcp_rn(double a)           { return 1.0 / a; }

#define CHEMSIM_HOST_ARITH 1
#include "../../chemsim_b200/csrc/consts.hpp"

using namespace chemsim;

template <typename T>
static void pull_cell(const T *src, const uint8_t *solid, int W, int H, int periodic, int y, int x, T (&c)[Q])
{
    const size_t plane = (size_t)W * H;
    for (int q = 0; q < Q; ++q) {
        int sy = y - ey_of(q), sx = x - ex_of(q);
        bool inside = sy >= 0 && sy < H && sx >= 0 && sx < W;
        if (!inside && periodic) { sy = (sy + H) % H; sx = (sx + W) % W; inside = true; }
        c[q] = inside ? src[q * plane + (size_t)sy * W + sx] : T(0);
    }
    bounce_back(c, solid && solid[(size_t)y * W + x] != 0);
}

// PAIRS: two neighbouring cells at a time through collide2<COL, true> — for f32 the F32x2
// instantiation of the collision operators that the f32 step kernels run (its packed additions
// are two lane-wise IEEE additions here, see d2q9.cuh)
template <typename T, int COL, bool PAIRS>
static void step_col(const T *src, T *dst, const uint8_t *solid, int W, int H, int periodic, const Consts<T> &k)
{
    const size_t plane = (size_t)W * H;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; x += PAIRS ? 2 : 1) {
            T c[Q];
            pull_cell(src, solid, W, H, periodic, y, x, c);
            if (PAIRS) {
                const int x1 = x + 1 < W ? x + 1 : x;      // odd width: the last cell is paired with itself
                T d[Q];
                pull_cell(src, solid, W, H, periodic, y, x1, d);
                collide2<COL, true>(c, d, k);
                for (int q = 0; q < Q; ++q) dst[q * plane + (size_t)y * W + x1] = d[q];
            } else {
                collide<COL>(c, k);
            }
            for (int q = 0; q < Q; ++q) dst[q * plane + (size_t)y * W + x] = c[q];
        }
}

template <typename T, bool PAIRS>
static int step_any(const T *src, T *dst, const uint8_t *solid, int W, int H, int periodic, double dx, double dt,
                    int kind, double tau, double tau_plus, double tau_minus, double viscosity)
{
    CollisionParams c;
    c.kind = kind; c.tau = tau; c.tau_plus = tau_plus; c.tau_minus = tau_minus; c.viscosity = viscosity;
    const Consts<T> k = make_consts<T>(dx, dt, c);
    switch (kind) {
    case COL_BGK:         step_col<T, COL_BGK, PAIRS>(src, dst, solid, W, H, periodic, k); return 0;
    case COL_TRT:         step_col<T, COL_TRT, PAIRS>(src, dst, solid, W, H, periodic, k); return 0;
    case COL_REGULARIZED: step_col<T, COL_REGULARIZED, PAIRS>(src, dst, solid, W, H, periodic, k); return 0;
    case COL_KBC:         step_col<T, COL_KBC, PAIRS>(src, dst, solid, W, H, periodic, k); return 0;
    }
    return 1;
}

extern "C" {
int host_step_f32(const float *src, float *dst, const uint8_t *solid, int W, int H, int periodic, double dx, double dt,
                  int kind, double tau, double tau_plus, double tau_minus, double viscosity)
{
    return step_any<float, false>(src, dst, solid, W, H, periodic, dx, dt, kind, tau, tau_plus, tau_minus, viscosity);
}
// the F32x2 ("packed") instantiation, two cells at a time
int host_step_f32_pairs(const float *src, float *dst, const uint8_t *solid, int W, int H, int periodic, double dx,
                        double dt, int kind, double tau, double tau_plus, double tau_minus, double viscosity)
{
    return step_any<float, true>(src, dst, solid, W, H, periodic, dx, dt, kind, tau, tau_plus, tau_minus, viscosity);
}
int host_step_f64(const double *src, double *dst, const uint8_t *solid, int W, int H, int periodic, double dx, double dt,
                  int kind, double tau, double tau_plus, double tau_minus, double viscosity)
{
    return step_any<double, false>(src, dst, solid, W, H, periodic, dx, dt, kind, tau, tau_plus, tau_minus, viscosity);
}
}
