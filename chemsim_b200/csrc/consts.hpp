// consts.hpp — the host scalars of the D2Q9 path, computed in the lattice dtype exactly as
// the reference computes them (src/lbm.rs:54-56, :64-66, :84, :209-219, :357, :428-439,
// :547-550, :638-656).  Host-only header: lattice.cu uses it for the device constants, the
// CPU arithmetic check (tests/host_arith/) for the same numbers.
#pragma once

#include <cmath>

#include "d2q9.cuh"

namespace chemsim {

struct CollisionParams {   // as given by the caller, converted to the lattice dtype in make_consts
    int kind = COL_NONE;
    double tau = 0.0;                      // BGK
    double tau_plus = 0.0, tau_minus = 0.0;   // TRT
    double viscosity = 0.0;                // KBC (and the viscosity a Regularized wrapper reports)
};

template <typename T>
Consts<T> make_consts(double dx_, double dt_, const CollisionParams &c)
{
    Consts<T> k{};
    static const int num[Q] = {16, 4, 4, 4, 4, 1, 1, 1, 1};
    for (int i = 0; i < Q; ++i) k.w[i] = (T)num[i] / (T)36.0;            // src/lbm.rs:209-219
    const T dx = (T)dx_, dt = (T)dt_;
    const T cs = dx / (std::sqrt((T)3.0) * dt);                           // src/lbm.rs:84
    k.cs2 = cs * cs;                                                      // :55
    const T cs4 = k.cs2 * k.cs2;                                          // :56
    k.k1 = (T)1.0 / k.cs2;                                                // :64
    k.k2 = (T)1.0 / ((T)2.0 * cs4);                                       // :65
    k.k3 = (T)-1.0 / ((T)2.0 * k.cs2);                                    // :66
    k.factor = c.tau != 0.0 ? -dt / (T)c.tau : (T)0;                      // :357
    // TRT, src/lbm.rs:428-439
    k.omega_m = c.tau_minus != 0.0 ? (T)1.0 / (T)c.tau_minus : (T)0;
    k.omega_p = c.tau_plus != 0.0 ? (T)1.0 / (T)c.tau_plus : (T)0;
    k.half = -dt * (T)0.5;
    // Regularized, src/lbm.rs:638-656
    for (int i = 0; i < Q; ++i) {
        const T cx = (T)cx_of(i), cy = (T)cy_of(i);
        const T qxx = cx * cx - k.cs2, qxy = cx * cy, qyx = cy * cx, qyy = cy * cy - k.cs2;
        const T sf = k.w[i] / ((T)2.0 * cs4);
        k.axx[i] = qxx * sf; k.axy[i] = qxy * sf; k.ayx[i] = qyx * sf; k.ayy[i] = qyy * sf;
    }
    // KBC, src/lbm.rs:478-571
    k.dx = dx;
    k.dx2 = dx * dx;
    k.dx_4 = dx * (T)4.0;
    k.four_dx = (T)4.0 * dx;
    k.two_dx2 = (T)2.0 * dx * dx;
    k.neg_dx = -dx;
    const T beta = (T)1.0 / (((T)2.0 * (T)c.viscosity / (cs * cs)) + (T)1.0);   // :547-550
    k.neg_beta = -beta;
    k.two_neg_beta = (T)2.0 * -beta;
    k.gamma_scale = (T)2.0 - (T)1.0 / beta;
    k.gamma_shift = (T)-1.0 / beta;
    return k;
}

}  // namespace chemsim
