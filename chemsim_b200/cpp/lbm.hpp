// lbm.hpp — C++ host-side mirror of the reference's `chemsim::lbm` module over the
// C ABI of include/chemsim_lbm.h.
//
// The reference's host code is Rust (/root/reference/src/lbm.rs, matrix.rs); no Rust
// toolchain exists in this image, so the compiled host layer above the C ABI is
// written in C++ with the same names, argument meaning and error behaviour, and the
// Rust shim itself is given as source in INTEGRATION.md.  Header-only; every
// floating-point result comes from libchemsim_lbm.so (CUDA), nothing is computed here.
//
//   lbm.rs item                         here
//   ----------------------------------  -----------------------------------------
//   Scalar (:13)                        chemsim::lbm::Scalar
//   Matrix, Error::InvalidSliceSize     chemsim::lbm::Matrix, InvalidSliceSize   (matrix.rs:10-44, :120-126)
//   Discretization (:75-86)             Discretization
//   Direction, D2Q9 (:90-95, :180-323)  Direction, D2Q9
//   compute_equilibrium (:43-71)        compute_equilibrium
//   BGK / TRT / KBC / Regularized<C>    BGK, TRT, KBC, Regularized<C>            (:345-666)
//   State<D2Q9> (:670-819)              State
//   render_* (render.rs:7-178)          State::render (device-side)
#pragma once

#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/chemsim_lbm.h"

namespace chemsim {
namespace lbm {

using Scalar = float;   // `pub type Scalar = f32`, src/lbm.rs:13

// Any non-zero status of the C ABI.  The reference panics (and aborts in release
// builds, Cargo.toml:128); a C++ exception is the closest recoverable equivalent.
struct LbmError : std::runtime_error {
    int status;
    LbmError(int status_, const std::string &msg) : std::runtime_error(msg), status(status_) {}
};
// matrix::Error::InvalidSliceSize, src/matrix.rs:15-19
struct InvalidSliceSize : LbmError {
    explicit InvalidSliceSize(const std::string &msg) : LbmError(CHEMSIM_LBM_ERR_INVALID_SLICE_SIZE, msg) {}
};

inline void check(int status, const chemsim_lbm_t *h)
{
    if (status == CHEMSIM_LBM_OK) return;
    const std::string msg = chemsim_lbm_last_error(h);
    if (status == CHEMSIM_LBM_ERR_INVALID_SLICE_SIZE) throw InvalidSliceSize(msg);
    throw LbmError(status, msg);
}

// matrix::Matrix as a host value: shape (w, h), element (y, x) at data[y*w + x].
class Matrix {
public:
    Matrix() = default;
    // Matrix::new(slice, (w, h)) -> Result<Matrix, InvalidSliceSize>, src/matrix.rs:24-30
    static Matrix make(const std::vector<Scalar> &slice, std::pair<size_t, size_t> dims)
    {
        if (slice.size() != dims.first * dims.second)
            throw InvalidSliceSize("slice has " + std::to_string(slice.size()) + " elements");
        Matrix m; m.w_ = dims.first; m.h_ = dims.second; m.data_ = slice; return m;
    }
    // Matrix::new_filled(value, (w, h)), src/matrix.rs:40-44 (with the intended (w,h) meaning)
    static Matrix new_filled(Scalar value, std::pair<size_t, size_t> dims)
    {
        Matrix m; m.w_ = dims.first; m.h_ = dims.second; m.data_.assign(dims.first * dims.second, value); return m;
    }
    size_t get_width() const { return w_; }
    size_t get_height() const { return h_; }
    std::pair<size_t, size_t> get_shape() const { return {w_, h_}; }
    const std::vector<Scalar> &get_underlying() const { return data_; }   // row-major y*w+x, src/matrix.rs:120-126
    std::vector<Scalar> &data() { return data_; }
private:
    size_t w_ = 0, h_ = 0;
    std::vector<Scalar> data_;
};

struct Discretization {   // src/lbm.rs:75-86
    Scalar delta_x = 1.0f, delta_t = 1.0f;
    Scalar isothermal_speed_of_sound() const { return delta_x / (std::sqrt(3.0f) * delta_t); }
};

struct Direction {        // src/lbm.rs:90-95
    Scalar w_scalar;
    std::pair<Scalar, Scalar> c_vector;
    int stencil[9];
};

// CollisionOperator impls (src/lbm.rs:327-666).  Each knows how to select itself on a
// handle (`apply`) and reports its viscosities in Scalar like the reference.
struct BGK {              // src/lbm.rs:345-370
    Scalar tau;
    Scalar kinematic_shear_viscosity(const Discretization &d) const
    {
        return (d.delta_x * d.delta_x / (3.0f * d.delta_t * d.delta_t)) * (tau - d.delta_t / 2.0f);
    }
    Scalar kinematic_bulk_viscosity(const Discretization &d) const { return 2.0f * kinematic_shear_viscosity(d) / 3.0f; }
    int apply(chemsim_lbm_t *h, const Discretization &) const { return chemsim_lbm_set_bgk(h, tau); }
};

struct TRT {              // src/lbm.rs:374-451
    Scalar tau_minus, tau_plus;
    static TRT make(Scalar lambda, Scalar ks_viscosity, const Discretization &d)   // TRT::new, :380-390
    {
        const Scalar dt = d.delta_t, cs = d.isothermal_speed_of_sound();
        const Scalar tau_plus = dt * ((ks_viscosity / (cs * cs)) + 0.5f);
        const Scalar tau_minus = dt * ((lambda / ((tau_plus / dt) - 0.5f)) + 0.5f);
        return TRT{tau_minus, tau_plus};
    }
    Scalar kinematic_shear_viscosity(const Discretization &d) const
    {
        const Scalar cs = d.isothermal_speed_of_sound();
        return cs * cs * (tau_plus / d.delta_t - 0.5f);
    }
    Scalar kinematic_bulk_viscosity(const Discretization &d) const { return 2.0f * kinematic_shear_viscosity(d) / 3.0f; }
    int apply(chemsim_lbm_t *h, const Discretization &) const { return chemsim_lbm_set_trt(h, tau_plus, tau_minus); }
};

struct KBC {              // src/lbm.rs:455-590
    Scalar ks_viscosity;
    Scalar kinematic_shear_viscosity(const Discretization &) const { return ks_viscosity; }
    Scalar kinematic_bulk_viscosity(const Discretization &d) const { return 2.0f * kinematic_shear_viscosity(d) / 3.0f; }
    int apply(chemsim_lbm_t *h, const Discretization &) const { return chemsim_lbm_set_kbc(h, ks_viscosity); }
};

template <typename C>
struct Regularized {      // src/lbm.rs:596-666: only the underlying operator's viscosity is ever used
    C underlying;
    Scalar kinematic_shear_viscosity(const Discretization &d) const { return underlying.kinematic_shear_viscosity(d); }
    Scalar kinematic_bulk_viscosity(const Discretization &d) const { return 2.0f * kinematic_shear_viscosity(d) / 3.0f; }
    // Regularized::kinematic_shear_viscosity(disc) = underlying's, with the State's discretization (:663-665)
    int apply(chemsim_lbm_t *h, const Discretization &d) const { return chemsim_lbm_set_regularized(h, underlying.kinematic_shear_viscosity(d)); }
};

// Value of compute_equilibrium: kept as the generating fields and evaluated on the GPU
// when the State is built (chemsim_lbm_init_equilibrium).
struct Populations {
    bool from_equilibrium = false;
    Matrix density, vx, vy;
    std::vector<Matrix> explicit_pops;   // nine arrays when built by hand
    size_t len() const { return 9; }
};

struct D2Q9 {             // src/lbm.rs:180-323
    std::pair<size_t, size_t> size;
    Populations populations;

    static std::vector<Direction> directions()   // :202-282
    {
        static const int num[9] = {16, 4, 4, 4, 4, 1, 1, 1, 1};
        static const int cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
        static const int one_at[9] = {4, 3, 7, 5, 1, 6, 8, 2, 0};
        std::vector<Direction> out(9);
        for (int i = 0; i < 9; ++i) {
            out[i].w_scalar = (Scalar)num[i] / 36.0f;
            out[i].c_vector = {(Scalar)cx[i], (Scalar)cy[i]};
            for (int j = 0; j < 9; ++j) out[i].stencil[j] = j == one_at[i] ? 1 : 0;
        }
        return out;
    }
    static D2Q9 make(const Populations &pops)   // D2Q9::new, :187-200
    {
        D2Q9 l; l.populations = pops;
        if (pops.from_equilibrium) l.size = pops.density.get_shape();
        else {
            if (pops.explicit_pops.size() != 9) throw LbmError(CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "need 9 populations");
            l.size = pops.explicit_pops[0].get_shape();
            for (const Matrix &m : pops.explicit_pops)
                if (m.get_shape() != l.size) throw LbmError(CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "population shapes differ");
        }
        return l;
    }
};

// lbm::compute_equilibrium, src/lbm.rs:43-71
inline Populations compute_equilibrium(const Matrix &density, const std::pair<Matrix, Matrix> &velocity,
                                       const std::vector<Direction> &directions, const Discretization &)
{
    if (density.get_shape() != velocity.first.get_shape() || density.get_shape() != velocity.second.get_shape())
        throw LbmError(CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "assert_eq!(size, v.get_shape()) failed");   // :51-52
    if (directions.size() != 9) throw LbmError(CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "need 9 directions");
    Populations p; p.from_equilibrium = true; p.density = density; p.vx = velocity.first; p.vy = velocity.second;
    return p;
}

using Geometry = std::vector<uint8_t>;   // row-major bool[y*w+x], main.rs:269-312

// lbm::State<D2Q9>, src/lbm.rs:670-819 — the device-resident lattice.
class State {
public:
    // State::initial(Box<L>, Geometry, Box<CollisionOperator<L>>, Discretization), :679-692.
    // `edge` is the one extension (the reference is always zero-fill).
    template <typename Collision>
    static State initial(const D2Q9 &lattice, const Geometry &geometry, const Collision &collision,
                         const Discretization &disc, int edge = CHEMSIM_LBM_EDGE_ZEROFILL, int device = -1)
    {
        State s;
        chemsim_lbm_t *h = nullptr;
        check(chemsim_lbm_create((int)lattice.size.first, (int)lattice.size.second, CHEMSIM_LBM_F32, edge, device, &h), nullptr);
        s.h_.reset(h, [](chemsim_lbm_t *p) { chemsim_lbm_destroy(p); });
        s.size_ = lattice.size; s.discretization = disc;
        check(chemsim_lbm_set_discretization(h, disc.delta_x, disc.delta_t), h);
        check(collision.apply(h, disc), h);             // Box<dyn CollisionOperator<L>>, :674
        const Populations &p = lattice.populations;
        const size_t n = s.size_.first * s.size_.second;
        if (p.from_equilibrium)
            check(chemsim_lbm_init_equilibrium(h, p.density.get_underlying().data(), p.vx.get_underlying().data(),
                                               p.vy.get_underlying().data(), n), h);
        else
            for (int q = 0; q < 9; ++q)
                check(chemsim_lbm_set_population(h, q, p.explicit_pops[q].get_underlying().data(),
                                                 p.explicit_pops[q].get_underlying().size()), h);
        s.set_geometry(geometry);
        return s;
    }

    void step() { check(chemsim_lbm_step(h_.get(), 1), h_.get()); }                 // :694-714
    void step(int n) { check(chemsim_lbm_step(h_.get(), n), h_.get()); }
    Scalar time() const { double t = 0; check(chemsim_lbm_time(h_.get(), &t), h_.get()); return (Scalar)t; }   // pub time, :671

    std::pair<size_t, size_t> size() const { return size_; }                         // :753-756
    Scalar delta_x() const { return discretization.delta_x; }
    Scalar delta_t() const { return discretization.delta_t; }
    Scalar isothermal_speed_of_sound() const { return discretization.isothermal_speed_of_sound(); }

    Matrix density() const { return get1(chemsim_lbm_get_density); }                 // :779
    Matrix pressure() const { return get1(chemsim_lbm_get_pressure); }               // :784
    Matrix speed() const { return get1(chemsim_lbm_get_speed); }                     // :800
    std::pair<Matrix, Matrix> velocity() const { return get2(chemsim_lbm_get_velocity); }                  // :795
    std::pair<Matrix, Matrix> momentum_density() const { return get2(chemsim_lbm_get_momentum_density); }  // :790
    Matrix population(int q) const                                                   // populations()[q].1, :769
    {
        Matrix m = Matrix::new_filled(0.0f, size_);
        check(chemsim_lbm_get_population(h_.get(), q, m.data().data(), m.data().size()), h_.get());
        return m;
    }
    bool is_unstable() const { int f = 0; check(chemsim_lbm_is_unstable(h_.get(), &f), h_.get()); return f != 0; }   // :815
    double total_mass() const { double m = 0; check(chemsim_lbm_total_mass(h_.get(), &m), h_.get()); return m; }

    // `pub geometry` field (:673): reassignable between steps, main.rs:77-89
    void set_geometry(const Geometry &g) { check(chemsim_lbm_set_geometry(h_.get(), g.data(), g.size()), h_.get()); }
    // main.rs:71-91 (the mouse handler: the geometry becomes the 9x9 block around the cursor) without its host
    // round trip: row = floor(pos[1]), column = floor(pos[0]); two small kernels on the device
    void paint_brush(double pos_x, double pos_y)
    {
        const long row = (long)std::floor(pos_y), col = (long)std::floor(pos_x);
        if (row < 0 || col < 0 || row >= (long)size_.second || col >= (long)size_.first) return;
        check(chemsim_lbm_fill_geometry(h_.get(), 0), h_.get());
        check(chemsim_lbm_paint_rect(h_.get(), (int)col - 4, (int)row - 4, 9, 9, 1), h_.get());
    }
    // the State as bytes (populations, geometry, time, step counter) and back
    std::vector<unsigned char> checkpoint() const
    {
        size_t n = 0;
        check(chemsim_lbm_checkpoint_bytes(h_.get(), &n), h_.get());
        std::vector<unsigned char> blob(n);
        check(chemsim_lbm_checkpoint(h_.get(), blob.data(), blob.size()), h_.get());
        return blob;
    }
    void restore(const std::vector<unsigned char> &blob) { check(chemsim_lbm_restore(h_.get(), blob.data(), blob.size()), h_.get()); }
    Geometry geometry() const
    {
        Geometry g(size_.first * size_.second);
        check(chemsim_lbm_get_geometry(h_.get(), g.data(), g.size()), h_.get());
        return g;
    }

    // render_scalar_field / render_vector_field + render_geometry (src/render.rs) on the device:
    // RGBA8, row-major y*w+x.  mode: 0 density, 1 speed, 2 velocity, 3 momentum density (main.rs:157-174)
    std::vector<uint8_t> render(int mode, bool overlay_geometry = true) const
    {
        std::vector<uint8_t> rgba(size_.first * size_.second * 4);
        check(chemsim_lbm_render(h_.get(), mode, overlay_geometry ? 1 : 0, rgba.data(), size_.first * size_.second), h_.get());
        return rgba;
    }

    Discretization discretization;
    chemsim_lbm_t *handle() const { return h_.get(); }

private:
    template <typename F> Matrix get1(F fn) const
    {
        Matrix m = Matrix::new_filled(0.0f, size_);
        check(fn(h_.get(), m.data().data(), m.data().size()), h_.get());
        return m;
    }
    template <typename F> std::pair<Matrix, Matrix> get2(F fn) const
    {
        Matrix a = Matrix::new_filled(0.0f, size_), b = Matrix::new_filled(0.0f, size_);
        check(fn(h_.get(), a.data().data(), b.data().data(), a.data().size()), h_.get());
        return {a, b};
    }
    std::shared_ptr<chemsim_lbm_t> h_;
    std::pair<size_t, size_t> size_{0, 0};
};

}  // namespace lbm
}  // namespace chemsim
