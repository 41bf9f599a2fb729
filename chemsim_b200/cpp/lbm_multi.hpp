// lbm_multi.hpp — one lattice on several GPUs of the box, driven from ONE process (what the
// reference's single-process main.rs needs to use more than one GPU).  The C++ twin of
// chemsim_b200/lbm.py: MultiState.  One y-slab handle per device; the collective entry points
// of the C ABI (slab creation, enable_p2p_halo, the first step after an upload, sharded render)
// are called from one host thread per slab, everything else is a loop over the slabs.
#pragma once

#include <functional>
#include <thread>

#include "lbm.hpp"

namespace chemsim {
namespace lbm {

class MultiState {
public:
    // State::initial for a lattice sharded over `devices` (src/lbm.rs:679-692 + SURVEY.md §8e).
    template <typename Collision>
    static MultiState initial(const D2Q9 &lattice, const Geometry &geometry, const Collision &collision,
                              const Discretization &disc, int edge, const std::vector<int> &devices, bool p2p = true)
    {
        if (!lattice.populations.from_equilibrium)
            throw LbmError(CHEMSIM_LBM_ERR_UNSUPPORTED, "MultiState::initial takes the value of compute_equilibrium");
        MultiState m;
        m.size_ = lattice.size;
        const int n = (int)devices.size(), w = (int)lattice.size.first, h = (int)lattice.size.second;
        unsigned char id[CHEMSIM_LBM_NCCL_ID_BYTES] = {0};
        if (n > 1) check(chemsim_lbm_nccl_unique_id(id), nullptr);
        m.slabs_.assign(n, nullptr);
        m.row0_.assign(n, 0);
        m.rows_.assign(n, 0);
        const Populations &p = lattice.populations;
        m.each([&](int r) {
            chemsim_lbm_t *s = nullptr;
            check(chemsim_lbm_create_slab(w, h, CHEMSIM_LBM_F32, edge, devices[r], r, n, id, &s), nullptr);
            m.slabs_[r] = s;
            check(chemsim_lbm_shape(s, nullptr, &m.rows_[r], nullptr, &m.row0_[r]), s);
            check(chemsim_lbm_set_discretization(s, disc.delta_x, disc.delta_t), s);
            check(collision.apply(s, disc), s);
            const size_t off = (size_t)m.row0_[r] * w, cnt = (size_t)m.rows_[r] * w;   // slab rows are contiguous
            check(chemsim_lbm_init_equilibrium(s, p.density.get_underlying().data() + off,
                                               p.vx.get_underlying().data() + off, p.vy.get_underlying().data() + off,
                                               cnt), s);
            check(chemsim_lbm_set_geometry(s, geometry.data() + off, cnt), s);
            if (p2p && n > 1) check(chemsim_lbm_enable_p2p_halo(s), s);
        });
        return m;
    }

    MultiState() = default;
    MultiState(const MultiState &) = delete;
    MultiState &operator=(const MultiState &) = delete;
    MultiState(MultiState &&o) noexcept { *this = std::move(o); }
    MultiState &operator=(MultiState &&o) noexcept
    {
        destroy();
        slabs_ = std::move(o.slabs_); row0_ = std::move(o.row0_); rows_ = std::move(o.rows_); size_ = o.size_;
        o.slabs_.clear();
        return *this;
    }
    ~MultiState() { destroy(); }

    void step(int nsteps = 1) { each([&](int r) { check(chemsim_lbm_step(slabs_[r], nsteps), slabs_[r]); }); }
    void synchronize() { each([&](int r) { check(chemsim_lbm_synchronize(slabs_[r]), slabs_[r]); }); }
    std::pair<size_t, size_t> size() const { return size_; }
    int halo_mode() const { int m = 0; check(chemsim_lbm_halo_mode(slabs_[0], &m), slabs_[0]); return m; }
    Scalar time() const { double t = 0; check(chemsim_lbm_time(slabs_[0], &t), slabs_[0]); return (Scalar)t; }

    Matrix density() const { return gather(chemsim_lbm_get_density); }
    Matrix speed() const { return gather(chemsim_lbm_get_speed); }
    Matrix pressure() const { return gather(chemsim_lbm_get_pressure); }
    double total_mass() const
    {
        double total = 0.0;
        for (chemsim_lbm_t *s : slabs_) { double m = 0; check(chemsim_lbm_total_mass(s, &m), s); total += m; }
        return total;
    }
    bool is_unstable() const
    {
        bool any = false;
        for (chemsim_lbm_t *s : slabs_) { int f = 0; check(chemsim_lbm_is_unstable(s, &f), s); any = any || f != 0; }
        return any;
    }

private:
    // one host thread per slab (the collective calls must be issued concurrently)
    void each(const std::function<void(int)> &fn) const
    {
        std::vector<std::thread> threads;
        std::vector<std::exception_ptr> errors(slabs_.size());
        for (size_t r = 0; r < slabs_.size(); ++r)
            threads.emplace_back([&, r] { try { fn((int)r); } catch (...) { errors[r] = std::current_exception(); } });
        for (std::thread &t : threads) t.join();
        for (const std::exception_ptr &e : errors) if (e) std::rethrow_exception(e);
    }
    template <typename F> Matrix gather(F fn) const
    {
        Matrix out = Matrix::new_filled(0.0f, size_);
        const size_t w = size_.first;
        for (size_t r = 0; r < slabs_.size(); ++r)
            check(fn(slabs_[r], out.data().data() + (size_t)row0_[r] * w, (size_t)rows_[r] * w), slabs_[r]);
        return out;
    }
    void destroy()
    {
        if (slabs_.empty()) return;
        std::vector<std::thread> threads;
        for (chemsim_lbm_t *s : slabs_) threads.emplace_back([s] { if (s) chemsim_lbm_destroy(s); });
        for (std::thread &t : threads) t.join();
        slabs_.clear();
    }
    std::vector<chemsim_lbm_t *> slabs_;
    std::vector<int> row0_, rows_;
    std::pair<size_t, size_t> size_{0, 0};
};

}  // namespace lbm
}  // namespace chemsim
