// main_rs_harness.cpp — drives the C ABI through the C++ mirror exactly as the
// reference's binary drives lbm.rs: `initial_state` (src/main.rs:180-328) followed by
// the LBMSim frame loop handle -> step x speed_factor -> render (src/main.rs:66-177,
// src/display.rs:121-147), headless like display::record (src/display.rs:157-185).
//
//   main_rs_harness W H FRAMES [paint_frame] [bgk|regkbc]
//
// The collision operator is the BGK{tau: 15.0} alternative of main.rs:187 (default) or the
// active Regularized<KBC(10)> of main.rs:198-199 ("regkbc").  Prints one line per frame for the parity test to compare
// with the oracle: frame, state.time, total mass, density[probe], speed[probe].
#include <cinttypes>
#include <cstdio>
#include <cstdlib>

#include "lbm.hpp"

using namespace chemsim::lbm;

struct LBMSim {                       // src/main.rs:55-61
    size_t speed_factor = 2;          // :324
    std::pair<size_t, size_t> size;
    State state;
};

static LBMSim initial_state(std::pair<size_t, size_t> size, bool regkbc)   // src/main.rs:180
{
    const size_t w = size.first, h = size.second;
    const Discretization disc{1.0f, 1.0f};                      // :185
    const BGK collision{15.0f};                                 // :187
    std::vector<Scalar> vec_x(w * h, 0.0f), vec_y(w * h, 0.0f); // :201-214
    for (size_t x = 0; x < w; ++x)
        for (size_t y = 0; y < h; ++y) { vec_x[y * w + x] = 0.02f; vec_y[y * w + x] = 0.0f; }
    const Matrix vx = Matrix::make(vec_x, size), vy = Matrix::make(vec_y, size);   // :215-216
    const Matrix density = Matrix::new_filled(1.0f, size);                          // :223
    const Populations pops = compute_equilibrium(density, {vx, vy}, D2Q9::directions(), disc);   // :258-263
    const D2Q9 lattice = D2Q9::make(pops);                                          // :267
    Geometry geometry(w * h, 0);                                                    // :269-295
    for (size_t x = 0; x < w; ++x)
        for (size_t y = 0; y < h; ++y) {
            double r = 0.0;
            r += std::pow((double)x - ((double)w / 2.0), 2);
            r += std::pow((double)y - ((double)h / 2.0), 2);
            r = std::sqrt(r);
            if (r < 25.0) geometry[y * w + x] = 1;
            if (x == 0 || y == 0 || x == w - 1 || y == h - 1) geometry[y * w + x] = 1;
        }
    LBMSim sim;
    sim.size = size;
    if (regkbc) sim.state = State::initial(lattice, geometry, Regularized<KBC>{KBC{10.0f}}, disc);   // :198-199
    else        sim.state = State::initial(lattice, geometry, collision, disc);     // :314-319
    return sim;
}

// The mouse handler of src/main.rs:71-91 with the cursor at (x, y): rewrites the whole
// geometry so that only the 9x9 block around the cursor is solid.
static void paint(LBMSim &sim, size_t x, size_t y)
{
    Geometry vec = sim.state.geometry();                                            // geometry.host(), :81
    for (size_t a = 0; a < sim.size.first; ++a)
        for (size_t b = 0; b < sim.size.second; ++b) {
            const long dx = std::labs((long)a - (long)x), dy = std::labs((long)b - (long)y);
            vec[b * sim.size.first + a] = (dx < 5) && (dy < 5);                     // :84-86
        }
    sim.state.set_geometry(vec);                                                    // :89
}

int main(int argc, char **argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: %s W H FRAMES [paint_frame]\n", argv[0]); return 2; }
    const size_t w = std::strtoul(argv[1], nullptr, 10), h = std::strtoul(argv[2], nullptr, 10);
    const int frames = std::atoi(argv[3]);
    const int paint_frame = argc > 4 ? std::atoi(argv[4]) : -1;
    const bool regkbc = argc > 5 && std::string(argv[5]) == "regkbc";
    try {
        LBMSim sim = initial_state({w, h}, regkbc);
        const size_t probe = (h / 2) * w + w / 4;
        for (int f = 0; f < frames; ++f) {
            if (f == paint_frame) paint(sim, w / 4, h / 2);                         // Simulation::handle
            for (size_t s = 0; s < sim.speed_factor; ++s) sim.state.step();         // Simulation::step, :128-136
            const Matrix rho = sim.state.density();                                 // Simulation::render, :157-160
            const Matrix spd = sim.state.speed();
            const std::vector<uint8_t> image = sim.state.render(0);                 // device-side render_scalar_field
            if (image.size() != w * h * 4 || image[3] != 255) { std::printf("render FAILED\n"); return 1; }
            std::printf("frame %d time %.9g mass %.17g rho %.9g speed %.9g unstable %d\n", f, (double)sim.state.time(),
                        sim.state.total_mass(), (double)rho.get_underlying()[probe], (double)spd.get_underlying()[probe],
                        (int)sim.state.is_unstable());
        }
        // error behaviour: a wrong-sized slice is the reference's Err(InvalidSliceSize)
        try { Matrix::make(std::vector<Scalar>(3), {2, 2}); std::printf("error-check FAILED\n"); return 1; }
        catch (const InvalidSliceSize &) { std::printf("error-check ok\n"); }
    } catch (const LbmError &e) {
        std::fprintf(stderr, "LbmError %d: %s\n", e.status, e.what());
        return 1;
    }
    return 0;
}
