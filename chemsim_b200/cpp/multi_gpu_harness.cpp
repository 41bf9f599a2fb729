// multi_gpu_harness.cpp — main.rs's scenario on N GPUs from ONE process through lbm_multi.hpp.
//   multi_gpu_harness W H FRAMES NGPUS [p2p|nccl]
// Prints one line per frame (two steps each, speed_factor = 2) for the parity test.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "lbm_multi.hpp"

using namespace chemsim::lbm;

int main(int argc, char **argv)
{
    if (argc < 5) { std::fprintf(stderr, "usage: %s W H FRAMES NGPUS [p2p|nccl]\n", argv[0]); return 2; }
    const size_t w = std::strtoul(argv[1], nullptr, 10), h = std::strtoul(argv[2], nullptr, 10);
    const int frames = std::atoi(argv[3]), ngpus = std::atoi(argv[4]);
    const bool p2p = !(argc > 5 && std::string(argv[5]) == "nccl");
    try {
        const Discretization disc{1.0f, 1.0f};
        const Matrix vx = Matrix::new_filled(0.02f, {w, h}), vy = Matrix::new_filled(0.0f, {w, h});
        const Populations pops = compute_equilibrium(Matrix::new_filled(1.0f, {w, h}), {vx, vy}, D2Q9::directions(), disc);
        Geometry geometry(w * h, 0);
        for (size_t x = 0; x < w; ++x)
            for (size_t y = 0; y < h; ++y) {
                const double r = std::sqrt(std::pow((double)x - w / 2.0, 2) + std::pow((double)y - h / 2.0, 2));
                geometry[y * w + x] = r < 25.0 || x == 0 || y == 0 || x == w - 1 || y == h - 1;
            }
        std::vector<int> devices;
        for (int d = 0; d < ngpus; ++d) devices.push_back(d);
        MultiState sim = MultiState::initial(D2Q9::make(pops), geometry, BGK{15.0f}, disc, CHEMSIM_LBM_EDGE_ZEROFILL,
                                             devices, p2p);
        std::printf("halo %s\n", sim.halo_mode() == CHEMSIM_LBM_HALO_P2P ? "p2p" : "nccl");
        const size_t probe = (h / 2) * w + w / 4;
        for (int f = 0; f < frames; ++f) {
            sim.step(2);
            const Matrix rho = sim.density();
            std::printf("frame %d time %.9g mass %.17g rho %.9g unstable %d\n", f, (double)sim.time(), sim.total_mass(),
                        (double)rho.get_underlying()[probe], (int)sim.is_unstable());
        }
    } catch (const LbmError &e) {
        std::fprintf(stderr, "LbmError %d: %s\n", e.status, e.what());
        return 1;
    }
    return 0;
}
