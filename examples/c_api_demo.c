/*
 * c_api_demo.c — the C ABI of include/chemsim_lbm.h from plain C99.
 *
 * Builds the setup of the reference's main.rs (src/main.rs:180-328) at W x H with
 * BGK{tau: 15.0} (main.rs:187), runs FRAMES frames of two steps each (speed_factor,
 * main.rs:324) and prints total mass and the centre-line density after every frame.
 *
 *   gcc -std=c99 -Iinclude examples/c_api_demo.c -Lchemsim_b200 -lchemsim_lbm \
 *       -Wl,-rpath,$PWD/chemsim_b200 -lm -o c_api_demo && ./c_api_demo 256 256 5
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "chemsim_lbm.h"

#define CHECK(call)                                                                        \
    do {                                                                                   \
        int st_ = (call);                                                                  \
        if (st_ != CHEMSIM_LBM_OK) {                                                       \
            fprintf(stderr, "%s -> status %d: %s\n", #call, st_, chemsim_lbm_last_error(h)); \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

int main(int argc, char **argv)
{
    const int w = argc > 1 ? atoi(argv[1]) : 256, hgt = argc > 2 ? atoi(argv[2]) : 256;
    const int frames = argc > 3 ? atoi(argv[3]) : 5;
    const size_t n = (size_t)w * (size_t)hgt;
    chemsim_lbm_t *h = NULL;

    float *rho = malloc(n * sizeof(float)), *vx = malloc(n * sizeof(float)), *vy = malloc(n * sizeof(float));
    uint8_t *solid = malloc(n);
    if (!rho || !vx || !vy || !solid) return 1;
    for (int y = 0; y < hgt; ++y)
        for (int x = 0; x < w; ++x) {
            const size_t i = (size_t)y * w + x;               /* row-major y*w + x, src/matrix.rs:24-30 */
            rho[i] = 1.0f; vx[i] = 0.02f; vy[i] = 0.0f;       /* main.rs:208-223 */
            const double r = sqrt(pow(x - w / 2.0, 2) + pow(y - hgt / 2.0, 2));
            solid[i] = r < 25.0 || x == 0 || y == 0 || x == w - 1 || y == hgt - 1;   /* main.rs:283-293 */
        }

    CHECK(chemsim_lbm_create(w, hgt, CHEMSIM_LBM_F32, CHEMSIM_LBM_EDGE_ZEROFILL, -1, &h));
    CHECK(chemsim_lbm_set_discretization(h, 1.0, 1.0));
    CHECK(chemsim_lbm_set_bgk(h, 15.0));
    CHECK(chemsim_lbm_init_equilibrium(h, rho, vx, vy, n));
    CHECK(chemsim_lbm_set_geometry(h, solid, n));

    for (int f = 0; f < frames; ++f) {
        double mass = 0.0, t = 0.0;
        int unstable = 0;
        CHECK(chemsim_lbm_step(h, 2));
        CHECK(chemsim_lbm_get_density(h, rho, n));
        CHECK(chemsim_lbm_total_mass(h, &mass));
        CHECK(chemsim_lbm_is_unstable(h, &unstable));
        CHECK(chemsim_lbm_time(h, &t));
        printf("frame %d time %g mass %.10f rho(centre line, x=w/4) %.7f unstable %d\n", f, t, mass,
               rho[(size_t)(hgt / 2) * w + w / 4], unstable);
    }
    /* a wrong-sized slice is reported, not crashed on (matrix::Error::InvalidSliceSize) */
    if (chemsim_lbm_get_density(h, rho, n - 1) != CHEMSIM_LBM_ERR_INVALID_SLICE_SIZE) return 1;
    CHECK(chemsim_lbm_destroy(h));
    free(rho); free(vx); free(vy); free(solid);
    return 0;
}
