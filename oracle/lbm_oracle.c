/*
 * oracle/lbm_oracle.c — instantiates lbm_oracle_impl.h for float and double.
 * TEST INFRASTRUCTURE ONLY; see lbm_oracle.h (parity unpinned).
 */
#include "lbm_oracle.h"

#include <math.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* lattice velocities c_i, src/lbm.rs:221-231 */
const int ORACLE_CX[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
const int ORACLE_CY[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
/* shift of population i under State::stream in [y][x] memory terms:
 * (EY, EX) = (-c_ix, +c_iy)  (derived in oracle/lbm_numpy.py
 * from the literal convolve2d calls)                                          */
const int ORACLE_EY[9] = {0, -1, 0, 1, 0, -1, 1, 1, -1};
const int ORACLE_EX[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
/* opposite directions, src/lbm.rs:298-309 */
const int ORACLE_OPP[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};

void lbm_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int lbm_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define REAL float
#define FN(name) name##_f32
#define SQRT sqrtf
#include "lbm_oracle_impl.h"
#undef REAL
#undef FN
#undef SQRT

#define REAL double
#define FN(name) name##_f64
#define SQRT sqrt
#include "lbm_oracle_impl.h"
#undef REAL
#undef FN
#undef SQRT
