/*
 * oracle/lbm_oracle_impl.h — type-generic body of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY (see lbm_oracle.h).  Included twice by lbm_oracle.c,
 * once with REAL=float (the reference's `Scalar`, /root/reference/src/lbm.rs:13)
 * and once with REAL=double (BASELINE.json's f64 extension: same operation
 * order evaluated in f64).
 *
 * Everything here is a restatement of what /root/reference/src/lbm.rs asks
 * ArrayFire to do, element by element, IN THE REFERENCE'S OPERATION ORDER.
 * Build with -ffp-contract=off: the ArrayFire CPU backend evaluates each JIT
 * node as a separate rounded IEEE operation.
 *
 * Conventions: a field is h rows of w values, element (y,x) at
 * [y*w + x] — the order Matrix::new takes and Matrix::get_underlying returns
 * (src/matrix.rs:24-30, :120-126).  The nine populations are stored as
 * f[q*h*w + y*w + x] ("SoA", which is what the reference's Vec<(Direction,
 * Population)> of nine separate arrays is, src/lbm.rs:103-107).
 */

#ifndef REAL
#error "include from lbm_oracle.c"
#endif

/* ---- host scalars, computed in REAL exactly as the reference does ---------- */

/* src/lbm.rs:81-86  cs = dx / (sqrt(3) * dt)            */
static REAL FN(cs)(REAL dx, REAL dt) { return dx / (SQRT((REAL)3.0) * dt); }

typedef struct {
    REAL w[9];      /* src/lbm.rs:209-219 */
    REAL cx[9];     /* src/lbm.rs:221-231 */
    REAL cy[9];
    REAL cs2, cs4;  /* src/lbm.rs:54-56   */
    REAL k1;        /* 1/cs2              src/lbm.rs:64 */
    REAL k2;        /* 1/(2*cs4)          src/lbm.rs:65 */
    REAL k3;        /* -1/(2*cs2)         src/lbm.rs:66 */
} FN(consts_t);

static void FN(make_consts)(FN(consts_t) *k, REAL dx, REAL dt)
{
    static const int num[9] = {16, 4, 4, 4, 4, 1, 1, 1, 1};
    for (int i = 0; i < 9; ++i) {
        k->w[i]  = (REAL)num[i] / (REAL)36.0;
        k->cx[i] = (REAL)ORACLE_CX[i];
        k->cy[i] = (REAL)ORACLE_CY[i];
    }
    REAL cs = FN(cs)(dx, dt);
    k->cs2 = cs * cs;
    k->cs4 = k->cs2 * k->cs2;
    k->k1 = (REAL)1.0 / k->cs2;
    k->k2 = (REAL)1.0 / ((REAL)2.0 * k->cs4);
    k->k3 = (REAL)-1.0 / ((REAL)2.0 * k->cs2);
}

void FN(lbm_oracle_constants)(REAL dx, REAL dt, REAL *out /* cs2,cs4,k1,k2,k3 */)
{
    FN(consts_t) k; FN(make_consts)(&k, dx, dt);
    out[0] = k.cs2; out[1] = k.cs4; out[2] = k.k1; out[3] = k.k2; out[4] = k.k3;
}

/* ---- array-at-a-time building blocks (the reference's structure) ----------- */

/* Lattice::density  src/lbm.rs:117-121 : result = 0; for pop: result += pop   */
void FN(lbm_oracle_density)(const REAL *f, size_t n, REAL *rho)
{
    for (size_t c = 0; c < n; ++c) rho[c] = (REAL)0.0;
    for (int i = 0; i < 9; ++i) {
        const REAL *fi = f + (size_t)i * n;
        for (size_t c = 0; c < n; ++c) rho[c] = rho[c] + fi[c];
    }
}

/* Lattice::momentum_density  src/lbm.rs:123-131 : md = md + f_i.scale(c_i)
 * (every product is formed, including *0.0 and *-1.0)                          */
void FN(lbm_oracle_momentum)(const REAL *f, size_t n, REAL *mx, REAL *my)
{
    for (size_t c = 0; c < n; ++c) { mx[c] = (REAL)0.0; my[c] = (REAL)0.0; }
    for (int i = 0; i < 9; ++i) {
        const REAL *fi = f + (size_t)i * n;
        const REAL cx = (REAL)ORACLE_CX[i], cy = (REAL)ORACLE_CY[i];
        for (size_t c = 0; c < n; ++c) {
            mx[c] = mx[c] + fi[c] * cx;
            my[c] = my[c] + fi[c] * cy;
        }
    }
}

/* Lattice::velocity  src/lbm.rs:133-138 ; Matrix::recip src/matrix.rs:133-136:
 * r = 1/rho (a real division), then v = r (*) m — reciprocal-then-multiply.    */
void FN(lbm_oracle_velocity)(const REAL *f, size_t n, REAL *vx, REAL *vy)
{
    REAL *rho = (REAL *)malloc(n * sizeof(REAL));
    FN(lbm_oracle_density)(f, n, rho);
    FN(lbm_oracle_momentum)(f, n, vx, vy);
    for (size_t c = 0; c < n; ++c) {
        REAL r = (REAL)1.0 / rho[c];
        vx[c] = r * vx[c];
        vy[c] = r * vy[c];
    }
    free(rho);
}

/* Lattice::speed  src/lbm.rs:151-154 */
void FN(lbm_oracle_speed)(const REAL *f, size_t n, REAL *speed)
{
    REAL *vx = (REAL *)malloc(n * sizeof(REAL));
    REAL *vy = (REAL *)malloc(n * sizeof(REAL));
    FN(lbm_oracle_velocity)(f, n, vx, vy);
    for (size_t c = 0; c < n; ++c) speed[c] = SQRT(vx[c] * vx[c] + vy[c] * vy[c]);
    free(vx); free(vy);
}

/* State::pressure  src/lbm.rs:784-787 : density().scale(cs*cs) */
void FN(lbm_oracle_pressure)(const REAL *f, size_t n, REAL dx, REAL dt, REAL *p)
{
    REAL cs = FN(cs)(dx, dt);
    REAL cs2 = cs * cs;
    FN(lbm_oracle_density)(f, n, p);
    for (size_t c = 0; c < n; ++c) p[c] = p[c] * cs2;
}

/* compute_equilibrium  src/lbm.rs:43-71 */
void FN(lbm_oracle_equilibrium)(const REAL *rho, const REAL *vx, const REAL *vy,
                                size_t n, REAL dx, REAL dt, REAL *feq)
{
    FN(consts_t) k; FN(make_consts)(&k, dx, dt);
    REAL *v2 = (REAL *)malloc(n * sizeof(REAL));
    for (size_t c = 0; c < n; ++c) v2[c] = vx[c] * vx[c] + vy[c] * vy[c];   /* :53 */
    for (int i = 0; i < 9; ++i) {
        REAL *out = feq + (size_t)i * n;
        for (size_t c = 0; c < n; ++c) {
            REAL vc  = vx[c] * k.cx[i] + vy[c] * k.cy[i];                    /* :60 */
            REAL vc2 = vc * vc;                                              /* :61 */
            REAL sum = (((REAL)1.0 + vc * k.k1) + vc2 * k.k2) + v2[c] * k.k3;/* :62-66 */
            out[c] = (rho[c] * k.w[i]) * sum;                                /* :67 */
        }
    }
    free(v2);
}

/* Lattice::equilibrium  src/lbm.rs:156-160 */
static void FN(lattice_equilibrium)(const REAL *f, size_t n, REAL dx, REAL dt, REAL *feq)
{
    REAL *rho = (REAL *)malloc(n * sizeof(REAL));
    REAL *vx  = (REAL *)malloc(n * sizeof(REAL));
    REAL *vy  = (REAL *)malloc(n * sizeof(REAL));
    FN(lbm_oracle_density)(f, n, rho);
    FN(lbm_oracle_velocity)(f, n, vx, vy);
    FN(lbm_oracle_equilibrium)(rho, vx, vy, n, dx, dt, feq);
    free(rho); free(vx); free(vy);
}

void FN(lbm_oracle_lattice_equilibrium)(const REAL *f, size_t n, REAL dx, REAL dt, REAL *feq)
{
    FN(lattice_equilibrium)(f, n, dx, dt, feq);
}

/* State::is_unstable  src/lbm.rs:815-818 : min(feq_0) < 0 */
int FN(lbm_oracle_is_unstable)(const REAL *f, size_t n, REAL dx, REAL dt)
{
    REAL *feq = (REAL *)malloc(9 * n * sizeof(REAL));
    FN(lattice_equilibrium)(f, n, dx, dt, feq);
    int bad = 0;
    for (size_t c = 0; c < n; ++c) if (feq[c] < (REAL)0.0) { bad = 1; break; }
    free(feq);
    return bad;
}

/* total mass: sum of all nine populations accumulated in double (the analogue
 * of Matrix::sum -> af::sum_all's f64 result, src/matrix.rs:138-140).
 * Pairwise-by-row so that the value does not depend on thread count.          */
double FN(lbm_oracle_total_mass)(const REAL *f, int w, int h)
{
    size_t n = (size_t)w * h;
    double total = 0.0;
    for (int y = 0; y < h; ++y) {
        double row = 0.0;
        for (int x = 0; x < w; ++x) {
            size_t c = (size_t)y * w + x;
            double cell = 0.0;
            for (int i = 0; i < 9; ++i) cell += (double)f[(size_t)i * n + c];
            row += cell;
        }
        total += row;
    }
    return total;
}

/* State::stream  src/lbm.rs:716-729.
 * af::convolve2(f_i, stencil_i^T, DEFAULT, SPATIAL) with a one-hot 3x3 filter:
 * out(y,x) = in(y - EY_i, x - EX_i) where (EY,EX) = (-c_ix, +c_iy), zero outside
 * the array (ArrayFire 3.6.1 semantics, unpinned).
 * edge==PERIODIC wraps instead (an extension the reference cannot express).   */
static void FN(stream)(const REAL *in, REAL *out, int w, int h, int edge)
{
    size_t n = (size_t)w * h;
    for (int i = 0; i < 9; ++i) {
        const REAL *src = in + (size_t)i * n;
        REAL *dst = out + (size_t)i * n;
        const int ey = ORACLE_EY[i], ex = ORACLE_EX[i];
        for (int y = 0; y < h; ++y) {
            int sy = y - ey;
            int yin = (sy >= 0 && sy < h);
            if (!yin && edge == ORACLE_EDGE_PERIODIC) { sy = (sy + h) % h; yin = 1; }
            for (int x = 0; x < w; ++x) {
                int sx = x - ex;
                int xin = (sx >= 0 && sx < w);
                if (!xin && edge == ORACLE_EDGE_PERIODIC) { sx = (sx + w) % w; xin = 1; }
                /* the convolution multiplies by the filter tap 1.0 and adds
                 * eight 0.0 products: exact for finite inputs                  */
                dst[(size_t)y * w + x] = (yin && xin) ? src[(size_t)sy * w + sx] : (REAL)0.0;
            }
        }
    }
}

/* D2Q9::swap_populations src/lbm.rs:298-309 + State::bounce_back :741-751:
 * af::replace(sw_i, geometry, f_i) keeps sw_i where geometry is true.          */
static void FN(bounce_back)(const REAL *in, REAL *out, const uint8_t *solid, size_t n)
{
    for (int i = 0; i < 9; ++i) {
        const REAL *fi = in + (size_t)i * n;
        const REAL *fo = in + (size_t)ORACLE_OPP[i] * n;
        REAL *dst = out + (size_t)i * n;
        for (size_t c = 0; c < n; ++c) dst[c] = (solid && solid[c]) ? fo[c] : fi[c];
    }
}

/* ---- collision operators (src/lbm.rs:345-666), array-at-a-time ------------- */

/* BGK::evaluate src/lbm.rs:349-364 : f + (f - feq).scale(-dt/tau) */
static void FN(collide_bgk)(const REAL *f, const REAL *feq, size_t n, REAL dt, REAL tau, REAL *out)
{
    const REAL factor = -dt / tau;
    for (size_t c = 0; c < 9 * n; ++c) out[c] = f[c] + (f[c] - feq[c]) * factor;
}

/* TRT::evaluate src/lbm.rs:401-444.  NOTE swap_equilibrium (src/lbm.rs:311-322)
 * overwrites slots 1..8 of the equilibrium with the opposite *population*, not
 * the opposite equilibrium; reproduced as written.                             */
static void FN(collide_trt)(const REAL *f, const REAL *feq, size_t n, REAL dt,
                            REAL tau_plus, REAL tau_minus, REAL *out)
{
    const REAL omega_m = (REAL)1.0 / tau_minus;
    const REAL omega_p = (REAL)1.0 / tau_plus;
    const REAL half = -dt * (REAL)0.5;
    for (int i = 0; i < 9; ++i) {
        const REAL *fi = f + (size_t)i * n, *fs = f + (size_t)ORACLE_OPP[i] * n;
        const REAL *ei = feq + (size_t)i * n;
        /* slot 0 keeps feq_0; slots 1..8 hold the opposite population */
        const REAL *es = (i == 0) ? ei : fs;
        REAL *dst = out + (size_t)i * n;
        for (size_t c = 0; c < n; ++c) {
            REAL f_p = fi[c] + fs[c], f_m = fi[c] - fs[c];
            REAL e_p = ei[c] + es[c], e_m = ei[c] - es[c];
            REAL omega = ((f_p - e_p) * omega_p + (f_m - e_m) * omega_m) * half;
            dst[c] = fi[c] + omega;
        }
    }
}

/* Regularized::evaluate src/lbm.rs:606-661 (never calls underlying.evaluate) */
static void FN(collide_regularized)(const REAL *f, const REAL *feq, size_t n,
                                    REAL dx, REAL dt, REAL *out)
{
    FN(consts_t) k; FN(make_consts)(&k, dx, dt);
    REAL *sxx = (REAL *)calloc(n, sizeof(REAL)), *sxy = (REAL *)calloc(n, sizeof(REAL));
    REAL *syx = (REAL *)calloc(n, sizeof(REAL)), *syy = (REAL *)calloc(n, sizeof(REAL));
    for (int i = 0; i < 9; ++i) {                                     /* :625-632 */
        const REAL *fi = f + (size_t)i * n, *ei = feq + (size_t)i * n;
        const REAL cxx = k.cx[i] * k.cx[i], cxy = k.cx[i] * k.cy[i];
        const REAL cyx = k.cy[i] * k.cx[i], cyy = k.cy[i] * k.cy[i];
        for (size_t c = 0; c < n; ++c) {
            REAL fneq = fi[c] - ei[c];                                /* :162-173 */
            sxx[c] = sxx[c] + fneq * cxx;
            sxy[c] = sxy[c] + fneq * cxy;
            syx[c] = syx[c] + fneq * cyx;
            syy[c] = syy[c] + fneq * cyy;
        }
    }
    for (int i = 0; i < 9; ++i) {                                     /* :638-658 */
        const REAL qxx = k.cx[i] * k.cx[i] - k.cs2, qxy = k.cx[i] * k.cy[i];
        const REAL qyx = k.cy[i] * k.cx[i],         qyy = k.cy[i] * k.cy[i] - k.cs2;
        const REAL sf = k.w[i] / ((REAL)2.0 * k.cs4);
        const REAL axx = qxx * sf, axy = qxy * sf, ayx = qyx * sf, ayy = qyy * sf;
        const REAL *ei = feq + (size_t)i * n;
        REAL *dst = out + (size_t)i * n;
        for (size_t c = 0; c < n; ++c) {
            REAL reg = ei[c];
            reg = reg + sxx[c] * axx;
            reg = reg + sxy[c] * axy;
            reg = reg + syx[c] * ayx;
            reg = reg + syy[c] * ayy;
            dst[c] = reg;
        }
    }
    free(sxx); free(sxy); free(syx); free(syy);
}

/* KBC::evaluate src/lbm.rs:468-585 (entropic; the DEBUG residual print at
 * :575-582 has no effect on the result and is not restated)                    */
static void FN(collide_kbc)(const REAL *f, const REAL *feq, size_t n,
                            REAL dx, REAL dt, REAL visc, REAL *out)
{
    REAL *rho = (REAL *)malloc(n * sizeof(REAL));
    REAL *u = (REAL *)malloc(n * sizeof(REAL)), *v = (REAL *)malloc(n * sizeof(REAL));
    REAL *ds = (REAL *)malloc(9 * n * sizeof(REAL)), *dh = (REAL *)malloc(9 * n * sizeof(REAL));
    FN(lbm_oracle_density)(f, n, rho);
    FN(lbm_oracle_velocity)(f, n, u, v);
    const REAL dx2 = dx * dx;
    const REAL cs = FN(cs)(dx, dt);
    const REAL beta = (REAL)1.0 / (((REAL)2.0 * visc / (cs * cs)) + (REAL)1.0);   /* :547-550 */
    for (size_t c = 0; c < n; ++c) {
        REAL uv = u[c] * v[c], u2 = u[c] * u[c], v2 = v[c] * v[c];                 /* :483-485 */
        REAL temp = (REAL)0.0;
        for (int i = 0; i < 9; ++i) temp = temp + f[(size_t)i * n + c] * dx2;      /* :487-490 */
        REAL pi_t = temp - uv;                                                     /* :492 */
        REAL n_t  = v2 - u2;                                                       /* :493 */
        REAL uv8 = uv * (REAL)8.0;
        REAL s0  = ((uv8 * pi_t + n_t * n_t) * rho[c]) * (REAL)0.5;               /* :496-501 */
        REAL s13 = (((((u[c] * dx - n_t) + (REAL)1.0) * n_t)
                     - (v[c] * (dx * (REAL)4.0) + uv8) * pi_t) * rho[c]) * (REAL)0.25;   /* :502-507 */
        REAL s24 = (((((v[c] * (-dx) - n_t) + (REAL)-1.0) * n_t)
                     - (u[c] * (dx * (REAL)4.0) + uv8) * pi_t) * rho[c]) * (REAL)0.25;   /* :508-513 */
        REAL s58 = (((((uv8 + u[c] * ((REAL)4.0 * dx)) + ((REAL)2.0 * dx * dx)) * pi_t)
                     + (n_t - (v[c] - u[c]) * dx) * n_t) * rho[c]) * (REAL)0.125;        /* :514-519 */
        REAL s[9] = {s0, s13, s24, s13, s24, s58, s58, s58, s58};                  /* :521-532 */
        REAL num = (REAL)0.0, den = (REAL)0.0;
        for (int i = 0; i < 9; ++i) {
            size_t q = (size_t)i * n + c;
            ds[q] = s[i];
            dh[q] = (f[q] - feq[q]) - s[i];                                        /* :541 */
            num = num + (s[i] * dh[q]) / feq[q];                                   /* :557 */
            den = den + (dh[q] * dh[q]) / feq[q];                                  /* :558 */
        }
        REAL gamma = ((((num / den) * ((REAL)2.0 - (REAL)1.0 / beta))
                       + ((REAL)-1.0 / beta))) * (REAL)-1.0;                       /* :560-564 */
        for (int i = 0; i < 9; ++i) {
            size_t q = (size_t)i * n + c;
            REAL omega = ds[q] * ((REAL)2.0 * -beta) + (dh[q] * gamma) * (-beta);  /* :569-571 */
            out[q] = f[q] + omega;                                                 /* :572 */
        }
    }
    free(rho); free(u); free(v); free(ds); free(dh);
}

/* ---- State::step, reference-structured (src/lbm.rs:694-714) ----------------
 * Three full passes with full-array temporaries, single thread.               */
void FN(lbm_oracle_step_ref)(REAL *f, const uint8_t *solid, int w, int h, int edge,
                             REAL dx, REAL dt, const lbm_oracle_collision_t *col, int nsteps)
{
    size_t n = (size_t)w * h;
    REAL *a = (REAL *)malloc(9 * n * sizeof(REAL));
    REAL *b = (REAL *)malloc(9 * n * sizeof(REAL));
    REAL *feq = (REAL *)malloc(9 * n * sizeof(REAL));
    for (int s = 0; s < nsteps; ++s) {
        FN(stream)(f, a, w, h, edge);                     /* :697 */
        FN(bounce_back)(a, b, solid, n);                  /* :703 */
        FN(lattice_equilibrium)(b, n, dx, dt, feq);       /* :735 */
        switch (col->kind) {                              /* :733 (dyn dispatch) */
        case ORACLE_COLLISION_BGK:
            FN(collide_bgk)(b, feq, n, dt, (REAL)col->tau, f); break;
        case ORACLE_COLLISION_TRT:
            FN(collide_trt)(b, feq, n, dt, (REAL)col->tau_plus, (REAL)col->tau_minus, f); break;
        case ORACLE_COLLISION_REGULARIZED:
            FN(collide_regularized)(b, feq, n, dx, dt, f); break;
        case ORACLE_COLLISION_KBC:
            FN(collide_kbc)(b, feq, n, dx, dt, (REAL)col->viscosity, f); break;
        }
    }
    free(a); free(b); free(feq);
}

/* ---- fused one-pass-per-step variant (BGK only), OpenMP over rows ----------
 * Same per-cell arithmetic in the same order as the three-pass version, so the
 * two are bit-identical (tests/test_oracle.py); this is the all-cores CPU
 * baseline that bench.py times, so it is written the way one would write the
 * CPU code for speed: row pointers hoisted, branch-free interior loop that the
 * compiler vectorises (still one rounded IEEE operation per reference operation:
 * -ffp-contract=off).  With ghost==1 the planes carry one ghost row above and
 * below (rows -1 and h at plane rows 0 and h+1): a caller can emulate a y-slab. */
static inline void FN(cell_bgk)(REAL *gq, int is_solid, const FN(consts_t) *k, REAL factor)
{
    /* State::bounce_back: g_i <- solid ? g_opp(i) : g_i */
    const REAL g1 = is_solid ? gq[3] : gq[1], g3 = is_solid ? gq[1] : gq[3];
    const REAL g2 = is_solid ? gq[4] : gq[2], g4 = is_solid ? gq[2] : gq[4];
    const REAL g5 = is_solid ? gq[7] : gq[5], g7 = is_solid ? gq[5] : gq[7];
    const REAL g6 = is_solid ? gq[8] : gq[6], g8 = is_solid ? gq[6] : gq[8];
    gq[1] = g1; gq[2] = g2; gq[3] = g3; gq[4] = g4; gq[5] = g5; gq[6] = g6; gq[7] = g7; gq[8] = g8;
    REAL rho = (REAL)0.0, mx = (REAL)0.0, my = (REAL)0.0;
    for (int i = 0; i < 9; ++i) rho = rho + gq[i];
    for (int i = 0; i < 9; ++i) { mx = mx + gq[i] * k->cx[i]; my = my + gq[i] * k->cy[i]; }
    const REAL r = (REAL)1.0 / rho;
    const REAL vx = r * mx, vy = r * my;
    const REAL v2 = vx * vx + vy * vy;
    for (int i = 0; i < 9; ++i) {
        const REAL vc = vx * k->cx[i] + vy * k->cy[i];
        const REAL vc2 = vc * vc;
        const REAL sum = (((REAL)1.0 + vc * k->k1) + vc2 * k->k2) + v2 * k->k3;
        const REAL fe = (rho * k->w[i]) * sum;
        gq[i] = gq[i] + (gq[i] - fe) * factor;
    }
}

void FN(lbm_oracle_step_fused)(const REAL *src, REAL *dst, const uint8_t *solid,
                               int w, int h, int edge, int ghost,
                               REAL dx, REAL dt, REAL tau)
{
    FN(consts_t) k; FN(make_consts)(&k, dx, dt);
    const REAL factor = -dt / tau;
    const int g = ghost ? 1 : 0;
    const size_t plane = (size_t)w * (h + 2 * g);
    REAL *zero_row = (REAL *)calloc((size_t)w, sizeof(REAL));
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        const REAL *row[9];
        for (int i = 0; i < 9; ++i) {
            int sy = y - ORACLE_EY[i];
            if (!ghost && (sy < 0 || sy >= h)) {
                if (edge == ORACLE_EDGE_PERIODIC) sy = (sy + h) % h;
                else { row[i] = zero_row; continue; }
            }
            row[i] = src + (size_t)i * plane + (size_t)(sy + g) * w;
        }
        const uint8_t *srow = solid ? solid + (size_t)y * w : NULL;
        REAL *out = dst + (size_t)(y + g) * w;
        /* the two edge columns (and everything, for w < 3): generic wrap / zero-fill */
        for (int pass = 0; pass < 2; ++pass) {
            const int x = pass == 0 ? 0 : w - 1;
            if (pass == 1 && w == 1) break;
            REAL gq[9];
            for (int i = 0; i < 9; ++i) {
                int sx = x - ORACLE_EX[i];
                int ok = 1;
                if (sx < 0 || sx >= w) { if (edge == ORACLE_EDGE_PERIODIC) sx = (sx + w) % w; else ok = 0; }
                gq[i] = ok ? row[i][sx] : (REAL)0.0;
            }
            FN(cell_bgk)(gq, srow && srow[x], &k, factor);
            for (int i = 0; i < 9; ++i) out[(size_t)i * plane + x] = gq[i];
        }
        /* interior columns: every source is in range */
#pragma omp simd
        for (int x = 1; x < w - 1; ++x) {
            REAL gq[9];
            gq[0] = row[0][x];     gq[1] = row[1][x];     gq[2] = row[2][x - 1];
            gq[3] = row[3][x];     gq[4] = row[4][x + 1]; gq[5] = row[5][x - 1];
            gq[6] = row[6][x - 1]; gq[7] = row[7][x + 1]; gq[8] = row[8][x + 1];
            FN(cell_bgk)(gq, srow ? srow[x] : 0, &k, factor);
            out[x] = gq[0];
            out[plane + x] = gq[1];     out[2 * plane + x] = gq[2]; out[3 * plane + x] = gq[3];
            out[4 * plane + x] = gq[4]; out[5 * plane + x] = gq[5]; out[6 * plane + x] = gq[6];
            out[7 * plane + x] = gq[7]; out[8 * plane + x] = gq[8];
        }
    }
    free(zero_row);
}
