// kernels.cu — sm_100a kernels of the D2Q9 path.
//
//  step_vec_kernel     the hot path: ONE pass per time step that pull-streams the
//                      nine populations (State::stream, src/lbm.rs:716-729),
//                      reverses them on solid cells (State::bounce_back, :741-751)
//                      and relaxes them (State::collide + BGK, :731-739, :349-364),
//                      with 128-bit loads/stores over the SoA layout.
//  step_scalar_kernel  the same update, one cell per thread, for widths that are
//                      not a multiple of the vector width.
//  readout / mass / unstable / init kernels for the macroscopic surface of
//                      lbm.rs (:117-160, :779-818, :43-71).
//
// HBM-bound integer-free streaming work: no tensor cores, no shared-memory
// tiling (every population value is read once and written once per step).
#include "kernels.cuh"

#include <cstdlib>

namespace chemsim {

namespace {

template <typename T> struct VecOf;
template <> struct VecOf<float>  { using type = float4;  static constexpr int N = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int N = 2; };

#ifdef CHEMSIM_LOAD_NOALLOC
#define CHEMSIM_LD_HINT ".L1::no_allocate"
#else
#define CHEMSIM_LD_HINT ""
#endif

// Predicated, branch-free global loads.  NC=true takes the read-only path: the
// source buffer is never written by the kernel that reads it (A-B buffering), so
// .nc is legal.  NC=false (coherent) is used by the P2P face kernel, whose ghost
// rows are written by the neighbouring GPU while the kernel may already be resident.
// Predication instead of `if` keeps every load of a thread in ONE straight-line
// batch: all of them are in flight before the first use.  The asm statements carry no
// memory dependence of their own: what keeps them behind griddepcontrol.wait and the halo
// flag wait is a data dependence — every address is derived from an "order token" (always
// 0, but opaque to the compiler) that those waits produce (see order_after_*).
#define CHEMSIM_LDG_BODY(NCSTR)                                                                       \
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"                                             \
        "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"                    \
        "@q ld.global" NCSTR ".v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"                                   \
        : "=&f"(v[0]), "=&f"(v[1]), "=&f"(v[2]), "=&f"(v[3]) : "l"(p), "r"((int)pred))
template <bool NC>
__device__ __forceinline__ void ldg_vec(const float *p, bool pred, float (&v)[4])
{
    if (NC) CHEMSIM_LDG_BODY(".nc" CHEMSIM_LD_HINT); else CHEMSIM_LDG_BODY("");
}
#undef CHEMSIM_LDG_BODY
#define CHEMSIM_LDG_BODY(NCSTR)                                                                       \
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t"                                             \
        "mov.b64 %0, 0;\n\tmov.b64 %1, 0;\n\t"                                                        \
        "@q ld.global" NCSTR ".v2.f64 {%0, %1}, [%2];\n\t}"                                           \
        : "=&d"(v[0]), "=&d"(v[1]) : "l"(p), "r"((int)pred))
template <bool NC>
__device__ __forceinline__ void ldg_vec(const double *p, bool pred, double (&v)[2])
{
    if (NC) CHEMSIM_LDG_BODY(".nc" CHEMSIM_LD_HINT); else CHEMSIM_LDG_BODY("");
}
#undef CHEMSIM_LDG_BODY
template <bool NC>
__device__ __forceinline__ float ldg_one(const float *p, bool pred)
{
    float v;
    if (NC) asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.b32 %0, 0;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}"
                : "=&f"(v) : "l"(p), "r"((int)pred));
    else    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.b32 %0, 0;\n\t@q ld.global.f32 %0, [%1];\n\t}"
                : "=&f"(v) : "l"(p), "r"((int)pred));
    return v;
}
template <bool NC>
__device__ __forceinline__ double ldg_one(const double *p, bool pred)
{
    double v;
    if (NC) asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.b64 %0, 0;\n\t@q ld.global.nc.f64 %0, [%1];\n\t}"
                : "=&d"(v) : "l"(p), "r"((int)pred));
    else    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.b64 %0, 0;\n\t@q ld.global.f64 %0, [%1];\n\t}"
                : "=&d"(v) : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ void store_vec(float *p, const float (&v)[4])
{
#ifdef CHEMSIM_STORE_CS
    __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
#else
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
#endif
}
__device__ __forceinline__ void store_vec(double *p, const double (&v)[2])
{
#ifdef CHEMSIM_STORE_CS
    __stcs(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1]));
#else
    *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
#endif
}
// V mask bytes as one 32-/16-bit word (0 when !pred)
__device__ __forceinline__ unsigned ldg_mask(const uint8_t *p, bool pred, const float *)
{
    unsigned v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.b32 %0, 0;\n\t@q ld.global.nc.u32 %0, [%1];\n\t}"
        : "=&r"(v) : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ unsigned ldg_mask(const uint8_t *p, bool pred, const double *)
{
    unsigned short v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.b16 %0, 0;\n\t@q ld.global.nc.u16 %0, [%1];\n\t}"
        : "=&h"(v) : "l"(p), "r"((int)pred));
    return v;
}

// Build-time tunables (defaults are the measured best; tools/variants.py sweeps them).
#ifndef CHEMSIM_STEP_THREADS
#define CHEMSIM_STEP_THREADS 256
#endif
#ifndef CHEMSIM_STEP_MIN_BLOCKS
#define CHEMSIM_STEP_MIN_BLOCKS 4   // <= 64 registers: 4 x 256 threads per SM (ptxas otherwise takes 88 for f64)
#endif
constexpr int STEP_THREADS = CHEMSIM_STEP_THREADS;
// resident blocks per SM the step kernels are compiled for: BGK fits 64 registers in both
// precisions; the f64 TRT / Regularized bodies need ~80 (3 blocks), KBC is left unconstrained
template <typename T, int COL>
constexpr int step_min_blocks()
{
    return COL == COL_KBC ? 1 : (COL != COL_BGK && sizeof(T) == 8 ? 3 : CHEMSIM_STEP_MIN_BLOCKS);
}

// ---- the fused step, vector form ---------------------------------------------
// Thread (tx, ty) of block (bx, by) updates the V = 16/sizeof(T) cells
// x0 … x0+V−1 of row y.  blockDim.x is a multiple of 32, so a warp always lies
// inside one row and the two neighbouring lanes hold the neighbouring vectors:
// populations that stream along x (dx = ±1) are assembled from the thread's own
// aligned vector plus ONE element shuffled in from the adjacent lane; only the
// first/last lane of a warp (or of the row) issues an extra scalar load, which
// also implements the x edge (wrap or zero-fill).
// The update of V cells of row y by one thread (see the kernel comment above).
// P2P=true additionally stores the populations that leave the slab through this face
// row straight into the neighbouring GPU's ghost row (peer-mapped memory, NVLink).
// griddepcontrol.wait as an opaque producer of 0: adding the result to a base pointer
// orders every load derived from it after the wait.
__device__ __forceinline__ int order_after_grid_dependency()
{
    int tok;
    asm volatile("griddepcontrol.wait;\n\tmov.u32 %0, 0;" : "=r"(tok) : : "memory");
    return tok;
}

template <typename T, bool PERIODIC_X, bool HAS_MASK, int COL, bool P2P>
__device__ __forceinline__ void step_vec_body(const StepArgs<T> &a, const int y, const int xv, const int lane,
                                              const int halo_tok)
{
    constexpr int V = VecOf<T>::N;
    constexpr bool NC = !P2P;
    const int nvec = a.W / V;
    if (xv - lane >= nvec) return;                   // whole warp beyond the row
    const bool active = xv < nvec;
    const int x0 = xv * V;
    // Programmatic dependent launch: the blocks of this step may already be resident while
    // the previous kernel in the stream drains; nothing is read before it has completed and
    // flushed (a no-op when the kernel was not launched as a dependent).  `tok` is 0.
    const int tok = order_after_grid_dependency() + halo_tok;
    const T *src = a.src + tok;
    const uint8_t *mask = a.mask + tok, *mask_flags = a.mask_flags + tok;
    // any solid cell in the 32*V cells of this warp?  (one or two 64-cell segments)
    unsigned seg_flags = 0;
    if (HAS_MASK) {
        const uint8_t *fl = mask_flags + (size_t)y * a.flag_pitch + ((xv - lane) * V) / MASK_SEGMENT;
        seg_flags = (V == 4) ? *reinterpret_cast<const unsigned short *>(fl) : *fl;
    }
    // which lanes must fetch the element their neighbour lane cannot supply
    const bool first = xv == 0, last = xv == nvec - 1;
    const bool need_left  = active && (lane == 0 || first)  && (PERIODIC_X || !first);
    const bool need_right = active && (lane == 31 || last) && (PERIODIC_X || !last);
    const int left_x  = first ? a.W - 1 : x0 - 1;    // wrap (periodic) or the previous warp's last element
    const int right_x = last ? 0 : x0 + V;

    // ---- phase 1: every load of this thread, back to back ----------------------
    T v[Q][V];      // the aligned vector of each population's source row
    T e[Q];         // the one extra element for populations that stream along x
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        int sy = y - ey_of(q);
        if (a.wrap_y) { if (sy < 0) sy = a.H - 1; else if (sy >= a.H) sy = 0; }
        const T *row = src + (size_t)q * a.plane + (size_t)(sy + 1) * a.pitch;
        ldg_vec<NC>(row + x0, active, v[q]);
        if (ex_of(q) == 1)       e[q] = ldg_one<NC>(row + left_x, need_left);
        else if (ex_of(q) == -1) e[q] = ldg_one<NC>(row + right_x, need_right);
        else                     e[q] = T(0);
    }
    unsigned maskw = 0;
    if (HAS_MASK && seg_flags != 0)                  // warp-uniform
        maskw = ldg_mask(mask + (size_t)y * a.mask_pitch + x0, active, (const T *)nullptr);

    // ---- phase 2: shift the x-streaming populations by one element -------------
    T g[Q][V];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        if (ex_of(q) == 0) {
#pragma unroll
            for (int j = 0; j < V; ++j) g[q][j] = v[q][j];
        } else if (ex_of(q) == 1) {                  // value at x comes from x−1
            const T nb = __shfl_up_sync(0xffffffffu, v[q][V - 1], 1);
            g[q][0] = (lane == 0 || first) ? e[q] : nb;
#pragma unroll
            for (int j = 1; j < V; ++j) g[q][j] = v[q][j - 1];
        } else {                                     // value at x comes from x+1
            const T nb = __shfl_down_sync(0xffffffffu, v[q][0], 1);
#pragma unroll
            for (int j = 0; j < V - 1; ++j) g[q][j] = v[q][j + 1];
            g[q][V - 1] = (lane == 31 || last) ? e[q] : nb;
        }
    }
    if (!active) return;

    // ---- phase 3: bounce-back + collide, cell by cell ---------------------------
#pragma unroll
    for (int j = 0; j < V; ++j) {
        T c[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) c[q] = g[q][j];
        if (HAS_MASK) bounce_back(c, ((maskw >> (8 * j)) & 0xffu) != 0);
        collide<COL>(c, a.k);
#pragma unroll
        for (int q = 0; q < Q; ++q) g[q][j] = c[q];
    }

    // ---- phase 4: nine aligned vector stores ------------------------------------
    T *out = a.dst + (size_t)(y + 1) * a.pitch + x0;
#pragma unroll
    for (int q = 0; q < Q; ++q) store_vec(out + (size_t)q * a.plane, g[q]);

    // ---- phase 5 (P2P face rows): the halo, written where the neighbour reads it --
    if (P2P) {
        const HaloP2P &p = a.halo;
        if (y == a.H - 1 && p.down_dst) {            // dy=+1 movers -> lower neighbour's ghost row −1 (plane row 0)
            T *peer = (T *)p.down_dst + x0;
#pragma unroll
            for (int q = 0; q < Q; ++q) if (ey_of(q) == 1) store_vec(peer + (size_t)q * p.down_plane, g[q]);
        }
        if (y == 0 && p.up_dst) {                    // dy=−1 movers -> upper neighbour's ghost row H_up
            T *peer = (T *)p.up_dst + (size_t)p.up_ghost_row * a.pitch + x0;
#pragma unroll
            for (int q = 0; q < Q; ++q) if (ey_of(q) == -1) store_vec(peer + (size_t)q * p.up_plane, g[q]);
        }
    }
}

template <typename T, bool PERIODIC_X, bool HAS_MASK, int COL, bool MULTIROW>
__global__ void __launch_bounds__(STEP_THREADS, step_min_blocks<T, COL>())
step_vec_kernel(const __grid_constant__ StepArgs<T> a)
{
    // MULTIROW=false: one row per block (blockDim.y == 1), so the row index and all
    // nine source-row addresses are block-uniform and live in uniform registers.
    // let the next step's kernel start filling SM slots as soon as every block of this
    // one has been scheduled (its blocks then wait in griddepcontrol.wait)
    asm volatile("griddepcontrol.launch_dependents;");
    // 1-D grid, x-chunk fastest: blocks that are scheduled together work on neighbouring
    // chunks of the same rows, so the 18 streams advance through DRAM pages in order
    // (measured +4 % over row-fastest block order).
    const int rg = blockIdx.x / a.xchunks, xc = blockIdx.x - rg * a.xchunks;
    const int yi = MULTIROW ? rg * blockDim.y + threadIdx.y : rg;
    if (yi >= a.y_count) return;                     // warp-uniform
    step_vec_body<T, PERIODIC_X, HAS_MASK, COL, false>(a, a.y_begin + yi * a.y_stride,
                                                       xc * blockDim.x + threadIdx.x, threadIdx.x & 31, 0);
}

// ---- fused face update + halo exchange over peer memory ------------------------
// One launch updates the two face rows {0, H−1} of a slab and delivers the
// populations that cross each face into the neighbouring GPUs' ghost rows with plain
// stores through NVLink-mapped pointers (cudaIpc): compute and exchange are ONE kernel,
// there is no pack buffer and no separate communication kernel.
// Flow control is a step counter per face in each GPU's memory:
//   wait   : ghost rows of step t are valid once the neighbour published flag >= t
//   signal : after every block has stored (and fenced) its rows, the last block to
//            finish publishes t+1 into both neighbours' flags
// A-B buffering makes one step of slack enough: a neighbour that is one step ahead
// writes into the buffer this GPU is not reading.
__device__ __forceinline__ void wait_flag(const unsigned *flag, unsigned want, int *error)
{
    if (!flag) return;
    unsigned spins = 0;
    while ((int)(*reinterpret_cast<const volatile unsigned *>(flag) - want) < 0) {
        __nanosleep(128);
        if (++spins > (1u << 27)) { atomicExch(error, 1); break; }   // ~20 s: report instead of hanging
    }
    __threadfence_system();
}

// One thread waits for both neighbours' step flags, the block follows through a barrier.
// Returns 0 through shared memory: an order token for the ghost-row loads (see ldg_*).
__device__ __forceinline__ int order_after_halo_flags(const HaloP2P &p)
{
    __shared__ int token;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        wait_flag(p.wait_up, p.step, p.error);
        wait_flag(p.wait_down, p.step, p.error);
        token = 0;
    }
    __syncthreads();
    return *reinterpret_cast<volatile int *>(&token);
}

template <typename T, bool PERIODIC_X, bool HAS_MASK, int COL>
__global__ void __launch_bounds__(STEP_THREADS, 1)
step_face_p2p_kernel(const __grid_constant__ StepArgs<T> a)
{
    const HaloP2P &p = a.halo;
    const int halo_tok = order_after_halo_flags(p);
    const int yi = blockIdx.x * blockDim.y + threadIdx.y;
    if (yi < a.y_count)
        step_vec_body<T, PERIODIC_X, HAS_MASK, COL, true>(a, a.y_begin + yi * a.y_stride,
                                                          blockIdx.y * blockDim.x + threadIdx.x, threadIdx.x & 31,
                                                          halo_tok);
    __threadfence_system();                          // my stores (local and peer) are visible system-wide ...
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const unsigned total = gridDim.x * gridDim.y;
        if (atomicAdd(p.done, 1u) == total - 1) {    // ... before the last block publishes the step
            *p.done = 0;
            __threadfence_system();
            if (p.signal_down) *reinterpret_cast<volatile unsigned *>(p.signal_down) = p.step + 1;
            if (p.signal_up)   *reinterpret_cast<volatile unsigned *>(p.signal_up) = p.step + 1;
        }
    }
}

// ---- the fused step, one cell per thread (any width) -------------------------
template <typename T, int COL>
__global__ void __launch_bounds__(STEP_THREADS)
step_scalar_kernel(const __grid_constant__ StepArgs<T> a)
{
    const int x = blockIdx.y * blockDim.x + threadIdx.x;
    const int yi = blockIdx.x * blockDim.y + threadIdx.y;
    if (yi >= a.y_count || x >= a.W) return;
    const int y = a.y_begin + yi * a.y_stride;
    T c[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        int sy = y - ey_of(q);
        if (a.wrap_y) { if (sy < 0) sy = a.H - 1; else if (sy >= a.H) sy = 0; }
        int sx = x - ex_of(q);
        bool inside = true;
        if (sx < 0)         { if (a.periodic_x) sx = a.W - 1; else inside = false; }
        else if (sx >= a.W) { if (a.periodic_x) sx = 0;       else inside = false; }
        c[q] = inside ? a.src[(size_t)q * a.plane + (size_t)(sy + 1) * a.pitch + sx] : T(0);
    }
    if (a.has_mask) bounce_back(c, a.mask[(size_t)y * a.mask_pitch + x] != 0);
    collide<COL>(c, a.k);
#pragma unroll
    for (int q = 0; q < Q; ++q) a.dst[(size_t)q * a.plane + (size_t)(y + 1) * a.pitch + x] = c[q];
}

// ---- compute_equilibrium on the device (src/lbm.rs:43-71) --------------------
template <typename T>
__global__ void init_equilibrium_kernel(const T *rho, const T *vx, const T *vy, T *dst, size_t plane, int pitch,
                                        int W, int row_begin, int rows, const __grid_constant__ Consts<T> k)
{
    const int x = blockIdx.y * blockDim.x + threadIdx.x;
    const int yr = blockIdx.x;
    if (x >= W || yr >= rows) return;
    const int y = row_begin + yr;
    const size_t c = (size_t)yr * W + x;
    const T r = rho[c], ux = vx[c], uy = vy[c];
    const T v2 = add(mul(ux, ux), mul(uy, uy));
#pragma unroll
    for (int q = 0; q < Q; ++q)
        dst[(size_t)q * plane + (size_t)(y + 1) * pitch + x] = equilibrium_i(q, r, ux, uy, v2, k);
}

// ---- macroscopic readout (src/lbm.rs:117-173, :779-812) ----------------------
template <typename T>
__global__ void readout_kernel(const __grid_constant__ ReadoutArgs<T> a)
{
    const int x = blockIdx.y * blockDim.x + threadIdx.x;
    const int y = blockIdx.x;
    if (x >= a.W || y >= a.H) return;
    T g[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) g[q] = a.src[(size_t)q * a.plane + (size_t)(y + 1) * a.pitch + x];
    const size_t c = (size_t)y * a.W + x;
    switch (a.kind) {
    case READ_DENSITY:  a.out0[c] = density(g); break;
    case READ_PRESSURE: a.out0[c] = mul(density(g), a.k.cs2); break;       // :784-787
    case READ_MOMENTUM: { T mx, my; momentum(g, mx, my); a.out0[c] = mx; a.out1[c] = my; break; }
    case READ_VELOCITY: { const Moments<T> m = moments(g); a.out0[c] = m.vx; a.out1[c] = m.vy; break; }
    case READ_SPEED: {                                                      // :151-154
        const Moments<T> m = moments(g);
        a.out0[c] = root(add(mul(m.vx, m.vx), mul(m.vy, m.vy)));
        break;
    }
    case READ_EQUILIBRIUM:
    case READ_NON_EQUILIBRIUM: {                                            // :156-173
        const Moments<T> m = moments(g);
        const T v2 = add(mul(m.vx, m.vx), mul(m.vy, m.vy));
        T fe = T(0), f = T(0);
#pragma unroll
        for (int q = 0; q < Q; ++q)
            if (q == a.q) { fe = equilibrium_i(q, m.rho, m.vx, m.vy, v2, a.k); f = g[q]; }
        a.out0[c] = (a.kind == READ_EQUILIBRIUM) ? fe : sub(f, fe);
        break;
    }
    }
}

// ---- reductions: warp shuffle -> block -> per-block partial -> final block ----
constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 148 * 8;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double warp_part[RED_THREADS / 32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < RED_THREADS / 32 ? warp_part[threadIdx.x] : 0.0;
        r = warp_sum(r);
    }
    return r;   // valid in thread 0
}

// total mass: every population of every cell, accumulated in f64
// (Matrix::sum -> af::sum_all, src/matrix.rs:138-140)
template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
mass_partial_kernel(const T *src, size_t plane, int pitch, int W, int H, double *partials)
{
    double acc = 0.0;
    const size_t cells = (size_t)W * H;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(c / W), x = (int)(c % W);
        const T *p = src + (size_t)(y + 1) * pitch + x;
        double cell = 0.0;
#pragma unroll
        for (int q = 0; q < Q; ++q) cell += (double)p[(size_t)q * plane];
        acc += cell;
    }
    const double b = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = b;
}

__global__ void __launch_bounds__(RED_THREADS)
mass_final_kernel(const double *partials, int n, double *out)
{
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];   // fixed order: deterministic
    const double b = block_sum(acc);
    if (threadIdx.x == 0) *out = b;
}

// State::is_unstable (src/lbm.rs:815-818): any f_eq,0 < 0
template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
unstable_kernel(const T *src, size_t plane, int pitch, int W, int H, const __grid_constant__ Consts<T> k, int *flag)
{
    bool bad = false;
    const size_t cells = (size_t)W * H;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(c / W), x = (int)(c % W);
        T g[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) g[q] = src[(size_t)q * plane + (size_t)(y + 1) * pitch + x];
        const Moments<T> m = moments(g);
        const T v2 = add(mul(m.vx, m.vx), mul(m.vy, m.vy));
        bad |= equilibrium_i(0, m.rho, m.vx, m.vy, v2, k) < T(0);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// One thread per (row, 64-cell segment): flags[y][seg] = any solid cell in the segment.
// Mask rows are padded with zeros to a multiple of 128 bytes, so whole 16-byte words
// can be read.
__global__ void __launch_bounds__(RED_THREADS)
mask_flags_kernel(const uint8_t *mask, int mask_pitch, int W, int row_begin, int rows, uint8_t *flags, int flag_pitch,
                  int *any)
{
    const int segs = (W + MASK_SEGMENT - 1) / MASK_SEGMENT;
    const size_t total = (size_t)rows * segs;
    bool found = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int y = row_begin + (int)(i / segs), seg = (int)(i % segs);
        const uint4 *p = reinterpret_cast<const uint4 *>(mask + (size_t)y * mask_pitch + (size_t)seg * MASK_SEGMENT);
        unsigned acc = 0;
#pragma unroll
        for (int j = 0; j < MASK_SEGMENT / 16; ++j) { const uint4 t = p[j]; acc |= t.x | t.y | t.z | t.w; }
        flags[(size_t)y * flag_pitch + seg] = acc != 0 ? 1 : 0;
        found |= acc != 0;
    }
    if (__any_sync(0xffffffffu, found) && (threadIdx.x & 31) == 0) atomicOr(any, 1);
}

// ---- device-side render.rs -----------------------------------------------------
// The scalar a render mode normalises: render_scalar_field (src/render.rs:23-89) uses
// the field itself, render_vector_field (:91-178) uses mag = vx*vx + vy*vy.
template <typename T>
__device__ __forceinline__ float render_scalar(const T (&g)[Q], int mode, float &vx, float &vy)
{
    vx = 0.f; vy = 0.f;
    if (mode == RENDER_DENSITY) return (float)density(g);
    if (mode == RENDER_MOMENTUM) {
        T mx, my; momentum(g, mx, my);
        vx = (float)mx; vy = (float)my;
    } else {
        const Moments<T> m = moments(g);
        vx = (float)m.vx; vy = (float)m.vy;
    }
    const float mag = __fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy));
    return mode == RENDER_SPEED ? __fsqrt_rn(mag) : mag;
}

template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
render_stats_partial_kernel(const T *src, size_t plane, int pitch, int W, int H, int mode, double *partials)
{
    double s1 = 0.0, s2 = 0.0;
    const size_t cells = (size_t)W * H;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(c / W), x = (int)(c % W);
        T g[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) g[q] = src[(size_t)q * plane + (size_t)(y + 1) * pitch + x];
        float vx, vy;
        const double v = (double)render_scalar(g, mode, vx, vy);
        s1 += v; s2 += v * v;
    }
    const double b1 = block_sum(s1);
    __syncthreads();
    const double b2 = block_sum(s2);
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = b1; partials[2 * blockIdx.x + 1] = b2; }
}

__global__ void __launch_bounds__(RED_THREADS)
render_stats_final_kernel(const double *partials, int n, double cells, double *stats)
{
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { s1 += partials[2 * i]; s2 += partials[2 * i + 1]; }
    const double b1 = block_sum(s1);
    __syncthreads();
    const double b2 = block_sum(s2);
    if (threadIdx.x == 0) { stats[0] = b1; stats[1] = b2; }   // raw sums (all-reduced over the slabs when sharded)
}

// sums[0..1] = (sum, sum of squares) over `cells` cells -> stats[0..1] = (mean, population stdev)
__global__ void render_stats_finish_kernel(const double *sums, double cells, double *stats)
{
    const double mean = sums[0] / cells;
    double var = sums[1] / cells - mean * mean;        // population variance (af::stdev_all)
    if (var < 0.0) var = 0.0;
    stats[0] = mean;
    stats[1] = sqrt(var);
}

// af::hsv2rgb on one pixel (h, s, v in [0,1])
__device__ __forceinline__ void hsv2rgb(float h, float s, float v, float &r, float &g, float &b)
{
    const float h6 = h * 6.0f;
    const int m = (int)h6;
    const float f = h6 - (float)m;
    const float p = v * (1.0f - s), q = v * (1.0f - s * f), t = v * (1.0f - s * (1.0f - f));
    switch (m) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    case 5: r = v; g = p; b = q; break;
    default: r = v; g = t; b = p; break;              // h == 1 wraps to red
    }
}

__device__ __forceinline__ unsigned char to_u8(float c)   // (256.0 * c).round().min(255.0).max(0.0) as u8
{
    return (unsigned char)fmaxf(fminf(roundf(256.0f * c), 255.0f), 0.0f);
}

template <typename T>
__global__ void render_image_kernel(const T *src, size_t plane, int pitch, int W, int H, int mode,
                                    const double *stats, const uint8_t *mask, int mask_pitch, uchar4 *rgba)
{
    const int x = blockIdx.y * blockDim.x + threadIdx.x;
    const int y = blockIdx.x;
    if (x >= W || y >= H) return;
    T g[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) g[q] = src[(size_t)q * plane + (size_t)(y + 1) * pitch + x];
    float vx, vy;
    const float field = render_scalar(g, mode, vx, vy);
    const float avg = (float)stats[0], inv_std = 1.0f / (float)stats[1];      // `avg as f32`, `1.0 / std as f32`
    const float z = (field - avg) * inv_std;
    float val = 1.0f / (1.0f + expf(-z));                                     // Matrix::logistic = af::sigmoid
    float hue = 0.0f, sat = 1.0f;                                             // render_scalar_field :31-37
    if (mode == RENDER_VELOCITY || mode == RENDER_MOMENTUM) {                 // render_vector_field :116-128
        hue = (atan2f(vy, vx) + 3.14159274f) * (0.318309886f * 0.5f);
        sat = 0.8f;
    }
    hue = fminf(fmaxf(hue, 0.0f), 1.0f);                                      // .clamp(0.0, 1.0)
    val = fminf(fmaxf(val, 0.0f), 1.0f);
    float r, gg, b;
    hsv2rgb(hue, sat, val, r, gg, b);
    uchar4 px = make_uchar4(to_u8(r), to_u8(gg), to_u8(b), 255);
    if (mask && mask[(size_t)y * mask_pitch + x]) px = make_uchar4(0, 0, 255, 255);   // render_geometry :7-21
    rgba[(size_t)y * W + x] = px;
}

inline int check_launch()
{
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

inline int reduction_blocks(size_t cells)
{
    size_t b = (cells + RED_THREADS - 1) / RED_THREADS;
    if (b > (size_t)RED_MAX_BLOCKS) b = RED_MAX_BLOCKS;
    if (b < 1) b = 1;
    return (int)b;
}

template <typename T>
bool use_vec(const StepArgs<T> &a) { return a.W % VecOf<T>::N == 0; }

}  // namespace

template <typename T>
const char *step_kernel_name(const StepArgs<T> &a)
{
    if (!use_vec(a)) return sizeof(T) == 4 ? "step_scalar_kernel<float>" : "step_scalar_kernel<double>";
    return sizeof(T) == 4 ? "step_vec_kernel<float>" : "step_vec_kernel<double>";
}

// Back-to-back step kernels are launched as programmatic dependents of each other
// (CHEMSIM_LBM_PDL=0 in the environment restores plain stream order).
inline bool pdl_enabled()
{
    static const bool on = [] { const char *e = getenv("CHEMSIM_LBM_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

template <typename T>
void launch_chained(void (*kernel)(const StepArgs<T>), dim3 grid, dim3 block, cudaStream_t s, const StepArgs<T> &a)
{
    if (!pdl_enabled()) { kernel<<<grid, block, 0, s>>>(a); return; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, a);
}

template <typename T, int COL>
void launch_step_col(const StepArgs<T> &a_in, cudaStream_t s)
{
    const int rows = a_in.y_count;
    if (use_vec(a_in)) {
        constexpr int V = VecOf<T>::N;
        const int nvec = a_in.W / V;
        int bx = ((nvec + 31) / 32) * 32;
        if (bx > STEP_THREADS) bx = STEP_THREADS;
        int by = STEP_THREADS / bx;
        if (by > rows) by = rows;
        const dim3 block(bx, by);
        StepArgs<T> a = a_in;
        a.xchunks = (nvec + bx - 1) / bx;
        const dim3 grid((unsigned)a.xchunks * (unsigned)((rows + by - 1) / by));
#define CHEMSIM_LAUNCH_VEC(PX, HM)                                                                   \
        do {                                                                                         \
            if (by == 1) launch_chained(step_vec_kernel<T, PX, HM, COL, false>, grid, block, s, a);  \
            else         launch_chained(step_vec_kernel<T, PX, HM, COL, true>, grid, block, s, a);   \
        } while (0)
        if (a.periodic_x) {
            if (a.has_mask) CHEMSIM_LAUNCH_VEC(true, true); else CHEMSIM_LAUNCH_VEC(true, false);
        } else {
            if (a.has_mask) CHEMSIM_LAUNCH_VEC(false, true); else CHEMSIM_LAUNCH_VEC(false, false);
        }
#undef CHEMSIM_LAUNCH_VEC
    } else {
        int bx = ((a_in.W + 31) / 32) * 32;
        if (bx > STEP_THREADS) bx = STEP_THREADS;
        int by = STEP_THREADS / bx;
        if (by > rows) by = rows;
        const dim3 block(bx, by);
        const dim3 grid((rows + by - 1) / by, (a_in.W + bx - 1) / bx);
        step_scalar_kernel<T, COL><<<grid, block, 0, s>>>(a_in);
    }
}

// ---- the whole slab step + halo in ONE kernel (peer-memory mode) ------------------
// 1-D grid of x-chunks x H blocks, x-chunk fastest; row slot r = 0 -> row 0, 1 -> row H−1,
// r >= 2 -> row r−1, so the two face rows are dispatched first: they wait for the neighbours' step flags, update
// their rows, store the outgoing populations into the neighbours' ghost rows and publish
// the next step early, while the remaining blocks stream through the interior.  One
// launch per step and GPU, chained with programmatic dependent launch; no events, no
// communication kernel, no second stream.
template <typename T, bool PERIODIC_X, bool HAS_MASK, int COL>
__global__ void __launch_bounds__(STEP_THREADS, step_min_blocks<T, COL>())
step_slab_p2p_kernel(const __grid_constant__ StepArgs<T> a)
{
    asm volatile("griddepcontrol.launch_dependents;");
    const int r = blockIdx.x / a.xchunks, xc = blockIdx.x - r * a.xchunks;
    const int y = r == 0 ? 0 : (r == 1 ? a.H - 1 : r - 1);
    const int xv = xc * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    if (r >= 2) {                                    // interior row: reads no ghost row
        step_vec_body<T, PERIODIC_X, HAS_MASK, COL, false>(a, y, xv, lane, 0);
        return;
    }
    const HaloP2P &p = a.halo;
    const int halo_tok = order_after_halo_flags(p);
    step_vec_body<T, PERIODIC_X, HAS_MASK, COL, true>(a, y, xv, lane, halo_tok);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned nface = (a.H > 1 ? 2u : 1u) * (unsigned)a.xchunks;
        if (atomicAdd(p.done, 1u) == nface - 1) {
            *p.done = 0;
            __threadfence_system();
            if (p.signal_down) *reinterpret_cast<volatile unsigned *>(p.signal_down) = p.step + 1;
            if (p.signal_up)   *reinterpret_cast<volatile unsigned *>(p.signal_up) = p.step + 1;
        }
    }
}

template <typename T, int COL>
void launch_slab_p2p_col(const StepArgs<T> &a_in, cudaStream_t s)
{
    constexpr int V = VecOf<T>::N;
    const int nvec = a_in.W / V;
    const dim3 block(STEP_THREADS, 1);
    StepArgs<T> a = a_in;
    a.xchunks = (nvec + STEP_THREADS - 1) / STEP_THREADS;
    const dim3 grid((unsigned)a.xchunks * (unsigned)a.H);
    if (a.periodic_x) {
        if (a.has_mask) launch_chained(step_slab_p2p_kernel<T, true, true, COL>, grid, block, s, a);
        else            launch_chained(step_slab_p2p_kernel<T, true, false, COL>, grid, block, s, a);
    } else {
        if (a.has_mask) launch_chained(step_slab_p2p_kernel<T, false, true, COL>, grid, block, s, a);
        else            launch_chained(step_slab_p2p_kernel<T, false, false, COL>, grid, block, s, a);
    }
}

// one block per row chunk needs full 256-thread rows; narrower lattices use the
// two-stream face/interior path instead
template <typename T>
bool slab_p2p_supported(const StepArgs<T> &a)
{
    return use_vec(a) && a.W / VecOf<T>::N >= STEP_THREADS;
}

template <typename T>
int launch_slab_p2p(const StepArgs<T> &a, cudaStream_t s)
{
    if (!slab_p2p_supported(a)) return -(int)cudaErrorInvalidValue;
    switch (a.collision) {
    case COL_BGK:         launch_slab_p2p_col<T, COL_BGK>(a, s); break;
    case COL_TRT:         launch_slab_p2p_col<T, COL_TRT>(a, s); break;
    case COL_REGULARIZED: launch_slab_p2p_col<T, COL_REGULARIZED>(a, s); break;
    case COL_KBC:         launch_slab_p2p_col<T, COL_KBC>(a, s); break;
    default: return -(int)cudaErrorInvalidValue;
    }
    const int e = check_launch();
    return e ? e : 1;
}

template <typename T, int COL>
void launch_face_p2p_col(const StepArgs<T> &a, cudaStream_t s)
{
    constexpr int V = VecOf<T>::N;
    const int rows = a.y_count, nvec = a.W / V;
    int bx = ((nvec + 31) / 32) * 32;
    if (bx > STEP_THREADS) bx = STEP_THREADS;
    int by = STEP_THREADS / bx;
    if (by > rows) by = rows;
    const dim3 block(bx, by);
    const dim3 grid((rows + by - 1) / by, (nvec + bx - 1) / bx);
    if (a.periodic_x) {
        if (a.has_mask) step_face_p2p_kernel<T, true, true, COL><<<grid, block, 0, s>>>(a);
        else            step_face_p2p_kernel<T, true, false, COL><<<grid, block, 0, s>>>(a);
    } else {
        if (a.has_mask) step_face_p2p_kernel<T, false, true, COL><<<grid, block, 0, s>>>(a);
        else            step_face_p2p_kernel<T, false, false, COL><<<grid, block, 0, s>>>(a);
    }
}

template <typename T>
bool face_p2p_supported(const StepArgs<T> &a) { return use_vec(a); }

template <typename T>
int launch_face_p2p(const StepArgs<T> &a, cudaStream_t s)
{
    if (a.y_count <= 0 || !use_vec(a)) return -(int)cudaErrorInvalidValue;
    switch (a.collision) {
    case COL_BGK:         launch_face_p2p_col<T, COL_BGK>(a, s); break;
    case COL_TRT:         launch_face_p2p_col<T, COL_TRT>(a, s); break;
    case COL_REGULARIZED: launch_face_p2p_col<T, COL_REGULARIZED>(a, s); break;
    case COL_KBC:         launch_face_p2p_col<T, COL_KBC>(a, s); break;
    default: return -(int)cudaErrorInvalidValue;
    }
    const int e = check_launch();
    return e ? e : 1;
}

template <typename T>
int launch_step(const StepArgs<T> &a, cudaStream_t s)
{
    if (a.y_count <= 0) return 0;
    switch (a.collision) {
    case COL_BGK:         launch_step_col<T, COL_BGK>(a, s); break;
    case COL_TRT:         launch_step_col<T, COL_TRT>(a, s); break;
    case COL_REGULARIZED: launch_step_col<T, COL_REGULARIZED>(a, s); break;
    case COL_KBC:         launch_step_col<T, COL_KBC>(a, s); break;
    default: return -(int)cudaErrorInvalidValue;
    }
    const int e = check_launch();
    return e ? e : 1;
}

template <typename T>
int launch_init_equilibrium(const T *rho, const T *vx, const T *vy, T *dst, size_t plane, int pitch, int W,
                            int row_begin, int rows, const Consts<T> &k, cudaStream_t s)
{
    const dim3 block(256), grid(rows, (W + 255) / 256);
    init_equilibrium_kernel<T><<<grid, block, 0, s>>>(rho, vx, vy, dst, plane, pitch, W, row_begin, rows, k);
    const int e = check_launch();
    return e ? e : 1;
}

template <typename T>
int launch_readout(const ReadoutArgs<T> &a, cudaStream_t s)
{
    const dim3 block(256), grid(a.H, (a.W + 255) / 256);
    readout_kernel<T><<<grid, block, 0, s>>>(a);
    const int e = check_launch();
    return e ? e : 1;
}

int mass_partials_capacity() { return RED_MAX_BLOCKS; }

template <typename T>
int launch_total_mass(const T *src, size_t plane, int pitch, int W, int H, double *partials, double *out,
                      cudaStream_t s)
{
    const int blocks = reduction_blocks((size_t)W * H);
    mass_partial_kernel<T><<<blocks, RED_THREADS, 0, s>>>(src, plane, pitch, W, H, partials);
    mass_final_kernel<<<1, RED_THREADS, 0, s>>>(partials, blocks, out);
    const int e = check_launch();
    return e ? e : 2;
}

template <typename T>
int launch_is_unstable(const T *src, size_t plane, int pitch, int W, int H, const Consts<T> &k, int *flag,
                       cudaStream_t s)
{
    const int blocks = reduction_blocks((size_t)W * H);
    unstable_kernel<T><<<blocks, RED_THREADS, 0, s>>>(src, plane, pitch, W, H, k, flag);
    const int e = check_launch();
    return e ? e : 1;
}

template <typename T>
int launch_render_stats(const T *src, size_t plane, int pitch, int W, int H, int mode, double *partials,
                        double *stats, cudaStream_t s)
{
    int blocks = reduction_blocks((size_t)W * H);
    if (blocks > RED_MAX_BLOCKS / 2) blocks = RED_MAX_BLOCKS / 2;     // two doubles per block
    render_stats_partial_kernel<T><<<blocks, RED_THREADS, 0, s>>>(src, plane, pitch, W, H, mode, partials);
    render_stats_final_kernel<<<1, RED_THREADS, 0, s>>>(partials, blocks, (double)W * (double)H, stats);
    const int e = check_launch();
    return e ? e : 2;
}

int launch_render_stats_finish(const double *sums, double cells, double *stats, cudaStream_t s)
{
    render_stats_finish_kernel<<<1, 1, 0, s>>>(sums, cells, stats);
    const int e = check_launch();
    return e ? e : 1;
}

template <typename T>
int launch_render_image(const T *src, size_t plane, int pitch, int W, int H, int mode, const double *stats,
                        const uint8_t *mask, int mask_pitch, uchar4 *rgba, cudaStream_t s)
{
    const dim3 block(256), grid(H, (W + 255) / 256);
    render_image_kernel<T><<<grid, block, 0, s>>>(src, plane, pitch, W, H, mode, stats, mask, mask_pitch, rgba);
    const int e = check_launch();
    return e ? e : 1;
}

int launch_mask_flags(const uint8_t *mask, int mask_pitch, int W, int row_begin, int rows, uint8_t *flags,
                      int flag_pitch, int *any, cudaStream_t s)
{
    const int segs = (W + MASK_SEGMENT - 1) / MASK_SEGMENT;
    const int blocks = reduction_blocks((size_t)rows * segs);
    mask_flags_kernel<<<blocks, RED_THREADS, 0, s>>>(mask, mask_pitch, W, row_begin, rows, flags, flag_pitch, any);
    const int e = check_launch();
    return e ? e : 1;
}

#define CHEMSIM_INSTANTIATE(T)                                                                                       \
    template int launch_step<T>(const StepArgs<T> &, cudaStream_t);                                                  \
    template int launch_face_p2p<T>(const StepArgs<T> &, cudaStream_t);                                              \
    template bool face_p2p_supported<T>(const StepArgs<T> &);                                                        \
    template int launch_slab_p2p<T>(const StepArgs<T> &, cudaStream_t);                                              \
    template bool slab_p2p_supported<T>(const StepArgs<T> &);                                                        \
    template const char *step_kernel_name<T>(const StepArgs<T> &);                                                   \
    template int launch_init_equilibrium<T>(const T *, const T *, const T *, T *, size_t, int, int, int, int,        \
                                            const Consts<T> &, cudaStream_t);                                        \
    template int launch_readout<T>(const ReadoutArgs<T> &, cudaStream_t);                                            \
    template int launch_total_mass<T>(const T *, size_t, int, int, int, double *, double *, cudaStream_t);           \
    template int launch_is_unstable<T>(const T *, size_t, int, int, int, const Consts<T> &, int *, cudaStream_t);  \
    template int launch_render_stats<T>(const T *, size_t, int, int, int, int, double *, double *, cudaStream_t);    \
    template int launch_render_image<T>(const T *, size_t, int, int, int, int, const double *, const uint8_t *, int, \
                                        uchar4 *, cudaStream_t);

CHEMSIM_INSTANTIATE(float)
CHEMSIM_INSTANTIATE(double)

}  // namespace chemsim
