// step2_impl.cuh — TWO time steps per pass over HBM (temporal blocking of State::step).
//
// The single-step kernels are bound by 72 B of HBM traffic per cell and step.
// This kernel halves that: a block loads the populations around a TY x TX tile once, performs
// step n+1 for the tile plus a one-cell rim into SHARED memory, and step n+2 for the tile from
// there — the intermediate lattice never touches HBM.  Per cell and per step the arithmetic is the
// very same call sequence (pull-stream, bounce_back, collide<COL>) with the same individually
// rounded operations, and the intermediate is held in the lattice dtype exactly as the A-B buffer
// would hold it, so two passes of the single-step kernel and one pass of this one are bit-identical.
//
//   phase A  (TY+2) x (TX+2) cells ("ext" region): one cell per thread and iteration, scalar
//            coalesced loads from HBM/L2; results go to smem[q][row][col + shift_q].  The rim is
//            redundant work (+14 % cells for the 16 x 128 tile) whose loads hit L2 (the
//            neighbouring tiles read the same lines).
//   phase B  one warp per tile row, V = 16/sizeof(T) cells per lane: nine aligned 128-bit
//            shared-memory loads (the per-population column shift_q makes every shifted read
//            start on a 16-byte boundary: conflict-free), collide, nine 128-bit global stores.
//
// Edges: cells outside a zero-fill lattice hold 0 in the ext region (they are never computed:
// src/lbm.rs:716-729 drops what leaves the array); periodic edges wrap the coordinates.  On a
// y-slab the rows beyond the slab are the neighbours' cells, read from the TWO ghost rows.
#pragma once

#include "step_decl.cuh"

namespace chemsim {

namespace {

template <typename T>
struct Step2Tile {
    static constexpr int V = VecOf<T>::N;          // cells per 16 bytes
    static constexpr int TX = 32 * V;              // one warp covers a tile row in phase B
    static constexpr int TY = 16;
    static constexpr int NT = 32 * TY;             // threads per block: one warp per tile row
    static constexpr int EX = TX + 2, EY = TY + 2; // tile + one-cell rim
    static constexpr int SP = ((EX + V - 1 + V - 1) / V) * V;   // shared row pitch (room for the column shift)
    static constexpr size_t SMEM = (size_t)Q * EY * SP * sizeof(T);
};

// column shift of population q in shared memory: makes (x + 1 - ex_q + shift_q) a multiple of V
template <int V> __host__ __device__ constexpr int shift_of(int q) { return (((ex_of(q) - 1) % V) + V) % V; }

template <typename T, bool PERIODIC_X, bool HAS_MASK, int COL>
__global__ void __launch_bounds__(Step2Tile<T>::NT, 2)
step2_kernel(const __grid_constant__ StepArgs<T> a)
{
    using TL = Step2Tile<T>;
    constexpr int V = TL::V, TX = TL::TX, TY = TL::TY, NT = TL::NT, EX = TL::EX, EY = TL::EY, SP = TL::SP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *const sm = reinterpret_cast<T *>(smem_raw);   // [Q][EY][SP]

    asm volatile("griddepcontrol.launch_dependents;");
    const int tx0 = blockIdx.x * TX;
    const int ty0 = a.y_begin + (blockIdx.z * gridDim.y + blockIdx.y) * TY;
    if (ty0 >= a.y_begin + a.y_count) return;
    const int y_end = a.y_begin + a.y_count;         // this launch owns tile rows [y_begin, y_end)
    const int tid = threadIdx.x;
    const int tok = order_after_grid_dependency();
    const T *src = a.src + tok;
    const uint8_t *mask = a.mask + tok;

    // ---- phase A: step n+1 on the ext region -> shared memory -----------------------------------
    // interior tile: every source cell of the ext region lies inside this lattice (no wrap, no
    // zero-fill, no ghost row of a zero-fill edge): block-uniform fast path
    const bool interior = ty0 >= 2 && ty0 + TY + 2 <= a.H && tx0 >= 2 && tx0 + TX + 2 <= a.W;
    if (interior) {
        // two ext cells per iteration, all eighteen loads issued before the first collision: the
        // HBM/L2 latency of one cell is covered by the arithmetic of the other
        constexpr int DY = NT / EX, DX = NT % EX;    // idx += NT  <=>  (ey, ex) += (DY, DX) with a carry
        int ey0 = tid / EX, ex0 = tid - ey0 * EX;
        // opaque base: keeps "pointer + precomputed offset" as two integer instructions per load
        unsigned long long base = reinterpret_cast<unsigned long long>(src) +
                                  ((size_t)(ty0 - 1 + a.ghost) * a.pitch + (tx0 - 1)) * sizeof(T);
        asm volatile("" : "+l"(base));
#pragma unroll 1
        for (int idx = tid; idx < EY * EX; idx += 2 * NT) {
            int ey1 = ey0 + DY, ex1 = ex0 + DX;
            if (ex1 >= EX) { ex1 -= EX; ey1 += 1; }
            const bool two = idx + NT < EY * EX;
            const char *p0 = reinterpret_cast<const char *>(base) + ((size_t)ey0 * a.pitch + ex0) * sizeof(T);
            const char *p1 = reinterpret_cast<const char *>(base) + ((size_t)(two ? ey1 : ey0) * a.pitch + (two ? ex1 : ex0)) * sizeof(T);
            T c0[Q], c1[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                c0[q] = __ldg(reinterpret_cast<const T *>(p0 + (a.ld_off[q] - ex_of(q) * (long long)sizeof(T))));
                c1[q] = __ldg(reinterpret_cast<const T *>(p1 + (a.ld_off[q] - ex_of(q) * (long long)sizeof(T))));
            }
            bool s0 = false, s1 = false;
            if (HAS_MASK) {
                s0 = __ldg(mask + (size_t)(ty0 - 1 + ey0) * a.mask_pitch + (tx0 - 1 + ex0)) != 0;
                s1 = __ldg(mask + (size_t)(ty0 - 1 + (two ? ey1 : ey0)) * a.mask_pitch + (tx0 - 1 + (two ? ex1 : ex0))) != 0;
            }
            if (HAS_MASK) bounce_back(c0, s0);
            collide<COL>(c0, a.k);
            T *d0 = sm + ey0 * SP + ex0;
#pragma unroll
            for (int q = 0; q < Q; ++q) d0[q * (EY * SP) + shift_of<V>(q)] = c0[q];
            if (two) {
                if (HAS_MASK) bounce_back(c1, s1);
                collide<COL>(c1, a.k);
                T *d1 = sm + ey1 * SP + ex1;
#pragma unroll
                for (int q = 0; q < Q; ++q) d1[q * (EY * SP) + shift_of<V>(q)] = c1[q];
            }
            ey0 = ey1 + DY; ex0 = ex1 + DX;
            if (ex0 >= EX) { ex0 -= EX; ey0 += 1; }
        }
    } else {
#pragma unroll 1
        for (int idx = tid; idx < EY * EX; idx += NT) {
            const int ey = idx / EX, ex = idx - ey * EX;
            int gy = ty0 - 1 + ey, gx = tx0 - 1 + ex;
            if (gy > a.H || gx > a.W) continue;      // beyond the rim of a partial tile: never read
            T c[Q];
            // is the ext cell itself outside the (global) lattice?  zero-fill: it holds 0
            bool outside = false;
            if (PERIODIC_X) { if (gx < 0) gx = a.W - 1; else if (gx >= a.W) gx = 0; }
            else            outside = gx < 0 || gx >= a.W;
            if (a.wrap_y) { if (gy < 0) gy = a.H - 1; else if (gy >= a.H) gy = 0; }
            else if (!a.periodic_y) outside = outside || a.row0 + gy < 0 || a.row0 + gy >= a.Hglobal;
            if (outside) {
#pragma unroll
                for (int q = 0; q < Q; ++q) c[q] = T(0);
            } else {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    int sy = gy - ey_of(q), sx = gx - ex_of(q);
                    bool in = true;
                    if (a.wrap_y) { if (sy < 0) sy = a.H - 1; else if (sy >= a.H) sy = 0; }
                    if (PERIODIC_X) { if (sx < 0) sx = a.W - 1; else if (sx >= a.W) sx = 0; }
                    else            in = sx >= 0 && sx < a.W;
                    // rows -ghost .. H+ghost-1 exist: ghost rows hold the neighbour slab's cells, or 0 at a zero-fill edge
                    c[q] = in ? __ldg(src + (size_t)q * a.plane + (size_t)(sy + a.ghost) * a.pitch + sx) : T(0);
                }
                // the mask of a ghost-row cell belongs to the neighbour slab: a.mask has `ghost` halo rows too
                if (HAS_MASK) bounce_back(c, __ldg(mask + (ptrdiff_t)gy * a.mask_pitch + gx) != 0);
                collide<COL>(c, a.k);
            }
            T *s = sm + ey * SP + ex;
#pragma unroll
            for (int q = 0; q < Q; ++q) s[q * (EY * SP) + shift_of<V>(q)] = c[q];
        }
    }
    __syncthreads();

    // ---- phase B: step n+2 on the tile, from shared memory --------------------------------------
    const int r = tid >> 5, lane = tid & 31;
    const int gy = ty0 + r, gx0 = tx0 + lane * V;
    if (gy >= y_end || gx0 >= a.W) return;
    T g[Q][V];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        // ext cell (r + 1 - ey_q, x + 1 - ex_q) of population q sits at column x + 1 - ex_q + shift_q
        const T *s = sm + q * (EY * SP) + (r + 1 - ey_of(q)) * SP + lane * V + (1 - ex_of(q) + shift_of<V>(q));
        const typename VecOf<T>::type v = *reinterpret_cast<const typename VecOf<T>::type *>(s);
        if constexpr (V == 4) { g[q][0] = v.x; g[q][1] = v.y; g[q][2] = v.z; g[q][3] = v.w; }
        else                  { g[q][0] = v.x; g[q][1] = v.y; }
    }
    unsigned maskw = 0;
    if (HAS_MASK) maskw = ldg_mask(mask + (size_t)gy * a.mask_pitch + gx0, true, (const T *)nullptr);
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
    char *out = reinterpret_cast<char *>(a.dst) + ((size_t)(gy + a.ghost) * a.pitch + gx0) * sizeof(T);
#pragma unroll
    for (int q = 0; q < Q; ++q) store_vec(reinterpret_cast<T *>(out + a.st_off[q]), g[q]);
}

}  // namespace

// rows [y_begin, y_begin + y_count) advance by TWO steps; y_count need not be a multiple of the tile height
template <typename T, int COL>
void launch_step2_col(const StepArgs<T> &a, cudaStream_t s)
{
    using TL = Step2Tile<T>;
    const dim3 block(TL::NT);
    const dim3 grid = row_grid((a.W + TL::TX - 1) / TL::TX, (a.y_count + TL::TY - 1) / TL::TY);
#define CHEMSIM_LAUNCH_STEP2(PX, HM)                                                                           \
    do {                                                                                                       \
        static bool opted_in_[64] = {};            /* per device: > 48 KB of dynamic shared memory */          \
        int dev_ = 0;                                                                                          \
        cudaGetDevice(&dev_);                                                                                  \
        if (!opted_in_[dev_ & 63]) {                                                                           \
            cudaFuncSetAttribute(step2_kernel<T, PX, HM, COL>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                 (int)TL::SMEM);                                                               \
            opted_in_[dev_ & 63] = true;                                                                       \
        }                                                                                                      \
        launch_chained(step2_kernel<T, PX, HM, COL>, grid, block, s, a, TL::SMEM);                             \
    } while (0)
    if (a.periodic_x) {
        if (a.has_mask) CHEMSIM_LAUNCH_STEP2(true, true); else CHEMSIM_LAUNCH_STEP2(true, false);
    } else {
        if (a.has_mask) CHEMSIM_LAUNCH_STEP2(false, true); else CHEMSIM_LAUNCH_STEP2(false, false);
    }
#undef CHEMSIM_LAUNCH_STEP2
}

}  // namespace chemsim
