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
//   phase A  (TY+2) x (TX+2) cells ("ext" region): f64 one cell per thread and iteration, scalar
//            coalesced loads; f32 one pair of horizontally adjacent cells (64-bit loads and
//            shared-memory stores where the pair is 8-byte aligned), collided together with packed
//            additions (F32x2 in d2q9.cuh); results go to smem[q][row][col + shift_q].
//            The rim is redundant work (+27 % cells for the 8 x 128 tile) whose loads hit L2 (the
//            neighbouring tiles read the same lines).
//   phase B  one warp per tile row, V = 16/sizeof(T) cells per lane: nine aligned 128-bit
//            shared-memory loads (the per-population column shift_q makes every shifted read
//            start on a 16-byte boundary: conflict-free), collide, nine 128-bit global stores.
//   prefetch every block first asks L2 for the source lines of the tile a quarter of a wave further
//            down the dispatch order (step2_prefetch_tile), so that block's first loads hit L2.
//
// Edges: cells outside a zero-fill lattice hold 0 in the ext region (they are never computed:
// src/lbm.rs:716-729 drops what leaves the array); periodic edges wrap the coordinates.  On a
// y-slab the rows beyond the slab are the neighbours' cells: their populations come from the TWO
// ghost rows, their solid flags from the mask's halo row (StepArgs::ghost_mask).
//
// step2_slab_p2p_kernel is the peer-memory form for a y-slab: the two FACE tile rows (rows
// 0..TY-1 and H-TY..H-1) are dispatched first, wait for the neighbours' step flags, and store
// rows 0, 1 / H-1, H-2 of the result straight into the neighbours' ghost rows before publishing
// step t+2 (same protocol as step_slab_p2p_kernel).
#pragma once

#include "step_decl.cuh"

namespace chemsim {

namespace {

// column shift of population q in shared memory: makes (x + 1 - ex_q + shift_q) a multiple of V
template <int V> __host__ __device__ constexpr int shift_of(int q) { return (((ex_of(q) - 1) % V) + V) % V; }

// L2 prefetch of a LATER tile's source lines (a hint, no effect on results).  A block lives for one
// "wave" (~10 us at 4096^2); its first loads would otherwise wait for HBM.  The tile that the block
// `prefetch_tiles` further down the dispatch order will work on is requested into L2 now, so that
// those loads find it there (the working set of one wave, ~26 MB, is far below the 126 MB L2).
// Only full tiles strictly inside the lattice / slab (rows pty0-1 .. pty0+TY all own rows) are
// requested: nothing outside the buffers, no ghost row.
template <typename T>
__device__ __forceinline__ void step2_prefetch_tile(const StepArgs<T> &a, const T *src, int ptx0, int pty0)
{
    using TL = Step2Tile<T>;
    constexpr int LINE = 128 / (int)sizeof(T), LINES = TL::TX / LINE;          // 128-byte lines per tile row
    if (pty0 < 1 || pty0 + TL::TY + 1 > a.H || ptx0 < 0 || ptx0 + TL::TX > a.W) return;
    static_assert(TL::EY * LINES <= TL::NT, "one thread per (row, line) of the tile");
    if (threadIdx.x >= TL::EY * LINES) return;       // the first EY*LINES threads: one line of all nine populations each
    const int r = threadIdx.x / LINES, l = threadIdx.x % LINES;
    const char *p = reinterpret_cast<const char *>(src + (size_t)(pty0 - 1 + GHOST + r) * a.pitch + ptx0 + l * LINE);
#pragma unroll
    for (int q = 0; q < Q; ++q) asm volatile("prefetch.global.L2 [%0];" : : "l"(p + a.st_off[q]));
}

// One tile: rows [ty0, ty0+TY) x columns [tx0, tx0+TX); rows >= y_end are not stored (tile rows
// that overlap the next region).  P2P: also deliver the face rows to the neighbours.
template <typename T, bool PERIODIC_X, int COL, bool P2P, bool USE_MASK>
__device__ __forceinline__ void step2_tile(const StepArgs<T> &a, T *const sm, const int tx0, const int ty0,
                                           const int y_end, const int tok)
{
    using TL = Step2Tile<T>;
    constexpr int V = TL::V, TX = TL::TX, TY = TL::TY, NT = TL::NT, EX = TL::EX, EY = TL::EY, SP = TL::SP;
    const int tid = threadIdx.x;
    const T *src = a.src + tok;
    const uint8_t *mask = a.mask + tok;
    // USE_MASK: does any cell this tile touches need the mask?  Own rows: if the slab has solids;
    // ghost rows (tiles at a slab face): the neighbour may have solids there even if this slab has
    // none, so face tiles of a y-slab always take the masked instantiation (see the launchers).
    constexpr bool use_mask = USE_MASK;

    // ---- phase A: step n+1 on the ext region -> shared memory -----------------------------------
    // interior tile: every source cell of the ext region lies inside this slab (no wrap, no
    // zero-fill, no ghost row): block-uniform fast path
    const bool interior = ty0 >= 2 && ty0 + TY + 2 <= a.H && tx0 >= 2 && tx0 + TX + 2 <= a.W;
    if (interior) {
        constexpr int DY = NT / EX, DX = NT % EX;    // idx += NT  <=>  (ey, ex) += (DY, DX) with a carry
        int ey0 = tid / EX, ex0 = tid - ey0 * EX;
        // opaque base: keeps "pointer + precomputed offset" as two integer instructions per load
        unsigned long long base = reinterpret_cast<unsigned long long>(src) +
                                  ((size_t)(ty0 - 1 + GHOST) * a.pitch + (tx0 - 1)) * sizeof(T);
        asm volatile("" : "+l"(base));
        if constexpr (sizeof(T) == 4 && CHEMSIM_STEP2_HPAIR != 0) {
            // f32: each thread takes two HORIZONTALLY adjacent ext cells (2 px, 2 px + 1).  The second cell's
            // addresses are the first one's + 4 bytes (no address arithmetic of its own), and because the ext region
            // starts at an odd column (tx0 - 1) the six populations that stream along x sit on an 8-byte boundary:
            // one 64-bit load each, straight into the register pair the packed collision works on.  In shared
            // memory the same six land on even columns (shift_q even): one 64-bit store each.
            static_assert(EX % 2 == 0 && SP % 2 == 0 && (EY * SP) % 2 == 0, "pairs must not straddle rows");
            constexpr int PX = EX / 2, DYP = NT / PX, DXP = NT % PX;
            int ey0 = tid / PX, px0 = tid - ey0 * PX;
#pragma unroll 1
            for (int idx = tid; idx < EY * PX; idx += NT) {
                const char *p0 = reinterpret_cast<const char *>(base) + ((size_t)ey0 * a.pitch + 2 * px0) * sizeof(T);
                T c0[Q], c1[Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const char *sp = p0 + (a.ld_off[q] - ex_of(q) * (long long)sizeof(T));
                    if (ex_of(q) != 0) {
                        const float2 v = __ldg(reinterpret_cast<const float2 *>(sp));
                        c0[q] = v.x; c1[q] = v.y;
                    } else {
                        c0[q] = __ldg(reinterpret_cast<const T *>(sp));
                        c1[q] = __ldg(reinterpret_cast<const T *>(sp + sizeof(T)));
                    }
                }
                if (use_mask) {
                    const uint8_t *mp = mask + (size_t)(ty0 - 1 + ey0) * a.mask_pitch + (tx0 - 1 + 2 * px0);
                    const bool s0 = __ldg(mp) != 0, s1 = __ldg(mp + 1) != 0;
                    bounce_back(c0, s0);
                    bounce_back(c1, s1);
                }
                collide2<COL, CHEMSIM_PACKED_STEP2 != 0>(c0, c1, a.k);
                T *d0 = sm + ey0 * SP + 2 * px0;
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    T *dq = d0 + q * (EY * SP) + shift_of<V>(q);
                    if (shift_of<V>(q) % 2 == 0) *reinterpret_cast<float2 *>(dq) = make_float2(c0[q], c1[q]);
                    else { dq[0] = c0[q]; dq[1] = c1[q]; }
                }
                ey0 += DYP; px0 += DXP;
                if (px0 >= PX) { px0 -= PX; ey0 += 1; }
            }
        } else if constexpr (sizeof(T) == 4) {
            // f32 with CHEMSIM_STEP2_HPAIR=0 (the earlier form, kept for A/B): two ext cells a block-width apart per
            // iteration, all eighteen loads issued before the first collision —
            // the HBM/L2 latency of one cell is covered by the arithmetic of the other
#pragma unroll 1
            for (int idx = tid; idx < EY * EX; idx += 2 * NT) {
                int ey1 = ey0 + DY, ex1 = ex0 + DX;
                if (ex1 >= EX) { ex1 -= EX; ey1 += 1; }
                const bool two = idx + NT < EY * EX;
                if (!two) { ey1 = ey0; ex1 = ex0; }  // no second cell: re-read the first (result discarded)
                const char *p0 = reinterpret_cast<const char *>(base) + ((size_t)ey0 * a.pitch + ex0) * sizeof(T);
                const char *p1 = reinterpret_cast<const char *>(base) + ((size_t)ey1 * a.pitch + ex1) * sizeof(T);
                T c0[Q], c1[Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    c0[q] = __ldg(reinterpret_cast<const T *>(p0 + (a.ld_off[q] - ex_of(q) * (long long)sizeof(T))));
                    c1[q] = __ldg(reinterpret_cast<const T *>(p1 + (a.ld_off[q] - ex_of(q) * (long long)sizeof(T))));
                }
                if (use_mask) {
                    const bool s0 = __ldg(mask + (size_t)(ty0 - 1 + ey0) * a.mask_pitch + (tx0 - 1 + ex0)) != 0;
                    const bool s1 = __ldg(mask + (size_t)(ty0 - 1 + ey1) * a.mask_pitch + (tx0 - 1 + ex1)) != 0;
                    bounce_back(c0, s0);
                    bounce_back(c1, s1);
                }
                collide2<COL, CHEMSIM_PACKED_STEP2 != 0>(c0, c1, a.k);              // without a second cell c1 repeats c0 (discarded)
                T *d0 = sm + ey0 * SP + ex0;
#pragma unroll
                for (int q = 0; q < Q; ++q) d0[q * (EY * SP) + shift_of<V>(q)] = c0[q];
                if (two) {
                    T *d1 = sm + ey1 * SP + ex1;
#pragma unroll
                    for (int q = 0; q < Q; ++q) d1[q * (EY * SP) + shift_of<V>(q)] = c1[q];
                }
                ey0 = ey1 + DY; ex0 = ex1 + DX;
                if (ex0 >= EX) { ex0 -= EX; ey0 += 1; }
            }
        } else {
            // f64: one cell per iteration (two would spill at 64 registers)
#pragma unroll 1
            for (int idx = tid; idx < EY * EX; idx += NT) {
                const char *p0 = reinterpret_cast<const char *>(base) + ((size_t)ey0 * a.pitch + ex0) * sizeof(T);
                T c0[Q];
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    c0[q] = __ldg(reinterpret_cast<const T *>(p0 + (a.ld_off[q] - ex_of(q) * (long long)sizeof(T))));
                if (use_mask) bounce_back(c0, __ldg(mask + (size_t)(ty0 - 1 + ey0) * a.mask_pitch + (tx0 - 1 + ex0)) != 0);
                collide<COL>(c0, a.k);
                T *d0 = sm + ey0 * SP + ex0;
#pragma unroll
                for (int q = 0; q < Q; ++q) d0[q * (EY * SP) + shift_of<V>(q)] = c0[q];
                ey0 += DY; ex0 += DX;
                if (ex0 >= EX) { ex0 -= EX; ey0 += 1; }
            }
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
                // the three source rows / columns of this cell, wrapped once (not once per population):
                // row[d], col[d] for a source at (gy + d - 1, gx + d - 1); population q reads (1 - ey_q, 1 - ex_q)
                size_t row[3];
                int col[3];
                bool col_in[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    int sy = gy + d - 1, sx = gx + d - 1;
                    if (a.wrap_y) { if (sy < 0) sy = a.H - 1; else if (sy >= a.H) sy = 0; }
                    // rows -GHOST .. H+GHOST-1 exist: ghost rows hold the neighbour slab's cells, or 0 at a zero-fill edge
                    row[d] = (size_t)(sy + GHOST) * a.pitch;
                    col_in[d] = true;
                    if (PERIODIC_X) { if (sx < 0) sx = a.W - 1; else if (sx >= a.W) sx = 0; }
                    else            { col_in[d] = sx >= 0 && sx < a.W; if (!col_in[d]) sx = gx; }
                    col[d] = sx;
                }
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const T *cell = src + (size_t)q * a.plane + row[1 - ey_of(q)] + col[1 - ex_of(q)];
                    // ghost rows are written by the neighbouring GPU while this kernel may be resident: coherent load
                    const T v = P2P ? *reinterpret_cast<const volatile T *>(cell) : __ldg(cell);
                    c[q] = col_in[1 - ex_of(q)] ? v : T(0);
                }
                // gy may be -1 or H on a slab: the mask's halo row holds the neighbour's face row
                if (use_mask) bounce_back(c, mask[(ptrdiff_t)gy * a.mask_pitch + gx] != 0);
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
    if (use_mask) maskw = ldg_mask(mask + (size_t)gy * a.mask_pitch + gx0, true, (const T *)nullptr);
#pragma unroll
    for (int j = 0; j < V; j += 2) {
        T c0[Q], c1[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) { c0[q] = g[q][j]; c1[q] = g[q][j + 1]; }
        if (use_mask) {
            bounce_back(c0, ((maskw >> (8 * j)) & 0xffu) != 0);
            bounce_back(c1, ((maskw >> (8 * j + 8)) & 0xffu) != 0);
        }
        collide2<COL, CHEMSIM_PACKED_STEP2 != 0>(c0, c1, a.k);
#pragma unroll
        for (int q = 0; q < Q; ++q) { g[q][j] = c0[q]; g[q][j + 1] = c1[q]; }
    }
    char *out = reinterpret_cast<char *>(a.dst) + ((size_t)(gy + GHOST) * a.pitch + gx0) * sizeof(T);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        if constexpr (CHEMSIM_STEP2_STORE_CS != 0) {
            if constexpr (V == 4) __stcs(reinterpret_cast<float4 *>(out + a.st_off[q]), make_float4(g[q][0], g[q][1], g[q][2], g[q][3]));
            else                  __stcs(reinterpret_cast<double2 *>(out + a.st_off[q]), make_double2(g[q][0], g[q][1]));
        } else {
            store_vec(reinterpret_cast<T *>(out + a.st_off[q]), g[q]);
        }
    }
    if (P2P) halo_store_row(a, gy, gx0, g);
}

// rows [y_begin, y_begin + y_count) advance by two steps (unsharded lattice, or one region of a slab)
template <typename T, bool PERIODIC_X, int COL, bool USE_MASK>
__global__ void __launch_bounds__(Step2Tile<T>::NT, Step2Tile<T>::BLOCKS)
step2_kernel(const __grid_constant__ StepArgs<T> a)
{
    using TL = Step2Tile<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    asm volatile("griddepcontrol.launch_dependents;");
    const int y_end = a.y_begin + a.y_count;
    const int trow = blockIdx.z * gridDim.y + blockIdx.y;
    const int ty0 = a.y_begin + trow * TL::TY;
    if (ty0 >= y_end) return;
    const int tok = order_after_grid_dependency();
    if (a.prefetch_tiles > 0) {
        int pcol = blockIdx.x + a.prefetch_cols, prow = trow + a.prefetch_rows;
        if (pcol >= (int)gridDim.x) { pcol -= gridDim.x; prow += 1; }
        const int pty0 = a.y_begin + prow * TL::TY;
        if (pty0 + TL::TY <= y_end) step2_prefetch_tile(a, a.src + tok, pcol * TL::TX, pty0);
    }
    step2_tile<T, PERIODIC_X, COL, false, USE_MASK>(a, reinterpret_cast<T *>(smem_raw), blockIdx.x * TL::TX, ty0, y_end, tok);
}

// a whole y-slab, H >= 2 TY: tile-row slot 0 -> rows [0, TY), slot 1 -> rows [H-TY, H) (the two face
// tile rows, dispatched first), slot s >= 2 -> rows [(s-1) TY, ...) clipped at H-TY
template <typename T, bool PERIODIC_X, int COL, bool USE_MASK>
__global__ void __launch_bounds__(Step2Tile<T>::NT, Step2Tile<T>::BLOCKS)
step2_slab_p2p_kernel(const __grid_constant__ StepArgs<T> a)
{
    using TL = Step2Tile<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    asm volatile("griddepcontrol.launch_dependents;");
    T *const sm = reinterpret_cast<T *>(smem_raw);
    const int slot = blockIdx.z * gridDim.y + blockIdx.y;
    const int tx0 = blockIdx.x * TL::TX;
    if (slot >= 2) {                                 // interior tile row: no ghost row, feeds no neighbour
        const int ty0 = (slot - 1) * TL::TY;
        if (ty0 >= a.H - TL::TY) return;
        const int tok = order_after_grid_dependency();
        if (a.prefetch_tiles > 0) {                  // a later interior tile (slot' >= 2 as well)
            int pcol = blockIdx.x + a.prefetch_cols, pslot = slot + a.prefetch_rows;
            if (pcol >= (int)gridDim.x) { pcol -= gridDim.x; pslot += 1; }
            const int pty0 = (pslot - 1) * TL::TY;
            if (pty0 + TL::TY <= a.H - TL::TY) step2_prefetch_tile(a, a.src + tok, pcol * TL::TX, pty0);
        }
        step2_tile<T, PERIODIC_X, COL, false, USE_MASK>(a, sm, tx0, ty0, a.H - TL::TY, tok);
        return;
    }
    const HaloP2P &p = a.halo;
    const int ty0 = slot == 0 ? 0 : a.H - TL::TY;
    const int tok = order_after_grid_dependency() + order_after_halo_flags(p, slot == 0, slot == 1);
    step2_tile<T, PERIODIC_X, COL, true, true>(a, sm, tx0, ty0, ty0 + TL::TY, tok);      // ghost rows: always masked
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) publish_step(p, 2u * gridDim.x, 2u);
}

// tiles ahead for the L2 prefetch: CHEMSIM_LBM_PREFETCH=<n> in the environment overrides the build's default
// (n > 0: tiles, n < 0: percent of one wave = resident blocks per SM x SMs of the device, 0: off)
template <typename T>
int step2_prefetch_tiles()
{
    static const int n = [] {
        if (const char *e = getenv("CHEMSIM_LBM_PREFETCH")) return atoi(e);
        return CHEMSIM_STEP2_PREFETCH_DEFAULT;
    }();
    if (n >= 0) return n;
    static const int sms = [] {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    return (int)((long long)-n * Step2Tile<T>::BLOCKS * sms / 100);
}

template <typename K>
void step2_opt_in(K kernel, size_t smem)
{
    // > 48 KB of dynamic shared memory needs an opt-in per function and device
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

}  // namespace

// rows [y_begin, y_begin + y_count) advance by TWO steps; y_count need not be a multiple of the tile height
template <typename T, int COL>
void launch_step2_col(const StepArgs<T> &a_in, cudaStream_t s)
{
    using TL = Step2Tile<T>;
    const dim3 block(TL::NT);
    const dim3 grid = row_grid((a_in.W + TL::TX - 1) / TL::TX, (a_in.y_count + TL::TY - 1) / TL::TY);
    StepArgs<T> a = a_in;
    a.prefetch_tiles = step2_prefetch_tiles<T>();
    a.prefetch_rows = a.prefetch_tiles / (int)grid.x;
    a.prefetch_cols = a.prefetch_tiles % (int)grid.x;
    // a region at a slab face recomputes ghost-row cells, whose solid flags the neighbour owns
    const bool masked = a.has_mask != 0 || (a.ghost_mask != 0 && (a.y_begin == 0 || a.y_begin + a.y_count >= a.H));
#define CHEMSIM_LAUNCH_STEP2(PX, M)                                                           \
    do {                                                                                      \
        step2_opt_in(step2_kernel<T, PX, COL, M>, TL::SMEM);                                  \
        launch_chained(step2_kernel<T, PX, COL, M>, grid, block, s, a, TL::SMEM);             \
    } while (0)
    if (a.periodic_x) { if (masked) CHEMSIM_LAUNCH_STEP2(true, true); else CHEMSIM_LAUNCH_STEP2(true, false); }
    else              { if (masked) CHEMSIM_LAUNCH_STEP2(false, true); else CHEMSIM_LAUNCH_STEP2(false, false); }
#undef CHEMSIM_LAUNCH_STEP2
}

// the whole slab, two steps, with the peer-memory halo (H >= 2 TY)
template <typename T, int COL>
void launch_slab_p2p2_col(const StepArgs<T> &a_in, cudaStream_t s)
{
    using TL = Step2Tile<T>;
    const dim3 block(TL::NT);
    const int interior_rows = a_in.H - 2 * TL::TY;
    const dim3 grid = row_grid((a_in.W + TL::TX - 1) / TL::TX, 2 + (interior_rows + TL::TY - 1) / TL::TY);
    StepArgs<T> a = a_in;
    a.prefetch_tiles = step2_prefetch_tiles<T>();
    a.prefetch_rows = a.prefetch_tiles / (int)grid.x;
    a.prefetch_cols = a.prefetch_tiles % (int)grid.x;
#define CHEMSIM_LAUNCH_STEP2(PX, M)                                                           \
    do {                                                                                      \
        step2_opt_in(step2_slab_p2p_kernel<T, PX, COL, M>, TL::SMEM);                         \
        launch_chained(step2_slab_p2p_kernel<T, PX, COL, M>, grid, block, s, a, TL::SMEM);    \
    } while (0)
    if (a.periodic_x) { if (a.has_mask) CHEMSIM_LAUNCH_STEP2(true, true); else CHEMSIM_LAUNCH_STEP2(true, false); }
    else              { if (a.has_mask) CHEMSIM_LAUNCH_STEP2(false, true); else CHEMSIM_LAUNCH_STEP2(false, false); }
#undef CHEMSIM_LAUNCH_STEP2
}

}  // namespace chemsim
