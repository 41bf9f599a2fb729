// step_impl.cuh — the fused step kernels (templates).  Included by one translation unit per
// collision operator, which instantiates launch_*_col<float|double, COL> (see step_decl.cuh).
//
//  step_vec_kernel       the hot path: ONE pass per time step that pull-streams the nine
//                        populations (State::stream, src/lbm.rs:716-729), reverses them on solid
//                        cells (State::bounce_back, :741-751) and relaxes them (State::collide,
//                        :731-739, with the operator COL), 128-bit loads/stores over the SoA layout.
//  step_slab_p2p_kernel  the same for a whole y-slab of a sharded lattice, fused with the halo
//                        exchange over peer memory; step_face_p2p_kernel for the two face rows only.
//  step_scalar_kernel    the same update, one cell per thread, for widths that are not a multiple
//                        of the vector width.
//
// HBM-bound streaming work: no tensor cores, no shared-memory tiling (every population value is
// read once and written once per step).
#pragma once

#include <cstdlib>

#include "step_decl.cuh"

namespace chemsim {

namespace {


#ifdef CHEMSIM_LOAD_NOALLOC
#define CHEMSIM_LD_HINT ".L1::no_allocate"
#else
#define CHEMSIM_LD_HINT ""
#endif

// Branch-free global loads (the one-element edge loads are predicated).  NC=true takes the read-only path: the
// source buffer is never written by the kernel that reads it (A-B buffering), so
// .nc is legal.  NC=false (coherent) is used by the P2P face kernel, whose ghost
// rows are written by the neighbouring GPU while the kernel may already be resident.
// Predication instead of `if` keeps every load of a thread in ONE straight-line
// batch: all of them are in flight before the first use.  The asm statements carry no
// memory dependence of their own: what keeps them behind griddepcontrol.wait and the halo
// flag wait is a data dependence — every address is derived from an "order token" (always
// 0, but opaque to the compiler) that those waits produce (see order_after_*).
#define CHEMSIM_LDG_BODY(NCSTR)                                                                       \
    asm("ld.global" NCSTR ".v4.f32 {%0, %1, %2, %3}, [%4];"                                           \
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p))
template <bool NC>
__device__ __forceinline__ void ldg_vec(const float *p, float (&v)[4])
{
    if (NC) CHEMSIM_LDG_BODY(".nc" CHEMSIM_LD_HINT); else CHEMSIM_LDG_BODY("");
}
#undef CHEMSIM_LDG_BODY
#define CHEMSIM_LDG_BODY(NCSTR)                                                                       \
    asm("ld.global" NCSTR ".v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p))
template <bool NC>
__device__ __forceinline__ void ldg_vec(const double *p, double (&v)[2])
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

// resident blocks per SM the step kernels are compiled for: BGK fits 64 registers in both
// precisions; the f64 TRT / Regularized bodies need ~80 (3 blocks)
template <typename T, int COL>
constexpr int step_min_blocks()
{
    // KBC: f32 capped at 64 registers too (28 B of spill, +1.6 % measured: 58.2 vs 57.3 GLUPS); f64 left
    // unconstrained (96 registers; the cap costs 6 %)
#ifndef CHEMSIM_KBC_MIN_BLOCKS
#define CHEMSIM_KBC_MIN_BLOCKS (sizeof(T) == 4 ? 4 : 1)    // tools/variants.py: kbc3 / kbc4 override
#endif
    return COL == COL_KBC ? CHEMSIM_KBC_MIN_BLOCKS : (COL != COL_BGK && sizeof(T) == 8 ? 3 : CHEMSIM_STEP_MIN_BLOCKS);
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

// Peer-memory halo: the V cells at (y, x0..) of all nine populations, just computed, go where the
// neighbouring GPU's next pass reads them (see HaloP2P): rows 0/1 to the upper neighbour's ghost
// rows H_up/H_up+1, rows H-1/H-2 to the lower neighbour's ghost rows -1/-2; the outer row of
// each pair only needs the populations that move towards the face.
template <typename T, int V>
__device__ __forceinline__ void halo_store_row(const StepArgs<T> &a, int y, int x0, const T (&g)[Q][V])
{
    const HaloP2P &p = a.halo;
    if (p.up_dst && y <= 1) {
        T *peer = (T *)p.up_dst + (size_t)(p.up_row0 + y) * a.pitch + x0;
#pragma unroll
        for (int q = 0; q < Q; ++q)
            if (y == 0 || ey_of(q) == -1) store_vec(peer + (size_t)q * p.up_plane, g[q]);
    }
    if (p.down_dst && y >= a.H - 2) {
        T *peer = (T *)p.down_dst + (size_t)(GHOST - (a.H - y)) * a.pitch + x0;    // row H-1 -> GHOST-1, H-2 -> GHOST-2
#pragma unroll
        for (int q = 0; q < Q; ++q)
            if (y == a.H - 1 || ey_of(q) == 1) store_vec(peer + (size_t)q * p.down_plane, g[q]);
    }
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
    // lanes beyond the end of the row (last, partially filled warp) re-read the warp's first vector
    // instead of being predicated off: the vector loads then need neither predicate nor zero-fill
    const int x0 = (active ? xv : xv - lane) * V;
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
    // Addresses.  ONE 64-bit per-thread byte pointer p0 (own row of population 0, first cell of the
    // vector); every load adds a 64-bit byte offset that the HOST precomputed per population
    // (a.ld_off[q] = (q*plane - ey_q*pitch)*sizeof(T), a kernel-parameter constant that SASS
    // uses as an immediate operand), so a load costs two integer instructions.  Rows 0 / H-1 of an
    // unsharded periodic lattice shift the base by +-H rows (block-uniform) instead of reading ghosts.
    const char *p0 = reinterpret_cast<const char *>(src) + ((size_t)(y + GHOST) * a.pitch + x0) * sizeof(T);
    const char *pu = p0, *pd = p0;                   // bases of the dy = +1 / dy = -1 movers' source rows
    if (a.wrap_y) { if (y == 0) pu = p0 + a.wrap_bytes; if (y == a.H - 1) pd = p0 - a.wrap_bytes; }
    // left / right edge element relative to the vector: x0-1 (wrapped: W-1, and first => x0 == 0) / x0+V (wrapped: 0)
    const int dl = (first ? a.W - 1 : -1) * (int)sizeof(T), dr = (last ? -x0 : V) * (int)sizeof(T);

    // ---- phase 1: every load of this thread, back to back ----------------------
    T v[Q][V];      // the aligned vector of each population's source row
    T e[Q];         // the one extra element for populations that stream along x
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const char *row = (ey_of(q) == 1 ? pu : (ey_of(q) == -1 ? pd : p0)) + a.ld_off[q];
        ldg_vec<NC>(reinterpret_cast<const T *>(row), v[q]);
        if (ex_of(q) == 1)       e[q] = ldg_one<NC>(reinterpret_cast<const T *>(row + dl), need_left);
        else if (ex_of(q) == -1) e[q] = ldg_one<NC>(reinterpret_cast<const T *>(row + dr), need_right);
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

    // ---- phase 3: bounce-back + collide, two cells at a time (f32: packed adds) ---
#pragma unroll
    for (int j = 0; j < V; j += 2) {
        T c0[Q], c1[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) { c0[q] = g[q][j]; c1[q] = g[q][j + 1]; }
        if (HAS_MASK) {
            bounce_back(c0, ((maskw >> (8 * j)) & 0xffu) != 0);
            bounce_back(c1, ((maskw >> (8 * j + 8)) & 0xffu) != 0);
        }
        collide2<COL, vec_packed<COL>()>(c0, c1, a.k);
#pragma unroll
        for (int q = 0; q < Q; ++q) { g[q][j] = c0[q]; g[q][j + 1] = c1[q]; }
    }

    // ---- phase 4: nine aligned vector stores ------------------------------------
    char *out = reinterpret_cast<char *>(a.dst) + ((size_t)(y + GHOST) * a.pitch + x0) * sizeof(T);
#pragma unroll
    for (int q = 0; q < Q; ++q) store_vec(reinterpret_cast<T *>(out + a.st_off[q]), g[q]);

    // ---- phase 5 (P2P face rows): the halo, written where the neighbour reads it --
    if (P2P) halo_store_row(a, y, x0, g);
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
    // Grid (x-chunks, row groups [, overflow of the row groups]): the x-chunk is the fastest block
    // index, so blocks that are scheduled together work on neighbouring chunks of the same rows and
    // the 18 streams advance through DRAM pages in order (measured +4 % over row-fastest block
    // order); the row comes straight from blockIdx, i.e. it is block-uniform without a division.
    const int rg = blockIdx.z * gridDim.y + blockIdx.y, xc = blockIdx.x;
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
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Spin (with back-off) until the neighbour has published step `want`.  A neighbour that does
// not show up within p.timeout_ns (host-side skew longer than that, a dead peer) is reported
// through the error words — one in device memory for the kernels, one in mapped host memory
// that every later API call checks — and the caller must then NOT publish its own step: the
// stale halo stays on this GPU.
__device__ __forceinline__ bool wait_flag(const unsigned *flag, unsigned want, const HaloP2P &p)
{
    if (!flag) return true;
    const volatile unsigned *f = reinterpret_cast<const volatile unsigned *>(flag);
    if ((int)(*f - want) >= 0) { __threadfence_system(); return true; }
    const unsigned long long t0 = global_timer_ns();
    unsigned spins = 0;
    while ((int)(*f - want) < 0) {
        __nanosleep(spins < 64 ? 32 : 256);
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > p.timeout_ns) {
            *reinterpret_cast<volatile int *>(p.error) = 1;          // device copy: checked by publish_step
            *reinterpret_cast<volatile int *>(p.error_host) = 1;     // host copy: checked by the API calls
            __threadfence_system();
            return false;
        }
    }
    __threadfence_system();
    return true;
}

// One thread waits for both neighbours' step flags, the block follows through a barrier.
// Returns 0 through shared memory: an order token for the ghost-row loads (see ldg_*).
__device__ __forceinline__ int order_after_halo_flags(const HaloP2P &p, bool need_up = true, bool need_down = true)
{
    __shared__ int token;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        if (need_up) wait_flag(p.wait_up, p.step, p);
        if (need_down) wait_flag(p.wait_down, p.step, p);
        token = 0;
    }
    __syncthreads();
    return *reinterpret_cast<volatile int *>(&token);
}

// The last face block to finish publishes step t+1 into both neighbours' flags — unless a wait
// timed out (now or in an earlier step): a poisoned lattice never tells its neighbours to go on,
// so they time out as well instead of consuming a stale halo.
__device__ __forceinline__ void publish_step(const HaloP2P &p, unsigned total_face_blocks, unsigned steps = 1)
{
    if (atomicAdd(p.done, 1u) == total_face_blocks - 1) {
        *p.done = 0;
        __threadfence_system();
        if (*reinterpret_cast<volatile int *>(p.error) != 0) return;
        if (p.signal_down) *reinterpret_cast<volatile unsigned *>(p.signal_down) = p.step + steps;
        if (p.signal_up)   *reinterpret_cast<volatile unsigned *>(p.signal_up) = p.step + steps;
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
        c[q] = inside ? a.src[(size_t)q * a.plane + (size_t)(sy + GHOST) * a.pitch + sx] : T(0);
    }
    if (a.has_mask) bounce_back(c, a.mask[(size_t)y * a.mask_pitch + x] != 0);
    collide<COL>(c, a.k);
#pragma unroll
    for (int q = 0; q < Q; ++q) a.dst[(size_t)q * a.plane + (size_t)(y + GHOST) * a.pitch + x] = c[q];
}

}  // namespace

// Back-to-back step kernels are launched as programmatic dependents of each other
// (CHEMSIM_LBM_PDL=0 in the environment restores plain stream order).
static inline bool pdl_enabled()
{
    static const bool on = [] { const char *e = getenv("CHEMSIM_LBM_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

template <typename T>
static void launch_chained(void (*kernel)(const StepArgs<T>), dim3 grid, dim3 block, cudaStream_t s, const StepArgs<T> &a,
                           size_t smem = 0)
{
    if (!pdl_enabled()) { kernel<<<grid, block, smem, s>>>(a); return; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, a);
}

// grid of x-chunks (fastest) x row groups; row groups beyond gridDim.y's limit spill into z
static inline dim3 row_grid(int xchunks, int rowgroups)
{
    const int gy = rowgroups < 32768 ? rowgroups : 32768;
    return dim3((unsigned)xchunks, (unsigned)gy, (unsigned)((rowgroups + gy - 1) / gy));
}

}  // namespace chemsim
#ifdef CHEMSIM_EXPERIMENT_BULK
#include "step_bulk_experiment.cuh"
#endif
namespace chemsim {

template <typename T, int COL>
void launch_step_col(const StepArgs<T> &a_in, cudaStream_t s)
{
#ifdef CHEMSIM_EXPERIMENT_BULK
    if (launch_step_bulk_experiment<T, COL>(a_in, s)) return;      // A/B experiment build only
#endif
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
        const dim3 grid = row_grid(a.xchunks, (rows + by - 1) / by);
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

namespace {

// ---- the whole slab step + halo in ONE kernel (peer-memory mode) ------------------
// Grid of x-chunks (fastest) x H row slots.  The four face rows come first (slots 0..3 -> rows
// 0, 1, H-1, H-2): they wait for the neighbours' step flags (row 0 / H-1 read a ghost row), update
// their rows, store them into the neighbours' ghost rows and publish the next step early, while
// the remaining blocks stream through the interior.  One launch per step and GPU, chained with
// programmatic dependent launch; no events, no communication kernel, no second stream.
// Needs H >= 4 (slab_p2p_supported).
template <typename T, bool PERIODIC_X, bool HAS_MASK, int COL>
__global__ void __launch_bounds__(STEP_THREADS, step_min_blocks<T, COL>())
step_slab_p2p_kernel(const __grid_constant__ StepArgs<T> a)
{
    asm volatile("griddepcontrol.launch_dependents;");
    const int r = blockIdx.z * gridDim.y + blockIdx.y, xc = blockIdx.x;
    if (r >= a.H) return;                            // padding of the row dimension (block-uniform)
    const int y = r == 0 ? 0 : (r == 1 ? 1 : (r == 2 ? a.H - 1 : (r == 3 ? a.H - 2 : r - 2)));
    const int xv = xc * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    if (r >= 4) {                                    // interior row: reads no ghost row, feeds no neighbour
        step_vec_body<T, PERIODIC_X, HAS_MASK, COL, false>(a, y, xv, lane, 0);
        return;
    }
    const HaloP2P &p = a.halo;
    // Rows 0 / H-1 read a ghost row the neighbour writes; rows 1 / H-2 read none, but all four STORE into
    // a neighbour's ghost rows of the buffer that neighbour's previous pass may still be reading: every face
    // block waits for the flag of the neighbour it exchanges with (published = its face blocks are done).
    const int halo_tok = order_after_halo_flags(p, r <= 1, r >= 2);
    step_vec_body<T, PERIODIC_X, HAS_MASK, COL, true>(a, y, xv, lane, halo_tok);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) publish_step(p, 4u * (unsigned)a.xchunks);
}

}  // namespace

template <typename T, int COL>
void launch_slab_p2p_col(const StepArgs<T> &a_in, cudaStream_t s)
{
    constexpr int V = VecOf<T>::N;
    const int nvec = a_in.W / V;
    const dim3 block(STEP_THREADS, 1);
    StepArgs<T> a = a_in;
    a.xchunks = (nvec + STEP_THREADS - 1) / STEP_THREADS;
    const dim3 grid = row_grid(a.xchunks, a.H);
    if (a.periodic_x) {
        if (a.has_mask) launch_chained(step_slab_p2p_kernel<T, true, true, COL>, grid, block, s, a);
        else            launch_chained(step_slab_p2p_kernel<T, true, false, COL>, grid, block, s, a);
    } else {
        if (a.has_mask) launch_chained(step_slab_p2p_kernel<T, false, true, COL>, grid, block, s, a);
        else            launch_chained(step_slab_p2p_kernel<T, false, false, COL>, grid, block, s, a);
    }
}

}  // namespace chemsim

#include "step2_impl.cuh"

namespace chemsim {

// the two-step kernels (not for KBC: compute-bound, the redundant rim makes it slower than two single steps)
#define CHEMSIM_INSTANTIATE_STEP2(COL)                                                             \
    template void launch_step2_col<float, COL>(const StepArgs<float> &, cudaStream_t);             \
    template void launch_step2_col<double, COL>(const StepArgs<double> &, cudaStream_t);           \
    template void launch_slab_p2p2_col<float, COL>(const StepArgs<float> &, cudaStream_t);         \
    template void launch_slab_p2p2_col<double, COL>(const StepArgs<double> &, cudaStream_t);

#define CHEMSIM_INSTANTIATE_STEP(COL)                                                              \
    template void launch_step_col<float, COL>(const StepArgs<float> &, cudaStream_t);              \
    template void launch_step_col<double, COL>(const StepArgs<double> &, cudaStream_t);            \
    template void launch_slab_p2p_col<float, COL>(const StepArgs<float> &, cudaStream_t);          \
    template void launch_slab_p2p_col<double, COL>(const StepArgs<double> &, cudaStream_t);

}  // namespace chemsim
