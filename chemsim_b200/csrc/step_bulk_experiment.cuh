// step_bulk_experiment.cuh — A/B EXPERIMENT, compiled only with -DCHEMSIM_EXPERIMENT_BULK
// (tools/variants.py builds libchemsim_lbm_bulk.so; CHEMSIM_LBM_BULK=1 selects it at run time).
//
// north_star: "TMA/shared-memory staging used only where ncu shows it improves".  This is the
// single-step kernel with its global loads replaced by TMA bulk copies (cp.async.bulk, SASS
// UBLKCP): one elected thread brings the nine source row segments of a 256-thread row chunk
// (TXB cells + one 16-byte pad per side for the x-shifted populations) into shared memory and
// signals an mbarrier with the transaction byte count; the threads then read aligned 128-bit
// vectors and the one neighbour element from shared memory instead of issuing 15 global loads +
// 6 shuffles each.  Stores stay 128-bit global stores.  Unsharded lattices, vector widths that
// are a multiple of the chunk, no mask — enough for the 4096^2 roofline configuration.
// Result and ncu pair: profiles/r02_tma_ab.md.
#pragma once

#include "step_decl.cuh"

namespace chemsim {
namespace {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <typename T, bool PERIODIC_X, int COL>
__global__ void __launch_bounds__(STEP_THREADS, 4)
step_bulk_kernel(const __grid_constant__ StepArgs<T> a)
{
    constexpr int V = VecOf<T>::N, TXB = STEP_THREADS * V, PAD = V;       // PAD elements = 16 bytes
    constexpr int SROW = TXB + 2 * PAD;                                    // shared row: [pad | chunk | pad]
    __shared__ __align__(128) T sm[Q][SROW];
    __shared__ __align__(8) unsigned long long mbar;
    asm volatile("griddepcontrol.launch_dependents;");
    const int y = blockIdx.z * gridDim.y + blockIdx.y;
    if (y >= a.H) return;
    const int x0 = blockIdx.x * TXB, t = threadIdx.x;
    const bool first = x0 == 0, last = x0 + TXB == a.W;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        // bytes: nine chunks + two pads for each of the six x-shifted populations (edges of a zero-fill lattice: no pad)
        unsigned bytes = Q * TXB * (unsigned)sizeof(T);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            if (ex_of(q) == 1 && (PERIODIC_X || !first)) bytes += PAD * sizeof(T);
            if (ex_of(q) == -1 && (PERIODIC_X || !last)) bytes += PAD * sizeof(T);
        }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(bytes) : "memory");
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            int sy = y - ey_of(q);
            if (a.wrap_y) { if (sy < 0) sy = a.H - 1; else if (sy >= a.H) sy = 0; }
            const T *row = a.src + (size_t)q * a.plane + (size_t)(sy + GHOST) * a.pitch;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(&sm[q][PAD])), "l"(row + x0), "r"((unsigned)(TXB * sizeof(T))), "r"(smem_u32(&mbar)) : "memory");
            if (ex_of(q) == 1 && (PERIODIC_X || !first)) {          // the V cells left of the chunk (wrapped at x0 == 0)
                const T *src = first ? row + a.W - PAD : row + x0 - PAD;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(&sm[q][0])), "l"(src), "r"((unsigned)(PAD * sizeof(T))), "r"(smem_u32(&mbar)) : "memory");
            }
            if (ex_of(q) == -1 && (PERIODIC_X || !last)) {          // the V cells right of the chunk
                const T *src = last ? row : row + x0 + TXB;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(&sm[q][PAD + TXB])), "l"(src), "r"((unsigned)(PAD * sizeof(T))), "r"(smem_u32(&mbar)) : "memory");
            }
        }
    }
    if (!PERIODIC_X) {                                               // zero-fill: what lies outside the lattice is 0
        if (first && t < PAD) { for (int q = 0; q < Q; ++q) if (ex_of(q) == 1) sm[q][t] = T(0); }
        if (last && t < PAD)  { for (int q = 0; q < Q; ++q) if (ex_of(q) == -1) sm[q][PAD + TXB + t] = T(0); }
        __syncthreads();
    }
    {   // wait for the transaction bytes (phase parity 0)
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    }
    T g[Q][V];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const typename VecOf<T>::type v = *reinterpret_cast<const typename VecOf<T>::type *>(&sm[q][PAD + t * V]);
        T w[V];
        if constexpr (V == 4) { w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; } else { w[0] = v.x; w[1] = v.y; }
        if (ex_of(q) == 0) {
#pragma unroll
            for (int j = 0; j < V; ++j) g[q][j] = w[j];
        } else if (ex_of(q) == 1) {                                  // value at x comes from x-1
            g[q][0] = sm[q][PAD + t * V - 1];
#pragma unroll
            for (int j = 1; j < V; ++j) g[q][j] = w[j - 1];
        } else {
#pragma unroll
            for (int j = 0; j < V - 1; ++j) g[q][j] = w[j + 1];
            g[q][V - 1] = sm[q][PAD + t * V + V];
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
        T c[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) c[q] = g[q][j];
        collide<COL>(c, a.k);
#pragma unroll
        for (int q = 0; q < Q; ++q) g[q][j] = c[q];
    }
    char *out = reinterpret_cast<char *>(a.dst) + ((size_t)(y + GHOST) * a.pitch + x0 + t * V) * sizeof(T);
#pragma unroll
    for (int q = 0; q < Q; ++q) store_vec(reinterpret_cast<T *>(out + a.st_off[q]), g[q]);
}

// true if the experiment handled the launch
template <typename T, int COL>
bool launch_step_bulk_experiment(const StepArgs<T> &a, cudaStream_t s)
{
    static const bool on = [] { const char *e = getenv("CHEMSIM_LBM_BULK"); return e && e[0] == '1'; }();
    constexpr int TXB = STEP_THREADS * VecOf<T>::N;
    if (!on || a.has_mask || a.y_begin != 0 || a.y_count != a.H || a.y_stride != 1 || a.W % TXB != 0) return false;
    const dim3 grid = row_grid(a.W / TXB, a.H), block(STEP_THREADS);
    if (a.periodic_x) launch_chained(step_bulk_kernel<T, true, COL>, grid, block, s, a);
    else              launch_chained(step_bulk_kernel<T, false, COL>, grid, block, s, a);
    return true;
}

}  // namespace
}  // namespace chemsim
