// step_decl.cuh — interface between the dispatch in kernels.cu and the per-operator
// translation units (step_bgk.cu, step_trt.cu, step_regularized.cu, step_kbc.cu) that hold the
// instantiations of the fused step kernels.  One TU per collision operator keeps the build
// parallel (the step kernels are ~95 % of the library's compile time).
#pragma once

#include "kernels.cuh"

namespace chemsim {

template <typename T> struct VecOf;
template <> struct VecOf<float>  { using type = float4;  static constexpr int N = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int N = 2; };

// Build-time tunables (defaults are the measured best; tools/variants.py sweeps them).
#ifndef CHEMSIM_STEP_THREADS
#define CHEMSIM_STEP_THREADS 256
#endif
#ifndef CHEMSIM_STEP_MIN_BLOCKS
#define CHEMSIM_STEP_MIN_BLOCKS 4   // <= 64 registers: 4 x 256 threads per SM (ptxas otherwise takes 88 for f64)
#endif
constexpr int STEP_THREADS = CHEMSIM_STEP_THREADS;
// f32: collide two cells at once with packed additions (F32x2, d2q9.cuh).  The two-step kernels
// always have two cells per thread in flight, so it costs them no registers; the single-step
// vector kernels are capped at 64 registers and would spill, and BGK is HBM-bound there anyway:
// a bit mask over the Collision enum says which operators' single-step kernels pack.
#ifndef CHEMSIM_PACKED_STEP2
#define CHEMSIM_PACKED_STEP2 1
#endif
#ifndef CHEMSIM_PACKED_VEC
#define CHEMSIM_PACKED_VEC 0
#endif
// two-step kernels, f32 phase A: 1 = each thread takes two horizontally adjacent cells (64-bit loads / stores where
// aligned, half the address arithmetic), 0 = two cells a block-width apart (tools/variants.py nohpair).  Measured
// (r02t, 4096^2 BGK): 157.0 vs 143.9 GLUPS in 20-step batches, 140.7 vs 131.5 sustained.
#ifndef CHEMSIM_STEP2_HPAIR
#define CHEMSIM_STEP2_HPAIR 1
#endif
// two-step kernels, phase B: 1 = streaming stores (st.global.cs: the result is not read again before the next pass,
// which leaves more of L2 to the prefetched source lines), 0 = default write-back stores
#ifndef CHEMSIM_STEP2_STORE_CS
#define CHEMSIM_STEP2_STORE_CS 1
#endif
// two-step kernels: L2 prefetch distance in tiles (> 0), or in percent of one wave of resident blocks
// (< 0: -100 = one wave, -25 = a quarter); 0 = off.  Measured on 4096^2 BGK f32 (r02o, TY = 8, wave = 592 tiles),
// GLUPS in 200-step batches (power-capped) / 20-step batches: off 128.9 / 137.3, 64 tiles 133.9, 148 tiles
// 133.7 / 151.0, 296 tiles 132.2 / 143.6, 444 tiles 130.2; f64: off 69.9, 148 tiles 75.0, 296 tiles 73.7.
// Further ahead the lines are evicted again before they are used (r02n, one wave at TY = 16: DRAM reads
// 604 -> 753 MB per launch, -1.3 %; two waves -10 %).
#ifndef CHEMSIM_STEP2_PREFETCH_DEFAULT
#define CHEMSIM_STEP2_PREFETCH_DEFAULT -25
#endif
template <int COL> __host__ __device__ constexpr bool vec_packed() { return ((CHEMSIM_PACKED_VEC >> COL) & 1) != 0; }

// widths that are a multiple of the vector width take the 128-bit kernels
template <typename T>
inline bool use_vec(const StepArgs<T> &a) { return a.W % VecOf<T>::N == 0; }

// Defined (and explicitly instantiated for float/double) in step_<operator>.cu:
// rows [y_begin, ...) of a lattice / the whole slab incl. the peer-memory halo / the two face rows.
template <typename T, int COL> void launch_step_col(const StepArgs<T> &a, cudaStream_t s);
template <typename T, int COL> void launch_slab_p2p_col(const StepArgs<T> &a, cudaStream_t s);
template <typename T, int COL> void launch_step2_col(const StepArgs<T> &a, cudaStream_t s);
template <typename T, int COL> void launch_slab_p2p2_col(const StepArgs<T> &a, cudaStream_t s);

// tile of the two-steps-per-pass kernel (step2_impl.cuh)
template <typename T>
struct Step2Tile {
    static constexpr int V = VecOf<T>::N;          // cells per 16 bytes
    static constexpr int TX = 32 * V;              // one warp covers a tile row in phase B
    // tile height (tools/variants.py s2ty16 / s2ty32 override it).  Measured on 4096^2 BGK, GLUPS f32 / f64:
    //   before the packed additions:  TY = 8: 125.9 / 69.6, TY = 16: 124.7 / 66.2, TY = 32: 115.7 / 58.9
    //   with them (r02n):             TY = 8: 132.7 (137.0 in 20-step batches, below the power cap), TY = 16: 128.9 (128.7)
    // Four 256-thread blocks per SM overlap their load / barrier / store phases better than two 512-thread
    // blocks; the taller rim of the flat tile (+27 % cells instead of +14 %) is cheaper than that.
#ifndef CHEMSIM_STEP2_TY
#define CHEMSIM_STEP2_TY 8
#endif
    static constexpr int TY = CHEMSIM_STEP2_TY;
    static constexpr int NT = 32 * TY;             // threads per block: one warp per tile row
    static constexpr int EX = TX + 2, EY = TY + 2; // tile + one-cell rim
    static constexpr int SP = ((EX + V - 1 + V - 1) / V) * V;   // shared row pitch (room for the column shift)
    static constexpr size_t SMEM = (size_t)Q * EY * SP * sizeof(T);
    // resident blocks per SM: shared memory (227 KB) and threads (2048) allow this many
    static constexpr int BLOCKS = (int)((227u * 1024u) / SMEM) < 2048 / NT ? (int)((227u * 1024u) / SMEM) : 2048 / NT;
};

}  // namespace chemsim
