// kernels.cuh — launch interface between the host runtime (lattice.cu) and the
// sm_100a kernels (kernels.cu).
#pragma once

#include "d2q9.cuh"

namespace chemsim {

// Device layout of one population set ("lattice buffer"):
//   value of population q at local row y (−GHOST … H+GHOST−1, the rows outside 0 … H−1 being
//   ghost rows) and column x lives at  base[q*plane + (y+GHOST)*pitch + x].
// Two ghost rows per side: a y-slab advances TWO steps per pass (step2_impl.cuh), which reads
// the neighbour's two face rows; the single-step kernels only touch the inner ghost row.
constexpr int GHOST = 2;
// pitch is W rounded up to 128 bytes so every row starts on a cache line and
// 128-bit vector accesses at x % VEC == 0 are aligned.
// Peer-memory halo of a y-slab (filled in only for the P2P face kernel).
struct HaloP2P {
    // The neighbours' DESTINATION population buffers (peer-mapped).  A slab delivers, per step or
    // double step, what the neighbour's next pass reads from its two ghost rows:
    //   to the lower neighbour (larger y): my row H-1, all nine populations -> its ghost row -1,
    //                                      my row H-2, the dy=+1 movers {3,6,7} -> its ghost row -2
    //   to the upper neighbour:            my row 0, all nine -> its ghost row H_up,
    //                                      my row 1, the dy=-1 movers {1,5,8} -> its ghost row H_up+1
    void *up_dst = nullptr, *down_dst = nullptr;
    size_t up_plane = 0, down_plane = 0;           // their plane strides, in elements
    int up_row0 = 0;                               // plane row of the upper neighbour's ghost row H_up (= H_up + GHOST)
    const unsigned *wait_up = nullptr, *wait_down = nullptr;   // local step flags the neighbours publish into
    unsigned *signal_up = nullptr, *signal_down = nullptr;     // the neighbours' flags this GPU publishes into
    unsigned *done = nullptr;                      // local counter of finished face blocks
    unsigned step = 0;                             // t: needs flags >= t, publishes t + (steps of this launch)
    int *error = nullptr;                          // device word: set to 1 if a wait timed out (sticky until re-upload)
    int *error_host = nullptr;                     // the same in mapped host memory, for the host API
    unsigned long long timeout_ns = 0;             // how long a face block waits for a neighbour's flag
};

template <typename T>
struct StepArgs {
    const T *src;
    T *dst;
    size_t plane;          // elements per population plane = (H+2)*pitch
    int pitch;             // elements per row
    int W, H;              // local lattice: W columns, H rows
    int y_begin, y_count;  // this launch updates rows y_begin + i*y_stride, i in [0, y_count)
    int y_stride;          // 1 for a band of rows; H−1 for the two face rows {0, H−1} of a slab
    int xchunks;           // blocks per row (set by the launcher): block b works on chunk b % xchunks of row group b / xchunks
    int ghost;             // ghost rows above and below the slab in the layout (row y lives at plane row y + ghost)
    int row0, Hglobal;     // this slab's first global row and the global height (zero-fill: what lies outside)
    int periodic_y;        // 1: the lattice is periodic in y (sharded: the ghost rows carry the wrap)
    int wrap_y;            // 1: rows −1/H alias rows H−1/0 (periodic, unsharded); 0: read the ghost rows
    int periodic_x;        // 1: wrap in x; 0: zero-fill (reference)
    const uint8_t *mask;   // H rows of mask_pitch bytes, non-zero = solid
    int mask_pitch;
    int has_mask;          // 0: no solid cell anywhere in this slab, mask not read
    int ghost_mask;        // 1: a y-slab whose mask carries one halo row per side (the neighbour's face row, refreshed
                           //    before every batch of steps): the two-step kernel applies it to ghost-row cells
    int collision;         // Collision enum: which CollisionOperator the step applies
    const uint8_t *mask_flags;  // per row, one byte per 64-cell segment: any solid cell in it?
    int flag_pitch;             // (a warp of the vector kernel covers whole segments and skips
                                //  the mask load when they are solid-free)
    // byte offsets from a thread's own-row pointer of population 0, precomputed by the host
    // (fill_offsets): loads of population q come from row y - ey_q, stores go to row y
    long long ld_off[Q];   // (q*plane - ey_q*pitch) * sizeof(T)
    long long st_off[Q];   // q*plane * sizeof(T)
    long long wrap_bytes;  // H*pitch*sizeof(T): what wrap_y adds/subtracts for the rows 0 / H-1
    int prefetch_tiles;    // two-step kernels: every block asks L2 for the source lines of the tile this many
                           // tiles further down the dispatch order (set by the launcher, 0 = off) ...
    int prefetch_rows, prefetch_cols;   // ... = prefetch_rows tile rows + prefetch_cols tile columns (no division on the device)
    HaloP2P halo;
    Consts<T> k;

    void fill_offsets()
    {
        for (int q = 0; q < Q; ++q) {
            st_off[q] = (long long)q * (long long)plane * (long long)sizeof(T);
            ld_off[q] = st_off[q] - (long long)ey_of(q) * pitch * (long long)sizeof(T);
        }
        wrap_bytes = (long long)H * pitch * (long long)sizeof(T);
    }
};

enum ReadoutKind : int {
    READ_DENSITY = 0, READ_PRESSURE, READ_SPEED, READ_VELOCITY, READ_MOMENTUM,
    READ_EQUILIBRIUM, READ_NON_EQUILIBRIUM
};

template <typename T>
struct ReadoutArgs {
    const T *src;
    size_t plane;
    int pitch;
    int W, H;
    int kind;
    int q;          // for READ_EQUILIBRIUM / READ_NON_EQUILIBRIUM
    T *out0, *out1; // dense (pitch == W) outputs
    Consts<T> k;
};

// Each launcher returns the number of kernels it launched (for the handle's
// launch counter) or a negative cudaError_t.
template <typename T> int launch_step(const StepArgs<T> &a, cudaStream_t s);
// the same rows advanced by TWO steps in one pass (temporal blocking through shared memory, step2_impl.cuh)
template <typename T> int launch_step2(const StepArgs<T> &a, cudaStream_t s);
template <typename T> bool step2_supported(const StepArgs<T> &a);
template <typename T> const char *step_kernel_name(const StepArgs<T> &a);
// the whole slab (face rows first) + the halo stores + step flags in ONE launch per step
template <typename T> int launch_slab_p2p(const StepArgs<T> &a, cudaStream_t s);
template <typename T> bool slab_p2p_supported(const StepArgs<T> &a);
// the same for TWO steps per launch (face tile rows first)
template <typename T> int launch_slab_p2p2(const StepArgs<T> &a, cudaStream_t s);
// rows a two-step pass needs at each face of a slab (the tile height of step2_impl.cuh)
template <typename T> int step2_tile_rows();
// rows [row_begin, row_begin + rows) of the lattice from dense (pitch == W) fields of `rows` rows
template <typename T> int launch_init_equilibrium(const T *rho, const T *vx, const T *vy, T *dst, size_t plane,
                                                  int pitch, int W, int row_begin, int rows, const Consts<T> &k,
                                                  cudaStream_t s);
template <typename T> int launch_readout(const ReadoutArgs<T> &a, cudaStream_t s);
// partials: at least mass_partials_capacity() doubles; out: one double
int mass_partials_capacity();

// Device-side render.rs: RGBA8 image of a macroscopic field.
enum RenderMode : int { RENDER_DENSITY = 0, RENDER_SPEED = 1, RENDER_VELOCITY = 2, RENDER_MOMENTUM = 3 };
// pass 1: sums[0] = sum, sums[1] = sum of squares of the displayed scalar over this slab (the field
// itself for DENSITY/SPEED, vx^2+vy^2 for the vector modes); partials: 2*capacity doubles.
// launch_render_stats_finish turns (all-reduced) sums into stats = (mean, population stdev).
template <typename T> int launch_render_stats(const T *src, size_t plane, int pitch, int W, int H, int mode,
                                              double *partials, double *stats, cudaStream_t s);
int launch_render_stats_finish(const double *sums, double cells, double *stats, cudaStream_t s);
// pass 2: colour mapping (z-score -> logistic -> HSV -> RGB) + geometry overlay -> rgba[y*W+x]
template <typename T> int launch_render_image(const T *src, size_t plane, int pitch, int W, int H, int mode,
                                              const double *stats, const uint8_t *mask, int mask_pitch,
                                              uchar4 *rgba, cudaStream_t s);
template <typename T> int launch_total_mass(const T *src, size_t plane, int pitch, int W, int H, double *partials,
                                            double *out, cudaStream_t s);
template <typename T> int launch_is_unstable(const T *src, size_t plane, int pitch, int W, int H, const Consts<T> &k,
                                             int *flag, cudaStream_t s);
constexpr int MASK_SEGMENT = 64;   // cells per mask-flag byte
// recompute the segment flags of rows [row_begin, row_begin + rows); *any |= 1 if a solid cell exists there
int launch_mask_flags(const uint8_t *mask, int mask_pitch, int W, int row_begin, int rows, uint8_t *flags,
                      int flag_pitch, int *any, cudaStream_t s);

// reverse a dense device buffer of n elements of elem_bytes (1, 4 or 8) bytes in place
int launch_reverse(void *data, size_t n, int elem_bytes, cudaStream_t s);
// set mask cells [y0, y0+h) x [x0, x0+w) (already clipped to the slab) to `value`
int launch_paint_rect(uint8_t *mask, int mask_pitch, int x0, int y0, int w, int h, uint8_t value, cudaStream_t s);

}  // namespace chemsim
