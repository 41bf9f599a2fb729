// kernels.cu — sm_100a kernels of the D2Q9 path other than the fused step itself, and the
// dispatch of the step launchers.
//
//  The step kernels (step_vec_kernel, step_slab_p2p_kernel, step_face_p2p_kernel,
//  step_scalar_kernel) live in step_impl.cuh and are instantiated per collision operator in
//  step_bgk.cu / step_trt.cu / step_regularized.cu / step_kbc.cu.
//  Here: readout / mass / unstable / init kernels for the macroscopic surface of lbm.rs
//  (:117-160, :779-818, :43-71), the segment flags of the geometry mask, the device-side render.
#include <cstdlib>

#include "step_decl.cuh"

namespace chemsim {

namespace {

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
        dst[(size_t)q * plane + (size_t)(y + GHOST) * pitch + x] = equilibrium_i(q, r, ux, uy, v2, k);
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
    for (int q = 0; q < Q; ++q) g[q] = a.src[(size_t)q * a.plane + (size_t)(y + GHOST) * a.pitch + x];
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
        const T *p = src + (size_t)(y + GHOST) * pitch + x;
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
        for (int q = 0; q < Q; ++q) g[q] = src[(size_t)q * plane + (size_t)(y + GHOST) * pitch + x];
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

// Live geometry edit (src/main.rs:71-91 rewrites the mask on the host for every mouse event):
// set the cells of a rectangle, one thread per cell.
__global__ void paint_rect_kernel(uint8_t *mask, int mask_pitch, int x0, int y0, int w, int h, uint8_t value)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < w && y < h) mask[(size_t)(y0 + y) * mask_pitch + x0 + x] = value;
}

// Element-wise reversal of a dense buffer of n elements of `size` bytes (1, 4, 8): the point
// reflection (y, x) <-> (H-1-y, W-1-x) of a row-major field is the reversal of its flat array.
// Used at the host boundary when the mirrored stream convention is selected.
template <typename E>
__global__ void reverse_kernel(E *data, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 2; i += (size_t)gridDim.x * blockDim.x) {
        const E a = data[i], b = data[n - 1 - i];
        data[i] = b;
        data[n - 1 - i] = a;
    }
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
        for (int q = 0; q < Q; ++q) g[q] = src[(size_t)q * plane + (size_t)(y + GHOST) * pitch + x];
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
    for (int q = 0; q < Q; ++q) g[q] = src[(size_t)q * plane + (size_t)(y + GHOST) * pitch + x];
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

}  // namespace

template <typename T>
const char *step_kernel_name(const StepArgs<T> &a)
{
    if (!use_vec(a)) return sizeof(T) == 4 ? "step_scalar_kernel<float>" : "step_scalar_kernel<double>";
    return sizeof(T) == 4 ? "step_vec_kernel<float>" : "step_vec_kernel<double>";
}

// one block per row chunk needs full 256-thread rows; narrower lattices use the
// two-stream face/interior path instead
template <typename T>
bool slab_p2p_supported(const StepArgs<T> &a)
{
    return use_vec(a) && a.W / VecOf<T>::N >= STEP_THREADS && a.H >= 4;
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

// Two steps per pass: vector widths only; the tile machinery pays off from a few tile rows on
// (CHEMSIM_LBM_STEP2=0 in the environment keeps every step on the single-step kernels).
template <typename T>
bool step2_supported(const StepArgs<T> &a)
{
    static const bool on = [] { const char *e = getenv("CHEMSIM_LBM_STEP2"); return !(e && e[0] == '0'); }();
    return on && a.collision != COL_KBC && use_vec(a) && a.H >= 4 && a.W >= 4 * VecOf<T>::N;
}

template <typename T>
int step2_tile_rows() { return Step2Tile<T>::TY; }

template <typename T>
int launch_slab_p2p2(const StepArgs<T> &a, cudaStream_t s)
{
    if (!slab_p2p_supported(a) || a.H < 2 * Step2Tile<T>::TY) return -(int)cudaErrorInvalidValue;
    switch (a.collision) {
    case COL_BGK:         launch_slab_p2p2_col<T, COL_BGK>(a, s); break;
    case COL_TRT:         launch_slab_p2p2_col<T, COL_TRT>(a, s); break;
    case COL_REGULARIZED: launch_slab_p2p2_col<T, COL_REGULARIZED>(a, s); break;
    default: return -(int)cudaErrorInvalidValue;
    }
    const int e = check_launch();
    return e ? e : 1;
}

template <typename T>
int launch_step2(const StepArgs<T> &a, cudaStream_t s)
{
    if (a.y_count <= 0) return 0;
    if (!use_vec(a)) return -(int)cudaErrorInvalidValue;
    switch (a.collision) {
    case COL_BGK:         launch_step2_col<T, COL_BGK>(a, s); break;
    case COL_TRT:         launch_step2_col<T, COL_TRT>(a, s); break;
    case COL_REGULARIZED: launch_step2_col<T, COL_REGULARIZED>(a, s); break;
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

int launch_reverse(void *data, size_t n, int elem_bytes, cudaStream_t s)
{
    if (n < 2) return 0;
    const int blocks = reduction_blocks(n / 2);
    if (elem_bytes == 1)      reverse_kernel<uint8_t><<<blocks, RED_THREADS, 0, s>>>((uint8_t *)data, n);
    else if (elem_bytes == 4) reverse_kernel<uint32_t><<<blocks, RED_THREADS, 0, s>>>((uint32_t *)data, n);
    else if (elem_bytes == 8) reverse_kernel<unsigned long long><<<blocks, RED_THREADS, 0, s>>>((unsigned long long *)data, n);
    else return -(int)cudaErrorInvalidValue;
    const int e = check_launch();
    return e ? e : 1;
}

int launch_paint_rect(uint8_t *mask, int mask_pitch, int x0, int y0, int w, int h, uint8_t value, cudaStream_t s)
{
    const dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    paint_rect_kernel<<<grid, block, 0, s>>>(mask, mask_pitch, x0, y0, w, h, value);
    const int e = check_launch();
    return e ? e : 1;
}

#define CHEMSIM_INSTANTIATE(T)                                                                                       \
    template int launch_step<T>(const StepArgs<T> &, cudaStream_t);                                                  \
    template int launch_step2<T>(const StepArgs<T> &, cudaStream_t);                                                 \
    template bool step2_supported<T>(const StepArgs<T> &);                                                           \
    template int launch_slab_p2p2<T>(const StepArgs<T> &, cudaStream_t);                                             \
    template int step2_tile_rows<T>();                                                                               \
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
