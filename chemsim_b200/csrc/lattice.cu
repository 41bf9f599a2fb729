// lattice.cu — host runtime behind the C ABI of include/chemsim_lbm.h.
//
// Owns the device-resident State (two population buffers for A-B stepping, the
// geometry mask, staging for readouts), the compute and halo streams and, for a
// sharded lattice, the NCCL communicator.  No CPU fallback: every compute entry
// point needs a CUDA device.
#include "../../include/chemsim_lbm.h"
#include "kernels.cuh"
#include "consts.hpp"
#include "nccl_dyn.h"

#include <cmath>
#include <ctime>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

using namespace chemsim;

namespace {

thread_local std::string g_create_error = "";

struct Scalars {   // host scalars in both dtypes, rebuilt whenever dx/dt/tau change
    Consts<float> f;
    Consts<double> d;
};

}  // namespace

struct chemsim_lbm {
    int W = 0, H = 0, Hglobal = 0, row0 = 0;
    int dtype = 0, edge = 0, device = 0;
    int rank = 0, nranks = 1;
    size_t esize = 4;
    double dx = 1.0, dt = 1.0;
    CollisionParams col;
    Scalars k;

    void *buf[2] = {nullptr, nullptr};
    int cur = 0;
    size_t plane = 0;   // elements
    int pitch = 0;      // elements
    uint8_t *mask_alloc = nullptr;   // H + 2 rows: one halo row above and below (a y-slab's neighbours' face rows)
    uint8_t *mask = nullptr;         // row 0 of the slab inside mask_alloc
    int mask_pitch = 0;
    int has_mask = 0;          // 0 = known solid-free (mask never read); 1 = consult the segment flags
    uint8_t *mask_flags = nullptr;
    int flag_pitch = 0;
    void *stage[3] = {nullptr, nullptr, nullptr};   // dense staging fields (grown on demand)
    size_t stage_bytes[3] = {0, 0, 0};
    void *snap[2] = {nullptr, nullptr};             // field snapshots (up to two planes each) for asynchronous readouts
    size_t snap_bytes[2] = {0, 0};
    int snap_next = 0;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_main = nullptr, ev_mask = nullptr, ev_snap_ready[2] = {nullptr, nullptr},
                ev_snap_done[2] = {nullptr, nullptr};
    double *d_partials = nullptr, *d_scalar = nullptr;
    int *d_flag = nullptr;
    double *h_scalar = nullptr;   // pinned
    int *h_flag = nullptr;        // pinned

    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t ev_face = nullptr, ev_interior = nullptr, ev_halo = nullptr;
    ncclComm_t comm = nullptr;
    bool ghosts_valid = false;
    // peer-memory halo (chemsim_lbm_enable_p2p_halo): the face kernel stores into the
    // neighbours' ghost rows directly; NCCL is then only used for the first exchange
    int halo_mode = CHEMSIM_LBM_HALO_NCCL;
    unsigned step_index = 0;                        // steps taken since creation (same on every rank)
    unsigned *p2p_flags = nullptr;                  // [0] from_up, [1] from_down, [2] done counter, [3] error (device copy)
    int *p2p_error_host = nullptr;                  // mapped host word the kernels set on a halo time-out
    int *p2p_error_dev = nullptr;                   // its device-side address
    double p2p_timeout_s = 30.0;                    // CHEMSIM_LBM_P2P_TIMEOUT_S / chemsim_lbm_set_p2p_timeout
    int stream_mirrored = 0;                        // chemsim_lbm_set_stream_convention
    void *peer_up_buf[2] = {nullptr, nullptr}, *peer_down_buf[2] = {nullptr, nullptr};
    unsigned *peer_up_flags = nullptr, *peer_down_flags = nullptr;
    int peer_up_H = 0, peer_down_H = 0;
    bool peer_same = false;                         // both neighbours are the same rank (nranks == 2)
    bool peer_up_ipc = false, peer_down_ipc = false;   // mapped through cudaIpc (other process) or raw (same process)
    bool have_populations = false;

    float time_f = 0.f;
    double time_d = 0.0;
    uint64_t launches = 0;
    std::string err = "";
};

namespace {

int fail(chemsim_lbm *h, int code, const std::string &msg)
{
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

#define CUDA_TRY(h, expr)                                                                      \
    do {                                                                                       \
        const cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess)                                                                 \
            return fail(h, CHEMSIM_LBM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

#define NCCL_TRY(h, expr)                                                                      \
    do {                                                                                       \
        const ncclResult_t r_ = (expr);                                                        \
        if (r_ != ncclSuccess)                                                                 \
            return fail(h, CHEMSIM_LBM_ERR_NCCL, std::string(#expr) + ": " + nccl_dyn().GetErrorString(r_)); \
    } while (0)

#define LAUNCH_TRY(h, expr)                                                                    \
    do {                                                                                       \
        const int n_ = (expr);                                                                 \
        if (n_ < 0)                                                                            \
            return fail(h, CHEMSIM_LBM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString((cudaError_t)(-n_))); \
        (h)->launches += (uint64_t)n_;                                                         \
    } while (0)

int bind_device(chemsim_lbm *h) { CUDA_TRY(h, cudaSetDevice(h->device)); return 0; }

#define BIND(h) do { const int b_ = bind_device(h); if (b_) return b_; } while (0)

int check_n(chemsim_lbm *h, size_t n)
{
    if (n != (size_t)h->W * (size_t)h->H)
        return fail(h, CHEMSIM_LBM_ERR_INVALID_SLICE_SIZE,
                    "slice has " + std::to_string(n) + " elements, lattice slab is " + std::to_string(h->W) + "x" +
                        std::to_string(h->H));
    return 0;
}

void rebuild_scalars(chemsim_lbm *h)
{
    h->k.f = make_consts<float>(h->dx, h->dt, h->col);
    h->k.d = make_consts<double>(h->dx, h->dt, h->col);
}

char *row_ptr(chemsim_lbm *h, int b, int q, int y)
{
    return (char *)h->buf[b] + ((size_t)q * h->plane + (size_t)(y + GHOST) * h->pitch) * h->esize;
}

// two-row halo (all slabs have >= 2 rows) or the one-row form; see build_halo_plan
bool deep_halo_of(int global_height, int nranks) { return nranks > 1 && global_height / nranks >= 2; }
bool deep_halo(const chemsim_lbm *h) { return deep_halo_of(h->Hglobal, h->nranks); }

template <typename T> const Consts<T> &consts_of(const chemsim_lbm *h);
template <> const Consts<float> &consts_of<float>(const chemsim_lbm *h) { return h->k.f; }
template <> const Consts<double> &consts_of<double>(const chemsim_lbm *h) { return h->k.d; }

template <typename T>
StepArgs<T> step_args(const chemsim_lbm *h, int y_begin, int y_count, int y_stride = 1)
{
    StepArgs<T> a;
    a.src = (const T *)h->buf[h->cur];
    a.dst = (T *)h->buf[h->cur ^ 1];
    a.plane = h->plane;
    a.pitch = h->pitch;
    a.W = h->W;
    a.H = h->H;
    a.y_begin = y_begin;
    a.y_count = y_count;
    a.y_stride = y_stride;
    a.xchunks = 1;
    a.ghost = GHOST;
    a.row0 = h->row0;
    a.Hglobal = h->Hglobal;
    a.periodic_y = h->edge == CHEMSIM_LBM_EDGE_PERIODIC ? 1 : 0;
    a.wrap_y = (h->edge == CHEMSIM_LBM_EDGE_PERIODIC && h->nranks == 1) ? 1 : 0;
    a.periodic_x = h->edge == CHEMSIM_LBM_EDGE_PERIODIC ? 1 : 0;
    a.mask = h->mask;
    a.mask_pitch = h->mask_pitch;
    a.has_mask = h->has_mask;
    a.ghost_mask = (h->nranks > 1 && deep_halo(h)) ? 1 : 0;
    a.collision = h->col.kind;
    a.mask_flags = h->mask_flags;
    a.flag_pitch = h->flag_pitch;
    a.prefetch_tiles = a.prefetch_rows = a.prefetch_cols = 0;   // set by the two-step launchers
    a.k = consts_of<T>(h);
    a.fill_offsets();
    return a;
}

// Halo plan.  What a slab delivers per exchange is what the neighbour's next pass —
// one step or two (step2_impl.cuh) — reads from its two ghost rows:
//   to the lower neighbour (larger y): my last row, all nine populations -> its ghost row -1, and the
//       dy=+1 movers q = 3,6,7 of my last-but-one row -> its ghost row -2;
//   to the upper neighbour: my first row, all nine -> its ghost row H, and the dy=-1 movers q = 1,5,8 of
//       my second row -> its ghost row H+1.
// When some slab has a single row (global_height / nranks < 2) the plan is the one-row form (three
// populations per face: all a single step needs) and two-step passes are off.
// Sends are issued [to-lower, to-upper] and receives [from-upper, from-lower], so that when both
// neighbours are the same rank (nranks == 2, periodic) the k-th send to a peer matches its k-th receive.
int build_halo_plan(int global_height, int rank, int nranks, int edge, chemsim_lbm_halo_msg *out)
{
    if (nranks <= 1) return 0;
    const bool periodic = edge == CHEMSIM_LBM_EDGE_PERIODIC, deep = deep_halo_of(global_height, nranks);
    const int up = (rank + nranks - 1) % nranks, down = (rank + 1) % nranks;
    const bool has_up = periodic || rank > 0, has_down = periodic || rank < nranks - 1;
    static const int to_down[3] = {3, 6, 7}, to_up[3] = {1, 5, 8};
    int n = 0;
    auto face = [&](int is_send, int peer, int row_all, int row_movers, const int (&movers)[3]) {
        if (deep) {
            for (int q = 0; q < Q; ++q) out[n++] = {is_send, peer, q, row_all};
            for (int q : movers) out[n++] = {is_send, peer, q, row_movers};
        } else {
            for (int q : movers) out[n++] = {is_send, peer, q, row_all};
        }
    };
    if (has_down) face(1, down, CHEMSIM_LBM_ROW_LAST, CHEMSIM_LBM_ROW_SECOND_LAST, to_down);
    if (has_up)   face(1, up, CHEMSIM_LBM_ROW_FIRST, CHEMSIM_LBM_ROW_SECOND, to_up);
    if (has_up)   face(0, up, CHEMSIM_LBM_ROW_GHOST_ABOVE, CHEMSIM_LBM_ROW_GHOST_ABOVE2, to_down);
    if (has_down) face(0, down, CHEMSIM_LBM_ROW_GHOST_BELOW, CHEMSIM_LBM_ROW_GHOST_BELOW2, to_up);
    return n;
}

int plan_row(const chemsim_lbm *h, int row)
{
    switch (row) {
    case CHEMSIM_LBM_ROW_FIRST: return 0;
    case CHEMSIM_LBM_ROW_SECOND: return 1;
    case CHEMSIM_LBM_ROW_LAST: return h->H - 1;
    case CHEMSIM_LBM_ROW_SECOND_LAST: return h->H - 2;
    case CHEMSIM_LBM_ROW_GHOST_ABOVE: return -1;
    case CHEMSIM_LBM_ROW_GHOST_ABOVE2: return -2;
    case CHEMSIM_LBM_ROW_GHOST_BELOW: return h->H;
    default: return h->H + 1;
    }
}

// Rows of one population are contiguous, so the sends/receives work directly on
// the lattice buffers (no packing); one ncclGroup = one fused NCCL kernel.
int exchange(chemsim_lbm *h, int b)
{
    chemsim_lbm_halo_msg plan[CHEMSIM_LBM_HALO_PLAN_MAX];
    const int count = build_halo_plan(h->Hglobal, h->rank, h->nranks, h->edge, plan);
    const size_t bytes = (size_t)h->W * h->esize;
    const NcclDyn &n = nccl_dyn();
    NCCL_TRY(h, n.GroupStart());
    for (int i = 0; i < count; ++i) {
        const chemsim_lbm_halo_msg &m = plan[i];
        const int y = plan_row(h, m.row);
        if (m.is_send) NCCL_TRY(h, n.Send(row_ptr(h, b, m.q, y), bytes, ncclChar, m.peer, h->comm, h->comm_stream));
        else           NCCL_TRY(h, n.Recv(row_ptr(h, b, m.q, y), bytes, ncclChar, m.peer, h->comm, h->comm_stream));
    }
    NCCL_TRY(h, n.GroupEnd());
    h->launches += 1;   // one fused NCCL send/recv kernel per group
    return 0;
}

// The geometry's halo rows: my first / last mask row goes to the upper / lower neighbour's halo row
// (what its two-step pass applies to the ghost-row cells it recomputes).  Runs before every batch of
// steps — the mask may have been edited on any rank since the last one, and only a collective that
// every rank always issues needs no agreement about that.
int exchange_mask(chemsim_lbm *h)
{
    const bool periodic = h->edge == CHEMSIM_LBM_EDGE_PERIODIC;
    const int up = (h->rank + h->nranks - 1) % h->nranks, down = (h->rank + 1) % h->nranks;
    const bool has_up = periodic || h->rank > 0, has_down = periodic || h->rank < h->nranks - 1;
    const NcclDyn &n = nccl_dyn();
    uint8_t *first = h->mask, *last = h->mask + (size_t)(h->H - 1) * h->mask_pitch;
    uint8_t *above = h->mask - h->mask_pitch, *below = h->mask + (size_t)h->H * h->mask_pitch;
    NCCL_TRY(h, n.GroupStart());
    if (has_down) NCCL_TRY(h, n.Send(last, h->W, ncclChar, down, h->comm, h->comm_stream));
    if (has_up)   NCCL_TRY(h, n.Send(first, h->W, ncclChar, up, h->comm, h->comm_stream));
    if (has_up)   NCCL_TRY(h, n.Recv(above, h->W, ncclChar, up, h->comm, h->comm_stream));
    if (has_down) NCCL_TRY(h, n.Recv(below, h->W, ncclChar, down, h->comm, h->comm_stream));
    NCCL_TRY(h, n.GroupEnd());
    h->launches += 1;
    return 0;
}

// Arguments of the P2P face kernel for the step that writes buffer cur^1.
void fill_halo(const chemsim_lbm *h, HaloP2P &p)
{
    const bool periodic = h->edge == CHEMSIM_LBM_EDGE_PERIODIC;
    const bool has_up = periodic || h->rank > 0, has_down = periodic || h->rank < h->nranks - 1;
    const int b = h->cur ^ 1;
    p = HaloP2P();
    if (has_up) {
        p.up_dst = h->peer_up_buf[b];
        p.up_plane = (size_t)(h->peer_up_H + 2 * GHOST) * h->pitch;
        p.up_row0 = h->peer_up_H + GHOST;
        p.wait_up = h->p2p_flags + 0;
        p.signal_up = h->peer_up_flags + 1;          // I am the upper neighbour's lower neighbour
    }
    if (has_down) {
        p.down_dst = h->peer_down_buf[b];
        p.down_plane = (size_t)(h->peer_down_H + 2 * GHOST) * h->pitch;
        p.wait_down = h->p2p_flags + 1;
        p.signal_down = h->peer_down_flags + 0;      // I am the lower neighbour's upper neighbour
    }
    p.done = h->p2p_flags + 2;
    p.error = (int *)(h->p2p_flags + 3);
    p.error_host = h->p2p_error_dev;
    p.timeout_ns = (unsigned long long)(h->p2p_timeout_s * 1e9);
    p.step = h->step_index;
}

// A halo time-out is sticky (until the next upload): every call that produces or consumes
// lattice data reports it instead of returning numbers computed from a stale halo.
int check_p2p_error(chemsim_lbm *h)
{
    if (h->p2p_error_host && *(volatile int *)h->p2p_error_host)
        return fail(h, CHEMSIM_LBM_ERR_CUDA,
                    "peer-memory halo: a neighbour did not publish its face rows within the time-out; the lattice "
                    "holds a stale halo (upload new populations on every rank to recover)");
    return 0;
}
#define P2P_CHECK(h) do { const int e_ = check_p2p_error(h); if (e_) return e_; } while (0)

struct P2PInfo {                 // what the ranks tell each other (all-gathered over NCCL)
    cudaIpcMemHandle_t buf[2];
    cudaIpcMemHandle_t flags;
    int H;
    int ok;
    // ranks that live in the SAME process (one host thread per GPU) cannot open each other's
    // IPC handles; they use the raw device pointers (unified addressing) + peer access
    long long pid;
    int device;
    int pad0;
    void *raw_buf[2];
    void *raw_flags;
    char pad[512 - 3 * sizeof(cudaIpcMemHandle_t) - 4 * sizeof(int) - sizeof(long long) - 3 * sizeof(void *)];
};
static_assert(sizeof(P2PInfo) == 512, "P2PInfo layout");

// Map one neighbour's population buffers and flag block.  Returns false on failure.
bool map_peer(chemsim_lbm *h, const P2PInfo &peer, void *(&buf)[2], unsigned *&flags, bool &via_ipc)
{
    if (peer.pid == (long long)getpid()) {          // same process: direct peer pointers
        via_ipc = false;
        // Two slabs on ONE device cannot use the flag handshake: with programmatic dependent launch the
        // next step's blocks of one slab may occupy the SM slots the other slab's face blocks need
        // while they spin on that slab's flag.  Such lattices keep the NCCL exchange.
        if (peer.device == h->device) return false;
        {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, h->device, peer.device) != cudaSuccess || !can) return false;
            const cudaError_t e = cudaDeviceEnablePeerAccess(peer.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return false; }
            cudaGetLastError();
        }
        buf[0] = peer.raw_buf[0]; buf[1] = peer.raw_buf[1];
        flags = (unsigned *)peer.raw_flags;
        return true;
    }
    via_ipc = true;
    const unsigned fl = cudaIpcMemLazyEnablePeerAccess;
    for (int b = 0; b < 2; ++b)
        if (cudaIpcOpenMemHandle(&buf[b], peer.buf[b], fl) != cudaSuccess) return false;
    return cudaIpcOpenMemHandle((void **)&flags, peer.flags, fl) == cudaSuccess;
}

void close_p2p(chemsim_lbm *h)
{
    for (int b = 0; b < 2; ++b) {
        if (h->peer_up_buf[b] && h->peer_up_ipc) cudaIpcCloseMemHandle(h->peer_up_buf[b]);
        if (h->peer_down_buf[b] && h->peer_down_ipc && !h->peer_same) cudaIpcCloseMemHandle(h->peer_down_buf[b]);
        h->peer_up_buf[b] = h->peer_down_buf[b] = nullptr;
    }
    if (h->peer_up_flags && h->peer_up_ipc) cudaIpcCloseMemHandle(h->peer_up_flags);
    if (h->peer_down_flags && h->peer_down_ipc && !h->peer_same) cudaIpcCloseMemHandle(h->peer_down_flags);
    h->peer_up_flags = h->peer_down_flags = nullptr;
    cudaGetLastError();
}

// Collective over all ranks of the lattice.  Every rank ends in the same mode.
int enable_p2p(chemsim_lbm *h)
{
    if (h->nranks == 1 || h->halo_mode == CHEMSIM_LBM_HALO_P2P) return 0;
    const NcclDyn &n = nccl_dyn();
    const bool periodic = h->edge == CHEMSIM_LBM_EDGE_PERIODIC;
    const int up = (h->rank + h->nranks - 1) % h->nranks, down = (h->rank + 1) % h->nranks;
    const bool has_up = periodic || h->rank > 0, has_down = periodic || h->rank < h->nranks - 1;
    h->peer_same = has_up && has_down && up == down;

    P2PInfo mine;
    std::memset(&mine, 0, sizeof(mine));
    mine.H = h->H;
    mine.ok = 1;
    mine.pid = (long long)getpid();
    mine.device = h->device;
    if (!h->p2p_flags) {
        if (cudaMalloc((void **)&h->p2p_flags, 4 * sizeof(unsigned)) != cudaSuccess) mine.ok = 0;
        else if (cudaMemset(h->p2p_flags, 0, 4 * sizeof(unsigned)) != cudaSuccess) mine.ok = 0;
    }
    if (!h->p2p_error_host) {
        if (cudaHostAlloc((void **)&h->p2p_error_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void **)&h->p2p_error_dev, h->p2p_error_host, 0) != cudaSuccess) {
            mine.ok = 0;
            cudaGetLastError();
        } else {
            *h->p2p_error_host = 0;
        }
        if (const char *e = getenv("CHEMSIM_LBM_P2P_TIMEOUT_S")) {
            const double t = atof(e);
            if (t > 0.0) h->p2p_timeout_s = t;
        }
    }
    // the fused slab kernels need full 256-thread row chunks, >= 4 rows and the two-row halo;
    // narrow, ragged or one-row slabs keep the NCCL exchange
    const bool slab_ok = h->dtype == CHEMSIM_LBM_F32 ? slab_p2p_supported(step_args<float>(h, 0, h->H))
                                                     : slab_p2p_supported(step_args<double>(h, 0, h->H));
    if (!slab_ok || !deep_halo(h)) mine.ok = 0;
    if (mine.ok && (cudaIpcGetMemHandle(&mine.buf[0], h->buf[0]) != cudaSuccess ||
                    cudaIpcGetMemHandle(&mine.buf[1], h->buf[1]) != cudaSuccess ||
                    cudaIpcGetMemHandle(&mine.flags, h->p2p_flags) != cudaSuccess)) {
        mine.ok = 0;
        cudaGetLastError();
    }
    mine.raw_buf[0] = h->buf[0]; mine.raw_buf[1] = h->buf[1]; mine.raw_flags = h->p2p_flags;
    // all-gather the handles
    char *d_all = nullptr;
    CUDA_TRY(h, cudaMalloc((void **)&d_all, sizeof(P2PInfo) * (h->nranks + 1)));
    CUDA_TRY(h, cudaMemcpyAsync(d_all + sizeof(P2PInfo) * h->nranks, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
    NCCL_TRY(h, n.AllGather(d_all + sizeof(P2PInfo) * h->nranks, d_all, sizeof(P2PInfo), ncclChar, h->comm, h->stream));
    std::string all(sizeof(P2PInfo) * h->nranks, '\0');
    CUDA_TRY(h, cudaMemcpyAsync(&all[0], d_all, all.size(), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const P2PInfo *info = reinterpret_cast<const P2PInfo *>(all.data());
    int ok = 1;
    for (int r = 0; r < h->nranks; ++r) ok &= info[r].ok;
    // map the neighbours' buffers
    if (ok) {
        if (has_up) {
            h->peer_up_H = info[up].H;
            if (!map_peer(h, info[up], h->peer_up_buf, h->peer_up_flags, h->peer_up_ipc)) ok = 0;
        }
        if (has_down && ok) {
            h->peer_down_H = info[down].H;
            if (h->peer_same) {
                h->peer_down_buf[0] = h->peer_up_buf[0]; h->peer_down_buf[1] = h->peer_up_buf[1];
                h->peer_down_flags = h->peer_up_flags;
                h->peer_down_ipc = h->peer_up_ipc;
            } else if (!map_peer(h, info[down], h->peer_down_buf, h->peer_down_flags, h->peer_down_ipc)) {
                ok = 0;
            }
        }
        if (!ok) cudaGetLastError();
    }
    // agree: P2P only if every rank mapped its neighbours
    int *d_ok = reinterpret_cast<int *>(d_all);
    CUDA_TRY(h, cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    NCCL_TRY(h, n.AllReduce(d_ok, d_ok + 1, 1, ncclInt, ncclMin, h->comm, h->stream));
    int all_ok = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&all_ok, d_ok + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaFree(d_all));
    if (all_ok) {
        // flags restart from the current step on every rank: the first P2P step reads ghost
        // rows that the NCCL exchange of begin_sharded() delivers
        const unsigned init[4] = {h->step_index, h->step_index, 0u, 0u};
        CUDA_TRY(h, cudaMemcpy(h->p2p_flags, init, sizeof(init), cudaMemcpyHostToDevice));
        h->ghosts_valid = false;
        h->halo_mode = CHEMSIM_LBM_HALO_P2P;
        // nobody may publish into my flags before they are initialised
        NCCL_TRY(h, n.AllReduce(h->d_scalar, h->d_scalar + 1, 1, ncclDouble, ncclSum, h->comm, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    } else {
        close_p2p(h);
    }
    return 0;
}

// Sharded stepping runs on two streams:
//   stream      (normal priority)  interior rows 1 … H−2, which read no ghost row
//   comm_stream (highest priority) the two face rows, then the NCCL exchange of them
// Step n on either stream depends only on step n−1 of the other, so the face-row
// kernel and the exchange of step n overlap with the interior kernel of step n:
//   face(n)     after interior(n−1)  (reads rows 1, H−2)   and exchange(n−1) (same stream)
//   interior(n) after face(n−1)      (reads rows 0, H−1)   and interior(n−1) (same stream)
// (interior(n+1) overwrites rows 1 … H−2 of the buffer exchange(n−1) works on, but
// the exchange only touches rows 0, H−1 and the ghost rows — disjoint.)
// The high priority lets the small face/NCCL kernels take SM slots as soon as blocks
// of the big interior kernel retire instead of queueing behind its whole grid.

// Make the ghost rows of the current buffer valid (after an upload) and order the
// halo stream after everything queued on the main stream so far.
int begin_sharded(chemsim_lbm *h)
{
    CUDA_TRY(h, cudaEventRecord(h->ev_interior, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->comm_stream, h->ev_interior, 0));
    if (deep_halo(h)) {
        const int r = exchange_mask(h);
        if (r) return r;
    }
    if (!h->ghosts_valid) {
        if (h->halo_mode == CHEMSIM_LBM_HALO_P2P) {
            // New populations (upload / restore) on every rank: restart the flag handshake from the
            // current step and clear a previous time-out.  My own kernels have finished (uploads
            // synchronise the stream); a neighbour cannot publish a NEWER step into these flags
            // before it has passed the NCCL exchange below, which needs this rank to get there too.
            const unsigned init[4] = {h->step_index, h->step_index, 0u, 0u};
            CUDA_TRY(h, cudaMemcpy(h->p2p_flags, init, sizeof(init), cudaMemcpyHostToDevice));
            *(volatile int *)h->p2p_error_host = 0;
        }
        const int r = exchange(h, h->cur);
        if (r) return r;
        h->ghosts_valid = true;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_face, h->comm_stream));
    return 0;
}

template <typename T>
int step_impl(chemsim_lbm *h, int nsteps)
{
    if (nsteps == 0) return 0;
    if (h->nranks == 1) {
        int left = nsteps;
        // pairs of steps in one pass over HBM (temporal blocking, step2_impl.cuh); an odd step on its own
        while (left >= 2 && step2_supported(step_args<T>(h, 0, h->H))) {
            LAUNCH_TRY(h, launch_step2<T>(step_args<T>(h, 0, h->H), h->stream));
            h->cur ^= 1;
            h->step_index += 2;
            left -= 2;
        }
        for (; left > 0; --left) {
            LAUNCH_TRY(h, launch_step<T>(step_args<T>(h, 0, h->H), h->stream));
            h->cur ^= 1;
            h->step_index += 1;
        }
        return 0;
    }
    const int r0 = begin_sharded(h);
    if (r0) return r0;
    // Two steps per pass when every slab is at least two tiles tall (the same decision on every rank:
    // it depends on the global shape only), an odd step on its own at the end.
    const int TY = step2_tile_rows<T>();
    const bool deep = deep_halo(h);
    const bool can2 = deep && h->Hglobal / h->nranks >= 2 * TY && step2_supported(step_args<T>(h, 0, h->H));
    int left = nsteps;
    if (h->halo_mode == CHEMSIM_LBM_HALO_P2P) {
        // peer-memory mode: the whole step (or double step) — face rows, halo stores into the
        // neighbours' ghost rows, step flags, interior — is ONE kernel on the main stream
        CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_face, 0));   // the first exchange (begin_sharded)
        while (left > 0) {
            const int n = (left >= 2 && can2) ? 2 : 1;
            StepArgs<T> all = step_args<T>(h, 0, h->H);
            fill_halo(h, all.halo);
            if (n == 2) LAUNCH_TRY(h, launch_slab_p2p2<T>(all, h->stream));
            else        LAUNCH_TRY(h, launch_slab_p2p<T>(all, h->stream));
            h->cur ^= 1;
            h->step_index += n;
            left -= n;
        }
        return 0;
    }
    while (left > 0) {
        const int n = (left >= 2 && can2) ? 2 : 1;
        // both waits refer to the events recorded for the previous pass (or by begin_sharded)
        CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_face, 0));
        CUDA_TRY(h, cudaStreamWaitEvent(h->comm_stream, h->ev_interior, 0));
        // halo stream: the face rows of this pass, then ship them; main stream: the interior
        if (n == 2) {
            LAUNCH_TRY(h, launch_step2<T>(step_args<T>(h, 0, TY), h->comm_stream));
            LAUNCH_TRY(h, launch_step2<T>(step_args<T>(h, h->H - TY, TY), h->comm_stream));
            CUDA_TRY(h, cudaEventRecord(h->ev_face, h->comm_stream));
            const int r = exchange(h, h->cur ^ 1);
            if (r) return r;
            if (h->H > 2 * TY) LAUNCH_TRY(h, launch_step2<T>(step_args<T>(h, TY, h->H - 2 * TY), h->stream));
        } else {
            const int face = deep ? 2 : 1;           // rows per face the neighbours receive
            if (h->H <= 2 * face) {                  // every row is a face row
                LAUNCH_TRY(h, launch_step<T>(step_args<T>(h, 0, h->H), h->comm_stream));
            } else {
                LAUNCH_TRY(h, launch_step<T>(step_args<T>(h, 0, face), h->comm_stream));
                LAUNCH_TRY(h, launch_step<T>(step_args<T>(h, h->H - face, face), h->comm_stream));
            }
            CUDA_TRY(h, cudaEventRecord(h->ev_face, h->comm_stream));
            const int r = exchange(h, h->cur ^ 1);
            if (r) return r;
            if (h->H > 2 * face) LAUNCH_TRY(h, launch_step<T>(step_args<T>(h, face, h->H - 2 * face), h->stream));
        }
        CUDA_TRY(h, cudaEventRecord(h->ev_interior, h->stream));
        h->cur ^= 1;
        h->step_index += n;
        left -= n;
    }
    // later work on the main stream (readouts, uploads) sees the finished halo work
    CUDA_TRY(h, cudaEventRecord(h->ev_halo, h->comm_stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
    return 0;
}

int ensure_stage(chemsim_lbm *h, int count, size_t bytes)
{
    for (int i = 0; i < count; ++i) {
        if (h->stage_bytes[i] >= bytes) continue;
        if (h->stage[i]) { CUDA_TRY(h, cudaStreamSynchronize(h->stream)); CUDA_TRY(h, cudaFree(h->stage[i])); }
        h->stage[i] = nullptr; h->stage_bytes[i] = 0;
        CUDA_TRY(h, cudaMalloc(&h->stage[i], bytes));
        h->stage_bytes[i] = bytes;
    }
    return 0;
}

template <typename T>
ReadoutArgs<T> readout_args(const chemsim_lbm *h, int kind, int q, void *out0, void *out1)
{
    ReadoutArgs<T> a;
    a.src = (const T *)h->buf[h->cur];
    a.plane = h->plane; a.pitch = h->pitch; a.W = h->W; a.H = h->H;
    a.kind = kind; a.q = q;
    a.out0 = (T *)out0; a.out1 = (T *)out1;
    a.k = consts_of<T>(h);
    return a;
}

template <typename T>
int readout_impl(chemsim_lbm *h, int kind, int q, void *dst0, void *dst1)
{
    const size_t bytes = (size_t)h->W * h->H * sizeof(T);
    const int rs = ensure_stage(h, dst1 ? 2 : 1, bytes);
    if (rs) return rs;
    LAUNCH_TRY(h, launch_readout<T>(readout_args<T>(h, kind, q, h->stage[0], h->stage[1]), h->stream));
    if (h->stream_mirrored)
        for (int i = 0; i < (dst1 ? 2 : 1); ++i)
            LAUNCH_TRY(h, launch_reverse(h->stage[i], (size_t)h->W * h->H, (int)sizeof(T), h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(dst0, h->stage[0], bytes, cudaMemcpyDeviceToHost, h->stream));
    if (dst1) CUDA_TRY(h, cudaMemcpyAsync(dst1, h->stage[1], bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    P2P_CHECK(h);
    return 0;
}

// Asynchronous field snapshot (record/render "on demand"): the readout kernel runs
// on the main stream into one of two snapshot buffers, the device->host copy on its own stream,
// so the next steps overlap with the transfer.  A population is snapshotted with a strided
// device copy (the lattice buffer itself is overwritten two steps later).
template <typename T>
int snapshot_async_impl(chemsim_lbm *h, int field, int q, void *dst0, void *dst1)
{
    const size_t bytes = (size_t)h->W * h->H * sizeof(T);
    const bool two = dst1 != nullptr;
    const int slot = h->snap_next;
    h->snap_next ^= 1;
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_snap_done[slot], 0));   // previous copy out of this slot
    if (h->snap_bytes[slot] < (two ? 2 : 1) * bytes) {
        if (h->snap[slot]) {
            CUDA_TRY(h, cudaStreamSynchronize(h->d2h_stream));
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            CUDA_TRY(h, cudaFree(h->snap[slot]));
            h->snap[slot] = nullptr; h->snap_bytes[slot] = 0;
        }
        CUDA_TRY(h, cudaMalloc(&h->snap[slot], 2 * bytes));
        h->snap_bytes[slot] = 2 * bytes;
    }
    char *s0 = (char *)h->snap[slot], *s1 = s0 + bytes;
    if (field == CHEMSIM_LBM_FIELD_POPULATION) {
        CUDA_TRY(h, cudaMemcpy2DAsync(s0, (size_t)h->W * sizeof(T), row_ptr(h, h->cur, q, 0), (size_t)h->pitch * sizeof(T),
                                      (size_t)h->W * sizeof(T), h->H, cudaMemcpyDeviceToDevice, h->stream));
    } else {
        static const int kind_of[] = {READ_DENSITY, READ_PRESSURE, READ_SPEED, READ_VELOCITY, READ_MOMENTUM, -1,
                                      READ_EQUILIBRIUM, READ_NON_EQUILIBRIUM};
        LAUNCH_TRY(h, launch_readout<T>(readout_args<T>(h, kind_of[field], q, s0, two ? s1 : nullptr), h->stream));
    }
    if (h->stream_mirrored)
        for (int i = 0; i < (two ? 2 : 1); ++i)
            LAUNCH_TRY(h, launch_reverse(i ? s1 : s0, (size_t)h->W * h->H, (int)sizeof(T), h->stream));
    CUDA_TRY(h, cudaEventRecord(h->ev_snap_ready[slot], h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->d2h_stream, h->ev_snap_ready[slot], 0));
    CUDA_TRY(h, cudaMemcpyAsync(dst0, s0, bytes, cudaMemcpyDeviceToHost, h->d2h_stream));
    if (two) CUDA_TRY(h, cudaMemcpyAsync(dst1, s1, bytes, cudaMemcpyDeviceToHost, h->d2h_stream));
    CUDA_TRY(h, cudaEventRecord(h->ev_snap_done[slot], h->d2h_stream));
    return 0;
}

int check_rows(chemsim_lbm *h, int row_begin, int row_count, size_t n)
{
    if (row_begin < 0 || row_count <= 0 || row_begin + row_count > h->H)
        return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "row range outside the slab");
    if (n != (size_t)h->W * (size_t)row_count)
        return fail(h, CHEMSIM_LBM_ERR_INVALID_SLICE_SIZE,
                    "slice has " + std::to_string(n) + " elements, row range is " + std::to_string(h->W) + "x" +
                        std::to_string(row_count));
    return 0;
}

int init_equilibrium_rows(chemsim_lbm *h, int row_begin, int row_count, const void *rho, const void *vx, const void *vy)
{
    const size_t bytes = (size_t)h->W * row_count * h->esize;
    const int rs = ensure_stage(h, 3, bytes);
    if (rs) return rs;
    CUDA_TRY(h, cudaMemcpyAsync(h->stage[0], rho, bytes, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->stage[1], vx, bytes, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->stage[2], vy, bytes, cudaMemcpyHostToDevice, h->stream));
    if (h->stream_mirrored) {           // host rows [b, b+n) are the reflected device rows [H-b-n, H-b), reversed
        for (int i = 0; i < 3; ++i)
            LAUNCH_TRY(h, launch_reverse(h->stage[i], (size_t)h->W * row_count, (int)h->esize, h->stream));
        row_begin = h->H - row_begin - row_count;
    }
    if (h->dtype == CHEMSIM_LBM_F32)
        LAUNCH_TRY(h, launch_init_equilibrium<float>((const float *)h->stage[0], (const float *)h->stage[1],
                                                     (const float *)h->stage[2], (float *)h->buf[h->cur], h->plane,
                                                     h->pitch, h->W, row_begin, row_count, h->k.f, h->stream));
    else
        LAUNCH_TRY(h, launch_init_equilibrium<double>((const double *)h->stage[0], (const double *)h->stage[1],
                                                      (const double *)h->stage[2], (double *)h->buf[h->cur], h->plane,
                                                      h->pitch, h->W, row_begin, row_count, h->k.d, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));   // the host buffers may be pageable / reused by the caller
    h->have_populations = true;
    h->ghosts_valid = false;
    return CHEMSIM_LBM_OK;
}

// Upload mask rows on stream `s`, refresh their segment flags; *d_flag |= any solid.
int upload_geometry_rows(chemsim_lbm *h, int row_begin, int row_count, const uint8_t *solid, cudaStream_t s)
{
    if (h->stream_mirrored) {           // through a dense staging copy, reversed (main stream only)
        const size_t n = (size_t)h->W * row_count;
        const int rs = ensure_stage(h, 1, n);
        if (rs) return rs;
        CUDA_TRY(h, cudaMemcpyAsync(h->stage[0], solid, n, cudaMemcpyHostToDevice, s));
        LAUNCH_TRY(h, launch_reverse(h->stage[0], n, 1, s));
        row_begin = h->H - row_begin - row_count;
        CUDA_TRY(h, cudaMemcpy2DAsync(h->mask + (size_t)row_begin * h->mask_pitch, h->mask_pitch, h->stage[0], h->W, h->W,
                                      row_count, cudaMemcpyDeviceToDevice, s));
    } else
    CUDA_TRY(h, cudaMemcpy2DAsync(h->mask + (size_t)row_begin * h->mask_pitch, h->mask_pitch, solid, h->W, h->W,
                                  row_count, cudaMemcpyHostToDevice, s));
    LAUNCH_TRY(h, launch_mask_flags(h->mask, h->mask_pitch, h->W, row_begin, row_count, h->mask_flags, h->flag_pitch,
                                    h->d_flag, s));
    return 0;
}

int readout(chemsim_lbm *h, int kind, int q, void *dst0, void *dst1, size_t n)
{
    if (!h || !dst0) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    BIND(h);
    const int c = check_n(h, n);
    if (c) return c;
    if (!h->have_populations) return fail(h, CHEMSIM_LBM_ERR_NOT_READY, "populations not set");
    return h->dtype == CHEMSIM_LBM_F32 ? readout_impl<float>(h, kind, q, dst0, dst1)
                                       : readout_impl<double>(h, kind, q, dst0, dst1);
}

int local_mass(chemsim_lbm *h)   // result left in h->d_scalar
{
    if (h->dtype == CHEMSIM_LBM_F32)
        LAUNCH_TRY(h, launch_total_mass<float>((const float *)h->buf[h->cur], h->plane, h->pitch, h->W, h->H,
                                               h->d_partials, h->d_scalar, h->stream));
    else
        LAUNCH_TRY(h, launch_total_mass<double>((const double *)h->buf[h->cur], h->plane, h->pitch, h->W, h->H,
                                                h->d_partials, h->d_scalar, h->stream));
    return 0;
}

int create_impl(int width, int global_height, int dtype, int edge, int device, int rank, int nranks,
                const void *nccl_id, chemsim_lbm_t **out)
{
    if (!out) return fail(nullptr, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    if (width <= 0 || global_height <= 0)
        return fail(nullptr, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "width and height must be positive");
    if (dtype != CHEMSIM_LBM_F32 && dtype != CHEMSIM_LBM_F64)
        return fail(nullptr, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "dtype must be CHEMSIM_LBM_F32 or CHEMSIM_LBM_F64");
    if (edge != CHEMSIM_LBM_EDGE_ZEROFILL && edge != CHEMSIM_LBM_EDGE_PERIODIC)
        return fail(nullptr, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "edge must be ZEROFILL or PERIODIC");
    if (nranks < 1 || rank < 0 || rank >= nranks)
        return fail(nullptr, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "need 0 <= rank < nranks");
    if (nranks > 1 && !nccl_id) return fail(nullptr, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "nccl_id is null");
    if (global_height < nranks)
        return fail(nullptr, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "fewer rows than ranks");

    if (device < 0) {
        const cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess) return fail(nullptr, CHEMSIM_LBM_ERR_CUDA, std::string("cudaGetDevice: ") + cudaGetErrorString(e));
    }
    chemsim_lbm *h = new (std::nothrow) chemsim_lbm;
    if (!h) return fail(nullptr, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "out of host memory");
    h->W = width;
    h->Hglobal = global_height;
    chemsim_lbm_slab_rows(global_height, rank, nranks, &h->row0, &h->H);
    h->dtype = dtype; h->edge = edge; h->device = device; h->rank = rank; h->nranks = nranks;
    h->esize = dtype == CHEMSIM_LBM_F32 ? 4 : 8;
    const int per_line = (int)(128 / h->esize);
    h->pitch = ((width + per_line - 1) / per_line) * per_line;
    h->plane = (size_t)(h->H + 2 * GHOST) * h->pitch;
    h->mask_pitch = ((width + 127) / 128) * 128;
    h->flag_pitch = (((width + MASK_SEGMENT - 1) / MASK_SEGMENT + 2 + 15) / 16) * 16;
    rebuild_scalars(h);

#define CREATE_TRY(expr)                                                                        \
    do {                                                                                        \
        const cudaError_t e_ = (expr);                                                          \
        if (e_ != cudaSuccess) {                                                                \
            const int rc_ = fail(nullptr, CHEMSIM_LBM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
            chemsim_lbm_destroy(h);                                                             \
            return rc_;                                                                         \
        }                                                                                       \
    } while (0)

    CREATE_TRY(cudaSetDevice(device));
    int prio_least = 0, prio_greatest = 0;
    CREATE_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    CREATE_TRY(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_least));
    CREATE_TRY(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, prio_greatest));
    CREATE_TRY(cudaEventCreateWithFlags(&h->ev_face, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&h->ev_interior, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
    const size_t buf_bytes = (size_t)Q * h->plane * h->esize;
    for (int b = 0; b < 2; ++b) {
        CREATE_TRY(cudaMalloc(&h->buf[b], buf_bytes));
        CREATE_TRY(cudaMemsetAsync(h->buf[b], 0, buf_bytes, h->stream));   // ghost rows of a zero-fill edge stay 0
    }
    CREATE_TRY(cudaMalloc((void **)&h->mask_alloc, (size_t)(h->H + 2) * h->mask_pitch));
    CREATE_TRY(cudaMemsetAsync(h->mask_alloc, 0, (size_t)(h->H + 2) * h->mask_pitch, h->stream));
    h->mask = h->mask_alloc + h->mask_pitch;
    CREATE_TRY(cudaMalloc((void **)&h->mask_flags, (size_t)h->H * h->flag_pitch));
    CREATE_TRY(cudaMemsetAsync(h->mask_flags, 0, (size_t)h->H * h->flag_pitch, h->stream));
    CREATE_TRY(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&h->ev_mask, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        CREATE_TRY(cudaEventCreateWithFlags(&h->ev_snap_ready[i], cudaEventDisableTiming));
        CREATE_TRY(cudaEventCreateWithFlags(&h->ev_snap_done[i], cudaEventDisableTiming));
    }
    CREATE_TRY(cudaMalloc((void **)&h->d_partials, sizeof(double) * mass_partials_capacity()));
    CREATE_TRY(cudaMalloc((void **)&h->d_scalar, 8 * sizeof(double)));   // [0..3] reductions, [4..5] barrier scratch
    CREATE_TRY(cudaMalloc((void **)&h->d_flag, sizeof(int)));
    CREATE_TRY(cudaMallocHost((void **)&h->h_scalar, 2 * sizeof(double)));
    CREATE_TRY(cudaMallocHost((void **)&h->h_flag, sizeof(int)));
    CREATE_TRY(cudaStreamSynchronize(h->stream));
#undef CREATE_TRY

    if (nranks > 1) {
        const NcclDyn &n = nccl_dyn();
        if (!n.ok) {
            const int rc = fail(nullptr, CHEMSIM_LBM_ERR_NCCL, "libnccl.so.2 could not be loaded: " + n.error);
            chemsim_lbm_destroy(h);
            return rc;
        }
        ncclUniqueId id;
        static_assert(sizeof(id) == CHEMSIM_LBM_NCCL_ID_BYTES, "ncclUniqueId size");
        std::memcpy(&id, nccl_id, sizeof(id));
        const ncclResult_t r = n.CommInitRank(&h->comm, nranks, id, rank);
        if (r != ncclSuccess) {
            const int rc = fail(nullptr, CHEMSIM_LBM_ERR_NCCL, std::string("ncclCommInitRank: ") + n.GetErrorString(r));
            chemsim_lbm_destroy(h);
            return rc;
        }
    }
    *out = h;
    return CHEMSIM_LBM_OK;
}

}  // namespace

// =============================== C ABI ========================================

extern "C" {

int chemsim_lbm_abi_version(void) { return CHEMSIM_LBM_ABI_VERSION; }

const char *chemsim_lbm_last_error(const chemsim_lbm_t *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int chemsim_lbm_create(int width, int height, int dtype, int edge, int device, chemsim_lbm_t **out)
{
    return create_impl(width, height, dtype, edge, device, 0, 1, nullptr, out);
}

int chemsim_lbm_create_slab(int width, int global_height, int dtype, int edge, int device, int rank, int nranks,
                            const void *nccl_id, chemsim_lbm_t **out)
{
    return create_impl(width, global_height, dtype, edge, device, rank, nranks, nccl_id, out);
}

int chemsim_lbm_nccl_unique_id(void *out_id)
{
    if (!out_id) return fail(nullptr, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "out_id is null");
    const NcclDyn &n = nccl_dyn();
    if (!n.ok) return fail(nullptr, CHEMSIM_LBM_ERR_NCCL, "libnccl.so.2 could not be loaded: " + n.error);
    ncclUniqueId id;
    const ncclResult_t r = n.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, CHEMSIM_LBM_ERR_NCCL, std::string("ncclGetUniqueId: ") + n.GetErrorString(r));
    std::memcpy(out_id, &id, sizeof(id));
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_slab_rows(int global_height, int rank, int nranks, int *row_offset, int *rows)
{
    if (global_height <= 0 || nranks < 1 || rank < 0 || rank >= nranks || global_height < nranks)
        return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    const int r0 = (int)(((long long)global_height * rank) / nranks);
    const int r1 = (int)(((long long)global_height * (rank + 1)) / nranks);
    if (row_offset) *row_offset = r0;
    if (rows) *rows = r1 - r0;
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_halo_plan(int global_height, int rank, int nranks, int edge, chemsim_lbm_halo_msg *out, int *count)
{
    if (!out || !count || nranks < 1 || rank < 0 || rank >= nranks || global_height < nranks)
        return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    *count = build_halo_plan(global_height, rank, nranks, edge, out);
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_enable_p2p_halo(chemsim_lbm_t *h)
{
    if (!h) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    BIND(h);
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->comm_stream));
    return enable_p2p(h);
}

int chemsim_lbm_halo_mode(const chemsim_lbm_t *h, int *mode)
{
    if (!h || !mode) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    *mode = h->halo_mode;
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_destroy(chemsim_lbm_t *h)
{
    if (!h) return CHEMSIM_LBM_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
    if (h->halo_mode == CHEMSIM_LBM_HALO_P2P) {
        // Neighbours may still be storing into my ghost rows: wait (bounded, local polling —
        // no collective, so a dead peer cannot hang the teardown) until both have published
        // the last step, i.e. their last face kernel has finished.
        const bool periodic = h->edge == CHEMSIM_LBM_EDGE_PERIODIC;
        const bool has_up = periodic || h->rank > 0, has_down = periodic || h->rank < h->nranks - 1;
        bool quiesced = false;
        const int max_spins = (int)(h->p2p_timeout_s * 1000.0) + 1;
        for (int spin = 0; spin < max_spins; ++spin) {
            unsigned f[2] = {0, 0};
            if (cudaMemcpy(f, h->p2p_flags, sizeof(f), cudaMemcpyDeviceToHost) != cudaSuccess) break;
            const bool up_done = !has_up || (int)(f[0] - h->step_index) >= 0;
            const bool down_done = !has_down || (int)(f[1] - h->step_index) >= 0;
            if (up_done && down_done) { quiesced = true; break; }
            struct timespec ts = {0, 1000000};
            nanosleep(&ts, nullptr);
        }
        close_p2p(h);
        if (!quiesced) {
            // A neighbour has not published the last step: its face kernel may still store into my
            // ghost rows and flags through its peer mapping.  Leak those allocations rather than
            // hand the neighbour a dangling pointer (device-side use-after-free).
            fprintf(stderr, "chemsim_lbm_destroy: rank %d: a neighbour has not finished step %u; "
                            "population buffers and halo flags are not freed\n", h->rank, h->step_index);
            h->buf[0] = h->buf[1] = nullptr;
            h->p2p_flags = nullptr;
        }
    }
    if (h->p2p_flags) cudaFree(h->p2p_flags);
    if (h->p2p_error_host) cudaFreeHost(h->p2p_error_host);
    if (h->comm) nccl_dyn().CommDestroy(h->comm);
    for (int b = 0; b < 2; ++b) if (h->buf[b]) cudaFree(h->buf[b]);
    if (h->h2d_stream) cudaStreamSynchronize(h->h2d_stream);
    if (h->d2h_stream) cudaStreamSynchronize(h->d2h_stream);
    for (int i = 0; i < 3; ++i) if (h->stage[i]) cudaFree(h->stage[i]);
    for (int i = 0; i < 2; ++i) {
        if (h->snap[i]) cudaFree(h->snap[i]);
        if (h->ev_snap_ready[i]) cudaEventDestroy(h->ev_snap_ready[i]);
        if (h->ev_snap_done[i]) cudaEventDestroy(h->ev_snap_done[i]);
    }
    if (h->ev_main) cudaEventDestroy(h->ev_main);
    if (h->ev_mask) cudaEventDestroy(h->ev_mask);
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    if (h->mask_alloc) cudaFree(h->mask_alloc);
    if (h->mask_flags) cudaFree(h->mask_flags);
    if (h->d_partials) cudaFree(h->d_partials);
    if (h->d_scalar) cudaFree(h->d_scalar);
    if (h->d_flag) cudaFree(h->d_flag);
    if (h->h_scalar) cudaFreeHost(h->h_scalar);
    if (h->h_flag) cudaFreeHost(h->h_flag);
    if (h->ev_face) cudaEventDestroy(h->ev_face);
    if (h->ev_interior) cudaEventDestroy(h->ev_interior);
    if (h->ev_halo) cudaEventDestroy(h->ev_halo);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    delete h;
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_shape(const chemsim_lbm_t *h, int *width, int *local_height, int *global_height, int *row_offset)
{
    if (!h) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    if (width) *width = h->W;
    if (local_height) *local_height = h->H;
    if (global_height) *global_height = h->Hglobal;
    if (row_offset) *row_offset = h->row0;
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_set_discretization(chemsim_lbm_t *h, double delta_x, double delta_t)
{
    if (!h) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    if (!(delta_x > 0.0) || !(delta_t > 0.0)) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "delta_x and delta_t must be positive");
    h->dx = delta_x; h->dt = delta_t;
    rebuild_scalars(h);
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_set_bgk(chemsim_lbm_t *h, double tau)
{
    if (!h) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    if (tau == 0.0 || tau != tau) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "tau must be a non-zero number");
    h->col = CollisionParams();
    h->col.kind = COL_BGK;
    h->col.tau = tau;
    rebuild_scalars(h);
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_set_trt(chemsim_lbm_t *h, double tau_plus, double tau_minus)
{
    if (!h) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    if (tau_plus == 0.0 || tau_minus == 0.0 || tau_plus != tau_plus || tau_minus != tau_minus)
        return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "tau_plus and tau_minus must be non-zero numbers");
    h->col = CollisionParams();
    h->col.kind = COL_TRT;
    h->col.tau_plus = tau_plus;
    h->col.tau_minus = tau_minus;
    rebuild_scalars(h);
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_set_regularized(chemsim_lbm_t *h, double underlying_viscosity)
{
    if (!h) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    h->col = CollisionParams();
    h->col.kind = COL_REGULARIZED;
    h->col.viscosity = underlying_viscosity;
    rebuild_scalars(h);
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_set_kbc(chemsim_lbm_t *h, double ks_viscosity)
{
    if (!h) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    if (ks_viscosity != ks_viscosity) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "viscosity is NaN");
    h->col = CollisionParams();
    h->col.kind = COL_KBC;
    h->col.viscosity = ks_viscosity;
    rebuild_scalars(h);
    return CHEMSIM_LBM_OK;
}

extern "C++" {
template <typename T>
static T shear_viscosity(const chemsim_lbm *h)
{
    const T dx = (T)h->dx, dt = (T)h->dt;
    switch (h->col.kind) {
    case COL_BGK: return (dx * dx / ((T)3.0 * dt * dt)) * ((T)h->col.tau - dt / (T)2.0);   // src/lbm.rs:366-369
    case COL_TRT: {                                                                          // :446-450
        const T cs = dx / (std::sqrt((T)3.0) * dt);
        return cs * cs * ((T)h->col.tau_plus / dt - (T)0.5);
    }
    default: return (T)h->col.viscosity;                                                     // :587-589, :663-665
    }
}
}  // extern "C++"

int chemsim_lbm_kinematic_shear_viscosity(const chemsim_lbm_t *h, double *out)
{
    if (!h || !out) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    if (h->col.kind == COL_NONE) return CHEMSIM_LBM_ERR_NOT_READY;
    *out = h->dtype == CHEMSIM_LBM_F32 ? (double)shear_viscosity<float>(h) : shear_viscosity<double>(h);
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_kinematic_bulk_viscosity(const chemsim_lbm_t *h, double *out)
{
    if (!h || !out) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    double nu = 0.0;
    const int r = chemsim_lbm_kinematic_shear_viscosity(h, &nu);
    if (r) return r;
    if (h->dtype == CHEMSIM_LBM_F32) *out = (double)(2.0f * (float)nu / 3.0f);   // src/lbm.rs:338-340
    else *out = 2.0 * nu / 3.0;
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_init_equilibrium(chemsim_lbm_t *h, const void *rho, const void *vx, const void *vy, size_t n)
{
    if (!h || !rho || !vx || !vy) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    BIND(h);
    const int c = check_n(h, n);
    if (c) return c;
    // bounded staging: at most ~64 MiB per field on the device at a time
    const size_t row_bytes = (size_t)h->W * h->esize;
    int chunk = (int)((64u << 20) / row_bytes);
    if (chunk < 1) chunk = 1;
    for (int y = 0; y < h->H; y += chunk) {
        const int rows = h->H - y < chunk ? h->H - y : chunk;
        const size_t off = (size_t)y * row_bytes;
        const int r = init_equilibrium_rows(h, y, rows, (const char *)rho + off, (const char *)vx + off,
                                            (const char *)vy + off);
        if (r) return r;
    }
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_init_equilibrium_rows(chemsim_lbm_t *h, int row_begin, int row_count, const void *rho, const void *vx,
                                      const void *vy, size_t n)
{
    if (!h || !rho || !vx || !vy) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    BIND(h);
    const int c = check_rows(h, row_begin, row_count, n);
    if (c) return c;
    return init_equilibrium_rows(h, row_begin, row_count, rho, vx, vy);
}

int chemsim_lbm_set_population(chemsim_lbm_t *h, int q, const void *src, size_t n)
{
    if (!h || !src) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    if (q < 0 || q >= Q) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "q must be in 0..9");
    BIND(h);
    const int c = check_n(h, n);
    if (c) return c;
    if (h->stream_mirrored) {
        const size_t bytes = n * h->esize;
        const int rs = ensure_stage(h, 1, bytes);
        if (rs) return rs;
        CUDA_TRY(h, cudaMemcpyAsync(h->stage[0], src, bytes, cudaMemcpyHostToDevice, h->stream));
        LAUNCH_TRY(h, launch_reverse(h->stage[0], n, (int)h->esize, h->stream));
        CUDA_TRY(h, cudaMemcpy2DAsync(row_ptr(h, h->cur, q, 0), (size_t)h->pitch * h->esize, h->stage[0], (size_t)h->W * h->esize,
                                      (size_t)h->W * h->esize, h->H, cudaMemcpyDeviceToDevice, h->stream));
    } else
    CUDA_TRY(h, cudaMemcpy2DAsync(row_ptr(h, h->cur, q, 0), (size_t)h->pitch * h->esize, src, (size_t)h->W * h->esize,
                                  (size_t)h->W * h->esize, h->H, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_populations = true;
    h->ghosts_valid = false;
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_set_geometry(chemsim_lbm_t *h, const uint8_t *solid, size_t n)
{
    if (!h || !solid) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    BIND(h);
    const int c = check_n(h, n);
    if (c) return c;
    CUDA_TRY(h, cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
    const int r = upload_geometry_rows(h, 0, h->H, solid, h->stream);
    if (r) return r;
    CUDA_TRY(h, cudaMemcpyAsync(h->h_flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->has_mask = *h->h_flag ? 1 : 0;   // a solid-free geometry takes the mask-free kernel
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_set_geometry_rows(chemsim_lbm_t *h, int row_begin, int row_count, const uint8_t *solid, size_t n)
{
    if (!h || !solid) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    BIND(h);
    const int c = check_rows(h, row_begin, row_count, n);
    if (c) return c;
    const int r = upload_geometry_rows(h, row_begin, row_count, solid, h->stream);
    if (r) return r;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->has_mask = 1;                    // other rows may hold solids: consult the segment flags
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_set_geometry_async(chemsim_lbm_t *h, const uint8_t *solid, size_t n)
{
    if (!h || !solid) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    BIND(h);
    const int c = check_n(h, n);
    if (c) return c;
    if (h->stream_mirrored) {           // staging + reversal live on the main stream: ordered, not overlapped
        const int r = upload_geometry_rows(h, 0, h->H, solid, h->stream);
        if (r) return r;
        h->has_mask = 1;
        return CHEMSIM_LBM_OK;
    }
    // the copy stream may not overwrite the mask while queued steps still read it
    CUDA_TRY(h, cudaEventRecord(h->ev_main, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->h2d_stream, h->ev_main, 0));
    const int r = upload_geometry_rows(h, 0, h->H, solid, h->h2d_stream);
    if (r) return r;
    CUDA_TRY(h, cudaEventRecord(h->ev_mask, h->h2d_stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_mask, 0));
    h->has_mask = 1;                    // not known on the host: the kernel consults the segment flags
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_fill_geometry(chemsim_lbm_t *h, int value)
{
    if (!h) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    BIND(h);
    const uint8_t v = value ? 1 : 0;
    CUDA_TRY(h, cudaMemset2DAsync(h->mask, h->mask_pitch, v, h->W, h->H, h->stream));
    const int segs = (h->W + MASK_SEGMENT - 1) / MASK_SEGMENT;
    CUDA_TRY(h, cudaMemset2DAsync(h->mask_flags, h->flag_pitch, v, segs, h->H, h->stream));
    h->has_mask = v;                    // all-fluid: the mask-free kernel; all-solid: consult the flags
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_paint_rect(chemsim_lbm_t *h, int x0, int y0, int width, int height, int value)
{
    if (!h) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    if (width < 0 || height < 0) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "negative rectangle size");
    BIND(h);
    if (h->stream_mirrored) { x0 = h->W - x0 - width; y0 = h->Hglobal - y0 - height; }
    // clip to the lattice in x and to this handle's slab in y (y is a GLOBAL row)
    long long xa = x0, xb = (long long)x0 + width, ya = (long long)y0 - h->row0, yb = ya + height;
    if (xa < 0) xa = 0;
    if (xb > h->W) xb = h->W;
    if (ya < 0) ya = 0;
    if (yb > h->H) yb = h->H;
    if (xa >= xb || ya >= yb) return CHEMSIM_LBM_OK;          // nothing of it on this slab
    LAUNCH_TRY(h, launch_paint_rect(h->mask, h->mask_pitch, (int)xa, (int)ya, (int)(xb - xa), (int)(yb - ya),
                                    value ? 1 : 0, h->stream));
    LAUNCH_TRY(h, launch_mask_flags(h->mask, h->mask_pitch, h->W, (int)ya, (int)(yb - ya), h->mask_flags, h->flag_pitch,
                                    h->d_flag, h->stream));
    if (value) h->has_mask = 1;
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_barrier(chemsim_lbm_t *h)
{
    if (!h) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    if (h->nranks == 1) return CHEMSIM_LBM_OK;
    BIND(h);
    NCCL_TRY(h, nccl_dyn().AllReduce(h->d_scalar + 4, h->d_scalar + 5, 1, ncclDouble, ncclSum, h->comm, h->stream));
    h->launches += 1;
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_set_stream_convention(chemsim_lbm_t *h, int mirrored)
{
    if (!h) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    if (mirrored && h->nranks > 1)
        return fail(h, CHEMSIM_LBM_ERR_UNSUPPORTED,
                    "mirrored stream convention on a sharded lattice: create the slabs in reversed rank order instead "
                    "(the reflection maps slab r onto slab nranks-1-r)");
    if (h->have_populations && (mirrored != 0) != (h->stream_mirrored != 0))
        return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "select the stream convention before the first upload");
    h->stream_mirrored = mirrored ? 1 : 0;
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_set_p2p_timeout(chemsim_lbm_t *h, double seconds)
{
    if (!h || !(seconds > 0.0)) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    h->p2p_timeout_s = seconds;
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_step(chemsim_lbm_t *h, int nsteps)
{
    if (!h) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    if (nsteps < 0) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "nsteps must be >= 0");
    if (!h->have_populations) return fail(h, CHEMSIM_LBM_ERR_NOT_READY, "populations not set (init_equilibrium / set_population)");
    if (h->col.kind == COL_NONE) return fail(h, CHEMSIM_LBM_ERR_NOT_READY, "collision operator not set (set_bgk / set_trt / set_regularized / set_kbc)");
    BIND(h);
    P2P_CHECK(h);
    const int r = h->dtype == CHEMSIM_LBM_F32 ? step_impl<float>(h, nsteps) : step_impl<double>(h, nsteps);
    if (r) return r;
    for (int s = 0; s < nsteps; ++s) {   // self.time += delta_t, in Scalar (src/lbm.rs:713)
        h->time_f += (float)h->dt;
        h->time_d += h->dt;
    }
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_synchronize(chemsim_lbm_t *h)
{
    if (!h) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    BIND(h);
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->comm_stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->h2d_stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->d2h_stream));
    P2P_CHECK(h);
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_time(const chemsim_lbm_t *h, double *out)
{
    if (!h || !out) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    *out = h->dtype == CHEMSIM_LBM_F32 ? (double)h->time_f : h->time_d;
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_get_density(chemsim_lbm_t *h, void *dst, size_t n) { return readout(h, READ_DENSITY, 0, dst, nullptr, n); }
int chemsim_lbm_get_async(chemsim_lbm_t *h, int field, int q, void *dst0, void *dst1, size_t n)
{
    if (!h || !dst0) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    if (field < CHEMSIM_LBM_FIELD_DENSITY || field > CHEMSIM_LBM_FIELD_NON_EQUILIBRIUM)
        return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "unknown field");
    const bool two = field == CHEMSIM_LBM_FIELD_VELOCITY || field == CHEMSIM_LBM_FIELD_MOMENTUM_DENSITY;
    if (two != (dst1 != nullptr))
        return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "dst1 is for (and required by) the two-component fields");
    const bool per_q = field >= CHEMSIM_LBM_FIELD_POPULATION;
    if (per_q && (q < 0 || q >= Q)) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "q must be in 0..9");
    BIND(h);
    const int c = check_n(h, n);
    if (c) return c;
    if (!h->have_populations) return fail(h, CHEMSIM_LBM_ERR_NOT_READY, "populations not set");
    P2P_CHECK(h);
    return h->dtype == CHEMSIM_LBM_F32 ? snapshot_async_impl<float>(h, field, q, dst0, dst1)
                                       : snapshot_async_impl<double>(h, field, q, dst0, dst1);
}

int chemsim_lbm_get_density_async(chemsim_lbm_t *h, void *dst, size_t n)
{
    return chemsim_lbm_get_async(h, CHEMSIM_LBM_FIELD_DENSITY, 0, dst, nullptr, n);
}

int chemsim_lbm_get_pressure(chemsim_lbm_t *h, void *dst, size_t n) { return readout(h, READ_PRESSURE, 0, dst, nullptr, n); }
int chemsim_lbm_get_speed(chemsim_lbm_t *h, void *dst, size_t n) { return readout(h, READ_SPEED, 0, dst, nullptr, n); }

int chemsim_lbm_get_velocity(chemsim_lbm_t *h, void *vx, void *vy, size_t n)
{
    if (!vy) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    return readout(h, READ_VELOCITY, 0, vx, vy, n);
}

int chemsim_lbm_get_momentum_density(chemsim_lbm_t *h, void *mx, void *my, size_t n)
{
    if (!my) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    return readout(h, READ_MOMENTUM, 0, mx, my, n);
}

int chemsim_lbm_get_population(chemsim_lbm_t *h, int q, void *dst, size_t n)
{
    if (!h || !dst) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    if (q < 0 || q >= Q) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "q must be in 0..9");
    BIND(h);
    const int c = check_n(h, n);
    if (c) return c;
    if (!h->have_populations) return fail(h, CHEMSIM_LBM_ERR_NOT_READY, "populations not set");
    if (h->stream_mirrored) {
        const size_t bytes = n * h->esize;
        const int rs = ensure_stage(h, 1, bytes);
        if (rs) return rs;
        CUDA_TRY(h, cudaMemcpy2DAsync(h->stage[0], (size_t)h->W * h->esize, row_ptr(h, h->cur, q, 0), (size_t)h->pitch * h->esize,
                                      (size_t)h->W * h->esize, h->H, cudaMemcpyDeviceToDevice, h->stream));
        LAUNCH_TRY(h, launch_reverse(h->stage[0], n, (int)h->esize, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(dst, h->stage[0], bytes, cudaMemcpyDeviceToHost, h->stream));
    } else
    CUDA_TRY(h, cudaMemcpy2DAsync(dst, (size_t)h->W * h->esize, row_ptr(h, h->cur, q, 0), (size_t)h->pitch * h->esize,
                                  (size_t)h->W * h->esize, h->H, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    P2P_CHECK(h);
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_get_equilibrium(chemsim_lbm_t *h, int q, void *dst, size_t n)
{
    if (q < 0 || q >= Q) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "q must be in 0..9");
    return readout(h, READ_EQUILIBRIUM, q, dst, nullptr, n);
}

int chemsim_lbm_get_non_equilibrium(chemsim_lbm_t *h, int q, void *dst, size_t n)
{
    if (q < 0 || q >= Q) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "q must be in 0..9");
    return readout(h, READ_NON_EQUILIBRIUM, q, dst, nullptr, n);
}

int chemsim_lbm_get_geometry(chemsim_lbm_t *h, uint8_t *dst, size_t n)
{
    if (!h || !dst) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    BIND(h);
    const int c = check_n(h, n);
    if (c) return c;
    if (h->stream_mirrored) {
        const int rs = ensure_stage(h, 1, n);
        if (rs) return rs;
        CUDA_TRY(h, cudaMemcpy2DAsync(h->stage[0], h->W, h->mask, h->mask_pitch, h->W, h->H, cudaMemcpyDeviceToDevice, h->stream));
        LAUNCH_TRY(h, launch_reverse(h->stage[0], n, 1, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(dst, h->stage[0], n, cudaMemcpyDeviceToHost, h->stream));
    } else
    CUDA_TRY(h, cudaMemcpy2DAsync(dst, h->W, h->mask, h->mask_pitch, h->W, h->H, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_total_mass(chemsim_lbm_t *h, double *out)
{
    if (!h || !out) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    if (!h->have_populations) return fail(h, CHEMSIM_LBM_ERR_NOT_READY, "populations not set");
    BIND(h);
    const int r = local_mass(h);
    if (r) return r;
    CUDA_TRY(h, cudaMemcpyAsync(h->h_scalar, h->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    P2P_CHECK(h);
    *out = h->h_scalar[0];
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_total_mass_global(chemsim_lbm_t *h, double *out)
{
    if (!h || !out) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    if (h->nranks == 1) return chemsim_lbm_total_mass(h, out);
    if (!h->have_populations) return fail(h, CHEMSIM_LBM_ERR_NOT_READY, "populations not set");
    BIND(h);
    const int r = local_mass(h);
    if (r) return r;
    NCCL_TRY(h, nccl_dyn().AllReduce(h->d_scalar, h->d_scalar + 1, 1, ncclDouble, ncclSum, h->comm, h->stream));
    h->launches += 1;
    CUDA_TRY(h, cudaMemcpyAsync(h->h_scalar, h->d_scalar + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    P2P_CHECK(h);
    *out = h->h_scalar[0];
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_render(chemsim_lbm_t *h, int mode, int overlay_geometry, uint8_t *rgba, size_t n_pixels)
{
    if (!h || !rgba) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    if (mode < RENDER_DENSITY || mode > RENDER_MOMENTUM) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "unknown render mode");
    BIND(h);
    const int c = check_n(h, n_pixels);
    if (c) return c;
    if (!h->have_populations) return fail(h, CHEMSIM_LBM_ERR_NOT_READY, "populations not set");
    const size_t bytes = n_pixels * 4;
    const int rs = ensure_stage(h, 1, bytes);
    if (rs) return rs;
    const uint8_t *mask = overlay_geometry ? h->mask : nullptr;
    double *sums = h->d_scalar, *stats = h->d_scalar + 2;
    // pass 1: sum and sum of squares of the displayed scalar over this slab ...
    if (h->dtype == CHEMSIM_LBM_F32)
        LAUNCH_TRY(h, launch_render_stats<float>((const float *)h->buf[h->cur], h->plane, h->pitch, h->W, h->H, mode,
                                                 h->d_partials, sums, h->stream));
    else
        LAUNCH_TRY(h, launch_render_stats<double>((const double *)h->buf[h->cur], h->plane, h->pitch, h->W, h->H, mode,
                                                  h->d_partials, sums, h->stream));
    // ... over the whole lattice when sharded: mean_all / stdev_all are global (collective call)
    if (h->nranks > 1) {
        NCCL_TRY(h, nccl_dyn().AllReduce(sums, sums, 2, ncclDouble, ncclSum, h->comm, h->stream));
        h->launches += 1;
    }
    LAUNCH_TRY(h, launch_render_stats_finish(sums, (double)h->W * (double)h->Hglobal, stats, h->stream));
    // pass 2: colour mapping of this slab
    if (h->dtype == CHEMSIM_LBM_F32)
        LAUNCH_TRY(h, launch_render_image<float>((const float *)h->buf[h->cur], h->plane, h->pitch, h->W, h->H, mode,
                                                 stats, mask, h->mask_pitch, (uchar4 *)h->stage[0], h->stream));
    else
        LAUNCH_TRY(h, launch_render_image<double>((const double *)h->buf[h->cur], h->plane, h->pitch, h->W, h->H, mode,
                                                  stats, mask, h->mask_pitch, (uchar4 *)h->stage[0], h->stream));
    if (h->stream_mirrored) LAUNCH_TRY(h, launch_reverse(h->stage[0], n_pixels, 4, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(rgba, h->stage[0], bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    P2P_CHECK(h);
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_is_unstable(chemsim_lbm_t *h, int *out)
{
    if (!h || !out) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    if (!h->have_populations) return fail(h, CHEMSIM_LBM_ERR_NOT_READY, "populations not set");
    BIND(h);
    CUDA_TRY(h, cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
    if (h->dtype == CHEMSIM_LBM_F32)
        LAUNCH_TRY(h, launch_is_unstable<float>((const float *)h->buf[h->cur], h->plane, h->pitch, h->W, h->H, h->k.f,
                                                h->d_flag, h->stream));
    else
        LAUNCH_TRY(h, launch_is_unstable<double>((const double *)h->buf[h->cur], h->plane, h->pitch, h->W, h->H, h->k.d,
                                                 h->d_flag, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    P2P_CHECK(h);
    *out = *h->h_flag ? 1 : 0;
    return CHEMSIM_LBM_OK;
}

// ---- checkpoint / restore -------------------------------------------------
// Layout: chemsim_lbm_checkpoint_header, then the nine populations of this handle's slab as dense
// row-major W x H planes in the lattice dtype, then the geometry (W x H bytes).
static size_t checkpoint_size(const chemsim_lbm *h)
{
    const size_t cells = (size_t)h->W * h->H;
    return sizeof(chemsim_lbm_checkpoint_header) + (size_t)Q * cells * h->esize + cells;
}

int chemsim_lbm_checkpoint_bytes(const chemsim_lbm_t *h, size_t *out)
{
    if (!h || !out) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    *out = checkpoint_size(h);
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_checkpoint(chemsim_lbm_t *h, void *dst, size_t bytes)
{
    if (!h || !dst) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    if (bytes < checkpoint_size(h)) return fail(h, CHEMSIM_LBM_ERR_INVALID_SLICE_SIZE, "checkpoint buffer too small");
    if (!h->have_populations) return fail(h, CHEMSIM_LBM_ERR_NOT_READY, "populations not set");
    BIND(h);
    chemsim_lbm_checkpoint_header hd;
    std::memset(&hd, 0, sizeof(hd));
    std::memcpy(hd.magic, CHEMSIM_LBM_CHECKPOINT_MAGIC, 8);
    hd.header_bytes = (uint32_t)sizeof(hd);
    hd.dtype = (uint32_t)h->dtype; hd.width = (uint32_t)h->W; hd.local_height = (uint32_t)h->H;
    hd.global_height = (uint32_t)h->Hglobal; hd.row_offset = (uint32_t)h->row0;
    hd.rank = (uint32_t)h->rank; hd.nranks = (uint32_t)h->nranks; hd.edge = (uint32_t)h->edge;
    hd.collision = (uint32_t)h->col.kind; hd.step_index = h->step_index;
    hd.time_f32 = h->time_f; hd.time_f64 = h->time_d;
    hd.delta_x = h->dx; hd.delta_t = h->dt;
    hd.tau = h->col.tau; hd.tau_plus = h->col.tau_plus; hd.tau_minus = h->col.tau_minus; hd.viscosity = h->col.viscosity;
    char *out = (char *)dst;
    std::memcpy(out, &hd, sizeof(hd));
    out += sizeof(hd);
    const size_t row = (size_t)h->W * h->esize, plane_bytes = row * h->H;
    for (int q = 0; q < Q; ++q)
        CUDA_TRY(h, cudaMemcpy2DAsync(out + (size_t)q * plane_bytes, row, row_ptr(h, h->cur, q, 0), (size_t)h->pitch * h->esize,
                                      row, h->H, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpy2DAsync(out + (size_t)Q * plane_bytes, h->W, h->mask, h->mask_pitch, h->W, h->H,
                                  cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    P2P_CHECK(h);
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_restore(chemsim_lbm_t *h, const void *src, size_t bytes)
{
    if (!h || !src) return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "null argument");
    if (bytes < sizeof(chemsim_lbm_checkpoint_header)) return fail(h, CHEMSIM_LBM_ERR_INVALID_SLICE_SIZE, "checkpoint truncated");
    chemsim_lbm_checkpoint_header hd;
    std::memcpy(&hd, src, sizeof(hd));
    if (std::memcmp(hd.magic, CHEMSIM_LBM_CHECKPOINT_MAGIC, 8) != 0 || hd.header_bytes != sizeof(hd))
        return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "not a chemsim_lbm checkpoint (magic / header size)");
    if ((int)hd.dtype != h->dtype || (int)hd.width != h->W || (int)hd.local_height != h->H ||
        (int)hd.global_height != h->Hglobal || (int)hd.row_offset != h->row0)
        return fail(h, CHEMSIM_LBM_ERR_INVALID_ARGUMENT, "checkpoint is of a different lattice (dtype / shape / slab)");
    if (bytes < checkpoint_size(h)) return fail(h, CHEMSIM_LBM_ERR_INVALID_SLICE_SIZE, "checkpoint truncated");
    BIND(h);
    const char *in = (const char *)src + sizeof(hd);
    const size_t row = (size_t)h->W * h->esize, plane_bytes = row * h->H;
    for (int q = 0; q < Q; ++q)
        CUDA_TRY(h, cudaMemcpy2DAsync(row_ptr(h, h->cur, q, 0), (size_t)h->pitch * h->esize, in + (size_t)q * plane_bytes, row,
                                      row, h->H, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
    const int r = upload_geometry_rows(h, 0, h->H, (const uint8_t *)in + (size_t)Q * plane_bytes, h->stream);
    if (r) return r;
    CUDA_TRY(h, cudaMemcpyAsync(h->h_flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->has_mask = *h->h_flag ? 1 : 0;
    h->have_populations = true;
    h->ghosts_valid = false;            // sharded: the next step exchanges the halo (collective, like any upload)
    h->step_index = hd.step_index;
    h->time_f = hd.time_f32;
    h->time_d = hd.time_f64;
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_cuda_stream(const chemsim_lbm_t *h, void **stream)
{
    if (!h || !stream) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    *stream = (void *)h->stream;
    return CHEMSIM_LBM_OK;
}

int chemsim_lbm_kernel_launches(const chemsim_lbm_t *h, uint64_t *out)
{
    if (!h || !out) return CHEMSIM_LBM_ERR_INVALID_ARGUMENT;
    *out = h->launches;
    return CHEMSIM_LBM_OK;
}

const char *chemsim_lbm_step_kernel_name(const chemsim_lbm_t *h)
{
    if (!h) return "";
    const bool f32 = h->dtype == CHEMSIM_LBM_F32;
    // what a batch of >= 2 steps launches: the two-step kernel where it applies (an odd last step and
    // the lattices it does not support take the single-step kernel)
    const bool two = f32 ? step2_supported(step_args<float>(h, 0, h->H)) : step2_supported(step_args<double>(h, 0, h->H));
    const int TY = f32 ? step2_tile_rows<float>() : step2_tile_rows<double>();
    if (h->nranks == 1) {
        if (two) return f32 ? "step2_kernel<float>" : "step2_kernel<double>";
    } else if (two && deep_halo(h) && h->Hglobal / h->nranks >= 2 * TY) {
        if (h->halo_mode == CHEMSIM_LBM_HALO_P2P) return f32 ? "step2_slab_p2p_kernel<float>" : "step2_slab_p2p_kernel<double>";
        return f32 ? "step2_kernel<float>" : "step2_kernel<double>";
    } else if (h->halo_mode == CHEMSIM_LBM_HALO_P2P) {
        return f32 ? "step_slab_p2p_kernel<float>" : "step_slab_p2p_kernel<double>";
    }
    return f32 ? step_kernel_name<float>(step_args<float>(h, 0, h->H)) : step_kernel_name<double>(step_args<double>(h, 0, h->H));
}

}  // extern "C"
