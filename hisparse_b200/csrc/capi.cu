// C ABI (include/hisparse_b200.h) over the CUDA runtime: the thin shim that replaces the
// reference's OpenCL host calls (sw/host.cpp:263-371; xrt/includes/xcl2).
#include "../../include/hisparse_b200.h"

#include <cuda.h>            // types of the stream memory operations only: the entry points are resolved at run time
#include <cuda_runtime.h>
#include <cuda_profiler_api.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "cpsr_decode.h"
#include "format_handle.h"
#include "gpu_format.h"
#include "spmv_kernels.cuh"
#include "tile_format.h"

namespace {

thread_local std::string g_err;
// Page-locked host buffers handed out by hsb_host_alloc: base -> bytes. Only these are ever used as mapped
// destinations of the drain's posted writes -- memory pinned by somebody else (torch, cudaHostRegister) can be
// unpinned behind our back and takes the copy-engine path.
std::mutex g_host_mu;
struct HostAlloc { size_t bytes; uintptr_t dev; };      // dev: device alias of the base (resolved once, at allocation)
std::map<uintptr_t, HostAlloc> g_host_allocs;

int set_err(int code, const std::string &m) { g_err = m; return code; }

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            char b__[512];                                                                          \
            std::snprintf(b__, sizeof b__, "CUDA error at %s:%d: %s (%s)", __FILE__, __LINE__,      \
                          cudaGetErrorString(e__), #expr);                                          \
            return set_err(HSB_ECUDA, b__);                                                         \
        }                                                                                           \
    } while (0)

// Stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32) let the copy streams and the
// SpMV kernels hand buffers to each other through sequence flags in device memory, so that the compute
// stream carries nothing but kernel launches (an event record or wait between two launches undoes their
// programmatic-dependent-launch overlap: 16 -> 25 us per SpMV on C2). Resolved through the runtime so
// that the library has no link-time dependency on libcuda.
typedef CUresult (*memop32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
memop32_fn g_write32 = nullptr, g_wait32 = nullptr;
bool load_memops() {
    static const bool ok = [] {
        if (std::getenv("HSB_NO_FLAGS")) return false;
        void *w = nullptr, *q = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &w, cudaEnableDefault, &st) != cudaSuccess ||
            st != cudaDriverEntryPointSuccess || !w) { cudaGetLastError(); return false; }
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &q, cudaEnableDefault, &st) != cudaSuccess ||
            st != cudaDriverEntryPointSuccess || !q) { cudaGetLastError(); return false; }
        g_write32 = (memop32_fn)w;
        g_wait32 = (memop32_fn)q;
        return true;
    }();
    return ok;
}
#define MEMOP_TRY(expr)                                                                             \
    do {                                                                                            \
        CUresult r__ = (expr);                                                                      \
        if (r__ != CUDA_SUCCESS) {                                                                  \
            char b__[256];                                                                          \
            std::snprintf(b__, sizeof b__, "stream memory operation failed at %s:%d (CUresult %d)", \
                          __FILE__, __LINE__, (int)r__);                                            \
            return set_err(HSB_ECUDA, b__);                                                         \
        }                                                                                           \
    } while (0)
enum { kFlagXReady = 0, kFlagYFree = 4, kFlagError = 6, kFlagDoneDev = 7, kFlagIterBarrier = 8, kNumFlags = 16, kXBuffers = 4, kAccBuffers = 4 };

struct DeviceMatrix {
    uint32_t *vals = nullptr;
    uint16_t *cols = nullptr;
        uint32_t *slice_rows = nullptr;
    void release() {
        cudaFree(vals); cudaFree(cols); cudaFree(slice_rows);
        vals = nullptr; cols = nullptr; slice_rows = nullptr;
    }
};

}  // namespace

struct hsb_ctx {
    int device = 0, impl = 0, arith = 0;
    hsb::ImplConfig cfg{};
    cudaStream_t stream = nullptr;
    int sm_count = 0, grid = 0;
    // resident matrix
    bool have_matrix = false;
    uint32_t rows = 0, cols = 0, x_words = 0;
    uint64_t nnz = 0;
    uint32_t rows_per_part = 0, n_row_parts = 0, n_col_tiles = 0, tile_cols = 0;
    uint64_t n_slices = 0, n_streams = 0, n_elems = 0, format_bytes = 0;
    std::vector<uint32_t> part_slice_begin;
    hsb::TiledMatrix meta;                // geometry of the resident matrix (tiles, slice list; no payload)
    std::vector<DeviceMatrix> mats;       // [0] + replicas
    size_t sz_vals = 0, sz_cols = 0, sz_rows = 0;
    unsigned next_replica = 0;
    // launch plans: slot 0 = whole matrix, slot 1 + j = row partition j
    uint32_t *d_cta_seg = nullptr;        // [slots][sm_count + 1]
    hsb::Segment *d_segs = nullptr;
    std::vector<uint32_t> plan_grid;      // CTAs used by each of those launches
    std::vector<uint32_t> plan_steps, plan_slices;   // slot 0: steps / slices started per CTA (profiling aid)
    uint32_t smem_bytes = 0;              // dynamic shared memory a launch needs: widest x tile + the zero words (+ combining tables)
    uint32_t comb_offset = 0;             // where the combining tables start in it, or 0: none
    // vectors
    // x is multi-buffered so that the upload of the next vector (copy stream) overlaps the SpMV that still
    // reads the current one; x_words words each (padded to whole tiles, zero filled). Event mode rotates two
    // buffers, the flag pipeline four, so that an upload never has to wait for the launch in flight.
    uint32_t *d_x[kXBuffers] = {nullptr, nullptr, nullptr, nullptr};   // slices of ONE allocation (d_x[0]): one IPC handle
    // multi-GPU iteration over peer memory (hsb_peer_connect): next-x buffers and arrival flags of every rank
    uint32_t *d_peer = nullptr;           // [0..15] arrival flags written by the ranks, [32] completion ticket
    int peer_world = 0, peer_rank = 0;
    uint32_t *peer_x[hsb::kMaxPeers] = {};     // rank g's d_x[0] (this process's own pointer for g == rank)
    uint32_t *peer_flags[hsb::kMaxPeers] = {}; // rank g's d_peer
    bool peer_ipc[hsb::kMaxPeers] = {};        // opened through CUDA IPC (another process): closed on disconnect
    size_t x_stride = 0;                  // words between consecutive x buffers
    uint32_t peer_seq = 0;
    int x_latest = 0;                     // buffer the next launch reads
    int x_next_buf = -1;                  // buffer hsb_axpb_to_vector wrote and hsb_vector_commit will make current
    bool x_next_from_peers = false;       // ... filled by all ranks (hsb_axpb_to_peers): the next SpMV polls their arrival flags
    bool x_dirty = false;                 // uploaded since the last launch: the launch must wait for the copy
    bool x_after_grid = false;            // the current x was written by the axpb kernel just in front on the stream: the next launch waits for that grid
    // y is double buffered too: the launch that follows a deferred download drains into d_y[y_cur], the
    // copy engine reads that buffer out, and later launches drain into the other one
    uint32_t *d_y[2] = {nullptr, nullptr}; // rows words each
    int y_cur = 0;                        // THE result vector: every drain writes d_y[y_cur]
    bool y_busy[2] = {false, false};      // an asynchronous download of d_y[b] may still be running
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaStream_t s_h2d_more[2] = {nullptr, nullptr};   // HSB_UPLOAD_STREAMS=3|4: further copy streams in the rotation (A/B aid)
    int n_upload_streams = 2;
    cudaStream_t s_h2d_b = nullptr;       // flag pipeline: uploads alternate between two streams, so that one copy runs while the other
                                          // stream is still busy with its flag write (a stream memory operation takes ~5 us there)
    cudaEvent_t ev_xready = nullptr, ev_xfree[2] = {nullptr, nullptr}, ev_yready = nullptr, ev_ydone = nullptr;
    // flag pipeline (see load_memops): sequence flags in device memory instead of events
    bool flags_mode = false;
    uint32_t *d_flags = nullptr;          // kNumFlags words
    // Launch numbers are 64-bit on the host and travel to the device as their low 32 bits: kernels and stream memory
    // operations compare flags cyclically ((int32_t)(flag - value) >= 0), which is exact while fewer than 2^31 launches
    // are in flight, and the host rebuilds the 64-bit number of the last finished launch from the 32-bit word
    // (done_launch). A context that has issued more than 2^32 launches (C1: five hours) keeps working.
    uint64_t launch_seq = 0;              // SpMV launches so far; launch n publishes n - 1 in done_seq when it starts
    uint64_t publish_sure = 0;            // highest sequence number that WILL appear in done_seq without further host action
    uint64_t d2h_wait_seq = 0;            // highest sequence number the download stream has been told to wait for
    uint64_t seq_base = 0;                // launch number this context started from (0; HSB_DEBUG_SEQ_BASE: the wrap-around tests)
    bool xwait_once = true;               // stop polling the x flag once a launch that polled it has completed
    bool host_drain = true;               // deferred downloads into page-locked memory are written by the drain itself
    uint64_t x_reader_seq[kXBuffers] = {0, 0, 0, 0};   // last launch that reads d_x[b] (0: none)
    // done_seq lives in mapped page-locked host memory: kernels / stream memory operations write it through
    // d_done, the host polls h_done directly (upload throttling) instead of queueing a stream wait
    volatile uint32_t *h_done = nullptr;
    uint32_t *d_done = nullptr;
    uint32_t x_seq = 0;                   // uploads so far == value written to x_ready[b] when the copy has landed
    int x_wait_buf = -1;                  // buffer whose x_ready flag the next launches wait for (-1: none)
    uint64_t x_wait_launch = 0;           // first launch that polled the current flag (0: none yet)
    uint32_t x_wait_val = 0;
    uint32_t dl_seq = 0, y_dl_seq[2] = {0, 0};   // downloads so far; value y_free[b] reaches when d_y[b] has been read out
    // deferred download; dev != null: `host` is page-locked and mapped, the next launch drains straight into it
    struct { void *host = nullptr; uint32_t *dev = nullptr; unsigned n = 0; bool active = false; } pending_dl;
    // gather of y across row-block shards (hsb_gather_export / hsb_gather_connect)
    uint32_t *d_gather_y = nullptr;       // this rank's gathered result (total rows of all shards), when it is a target
    uint32_t *d_gather_flags = nullptr;   // [0..15] arrival flags written by the ranks, [32] this rank's completion ticket
    hsb::GatherTargets *d_gather = nullptr;   // device table the drains read, or null: no gather
    uint32_t gather_total_rows = 0, gather_seq = 0;
    int gather_world = 0, gather_rank = 0;
    bool gather_is_target = false;
    void *gather_opened[2 * hsb::kMaxPeers] = {};   // IPC mappings to close
    int gather_n_opened = 0;
    bool acquire = true;                  // flag waits of the kernels end in an acquire fence (hsb_set_option "acquire")
    // "x has landed" flag written by a 4-byte copy from a ring of page-locked sequence words instead of a stream
    // memory operation (hsb_set_option "xflag_copy"): a pure copy-engine operation behind the vector's copy
    int xflag_copy = 1;
    bool iterate_persistent = true;       // hsb_iterate as one cooperative launch (hsb_set_option "iterate_persistent", HSB_ITERATE_PERSISTENT)
    uint32_t *h_seq_ring = nullptr;       // 256 page-locked words; slot (seq & 255) holds seq while its copy may be in flight
    // rows + 1 accumulators (uint64 fixed / fp32 float) x 4 in rotation: launch n adds into buffer n % 4 and,
    // at its very end, drains buffer (n - 1) % 4 (its predecessor's sums) into y and re-zeroes it. Four, because
    // consecutive launches overlap freely (see the kernel): buffer n % 4 is reused by launch n + 4, which
    // checks that launch n + 1 -- the one that re-zeroed it -- is complete.
    void *d_acc[kAccBuffers] = {nullptr, nullptr, nullptr, nullptr};
    int acc_cur = 0;
    int acc_bufs = kAccBuffers;           // buffers in rotation: 4, or 2 with a wait at launch start (accumulators >> L2)
    bool drain_pending = false;           // d_acc[acc_cur ^ 1] holds sums that are not in y yet
    uint32_t drain_begin = 0, drain_end = 0;
    unsigned long long *d_trace = nullptr; // optional per-warp clock stamps of the last launch
    unsigned long long *d_timeline = nullptr;   // optional [256][8] globaltimer stamps of the last 256 launches
    uint64_t launches = 0;
    double preprocess_s = 0;
};

namespace {

void close_peers(hsb_ctx *c) {
    for (int g = 0; g < hsb::kMaxPeers; g++) {
        if (c->peer_ipc[g]) {
            if (c->peer_x[g]) cudaIpcCloseMemHandle(c->peer_x[g]);
            if (c->peer_flags[g]) cudaIpcCloseMemHandle(c->peer_flags[g]);
        }
        c->peer_x[g] = nullptr; c->peer_flags[g] = nullptr; c->peer_ipc[g] = false;
    }
    c->peer_world = 0;
}

void close_gather(hsb_ctx *c) {
    for (int i = 0; i < c->gather_n_opened; i++) cudaIpcCloseMemHandle(c->gather_opened[i]);
    c->gather_n_opened = 0;
    cudaFree(c->d_gather); cudaFree(c->d_gather_y); cudaFree(c->d_gather_flags);
    c->d_gather = nullptr; c->d_gather_y = nullptr; c->d_gather_flags = nullptr;
    c->gather_world = 0; c->gather_seq = 0; c->gather_total_rows = 0; c->gather_is_target = false;
}

void free_matrix(hsb_ctx *c) {
    for (auto &m : c->mats) m.release();
    c->mats.clear();
    close_peers(c);
    close_gather(c);
    cudaFree(c->d_x[0]); cudaFree(c->d_peer); c->d_peer = nullptr;
    for (int b = 0; b < kXBuffers; b++) { c->d_x[b] = nullptr; c->x_reader_seq[b] = 0; }
    for (int b = 0; b < kAccBuffers; b++) { cudaFree(c->d_acc[b]); c->d_acc[b] = nullptr; }
    cudaFree(c->d_y[0]); cudaFree(c->d_y[1]); cudaFree(c->d_cta_seg); cudaFree(c->d_segs);
    c->d_y[0] = c->d_y[1] = nullptr; c->x_latest = 0; c->x_dirty = false; c->d_cta_seg = nullptr; c->d_segs = nullptr;
    c->y_cur = 0; c->y_busy[0] = c->y_busy[1] = false; c->pending_dl.active = false;
    c->x_wait_buf = -1; c->x_next_buf = -1;
    c->drain_pending = false; c->acc_cur = 0;
    c->have_matrix = false;
}

int install_matrix(hsb_ctx *c, const hsb::TiledMatrix &M, const DeviceMatrix &d, size_t n_elems, size_t n_slices);
int quiesce(hsb_ctx *c);

int upload_tiled(hsb_ctx *c, const hsb::TiledMatrix &M) {
    CUDA_TRY(cudaSetDevice(c->device));
    { int rc = quiesce(c); if (rc) return rc; }
    free_matrix(c);
    DeviceMatrix d;
    CUDA_TRY(cudaMalloc(&d.vals, M.vals.size() * 4 + 16));
    CUDA_TRY(cudaMalloc(&d.cols, M.cols16.size() * 2 + 16));
    CUDA_TRY(cudaMalloc(&d.slice_rows, M.slice_rows.size() * 4 + 16));
    CUDA_TRY(cudaMemcpyAsync(d.vals, M.vals.data(), M.vals.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d.cols, M.cols16.data(), M.cols16.size() * 2, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d.slice_rows, M.slice_rows.data(), M.slice_rows.size() * 4, cudaMemcpyHostToDevice, c->stream));
    return install_matrix(c, M, d, M.n_elems(), M.n_slices());
}

// take ownership of a device-resident tile-stream matrix; M supplies the geometry for the planner
int install_matrix(hsb_ctx *c, const hsb::TiledMatrix &M, const DeviceMatrix &d, size_t n_elems, size_t n_slices) {
    c->rows = M.rows; c->cols = M.cols; c->nnz = M.nnz;
    c->rows_per_part = M.rows_per_part; c->n_row_parts = M.n_row_parts;
    c->n_col_tiles = M.n_col_tiles; c->tile_cols = M.tile_cols;
    c->n_slices = n_slices; c->n_streams = M.n_streams; c->n_elems = n_elems;
    c->sz_vals = n_elems * 4; c->sz_cols = n_elems * 2;
    c->sz_rows = M.narrow ? 0 : n_slices * hsb::kLanes * 4;      // narrow layout: the row ids are in the stream
    c->format_bytes = c->sz_vals + c->sz_cols + c->sz_rows;
    c->part_slice_begin = M.part_slice_begin;
    c->meta = hsb::TiledMatrix();
    c->meta.rows = M.rows; c->meta.cols = M.cols; c->meta.nnz = M.nnz; c->meta.rows_per_part = M.rows_per_part;
    c->meta.n_row_parts = M.n_row_parts; c->meta.n_col_tiles = M.n_col_tiles; c->meta.tile_cols = M.tile_cols;
    c->meta.n_streams = M.n_streams; c->meta.slices = M.slices; c->meta.tiles = M.tiles; c->meta.narrow = M.narrow;
    c->meta.part_slice_begin = M.part_slice_begin;
    c->mats.push_back(d);
    // cost-balanced work plans for the whole-matrix launch and for each row partition
    const uint32_t G = (uint32_t)c->sm_count, T = M.n_col_tiles;
    std::vector<uint32_t> cta_seg_all;
    std::vector<hsb::Segment> segs;
    c->plan_grid.assign(1 + M.n_row_parts, 1);
    auto plan = [&](size_t slot, uint32_t tile0, uint32_t tile1) {
        uint64_t steps = 0;
        for (uint32_t t = tile0; t < tile1; t++) steps += M.tiles[t].slice_end - M.tiles[t].slice_begin;
        // at least kMinSteps steps per CTA (one per warp) when the launch is that small: a tiny launch on a quarter of
        // the SMs leaves room for its successor to start at once instead of waiting for an SM (HSB_MIN_STEPS_PER_CTA)
        static const uint64_t min_steps = [] { const char *e = std::getenv("HSB_MIN_STEPS_PER_CTA"); return (uint64_t)std::max(1, e ? std::atoi(e) : 1); }();
        uint64_t unit_steps = 0;
        for (uint32_t t = tile0; t < tile1; t++)
            if (M.tiles[t].slice_end > M.tiles[t].slice_begin) {
                const hsb::SliceDesc &last = M.slices[M.tiles[t].slice_end - 1];
                unit_steps += last.off + (last.tile_steps & 0xFFu) - M.tiles[t].step_begin;
            }
        uint32_t g = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(G, std::min<uint64_t>(steps, (unit_steps + min_steps - 1) / min_steps)));
        c->plan_grid[slot] = g;
        std::vector<uint32_t> cs;
        hsb::plan_launch(M, tile0, tile1, g, &cs, &segs);
        cs.resize(G + 1, cs.back());
        cta_seg_all.insert(cta_seg_all.end(), cs.begin(), cs.end());
    };
    plan(0, 0, M.n_row_parts * T);
    c->plan_steps.assign(G, 0); c->plan_slices.assign(G, 0);
    for (uint32_t b = 0; b < G; b++)
        for (uint32_t g = cta_seg_all[b]; g < cta_seg_all[b + 1]; g++) {
            c->plan_steps[b] += segs[g].t_hi - segs[g].t_lo;
            const hsb::TileDesc &td = M.tiles[segs[g].tile];
            for (uint32_t s = td.slice_begin; s < td.slice_end; s++) {
                uint32_t o = M.slices[s].off - td.step_begin;
                if (o >= segs[g].t_lo && o < segs[g].t_hi) c->plan_slices[b]++;
            }
        }
    for (uint32_t j = 0; j < M.n_row_parts; j++) plan(1 + j, j * T, (j + 1) * T);
    CUDA_TRY(cudaMalloc(&c->d_cta_seg, cta_seg_all.size() * 4));
    CUDA_TRY(cudaMemcpyAsync(c->d_cta_seg, cta_seg_all.data(), cta_seg_all.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMalloc(&c->d_segs, (segs.size() + 1) * sizeof(hsb::Segment)));
    CUDA_TRY(cudaMemcpyAsync(c->d_segs, segs.data(), segs.size() * sizeof(hsb::Segment), cudaMemcpyHostToDevice, c->stream));
    // x is padded to whole tiles so that every bulk copy of a tile stays inside the buffer
    c->x_words = M.n_col_tiles * M.tile_cols;
    c->x_stride = ((size_t)c->x_words + 8 + 31) & ~(size_t)31;
    CUDA_TRY(cudaMalloc(&c->d_x[0], c->x_stride * kXBuffers * 4));
    CUDA_TRY(cudaMemsetAsync(c->d_x[0], 0, c->x_stride * kXBuffers * 4, c->stream));
    for (int b = 0; b < kXBuffers; b++) {
        c->d_x[b] = c->d_x[0] + b * c->x_stride;
        if (b < 2) CUDA_TRY(cudaEventRecord(c->ev_xfree[b], c->stream));
    }
    CUDA_TRY(cudaMalloc(&c->d_peer, 64 * 4));
    CUDA_TRY(cudaMemsetAsync(c->d_peer, 0, 64 * 4, c->stream));
    for (int b = 0; b < 2; b++) {
        CUDA_TRY(cudaMalloc(&c->d_y[b], (size_t)std::max(c->rows, 1u) * 4));
        CUDA_TRY(cudaMemsetAsync(c->d_y[b], 0, (size_t)std::max(c->rows, 1u) * 4, c->stream));
    }
    const size_t esz = c->arith == hsb::kArithFixed ? 8 : 4;
    // Four buffers let consecutive launches overlap without any wait, but every launch then updates rows in a
    // buffer that was last touched three launches ago. When the buffers together no longer fit in L2 (C5
    // shards: 12.5 M rows) the row updates would all miss; such launches are long anyway, so they rotate two
    // buffers and wait for their predecessor before the first row update instead.
    c->acc_bufs = ((size_t)c->rows + 1) * esz * kAccBuffers > (size_t)96 << 20 ? 2 : kAccBuffers;
    if (const char *e = std::getenv("HSB_ACC_BUFFERS")) { const int v = std::atoi(e); c->acc_bufs = v == 1 ? 1 : v == 2 ? 2 : kAccBuffers; }
    for (int b = 0; b < c->acc_bufs; b++) {
        CUDA_TRY(cudaMalloc(&c->d_acc[b], ((size_t)c->rows + 1) * esz));
        CUDA_TRY(cudaMemsetAsync(c->d_acc[b], 0, ((size_t)c->rows + 1) * esz, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->grid = (int)c->plan_grid[0];
    uint32_t widest = 8;
    for (const auto &td : M.tiles) widest = std::max(widest, td.col_count);
    c->smem_bytes = widest * 4u + hsb::kXTileOffset;
    // the two combining tables of the wide kernel (Segment::comb_n) go behind the x tile when the CTA has room
    c->comb_offset = 0;
    const uint32_t comb_bytes = 2u * hsb::kCombineSlots * hsb::kLanes * (uint32_t)esz, at = (c->smem_bytes + 15u) & ~15u;
    if (!M.narrow && at + comb_bytes <= hsb::kSmemBytes) { c->comb_offset = at; c->smem_bytes = at + comb_bytes; }
    c->next_replica = 0;
    c->have_matrix = true;
    return HSB_OK;
}

int finish(hsb_ctx *c);

// NUMA node the GPU hangs off (sysfs, through its PCI bus id), or -1 when the platform does not say
int device_numa_node(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char *q = bus; *q; q++) if (*q >= 'A' && *q <= 'Z') *q = (char)(*q - 'A' + 'a');
    char path[128];
    std::snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = std::fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
    return node;
}

// Page-locked buffers are faulted in by the allocation itself, on the node the calling thread's memory policy
// prefers: prefer the node of the thread's CURRENT device while allocating (set_mempolicy through the raw
// system call: no libnuma in this image), so that every x upload and y download of a rank crosses only its
// own root complex.
struct PreferDeviceNode {
    bool active = false;
    PreferDeviceNode() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
        const int node = device_numa_node(dev);
        if (node < 0 || node >= 1024) return;
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        active = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, 1025ul) == 0;
    }
    ~PreferDeviceNode() { if (active) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul); }
};

// device alias of a page-locked, mapped host buffer from hsb_host_alloc (n_words must fit inside it), or null
uint32_t *mapped_alias(hsb_ctx *c, const void *host, size_t n_words) {
    (void)c;
    const uintptr_t h = (uintptr_t)host;
    std::lock_guard<std::mutex> lk(g_host_mu);
    auto it = g_host_allocs.upper_bound(h);
    if (it == g_host_allocs.begin()) return nullptr;
    --it;
    if (!it->second.dev || h + n_words * 4 > it->first + it->second.bytes) return nullptr;
    return (uint32_t *)(it->second.dev + (h - it->first));
}

// a kernel gave up waiting for a flag: report it once (the launch skipped its row updates, y is incomplete). Callers
// have synchronised the streams that could still raise it. With the flag pipeline the word is in mapped host memory
// (h_done[8]): a host load, not the 10 us of a blocking 4-byte copy on every download.
int check_error_flag(hsb_ctx *c) {
    uint32_t err = 0;
    if (c->h_done) {
        err = c->h_done[8];
        if (err) c->h_done[8] = 0;
    } else {
        CUDA_TRY(cudaMemcpy(&err, c->d_flags + kFlagError, 4, cudaMemcpyDeviceToHost));
        if (err) CUDA_TRY(cudaMemset(c->d_flags + kFlagError, 0, 4));
    }
    if (err)
        return set_err(HSB_ECUDA, "a kernel gave up waiting for a flag (vector upload, result download, peer slice or "
                                  "accumulator reuse) and skipped its row updates: the result is incomplete");
    return HSB_OK;
}

// A launch's completion is normally announced by its successor (after the griddepcontrol.wait at the end of
// its CTA 0). When there may be no successor (hsb_sync, a second upload in a row, ...) the compute stream
// announces everything launched so far itself, with a stream memory operation behind the last kernel.
// 64-bit number of the last launch known to be complete, from the 32-bit word the device publishes: it never
// exceeds launch_seq and trails it by less than 2^32.
inline uint64_t done_launch(const hsb_ctx *c) {
    return c->launch_seq - (uint32_t)((uint32_t)c->launch_seq - *c->h_done);
}

int publish_done(hsb_ctx *c) {
    if (c->publish_sure == c->launch_seq) return HSB_OK;
    MEMOP_TRY(g_write32((CUstream)c->stream, (CUdeviceptr)c->d_done, (uint32_t)c->launch_seq, 0));
    c->publish_sure = c->launch_seq;
    return HSB_OK;
}

// start the device -> host copy of the result vector; y is final on the compute stream in stream order
int issue_copy_after_main(hsb_ctx *c, void *host, unsigned n) {
    const int b = c->y_cur;
    CUDA_TRY(cudaEventRecord(c->ev_yready, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->s_d2h, c->ev_yready, 0));
    CUDA_TRY(cudaMemcpyAsync(host, c->d_y[b], (size_t)n * 4, cudaMemcpyDeviceToHost, c->s_d2h));
    if (c->flags_mode) {
        c->y_dl_seq[b] = ++c->dl_seq;
        MEMOP_TRY(g_write32((CUstream)c->s_d2h, (CUdeviceptr)(c->d_flags + kFlagYFree + b), c->y_dl_seq[b], 0));
    } else {
        CUDA_TRY(cudaEventRecord(c->ev_ydone, c->s_d2h));
    }
    c->y_busy[b] = true;
    return HSB_OK;
}

// one launch = one SpMV over the slices of `slot` (0: whole matrix, 1 + j: row partition j), rows [rb, re)
int run_slot(hsb_ctx *c, size_t slot, uint32_t rb, uint32_t re, cudaEvent_t k0, cudaEvent_t k1) {
    // a deferred download rides on a whole-matrix launch only (its successor's drain then rewrites every
    // row of the other y buffer); anything else resolves it the immediate way first
    if (c->pending_dl.active && slot != 0) { int rc = finish(c); if (rc) return rc; }
    // one accumulator buffer (A/B aid, HSB_ACC_BUFFERS=1): the previous sums are drained by the drain kernel before
    // this launch may add to the same buffer
    if (c->acc_bufs == 1 && c->drain_pending) { int rc = finish(c); if (rc) return rc; }
    const DeviceMatrix &m = c->mats[c->next_replica % c->mats.size()];
    c->next_replica++;
    const uint32_t G = (uint32_t)c->sm_count;
    const int grid = (int)c->plan_grid[slot];
    hsb::SpmvParams p;
    std::memset(&p, 0, sizeof p);
    p.vals = m.vals; p.cols = m.cols; p.slice_rows = m.slice_rows;
    p.cta_seg = c->d_cta_seg + slot * (size_t)(G + 1);
    p.segs = c->d_segs;
    const int yb = c->y_cur;
    if (c->flags_mode) {
        // nothing but the launch goes on the compute stream: the kernel itself waits for its x (a flag the
        // upload stream writes after the copy) and, before its prologue drain overwrites y, for the
        // download that still reads it
        // every launch that may start before the upload has landed polls its flag; once a launch that
        // polled is known to be complete (the host-visible done_seq) the vector is there for good
        if (c->x_wait_buf >= 0 && c->xwait_once && c->x_wait_launch && done_launch(c) >= c->x_wait_launch)
            c->x_wait_buf = -1;
        if (c->x_wait_buf >= 0) {
            p.wait_x_flag = c->d_flags + kFlagXReady + c->x_wait_buf; p.wait_x_val = c->x_wait_val; p.wait_x_count = 1;
            if (c->x_wait_buf >= kXBuffers) { p.wait_x_flag = c->d_peer; p.wait_x_count = (uint32_t)c->peer_world; }   // slices from all ranks
            if (!c->x_wait_launch) c->x_wait_launch = c->launch_seq + 1;
        }
        if (c->drain_pending && c->y_busy[yb]) {
            p.wait_y_flag = c->d_flags + kFlagYFree + yb; p.wait_y_val = c->y_dl_seq[yb];
            c->y_busy[yb] = false;                         // later writers are ordered behind this launch
        }
        p.done_seq = c->d_done;
        c->x_reader_seq[c->x_latest] = c->launch_seq + 1;
        c->publish_sure = std::max(c->publish_sure, c->launch_seq);
    } else {
        if (c->x_dirty) {                                   // the vector this launch reads is still being uploaded
            CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_xready, 0));
            c->x_dirty = false;
        }
        if (c->drain_pending && c->y_busy[yb]) {            // the prologue drain writes y: wait for its last reader
            CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_ydone, 0));
            c->y_busy[yb] = false;
        }
    }
    p.seq = (uint32_t)++c->launch_seq;
    p.done_dev = c->d_flags + kFlagDoneDev;
    p.error_flag = c->d_done ? c->d_done + 8 : c->d_flags + kFlagError;
    if (c->acc_bufs <= 2) p.sync_start = 1;
    else if (c->launch_seq - c->seq_base > 3) { p.guard_flag = p.done_dev; p.guard_val = p.seq - 3u; }
    p.x = c->d_x[c->x_latest]; p.y = c->d_y[yb];
    if (c->pending_dl.active && c->pending_dl.dev && c->drain_pending) {
        // page-locked destination: the prologue drain writes the result words to the host buffer as well
        // (posted PCIe writes that trickle out under the SpMV); no copy engine, no second y buffer
        p.y_host = c->pending_dl.dev; p.y_host_rows = c->pending_dl.n;
    }
    p.acc = c->d_acc[c->acc_cur];
    p.drain_acc = c->drain_pending ? c->d_acc[(c->acc_cur + c->acc_bufs - 1) % c->acc_bufs] : nullptr;
    p.acquire = c->acquire ? 1u : 0u;
    p.narrow = c->meta.narrow ? 1u : 0u;
    p.comb_offset = c->comb_offset;
    p.x_after_grid = c->x_after_grid ? 1u : 0u;
    c->x_after_grid = false;                               // (later launches are ordered behind this one)
    if (c->d_gather && c->drain_pending) { p.gather = c->d_gather; p.gather_seq = ++c->gather_seq; }
    p.drain_begin = c->drain_begin; p.drain_end = c->drain_end;
    p.trash_row = c->rows;
    p.trace = c->d_trace;
    if (c->d_timeline) {
        p.timeline = c->d_timeline;
    }
    if (k0) CUDA_TRY(cudaEventRecord(k0, c->stream));
    CUDA_TRY(hsb::launch_spmv(c->arith, p, grid, c->smem_bytes, c->stream));
    c->launches++;
    if (k1) CUDA_TRY(cudaEventRecord(k1, c->stream));
    if (c->pending_dl.active && !c->pending_dl.dev) {
        // the launch just issued drains the requested result into d_y[yb]; once it has completed (its
        // sequence number appears in done_seq) the download stream copies it out, and from now on the
        // drains go to the other buffer
        MEMOP_TRY(g_wait32((CUstream)c->s_d2h, (CUdeviceptr)c->d_done, p.seq, CU_STREAM_WAIT_VALUE_GEQ));
        c->d2h_wait_seq = c->launch_seq;
        CUDA_TRY(cudaMemcpyAsync(c->pending_dl.host, c->d_y[yb], (size_t)c->pending_dl.n * 4, cudaMemcpyDeviceToHost, c->s_d2h));
        c->y_dl_seq[yb] = ++c->dl_seq;
        MEMOP_TRY(g_write32((CUstream)c->s_d2h, (CUdeviceptr)(c->d_flags + kFlagYFree + yb), c->y_dl_seq[yb], 0));
        c->y_busy[yb] = true;
        c->y_cur ^= 1;
    }
    c->pending_dl.active = false;                           // (a mapped host buffer was handed to the launch itself)
    c->drain_pending = true;
    c->drain_begin = rb; c->drain_end = re;
    c->acc_cur = (c->acc_cur + 1) % c->acc_bufs;
    return HSB_OK;
}

// make y final: drain what the last launch accumulated (stream-ordered, no host sync), and start a
// deferred download if one is waiting
int finish(hsb_ctx *c) {
    if (c->drain_pending) {
        const int yb = c->y_cur;
        if (c->y_busy[yb]) {
            if (c->flags_mode)
                MEMOP_TRY(g_wait32((CUstream)c->stream, (CUdeviceptr)(c->d_flags + kFlagYFree + yb), c->y_dl_seq[yb],
                                   CU_STREAM_WAIT_VALUE_GEQ));   // (off the fast path)
            else
                CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_ydone, 0));
            c->y_busy[yb] = false;
        }
        const uint32_t gseq = c->d_gather ? ++c->gather_seq : 0;
        // a download that was waiting for this drain and goes to mapped page-locked memory rides on it: the drain
        // kernel writes the result words to the host buffer as well (posted PCIe writes), no copy engine, no event hop
        uint32_t *y_host = nullptr;
        uint32_t y_host_rows = 0;
        if (c->pending_dl.active && c->pending_dl.dev && c->drain_begin == 0 && c->drain_end == c->rows) {
            y_host = c->pending_dl.dev; y_host_rows = c->pending_dl.n;
            c->pending_dl.active = false;
        }
        CUDA_TRY(hsb::launch_drain(c->arith, c->d_acc[(c->acc_cur + c->acc_bufs - 1) % c->acc_bufs], c->d_y[yb], c->drain_begin, c->drain_end, c->rows,
                                   c->d_gather, gseq, y_host, y_host_rows, c->stream));
        c->launches++;
        c->drain_pending = false;
    }
    // the download stream may be parked on the last launch's number, which only a successor would announce
    if (c->flags_mode && c->d2h_wait_seq > c->publish_sure) { int rc = publish_done(c); if (rc) return rc; }
    if (c->pending_dl.active) {
        c->pending_dl.active = false;
        return issue_copy_after_main(c, c->pending_dl.host, c->pending_dl.n);
    }
    return HSB_OK;
}

// every stream idle, nothing deferred: the state in which buffers may be freed or replaced
int quiesce(hsb_ctx *c) {
    if (c->have_matrix) { int rc = finish(c); if (rc) return rc; }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->s_h2d));
    CUDA_TRY(cudaStreamSynchronize(c->s_h2d_b));
    for (int i = 0; i < 2; i++) CUDA_TRY(cudaStreamSynchronize(c->s_h2d_more[i]));
    CUDA_TRY(cudaStreamSynchronize(c->s_d2h));
    c->y_busy[0] = c->y_busy[1] = false;
    return check_error_flag(c);
}

int run_all(hsb_ctx *c, cudaEvent_t k0, cudaEvent_t k1) { return run_slot(c, 0, 0, c->rows, k0, k1); }

}  // namespace

extern "C" {

const char *hsb_version(void) { return "hisparse_b200 0.1 (sm_100a)"; }
const char *hsb_last_error(void) { return g_err.c_str(); }

int hsb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int hsb_get_config(int impl, hsb_config *out) {
    hsb::ImplConfig cfg;
    if (!out || !hsb::impl_config(impl, &cfg)) return set_err(HSB_EINVAL, "unknown impl");
    out->pack_size = HSB_PACK_SIZE; out->num_hbm_channels = HSB_NUM_HBM_CHANNELS;
    out->interleave_factor = cfg.interleave; out->logical_ob_size = cfg.ob_size; out->logical_vb_size = cfg.vb_size;
    return HSB_OK;
}

hsb_ctx *hsb_create(int device, int impl) {
    hsb::ImplConfig cfg;
    if (!hsb::impl_config(impl, &cfg)) { set_err(HSB_EINVAL, "unknown impl"); return nullptr; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || device < 0 || device >= n) {
        set_err(HSB_ECUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    hsb_ctx *c = new hsb_ctx;
    c->device = device; c->impl = impl; c->cfg = cfg;
    c->arith = impl == HSB_IMPL_FIXED ? hsb::kArithFixed : hsb::kArithFloat;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_err(HSB_ECUDA, std::string("device setup failed: ") + cudaGetErrorString(cudaGetLastError()));
        delete c;
        return nullptr;
    }
    if (prop.major != 10) {
        set_err(HSB_ECUDA, "this build contains sm_100a code only; device is not compute capability 10.x");
        cudaStreamDestroy(c->stream);
        delete c;
        return nullptr;
    }
    c->sm_count = prop.multiProcessorCount * hsb::kCtasPerSm;    // CTA slots the launch plans are cut for
    e = cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_h2d_b, cudaStreamNonBlocking);
    for (int i = 0; i < 2; i++) if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_h2d_more[i], cudaStreamNonBlocking);
    if (const char *v = std::getenv("HSB_ITERATE_PERSISTENT")) c->iterate_persistent = std::atoi(v) != 0;
    {   // the one-launch iteration needs cooperative launches (every CTA of the grid resident)
        int coop = 0;
        if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) != cudaSuccess || !coop) { cudaGetLastError(); c->iterate_persistent = false; }
    }
    if (const char *v = std::getenv("HSB_UPLOAD_STREAMS")) c->n_upload_streams = std::min(4, std::max(1, std::atoi(v)));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking);
    cudaEvent_t *evs[] = {&c->ev_xready, &c->ev_xfree[0], &c->ev_xfree[1], &c->ev_yready, &c->ev_ydone};
    for (cudaEvent_t *ev : evs)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = hsb::configure_kernels(prop.multiProcessorCount);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_flags, kNumFlags * 4);
    if (e == cudaSuccess) e = cudaMemset(c->d_flags, 0, kNumFlags * 4);
    if (e == cudaSuccess && load_memops()) {
        void *hd = nullptr;
        if (e == cudaSuccess) e = cudaHostAlloc(&hd, 64, cudaHostAllocMapped);
        if (e == cudaSuccess) { std::memset(hd, 0, 64); c->h_done = (volatile uint32_t *)hd; }
        if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&c->d_done, hd, 0);
        if (e == cudaSuccess) e = cudaHostAlloc((void **)&c->h_seq_ring, 256 * 4, cudaHostAllocDefault);
        c->flags_mode = e == cudaSuccess;
        // test aid: start every sequence counter (and the flag words they are compared with) just below a 32-bit
        // wrap-around instead of at 0, so that a short run crosses it
        if (const char *v = std::getenv("HSB_DEBUG_SEQ_BASE")) {
            const uint64_t base = std::strtoull(v, nullptr, 0);
            const uint32_t b32 = (uint32_t)base;
            uint32_t init[kNumFlags];
            for (int i = 0; i < kNumFlags; i++) init[i] = (i == kFlagError || i >= kFlagIterBarrier) ? 0u : b32;
            if (e == cudaSuccess) e = cudaMemcpy(c->d_flags, init, sizeof init, cudaMemcpyHostToDevice);
            c->h_done[0] = b32;
            c->seq_base = c->launch_seq = c->publish_sure = c->d2h_wait_seq = base;
            c->x_seq = c->dl_seq = c->x_wait_val = b32;
            c->y_dl_seq[0] = c->y_dl_seq[1] = b32;
        }
        if (const char *v = std::getenv("HSB_XFLAG_COPY")) c->xflag_copy = std::atoi(v);
    }
    if (e != cudaSuccess) {
        set_err(HSB_ECUDA, std::string("kernel configuration failed: ") + cudaGetErrorString(e));
        cudaStreamDestroy(c->stream);
        delete c;
        return nullptr;
    }
    return c;
}

void hsb_destroy(hsb_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    quiesce(c);
    free_matrix(c);
    cudaFree(c->d_trace);
    cudaFree(c->d_timeline);
    cudaFree(c->d_flags);
    if (c->h_done) cudaFreeHost((void *)c->h_done);
    if (c->h_seq_ring) cudaFreeHost(c->h_seq_ring);
    cudaEvent_t evs[] = {c->ev_xready, c->ev_xfree[0], c->ev_xfree[1], c->ev_yready, c->ev_ydone};
    for (cudaEvent_t ev : evs) if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(c->s_h2d);
    cudaStreamDestroy(c->s_h2d_b);
    for (int i = 0; i < 2; i++) cudaStreamDestroy(c->s_h2d_more[i]);
    cudaStreamDestroy(c->s_d2h);
    cudaStreamDestroy(c->stream);
    delete c;
}

void *hsb_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (!bytes) bytes = 1;
    {
        PreferDeviceNode numa;
        if (cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    void *dev = nullptr;
    if (cudaHostGetDevicePointer(&dev, p, 0) != cudaSuccess) { cudaGetLastError(); dev = nullptr; }
    std::lock_guard<std::mutex> lk(g_host_mu);
    g_host_allocs[(uintptr_t)p] = HostAlloc{bytes, (uintptr_t)dev};
    return p;
}
void hsb_host_free(void *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        g_host_allocs.erase((uintptr_t)p);
    }
    cudaFreeHost(p);
}

int hsb_upload_matrix_csr(hsb_ctx *c, uint32_t rows, uint32_t cols, const uint32_t *indptr,
                          const uint32_t *indices, const void *vals, uint32_t rows_per_partition) {
    if (!c || !indptr || (rows && indptr[rows] && (!indices || !vals))) return set_err(HSB_EINVAL, "null argument");
    // formatting runs on the device by default (gpu_format.cu); HSB_HOST_FORMAT=1 selects the host builder
    static const bool host_format = std::getenv("HSB_HOST_FORMAT") != nullptr;
    if (!host_format) return hsb_upload_matrix_csr_gpu(c, rows, cols, indptr, indices, vals, rows_per_partition);
    auto t0 = std::chrono::steady_clock::now();
    hsb::TiledMatrix M;
    std::string err;
    if (!hsb::build_tiled(rows, cols, indptr, indices, (const uint32_t *)vals, rows_per_partition,
                          hsb::choose_tile_cols(cols, rows, rows ? indptr[rows] : 0), 0, &M, &err))
        return set_err(HSB_EINVAL, "malformed CSR: " + err);
    c->preprocess_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return upload_tiled(c, M);
}

}  // extern "C"

namespace {
// device-resident CSR (d_indptr) or COO (d_coo_rows) -> tile streams on the device -> resident matrix
int upload_device_matrix(hsb_ctx *c, uint32_t rows, uint32_t cols, uint64_t nnz, const uint32_t *d_indptr, const uint32_t *d_coo_rows,
                         const uint32_t *d_indices, const void *d_vals, uint32_t rows_per_partition) {
    CUDA_TRY(cudaSetDevice(c->device));
    { int rc = quiesce(c); if (rc) return rc; }
    auto t0 = std::chrono::steady_clock::now();
    free_matrix(c);
    hsb::TiledMatrix M;
    hsb::DeviceFormat f;
    std::string err;
    cudaError_t e = hsb::build_tiled_gpu(rows, cols, nnz, d_indptr, d_coo_rows, d_indices, (const uint32_t *)d_vals,
                                         rows_per_partition, hsb::choose_tile_cols(cols, rows, nnz), c->stream, &M, &f, &err);
    if (e != cudaSuccess) {
        cudaFree(f.vals); cudaFree(f.cols); cudaFree(f.slice_rows);
        cudaGetLastError();
        if (!err.empty()) return set_err(HSB_EINVAL, "malformed matrix: " + err);
        return set_err(HSB_ECUDA, std::string("GPU formatting failed: ") + cudaGetErrorString(e));
    }
    c->preprocess_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    DeviceMatrix d;
    d.vals = f.vals; d.cols = f.cols; d.slice_rows = f.slice_rows;
    if (!d.vals) {                                       // empty matrix: the kernel still wants valid pointers
        CUDA_TRY(cudaMalloc(&d.vals, 16)); CUDA_TRY(cudaMalloc(&d.cols, 16)); CUDA_TRY(cudaMalloc(&d.slice_rows, 16));
    }
    return install_matrix(c, M, d, f.n_elems, f.n_slices);
}
}  // namespace

extern "C" {

int hsb_upload_matrix_csr_device(hsb_ctx *c, uint32_t rows, uint32_t cols, uint64_t nnz, const uint32_t *d_indptr,
                                 const uint32_t *d_indices, const void *d_vals, uint32_t rows_per_partition) {
    if (!c || !d_indptr || (nnz && (!d_indices || !d_vals))) return set_err(HSB_EINVAL, "null argument");
    return upload_device_matrix(c, rows, cols, nnz, d_indptr, nullptr, d_indices, d_vals, rows_per_partition);
}

int hsb_upload_matrix_csr_gpu(hsb_ctx *c, uint32_t rows, uint32_t cols, const uint32_t *indptr,
                              const uint32_t *indices, const void *vals, uint32_t rows_per_partition) {
    if (!c || !indptr || (rows && indptr[rows] && (!indices || !vals))) return set_err(HSB_EINVAL, "null argument");
    for (uint32_t r = 0; r < rows; r++)
        if (indptr[r + 1] < indptr[r]) return set_err(HSB_EINVAL, "malformed CSR: indptr is not monotone");
    if (rows && indptr[0] != 0) return set_err(HSB_EINVAL, "malformed CSR: indptr[0] must be 0");
    CUDA_TRY(cudaSetDevice(c->device));
    const uint64_t nnz = rows ? indptr[rows] : 0;
    struct Tmp {                                           // freed on every path out, the error returns included
        uint32_t *ip = nullptr, *ix = nullptr, *v = nullptr;
        ~Tmp() { cudaFree(ip); cudaFree(ix); cudaFree(v); }
    } t;
    CUDA_TRY(cudaMalloc(&t.ip, ((size_t)rows + 1) * 4));
    CUDA_TRY(cudaMalloc(&t.ix, std::max<uint64_t>(nnz, 1) * 4));
    CUDA_TRY(cudaMalloc(&t.v, std::max<uint64_t>(nnz, 1) * 4));
    CUDA_TRY(cudaMemcpyAsync(t.ip, indptr, ((size_t)rows + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    if (nnz) {
        CUDA_TRY(cudaMemcpyAsync(t.ix, indices, nnz * 4, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(t.v, vals, nnz * 4, cudaMemcpyHostToDevice, c->stream));
    }
    return hsb_upload_matrix_csr_device(c, rows, cols, nnz, t.ip, t.ix, t.v, rows_per_partition);
}

int hsb_upload_matrix_cpsr(hsb_ctx *c, const void *const ch[HSB_NUM_HBM_CHANNELS],
                           const size_t ch_packets[HSB_NUM_HBM_CHANNELS], unsigned num_row_partitions,
                           unsigned num_col_partitions, unsigned num_rows, unsigned num_cols) {
    if (!c || !ch || !ch_packets) return set_err(HSB_EINVAL, "null argument");
    if (num_rows % 128u || num_cols % 8u) return set_err(HSB_EINVAL, "rows must divide 128 and cols 8 (util_round_csr_matrix_dim)");
    if (num_row_partitions != (num_rows + c->cfg.ob_size - 1) / c->cfg.ob_size ||
        num_col_partitions != (num_cols + c->cfg.vb_size - 1) / c->cfg.vb_size)
        return set_err(HSB_EINVAL, "partition counts do not match the implementation's buffer sizes");
    auto t0 = std::chrono::steady_clock::now();
    const uint32_t *imgs[16];
    for (int i = 0; i < 16; i++) imgs[i] = (const uint32_t *)ch[i];
    static const bool host_format = std::getenv("HSB_HOST_FORMAT") != nullptr;
    if (!host_format) {
        // the channel images go to HBM as they are; decoding and re-formatting run on the device
        CUDA_TRY(cudaSetDevice(c->device));
        uint32_t *d_rows = nullptr, *d_ix = nullptr, *d_v = nullptr;     // a COO list in image order
        uint64_t nnz = 0;
        std::string derr;
        cudaError_t e = hsb::cpsr_decode_gpu(c->cfg, imgs, ch_packets, num_row_partitions, num_col_partitions, num_rows,
                                             num_cols, c->stream, &d_rows, &d_ix, &d_v, &nnz, &derr);
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (!derr.empty()) return set_err(HSB_EINVAL, "malformed CPSR image: " + derr);
            return set_err(HSB_ECUDA, std::string("GPU decoding failed: ") + cudaGetErrorString(e));
        }
        const auto t1 = std::chrono::steady_clock::now();
        int rc = upload_device_matrix(c, num_rows, num_cols, nnz, nullptr, d_rows, d_ix, d_v, c->cfg.ob_size);
        cudaFree(d_rows); cudaFree(d_ix); cudaFree(d_v);
        c->preprocess_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        static const bool debug = std::getenv("HSB_DEBUG_PLAN") != nullptr;
        if (debug)
            std::fprintf(stderr, "[hsb cpsr] copy + decode of the channel images %.1f ms, format + install %.1f ms\n",
                         1e3 * std::chrono::duration<double>(t1 - t0).count(),
                         1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
        return rc;
    }
    std::vector<uint32_t> rows_in(num_row_partitions);
    for (unsigned j = 0; j < num_row_partitions; j++)
        rows_in[j] = std::min<uint64_t>(c->cfg.ob_size, (uint64_t)num_rows - (uint64_t)j * c->cfg.ob_size);
    hsb::HostCsr csr;
    std::string err;
    if (!hsb::cpsr_decode(c->cfg, imgs, ch_packets, num_row_partitions, num_col_partitions, 0, num_row_partitions,
                          rows_in.data(), num_cols, &csr, &err))
        return set_err(HSB_EINVAL, "malformed CPSR image: " + err);
    int rc = hsb_upload_matrix_csr(c, csr.rows, csr.cols, csr.indptr.data(), csr.indices.data(), csr.vals.data(),
                                   c->cfg.ob_size);
    c->preprocess_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

int hsb_upload_vector(hsb_ctx *c, const void *x_packed, unsigned num_cols) {
    if (!c || !x_packed) return set_err(HSB_EINVAL, "null argument");
    if (!c->have_matrix) return set_err(HSB_ESTATE, "upload a matrix first");
    if (num_cols > c->x_words || num_cols < c->cols) return set_err(HSB_EINVAL, "num_cols does not match the matrix");
    CUDA_TRY(cudaSetDevice(c->device));
    // write the buffer no in-flight launch reads: its last readers were launched before the previous
    // upload, which is when ev_xfree[b] was recorded on the compute stream
    if (c->flags_mode) {
        // Four buffers in rotation: this one was last read three launches ago. Instead of queueing a
        // stream wait (a stream memory operation costs ~3 us on the copy stream) the host looks at
        // done_seq itself -- in steady state the number is already there -- and then queues the copy and
        // the flag that tells the next launch its x has landed.
        const int b = (c->x_latest + 1) % kXBuffers;
        const uint64_t need = c->x_reader_seq[b];
        if (need) {
            if (need > c->publish_sure) { int rc = publish_done(c); if (rc) return rc; }
            const auto t0 = std::chrono::steady_clock::now();
            // bounded back-off: a pause per poll (the mapped word is written over PCIe; hammering it from several
            // ranks' host threads on one socket slows everybody's posted writes), a yield every 64 K polls
            for (unsigned spins = 0; done_launch(c) < need; spins++) {
#if defined(__x86_64__) || defined(__i386__)
                __builtin_ia32_pause();
#endif
                if ((spins & 0xFFFFu) == 0xFFFFu) {
                    sched_yield();
                    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 20.0)
                        return set_err(HSB_ECUDA, "timed out waiting for the launch that still reads the x buffer");
                }
            }
        }
        // The "landed" flag behind the copy is itself a 4-byte copy from a ring of page-locked sequence words: a pure
        // copy-engine operation. A stream memory operation in its place costs 1.7 us more per SpMV in the host-buffer
        // pipeline (C2: 18.8 -> 16.9 us; hsb_set_option "xflag_copy" / HSB_XFLAG_COPY=0 for the A/B). Tried and
        // dropped: the vector as two halves on the two copy streams with two flags (upload 15.2 -> 19.6 us per vector).
        cudaStream_t all_up[4] = {c->s_h2d, c->s_h2d_b, c->s_h2d_more[0], c->s_h2d_more[1]};
        cudaStream_t up = all_up[c->x_seq % (unsigned)c->n_upload_streams];
        CUDA_TRY(cudaMemcpyAsync(c->d_x[b], x_packed, (size_t)num_cols * 4, cudaMemcpyHostToDevice, up));
        c->x_wait_val = ++c->x_seq;
        c->x_wait_buf = b;
        c->x_wait_launch = 0;
        uint32_t *flag = c->d_flags + kFlagXReady + b;
        if (c->xflag_copy) {
            uint32_t *word = c->h_seq_ring + (c->x_wait_val & 255u);     // rewritten 256 uploads later; at most four are in flight
            *word = c->x_wait_val;
            CUDA_TRY(cudaMemcpyAsync(flag, word, 4, cudaMemcpyHostToDevice, up));
        } else {
            MEMOP_TRY(g_write32((CUstream)up, (CUdeviceptr)flag, c->x_wait_val, 0));
        }
        c->x_latest = b;
        return HSB_OK;
    }
    const int b = c->x_latest ^ 1;
    CUDA_TRY(cudaStreamWaitEvent(c->s_h2d, c->ev_xfree[b], 0));
    CUDA_TRY(cudaMemcpyAsync(c->d_x[b], x_packed, (size_t)num_cols * 4, cudaMemcpyHostToDevice, c->s_h2d));
    CUDA_TRY(cudaEventRecord(c->ev_xready, c->s_h2d));
    CUDA_TRY(cudaEventRecord(c->ev_xfree[c->x_latest], c->stream));   // "every launch so far that reads the old buffer"
    c->x_latest = b;
    c->x_dirty = true;
    return HSB_OK;
}

int hsb_spmv_row_partition(hsb_ctx *c, unsigned row_part_id, unsigned part_len, unsigned num_col_partitions,
                           unsigned num_partitions, unsigned num_cols) {
    if (!c) return set_err(HSB_EINVAL, "null context");
    if (!c->have_matrix) return set_err(HSB_ESTATE, "upload a matrix first");
    if (row_part_id >= c->n_row_parts) return set_err(HSB_EINVAL, "row_part_id out of range");
    uint32_t rb = row_part_id * c->rows_per_part;
    uint32_t re = (uint32_t)std::min<uint64_t>(c->rows, (uint64_t)rb + c->rows_per_part);
    if ((uint64_t)part_len * HSB_NUM_HBM_CHANNELS != re - rb)
        return set_err(HSB_EINVAL, "part_len does not match the rows of this partition");
    if (num_cols < c->cols || num_cols > c->x_words || (num_col_partitions && num_partitions % num_col_partitions))
        return set_err(HSB_EINVAL, "num_cols / partition counts do not match the matrix");
    CUDA_TRY(cudaSetDevice(c->device));
    return run_slot(c, 1 + row_part_id, rb, re, nullptr, nullptr);
}

int hsb_spmv(hsb_ctx *c) {
    if (!c) return set_err(HSB_EINVAL, "null context");
    if (!c->have_matrix) return set_err(HSB_ESTATE, "upload a matrix first");
    CUDA_TRY(cudaSetDevice(c->device));
    return run_all(c, nullptr, nullptr);
}

int hsb_sync(hsb_ctx *c) {
    if (!c) return set_err(HSB_EINVAL, "null context");
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->have_matrix) { int rc = finish(c); if (rc) return rc; }
    CUDA_TRY(cudaStreamSynchronize(c->s_h2d));
    CUDA_TRY(cudaStreamSynchronize(c->s_h2d_b));
    for (int i = 0; i < 2; i++) CUDA_TRY(cudaStreamSynchronize(c->s_h2d_more[i]));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->s_d2h));
    c->y_busy[0] = c->y_busy[1] = false;
    return check_error_flag(c);
}

int hsb_download_result_async(hsb_ctx *c, void *y_packed, unsigned num_rows) {
    if (!c || !y_packed) return set_err(HSB_EINVAL, "null argument");
    if (!c->have_matrix) return set_err(HSB_ESTATE, "upload a matrix first");
    if (num_rows > c->rows) return set_err(HSB_EINVAL, "num_rows exceeds the matrix");
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->pending_dl.active) { int rc = finish(c); if (rc) return rc; }
    // (pageable destinations are never deferred: cudaMemcpyAsync into pageable memory blocks the calling
    // thread until the copy has run, and a deferred copy waits for a launch this thread has yet to issue)
    uint32_t *alias = c->flags_mode && c->drain_pending ? mapped_alias(c, y_packed, num_rows) : nullptr;
    if (alias && c->drain_begin == 0 && c->drain_end == c->rows) {
        // Deferred: the sums of the last SpMV are still in the row accumulators. The next whole-matrix
        // hsb_spmv drains them in its prologue anyway; the copy is attached to that launch (run_slot), or
        // issued behind an explicit drain by hsb_sync / the blocking download, whichever comes first.
        c->pending_dl.host = y_packed; c->pending_dl.n = num_rows; c->pending_dl.active = true;
        c->pending_dl.dev = c->host_drain ? alias : nullptr;
        return HSB_OK;
    }
    { int rc = finish(c); if (rc) return rc; }
    return issue_copy_after_main(c, y_packed, num_rows);
}

int hsb_download_result(hsb_ctx *c, void *y_packed, unsigned num_rows) {
    int rc = hsb_download_result_async(c, y_packed, num_rows);
    if (rc) return rc;
    rc = finish(c);                                       // a deferred download is resolved now: by the drain kernel itself
    if (rc) return rc;                                    // (mapped page-locked destination) or by the copy engine
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->s_d2h));
    c->y_busy[0] = c->y_busy[1] = false;
    return check_error_flag(c);      // a launch that gave up on a flag produced no row updates: never hand that out as y
}

int hsb_top_wrapper(int impl, const void *const matrix_hbm[HSB_NUM_HBM_CHANNELS], const void *x, void *y,
                    unsigned row_part_id, unsigned part_len, unsigned num_col_partitions,
                    unsigned num_partitions, unsigned num_cols) {
    hsb::ImplConfig cfg;
    if (!hsb::impl_config(impl, &cfg)) return set_err(HSB_EINVAL, "unknown impl");
    if (!matrix_hbm || !x || !y || !num_col_partitions || num_partitions % num_col_partitions)
        return set_err(HSB_EINVAL, "bad argument");
    const unsigned nrp = num_partitions / num_col_partitions;
    if (row_part_id >= nrp) return set_err(HSB_EINVAL, "row_part_id out of range");
    const uint32_t *imgs[16];
    size_t lens[16];
    for (int i = 0; i < 16; i++) {
        imgs[i] = (const uint32_t *)matrix_hbm[i];
        lens[i] = hsb::cpsr_image_packets(cfg, imgs[i], num_partitions);
    }
    uint32_t rows_here = part_len * HSB_NUM_HBM_CHANNELS;
    hsb::HostCsr csr;
    std::string err;
    if (!hsb::cpsr_decode(cfg, imgs, lens, nrp, num_col_partitions, row_part_id, row_part_id + 1, &rows_here,
                          num_cols, &csr, &err))
        return set_err(HSB_EINVAL, "malformed CPSR image: " + err);
    // the calling thread's current device (cudaSetDevice), or HSB_DEVICE
    int dev = 0;
    if (const char *e = std::getenv("HSB_DEVICE")) dev = std::atoi(e);
    else if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    hsb_ctx *c = hsb_create(dev, impl);
    if (!c) return HSB_ECUDA;
    int rc = hsb_upload_matrix_csr(c, csr.rows, csr.cols, csr.indptr.data(), csr.indices.data(), csr.vals.data(), 0);
    if (rc == HSB_OK) rc = hsb_upload_vector(c, x, num_cols);
    if (rc == HSB_OK) rc = hsb_spmv(c);
    // the result drain writes at row_part_id * LOGICAL_OB_SIZE (spmv_result_drain.cpp:36)
    if (rc == HSB_OK) rc = hsb_download_result(c, (uint32_t *)y + (size_t)row_part_id * cfg.ob_size, rows_here);
    std::string keep = g_err;
    hsb_destroy(c);
    g_err = keep;
    return rc;
}

int hsb_get_stats(hsb_ctx *c, hsb_stats *out) {
    if (!c || !out) return set_err(HSB_EINVAL, "null argument");
    std::memset(out, 0, sizeof *out);
    out->nnz = c->nnz; out->rows = c->rows; out->cols = c->cols;
    out->n_row_parts = c->n_row_parts; out->n_col_tiles = c->n_col_tiles; out->tile_cols = c->tile_cols;
    out->n_slices = c->n_slices; out->n_streams = c->n_streams; out->n_elems = c->n_elems;
    out->format_bytes = c->format_bytes;
    out->algorithmic_bytes = 8ull * c->nnz + 4ull * ((uint64_t)c->rows + 1) + 4ull * c->rows + 4ull * c->cols;
    out->kernel_launches = c->launches; out->sm_count = c->sm_count; out->grid = c->grid;
    out->replicas = (uint32_t)c->mats.size(); out->preprocess_seconds = c->preprocess_s;
    out->layout = c->meta.narrow ? 1u : 0u;
    return HSB_OK;
}

int hsb_set_replicas(hsb_ctx *c, int n) {
    if (!c || n < 1) return set_err(HSB_EINVAL, "bad argument");
    if (!c->have_matrix) return set_err(HSB_ESTATE, "upload a matrix first");
    CUDA_TRY(cudaSetDevice(c->device));
    while ((int)c->mats.size() > n) { c->mats.back().release(); c->mats.pop_back(); }
    while ((int)c->mats.size() < n) {
        DeviceMatrix d;
        const DeviceMatrix &s = c->mats[0];
        CUDA_TRY(cudaMalloc(&d.vals, c->sz_vals + 16));
        CUDA_TRY(cudaMalloc(&d.cols, c->sz_cols + 16));
            CUDA_TRY(cudaMalloc(&d.slice_rows, c->sz_rows + 16));
        CUDA_TRY(cudaMemcpyAsync(d.vals, s.vals, c->sz_vals, cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(d.cols, s.cols, c->sz_cols, cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(d.slice_rows, s.slice_rows, c->sz_rows, cudaMemcpyDeviceToDevice, c->stream));
        c->mats.push_back(d);
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return HSB_OK;
}

int hsb_time_spmv(hsb_ctx *c, int warmup, int steps, float *step_ms, float *kernel_ms) {
    if (!c || steps < 1 || warmup < 0) return set_err(HSB_EINVAL, "bad argument");
    if (!c->have_matrix) return set_err(HSB_ESTATE, "upload a matrix first");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    for (int i = 0; i < warmup; i++) { int rc = run_all(c, nullptr, nullptr); if (rc) return rc; }
    { int rc = finish(c); if (rc) return rc; }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventRecord(e0, c->stream));
    for (int i = 0; i < steps; i++) { int rc = run_all(c, nullptr, nullptr); if (rc) return rc; }
    { int rc = finish(c); if (rc) return rc; }            // the last drain belongs to the timed work
    CUDA_TRY(cudaEventRecord(e1, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (step_ms) *step_ms = ms / steps;
    if (kernel_ms) {
        std::vector<cudaEvent_t> ev(2 * (size_t)steps);
        for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e));
        for (int i = 0; i < steps; i++) { int rc = run_all(c, ev[2 * i], ev[2 * i + 1]); if (rc) return rc; }
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        double tot = 0;
        for (int i = 0; i < steps; i++) { float t = 0; CUDA_TRY(cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1])); tot += t; }
        for (auto &e : ev) cudaEventDestroy(e);
        *kernel_ms = (float)(tot / steps);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return HSB_OK;
}

int hsb_time_e2e(hsb_ctx *c, const void *const x_host[2], void *const y_host[2], unsigned num_cols, unsigned num_rows,
                 int iters, int async_download, double *seconds_per_spmv) {
    if (!c || !x_host || !y_host || !x_host[0] || !x_host[1] || !y_host[0] || !y_host[1] || iters < 1 || !seconds_per_spmv)
        return set_err(HSB_EINVAL, "bad argument");
    int rc = hsb_sync(c);
    if (rc) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < iters; k++) {
        if ((rc = hsb_upload_vector(c, x_host[k & 1], num_cols))) return rc;
        if ((rc = hsb_spmv(c))) return rc;
        rc = async_download ? hsb_download_result_async(c, y_host[k & 1], num_rows)
                            : hsb_download_result(c, y_host[k & 1], num_rows);
        if (rc) return rc;
    }
    if ((rc = hsb_sync(c))) return rc;
    *seconds_per_spmv = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / iters;
    return HSB_OK;
}

// ---- iterative callers: x <- alpha (*) y (+) beta on the device ------------------------------------------
int hsb_axpb_to_vector(hsb_ctx *c, uint32_t alpha_word, uint32_t beta_word, uint32_t col_offset) {
    if (!c) return set_err(HSB_EINVAL, "null context");
    if (!c->have_matrix) return set_err(HSB_ESTATE, "upload a matrix first");
    if (col_offset >= c->x_words) return set_err(HSB_EINVAL, "col_offset beyond the vector");
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->pending_dl.active) { int rc = finish(c); if (rc) return rc; }
    // the buffer being written must not be the target of an upload still in flight
    CUDA_TRY(cudaStreamSynchronize(c->s_h2d));
    CUDA_TRY(cudaStreamSynchronize(c->s_h2d_b));
    for (int i = 0; i < 2; i++) CUDA_TRY(cudaStreamSynchronize(c->s_h2d_more[i]));
    const int nb = c->flags_mode ? kXBuffers : 2;
    const int b = (c->x_latest + 1) % nb;
    const int yb = c->y_cur;
    if (c->y_busy[yb]) {                                   // a download still reads y: order the rewrite behind it
        if (c->flags_mode)
            MEMOP_TRY(g_wait32((CUstream)c->stream, (CUdeviceptr)(c->d_flags + kFlagYFree + yb), c->y_dl_seq[yb], CU_STREAM_WAIT_VALUE_GEQ));
        else
            CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_ydone, 0));
        c->y_busy[yb] = false;
    }
    void *acc = nullptr;
    if (c->drain_pending && c->drain_begin == 0 && c->drain_end == c->rows) {
        acc = c->d_acc[(c->acc_cur + c->acc_bufs - 1) % c->acc_bufs];       // fused: drain + update in one pass
        c->drain_pending = false;
    } else {
        int rc = finish(c);
        if (rc) return rc;
    }
    // an ordinary (non-programmatic) launch: it starts when every SpMV before it has completed, and the next
    // SpMV -- which reads the vector written here -- starts when it has completed
    CUDA_TRY(hsb::launch_axpb(c->arith, acc, c->d_y[yb], c->d_x[b], c->rows, c->x_words, alpha_word, beta_word, col_offset,
                              c->rows, c->stream));
    c->launches++;
    c->x_next_buf = b;
    return HSB_OK;
}

void *hsb_device_x_next(hsb_ctx *c) {
    if (!c || !c->have_matrix) return nullptr;
    const int nb = c->flags_mode ? kXBuffers : 2;
    return c->d_x[c->x_next_buf >= 0 ? c->x_next_buf : (c->x_latest + 1) % nb];
}

int hsb_vector_commit(hsb_ctx *c) {
    if (!c) return set_err(HSB_EINVAL, "null context");
    if (!c->have_matrix) return set_err(HSB_ESTATE, "upload a matrix first");
    const int nb = c->flags_mode ? kXBuffers : 2;
    // written by an axpb kernel on the compute stream (the peer form is announced by arrival flags instead, this rank's own slice included)
    c->x_after_grid = c->x_next_buf >= 0 && !c->x_next_from_peers;
    c->x_latest = c->x_next_buf >= 0 ? c->x_next_buf : (c->x_latest + 1) % nb;
    c->x_next_buf = -1;
    c->x_wait_buf = -1;                                    // written on the compute stream: plain stream order
    if (c->x_next_from_peers) {                            // ... except the other ranks' slices: arrival flags
        c->x_wait_buf = kXBuffers; c->x_wait_val = c->peer_seq; c->x_wait_launch = 0;
        c->x_next_from_peers = false;
    }
    c->x_dirty = false;
    if (!c->flags_mode) CUDA_TRY(cudaEventRecord(c->ev_xfree[c->x_latest ^ 1], c->stream));
    return HSB_OK;
}

// ---- the same across GPUs, over peer memory (one process per GPU, CUDA IPC) ------------------------------
namespace {
struct PeerBlob {                      // what hsb_peer_export hands to the other ranks (HSB_PEER_BLOB_BYTES)
    cudaIpcMemHandle_t x, flags;
    uint64_t x_stride;
    uint32_t x_words, x_latest;
    // contexts of ONE process (one host thread per GPU, or several contexts on one GPU) cannot open their own
    // IPC handles: they recognise each other by pid and use the raw pointers (peer access enabled on demand)
    uint64_t pid, raw_x, raw_flags;
    int32_t device;
};
static_assert(sizeof(PeerBlob) <= HSB_PEER_BLOB_BYTES, "blob size");

struct GatherBlob {                    // hsb_gather_export
    cudaIpcMemHandle_t y, flags;
    uint64_t pid, raw_y, raw_flags;
    uint32_t total_rows, has_buffer;
    int32_t device;
};
static_assert(sizeof(GatherBlob) <= HSB_PEER_BLOB_BYTES, "blob size");

// make memory of `peer_device` (same process) addressable from the current device
cudaError_t enable_peer(int my_device, int peer_device) {
    if (my_device == peer_device) return cudaSuccess;
    int can = 0;
    cudaError_t e = cudaDeviceCanAccessPeer(&can, my_device, peer_device);
    if (e != cudaSuccess) return e;
    if (!can) return cudaErrorPeerAccessUnsupported;
    e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    return e;
}
}  // namespace

int hsb_peer_export(hsb_ctx *c, void *blob) {
    if (!c || !blob) return set_err(HSB_EINVAL, "null argument");
    if (!c->have_matrix) return set_err(HSB_ESTATE, "upload a matrix first");
    CUDA_TRY(cudaSetDevice(c->device));
    { int rc = quiesce(c); if (rc) return rc; }
    PeerBlob b;
    std::memset(&b, 0, sizeof b);
    CUDA_TRY(cudaIpcGetMemHandle(&b.x, c->d_x[0]));
    CUDA_TRY(cudaIpcGetMemHandle(&b.flags, c->d_peer));
    b.x_stride = c->x_stride; b.x_words = c->x_words; b.x_latest = (uint32_t)c->x_latest;
    b.pid = (uint64_t)getpid(); b.raw_x = (uint64_t)(uintptr_t)c->d_x[0]; b.raw_flags = (uint64_t)(uintptr_t)c->d_peer;
    b.device = c->device;
    std::memset(blob, 0, HSB_PEER_BLOB_BYTES);
    std::memcpy(blob, &b, sizeof b);
    return HSB_OK;
}

int hsb_peer_connect(hsb_ctx *c, int world, int rank, const void *blobs) {
    if (!c || !blobs || world < 1 || world > hsb::kMaxPeers || rank < 0 || rank >= world)
        return set_err(HSB_EINVAL, "bad argument");
    if (!c->have_matrix) return set_err(HSB_ESTATE, "upload a matrix first");
    if (!c->flags_mode) return set_err(HSB_ESTATE, "the peer iteration needs the flag pipeline (stream memory operations)");
    CUDA_TRY(cudaSetDevice(c->device));
    { int rc = quiesce(c); if (rc) return rc; }
    close_peers(c);                        // a repeated connect replaces the old mappings instead of leaking them
    CUDA_TRY(cudaMemsetAsync(c->d_peer, 0, 64 * 4, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int g = 0; g < world; g++) {
        PeerBlob b;
        std::memcpy(&b, (const char *)blobs + (size_t)g * HSB_PEER_BLOB_BYTES, sizeof b);
        if (b.x_words != c->x_words || b.x_stride != c->x_stride || (int)b.x_latest != c->x_latest) {
            close_peers(c);
            return set_err(HSB_EINVAL, "rank " + std::to_string(g) + " has a different vector layout or upload history");
        }
        if (g == rank) { c->peer_x[g] = c->d_x[0]; c->peer_flags[g] = c->d_peer; continue; }
        cudaError_t e;
        if (b.pid == (uint64_t)getpid()) {
            e = enable_peer(c->device, b.device);
            c->peer_x[g] = (uint32_t *)(uintptr_t)b.raw_x; c->peer_flags[g] = (uint32_t *)(uintptr_t)b.raw_flags;
        } else {
            void *px = nullptr, *pf = nullptr;
            e = cudaIpcOpenMemHandle(&px, b.x, cudaIpcMemLazyEnablePeerAccess);
            if (e == cudaSuccess) {
                e = cudaIpcOpenMemHandle(&pf, b.flags, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) cudaIpcCloseMemHandle(px);
            }
            if (e == cudaSuccess) { c->peer_x[g] = (uint32_t *)px; c->peer_flags[g] = (uint32_t *)pf; c->peer_ipc[g] = true; }
        }
        if (e != cudaSuccess) {
            close_peers(c);
            cudaGetLastError();
            return set_err(HSB_ECUDA, "cannot map the vector buffers of rank " + std::to_string(g) + ": " + cudaGetErrorString(e));
        }
    }
    c->peer_world = world; c->peer_rank = rank; c->peer_seq = 0;
    return HSB_OK;
}

// ---- gather of y across row-block shards: the drain's epilogue stores into the targets' buffers --------------
int hsb_gather_export(hsb_ctx *c, uint32_t total_rows, int want_buffer, void *blob) {
    if (!c || !blob || !total_rows) return set_err(HSB_EINVAL, "bad argument");
    if (!c->have_matrix) return set_err(HSB_ESTATE, "upload a matrix first");
    CUDA_TRY(cudaSetDevice(c->device));
    { int rc = quiesce(c); if (rc) return rc; }
    close_gather(c);
    CUDA_TRY(cudaMalloc(&c->d_gather_flags, 64 * 4));
    CUDA_TRY(cudaMemset(c->d_gather_flags, 0, 64 * 4));
    CUDA_TRY(cudaMalloc(&c->d_gather_y, (want_buffer ? (size_t)total_rows : 1) * 4));
    CUDA_TRY(cudaMemset(c->d_gather_y, 0, (want_buffer ? (size_t)total_rows : 1) * 4));
    c->gather_total_rows = total_rows;
    c->gather_is_target = want_buffer != 0;
    GatherBlob b;
    std::memset(&b, 0, sizeof b);
    CUDA_TRY(cudaIpcGetMemHandle(&b.y, c->d_gather_y));
    CUDA_TRY(cudaIpcGetMemHandle(&b.flags, c->d_gather_flags));
    b.pid = (uint64_t)getpid(); b.raw_y = (uint64_t)(uintptr_t)c->d_gather_y; b.raw_flags = (uint64_t)(uintptr_t)c->d_gather_flags;
    b.total_rows = total_rows; b.has_buffer = want_buffer ? 1u : 0u; b.device = c->device;
    std::memset(blob, 0, HSB_PEER_BLOB_BYTES);
    std::memcpy(blob, &b, sizeof b);
    return HSB_OK;
}

int hsb_gather_connect(hsb_ctx *c, int world, int rank, uint32_t row_offset, const void *blobs) {
    if (!c || !blobs || world < 1 || world > hsb::kMaxPeers || rank < 0 || rank >= world)
        return set_err(HSB_EINVAL, "bad argument");
    if (!c->have_matrix || !c->d_gather_flags) return set_err(HSB_ESTATE, "call hsb_gather_export first");
    if ((uint64_t)row_offset + c->rows > c->gather_total_rows)
        return set_err(HSB_EINVAL, "this rank's row block does not fit in the gathered vector");
    CUDA_TRY(cudaSetDevice(c->device));
    { int rc = quiesce(c); if (rc) return rc; }
    hsb::GatherTargets t;
    std::memset(&t, 0, sizeof t);
    for (int g = 0; g < world; g++) {
        GatherBlob b;
        std::memcpy(&b, (const char *)blobs + (size_t)g * HSB_PEER_BLOB_BYTES, sizeof b);
        if (b.total_rows != c->gather_total_rows) return set_err(HSB_EINVAL, "rank " + std::to_string(g) + " gathers a different row count");
        if (!b.has_buffer) continue;                        // not a target
        uint32_t *py = nullptr, *pf = nullptr;
        if (g == rank) { py = c->d_gather_y; pf = c->d_gather_flags; }
        else if (b.pid == (uint64_t)getpid()) {
            cudaError_t e = enable_peer(c->device, b.device);
            if (e != cudaSuccess) { cudaGetLastError(); return set_err(HSB_ECUDA, std::string("no peer access to a target rank: ") + cudaGetErrorString(e)); }
            py = (uint32_t *)(uintptr_t)b.raw_y; pf = (uint32_t *)(uintptr_t)b.raw_flags;
        } else {
            void *a = nullptr, *f = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&a, b.y, cudaIpcMemLazyEnablePeerAccess);
            if (e == cudaSuccess) { c->gather_opened[c->gather_n_opened++] = a; e = cudaIpcOpenMemHandle(&f, b.flags, cudaIpcMemLazyEnablePeerAccess); }
            if (e == cudaSuccess) c->gather_opened[c->gather_n_opened++] = f;
            if (e != cudaSuccess) { cudaGetLastError(); return set_err(HSB_ECUDA, std::string("cannot map the gathered vector of a target rank: ") + cudaGetErrorString(e)); }
            py = (uint32_t *)a; pf = (uint32_t *)f;
        }
        t.y[t.n] = py + row_offset; t.flag[t.n] = pf + rank; t.n++;
    }
    t.ticket = c->d_gather_flags + 32;
    if (!c->d_gather) CUDA_TRY(cudaMalloc(&c->d_gather, sizeof t));
    CUDA_TRY(cudaMemcpy(c->d_gather, &t, sizeof t, cudaMemcpyHostToDevice));
    c->gather_world = world; c->gather_rank = rank; c->gather_seq = 0;
    return HSB_OK;
}

int hsb_gather_wait(hsb_ctx *c) {
    if (!c) return set_err(HSB_EINVAL, "null context");
    if (!c->d_gather || !c->gather_is_target) return set_err(HSB_ESTATE, "this rank is not a gather target");
    CUDA_TRY(cudaSetDevice(c->device));
    { int rc = finish(c); if (rc) return rc; }              // this rank's own block of the last SpMV
    CUDA_TRY(hsb::launch_wait_flags(c->d_gather_flags, (uint32_t)c->gather_world, c->gather_seq, c->d_done ? c->d_done + 8 : c->d_flags + kFlagError, c->stream));
    c->launches++;
    return HSB_OK;
}

void *hsb_device_y_gathered(hsb_ctx *c) { return c && c->gather_is_target ? c->d_gather_y : nullptr; }

int hsb_download_gathered(hsb_ctx *c, void *y_packed, uint32_t total_rows) {
    if (!c || !y_packed) return set_err(HSB_EINVAL, "null argument");
    if (total_rows > c->gather_total_rows) return set_err(HSB_EINVAL, "more rows than the gathered vector holds");
    int rc = hsb_gather_wait(c);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(y_packed, c->d_gather_y, (size_t)total_rows * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return check_error_flag(c);
}

int hsb_device_numa_node(int device) { return device_numa_node(device); }

size_t hsb_device_l2_bytes(int device) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device) != cudaSuccess) { cudaGetLastError(); return 0; }
    return (size_t)v;
}

int hsb_axpb_to_peers(hsb_ctx *c, uint32_t alpha_word, uint32_t beta_word, uint32_t col_offset) {
    if (!c) return set_err(HSB_EINVAL, "null context");
    if (!c->have_matrix || c->peer_world < 1) return set_err(HSB_ESTATE, "call hsb_peer_connect first");
    if (col_offset >= c->x_words) return set_err(HSB_EINVAL, "col_offset beyond the vector");
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->pending_dl.active) { int rc = finish(c); if (rc) return rc; }
    const int b = (c->x_latest + 1) % kXBuffers;
    const int yb = c->y_cur;
    if (c->y_busy[yb]) {
        MEMOP_TRY(g_wait32((CUstream)c->stream, (CUdeviceptr)(c->d_flags + kFlagYFree + yb), c->y_dl_seq[yb], CU_STREAM_WAIT_VALUE_GEQ));
        c->y_busy[yb] = false;
    }
    void *acc = nullptr;
    if (c->drain_pending && c->drain_begin == 0 && c->drain_end == c->rows) {
        acc = c->d_acc[(c->acc_cur + c->acc_bufs - 1) % c->acc_bufs];
        c->drain_pending = false;
    } else {
        int rc = finish(c);
        if (rc) return rc;
    }
    hsb::PeerTargets t;
    std::memset(&t, 0, sizeof t);
    t.world = c->peer_world;
    for (int g = 0; g < c->peer_world; g++) {
        t.x_next[g] = c->peer_x[g] + (size_t)b * c->x_stride;
        t.flag[g] = c->peer_flags[g] + c->peer_rank;
    }
    // Every rank runs the same sequence, so buffer b is the same buffer everywhere; a rank can be at most one
    // iteration ahead of the slowest one (its next SpMV waits for everybody's slice), and four buffers rotate.
    CUDA_TRY(hsb::launch_axpb_peers(c->arith, acc, c->d_y[yb], t, c->rows, c->x_words, alpha_word, beta_word, col_offset,
                                    c->rows, ++c->peer_seq, c->d_peer + 32, c->stream));
    c->launches++;
    c->x_next_buf = b;
    c->x_next_from_peers = true;
    return HSB_OK;
}

namespace {
// hsb_iterate as ONE cooperative launch (spmv_iterate_kernel): the grid stays resident and the two dependencies of an
// iteration are grid-wide barriers instead of kernel boundaries. HSB_ITERATE_PERSISTENT=0: the launch-per-step form.
int iterate_persistent(hsb_ctx *c, int iters, uint32_t alpha_word, uint32_t beta_word, bool peers, uint32_t col_offset) {
    CUDA_TRY(cudaSetDevice(c->device));
    { int rc = finish(c); if (rc) return rc; }             // y final, every accumulator buffer zero, deferred download issued
    // the vector the first iteration reads has landed, and no upload is writing the buffer the iterations alternate with
    CUDA_TRY(cudaStreamSynchronize(c->s_h2d));
    CUDA_TRY(cudaStreamSynchronize(c->s_h2d_b));
    for (int i = 0; i < 2; i++) CUDA_TRY(cudaStreamSynchronize(c->s_h2d_more[i]));
    // (multi-GPU: the current vector may still be arriving from the other ranks: the first iteration polls their flags)
    const bool wait_first = peers && c->x_wait_buf == kXBuffers;
    c->x_wait_buf = -1; c->x_dirty = false;
    const int yb = c->y_cur;
    if (c->y_busy[yb]) {                                   // a download still reads y: order the rewrite behind it
        if (c->flags_mode)
            MEMOP_TRY(g_wait32((CUstream)c->stream, (CUdeviceptr)(c->d_flags + kFlagYFree + yb), c->y_dl_seq[yb], CU_STREAM_WAIT_VALUE_GEQ));
        else
            CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_ydone, 0));
        c->y_busy[yb] = false;
    }
    const int nb = c->flags_mode ? kXBuffers : 2;
    const DeviceMatrix &m = c->mats[c->next_replica % c->mats.size()];
    c->next_replica++;
    const uint32_t G = (uint32_t)c->sm_count;
    const int grid = (int)c->plan_grid[0];
    const int iters_total = iters;
    while (iters > 0) {
        // iterations per launch: the barrier counter takes 2 * grid per iteration and stays below 2^31 (HSB_ITERATE_CHUNK:
        // a small value for the tests of the hand-over between two launches)
        static const int chunk = [] { const char *v = std::getenv("HSB_ITERATE_CHUNK"); return v && std::atoi(v) > 0 ? std::atoi(v) : (1 << 20); }();
        const int n = std::min(iters, chunk);
        const int b = (c->x_latest + 1) % nb;
        hsb::SpmvParams p;
        std::memset(&p, 0, sizeof p);
        p.vals = m.vals; p.cols = m.cols; p.slice_rows = m.slice_rows;
        p.cta_seg = c->d_cta_seg;
        p.segs = c->d_segs;
        p.seq = (uint32_t)++c->launch_seq;
        p.done_dev = c->d_flags + kFlagDoneDev;
        p.done_seq = c->flags_mode ? c->d_done : nullptr;
        p.error_flag = c->d_done ? c->d_done + 8 : c->d_flags + kFlagError;
        p.y = c->d_y[yb];
        p.acc = c->d_acc[c->acc_cur];
        p.narrow = c->meta.narrow ? 1u : 0u;
        p.comb_offset = c->comb_offset;
        p.trash_row = c->rows;
        hsb::IterateParams it;
        it.x0 = c->d_x[c->x_latest]; it.x1 = c->d_x[b];
        it.barrier = c->d_flags + kFlagIterBarrier;
        it.iters = (uint32_t)n; it.alpha = alpha_word; it.beta = beta_word;
        it.rows = c->rows; it.x_limit = c->x_words;
        CUDA_TRY(cudaMemsetAsync(it.barrier, 0, 4, c->stream));
        if (peers) {
            // Every rank runs the same sequence, so buffer b is the same buffer everywhere; a rank can be at most one
            // iteration ahead of the slowest one (it waits for everybody's slice), and four buffers rotate.
            hsb::IteratePeers pr;
            std::memset(&pr, 0, sizeof pr);
            pr.world = (uint32_t)c->peer_world;
            for (int g = 0; g < c->peer_world; g++) {
                pr.x_base[g] = c->peer_x[g];
                pr.flag[g] = c->peer_flags[g] + c->peer_rank;
            }
            pr.arrival = c->d_peer;
            pr.x_stride = c->x_stride;
            pr.buf0 = (uint32_t)c->x_latest; pr.seq0 = c->peer_seq; pr.col_offset = col_offset;
            // (a later chunk of a very long run starts on slices the other ranks stored at the end of the previous one)
            pr.wait_first = (wait_first || iters != iters_total) ? 1u : 0u;
            it.x0 = c->d_x[0]; it.x1 = nullptr;
            CUDA_TRY(hsb::launch_iterate(c->arith, p, it, &pr, grid, c->smem_bytes, c->stream));
            c->peer_seq += (uint32_t)n;
            for (int q = 0; q < kXBuffers; q++) c->x_reader_seq[q] = c->launch_seq;
            c->x_latest = (c->x_latest + n) % kXBuffers;
        } else {
            CUDA_TRY(hsb::launch_iterate(c->arith, p, it, nullptr, grid, c->smem_bytes, c->stream));
            // both buffers were read by this launch; its number appears in done_seq through the stream operation below
            c->x_reader_seq[c->x_latest] = c->x_reader_seq[b] = c->launch_seq;
            if (n & 1) c->x_latest = b;
        }
        c->launches++;
        iters -= n;
    }
    (void)G;
    if (c->flags_mode) { int rc = publish_done(c); if (rc) return rc; }
    else CUDA_TRY(cudaEventRecord(c->ev_xfree[c->x_latest ^ 1], c->stream));
    c->x_next_buf = -1;
    c->x_next_from_peers = false;
    if (peers) {
        // the last slices of the other ranks may still be on their way: the next SpMV polls the arrival flags
        c->x_wait_buf = kXBuffers; c->x_wait_val = c->peer_seq; c->x_wait_launch = 0;
    } else {
        // the next SpMV is a programmatic launch that stages x before its griddepcontrol.wait: it has to wait for this grid first
        c->x_after_grid = true;
    }
    return HSB_OK;
}
}  // namespace

int hsb_iterate(hsb_ctx *c, int iters, uint32_t alpha_word, uint32_t beta_word) {
    if (!c || iters < 0) return set_err(HSB_EINVAL, "bad argument");
    if (!c->have_matrix) return set_err(HSB_ESTATE, "upload a matrix first");
    if (c->rows > c->x_words) return set_err(HSB_EINVAL, "hsb_iterate needs rows <= columns (x <- f(A x))");
    // (a gather of y rides on the drains of the launch-per-step form; a pending update of the next vector belongs to it too)
    if (c->iterate_persistent && iters > 0 && !c->d_gather && c->x_next_buf < 0) return iterate_persistent(c, iters, alpha_word, beta_word, false, 0);
    for (int k = 0; k < iters; k++) {
        int rc = hsb_spmv(c);
        if (rc == HSB_OK) rc = hsb_axpb_to_vector(c, alpha_word, beta_word, 0);
        if (rc == HSB_OK) rc = hsb_vector_commit(c);
        if (rc) return rc;
    }
    return HSB_OK;
}

int hsb_iterate_peers(hsb_ctx *c, int iters, uint32_t alpha_word, uint32_t beta_word, uint32_t col_offset) {
    if (!c || iters < 0) return set_err(HSB_EINVAL, "bad argument");
    if (!c->have_matrix || c->peer_world < 1) return set_err(HSB_ESTATE, "call hsb_peer_connect first");
    if (col_offset >= c->x_words) return set_err(HSB_EINVAL, "col_offset beyond the vector");
    if (c->iterate_persistent && iters > 0 && !c->d_gather && c->x_next_buf < 0)
        return iterate_persistent(c, iters, alpha_word, beta_word, true, col_offset);
    for (int k = 0; k < iters; k++) {
        int rc = hsb_spmv(c);
        if (rc == HSB_OK) rc = hsb_axpb_to_peers(c, alpha_word, beta_word, col_offset);
        if (rc == HSB_OK) rc = hsb_vector_commit(c);
        if (rc) return rc;
    }
    return HSB_OK;
}

int hsb_set_option(hsb_ctx *c, const char *name, int value) {
    if (!c || !name) return set_err(HSB_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    { int rc = quiesce(c); if (rc) return rc; }
    const std::string n(name);
    if (n == "flags") {                 // 0: events between launches (the pre-pipeline behaviour), 1: flag pipeline
        if (value && !c->d_done) return set_err(HSB_ESTATE, "stream memory operations are not available");
        c->flags_mode = value != 0;
        if (!c->flags_mode && c->x_latest > 1 && c->have_matrix) {      // event mode rotates buffers 0 and 1 only
            CUDA_TRY(cudaMemcpy(c->d_x[0], c->d_x[c->x_latest], (size_t)c->x_words * 4, cudaMemcpyDeviceToDevice));
            c->x_latest = 0;
        }
        c->x_wait_buf = -1; c->x_dirty = false;
        for (int b = 0; b < kXBuffers; b++) c->x_reader_seq[b] = 0;
        for (int b = 0; b < 2; b++) CUDA_TRY(cudaEventRecord(c->ev_xfree[b], c->stream));
    } else if (n == "iterate_persistent") {
        c->iterate_persistent = value != 0;
    } else if (n == "xwait_once") {
        c->xwait_once = value != 0;
    } else if (n == "host_drain") {
        c->host_drain = value != 0;
    } else if (n == "acquire") {
        c->acquire = value != 0;
    } else if (n == "xflag_copy") {
        c->xflag_copy = value != 0;

    } else {
        return set_err(HSB_EINVAL, "unknown option: " + n);
    }
    return HSB_OK;
}

int hsb_debug_timeline(hsb_ctx *c, unsigned long long *out, size_t capacity) {
    if (!c) return set_err(HSB_EINVAL, "null context");
    const size_t n = 256 * 8;
    CUDA_TRY(cudaSetDevice(c->device));
    { int rc = quiesce(c); if (rc) return rc; }
    if (!out) {
        if (capacity) {                            // (re-)arm: covers the next 256 launches
            if (!c->d_timeline) CUDA_TRY(cudaMalloc(&c->d_timeline, n * 8));
            std::vector<unsigned long long> init(n, 0);
            for (size_t i = 0; i < n; i += 8) init[i] = ~0ull;
            CUDA_TRY(cudaMemcpy(c->d_timeline, init.data(), n * 8, cudaMemcpyHostToDevice));
        } else if (!capacity && c->d_timeline) {
            cudaFree(c->d_timeline);
            c->d_timeline = nullptr;
        }
        return (int)n;
    }
    if (!c->d_timeline || capacity < n) return set_err(HSB_ESTATE, "timeline not armed or buffer too small");
    CUDA_TRY(cudaMemcpy(out, c->d_timeline, n * 8, cudaMemcpyDeviceToHost));
    return (int)(c->launch_seq & 0x7FFFFFFF);
}

int hsb_debug_trace(hsb_ctx *c, unsigned long long *out, size_t capacity) {
    if (!c) return set_err(HSB_EINVAL, "null context");
    const size_t n = (size_t)c->sm_count * (hsb::kWarps + 2);
    CUDA_TRY(cudaSetDevice(c->device));
    if (!out) {                                   // arm (or disarm with capacity == 0)
        if (capacity && !c->d_trace) {
            CUDA_TRY(cudaMalloc(&c->d_trace, n * 8));
            CUDA_TRY(cudaMemset(c->d_trace, 0, n * 8));
        } else if (!capacity && c->d_trace) {
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            cudaFree(c->d_trace);
            c->d_trace = nullptr;
        }
        return (int)n;
    }
    if (!c->d_trace || capacity < n) return set_err(HSB_ESTATE, "trace not armed or buffer too small");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(out, c->d_trace, n * 8, cudaMemcpyDeviceToHost));
    return (int)n;
}

hsb_format *hsb_format_from_context(hsb_ctx *c) {
    if (!c || !c->have_matrix) { set_err(HSB_ESTATE, "no resident matrix"); return nullptr; }
    if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) return nullptr;
    hsb_format *f = new hsb_format;
    f->M = c->meta;
    f->M.vals.resize(c->n_elems);
    f->M.cols16.resize(c->n_elems);
    f->M.slice_rows.resize(c->meta.narrow ? 0 : c->n_slices * hsb::kLanes);
    const DeviceMatrix &d = c->mats[0];
    bool ok = cudaMemcpy(f->M.vals.data(), d.vals, c->n_elems * 4, cudaMemcpyDeviceToHost) == cudaSuccess &&
              cudaMemcpy(f->M.cols16.data(), d.cols, c->n_elems * 2, cudaMemcpyDeviceToHost) == cudaSuccess &&
              cudaMemcpy(f->M.slice_rows.data(), d.slice_rows, f->M.slice_rows.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
    if (!ok) { delete f; set_err(HSB_ECUDA, "download of the device format failed"); return nullptr; }
    return f;
}

int hsb_debug_profiler(hsb_ctx *c, int on) {
    if (!c) return set_err(HSB_EINVAL, "null context");
    CUDA_TRY(cudaSetDevice(c->device));
    { int rc = quiesce(c); if (rc) return rc; }
    CUDA_TRY(on ? cudaProfilerStart() : cudaProfilerStop());
    return HSB_OK;
}

int hsb_debug_plan(hsb_ctx *c, uint32_t *steps, uint32_t *slices, size_t capacity) {
    if (!c || !c->have_matrix || capacity < c->plan_steps.size()) return set_err(HSB_EINVAL, "bad argument");
    std::memcpy(steps, c->plan_steps.data(), c->plan_steps.size() * 4);
    std::memcpy(slices, c->plan_slices.data(), c->plan_slices.size() * 4);
    return (int)c->plan_steps.size();
}

void *hsb_device_x(hsb_ctx *c) { return c ? c->d_x[c->x_latest] : nullptr; }
void *hsb_device_y(hsb_ctx *c) { return c ? c->d_y[c->y_cur] : nullptr; }
void *hsb_stream(hsb_ctx *c) { return c ? (void *)c->stream : nullptr; }

}  // extern "C"
