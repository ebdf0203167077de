// sm_100a SpMV kernel over the tile-stream format (tile_format.h).
//
// One persistent CTA per SM replaces one HiSparse "cluster" pipeline
// (spmv/libfpga/spmv_cluster.h:196-373); a whole SpMV (all row partitions or one) is ONE launch:
//
//   reference unit (file:line)                          here
//   ------------------------------------------------    --------------------------------------------
//   spmv_vector_loader + axis_duplicate + vecbuf_writer x tile -> shared memory with cp.async.bulk
//     (spmv_vector_loader.cpp:7-79, stream_utils.h:8,     (TMA bulk copy, mbarrier complete_tx);
//      vecbuf_access_unit.h:92-136)                       one <= 224 KB tile (the reference: 128 KB = LOGICAL_VB_SIZE words)
//   CPSR_matrix_loader (spmv_cluster.h:34-107)          every warp streams ONE contiguous run of slice steps
//                                                         with 128-bit / 64-bit ld.global.nc loads (each
//                                                         warp-wide load a contiguous 512 / 256 B), a
//                                                         register ring keeps kPrefetch steps in flight
//                                                         across slice boundaries
//   shuffler<EDGE> + vecbuf_reader                      per-lane shared-memory gather xs[col]
//     (shuffle.h:380-468, vecbuf_access_unit.h:18-84)     (bank = col % 32, conflicts replayed by HW)
//   shuffler<UPDATE> + pe (shuffle.h, pe.h:22-90)       every lane owns a whole lane stream (a row
//                                                         segment) and accumulates it in registers: no
//                                                         routing by row, no read-after-write hazard
//   pe dump + result_packer + axis_merge + result_drain one red.global.add per lane stream (or per warp when
//     (pe.h:95-116, spmv_cluster.h:133-193,                a slice holds one row) into the row accumulator;
//      stream_utils.h:36-75, spmv_result_drain.cpp)        four accumulator buffers rotate and launch n drains
//                                                         the buffer of launch n-1 (clamp, store y, re-zero) at
//                                                         its very end, or a drain kernel does at sync time
//
// Consecutive launches overlap freely (griddepcontrol.launch_dependents first, griddepcontrol.wait last):
// see the comments in spmv_tiles_kernel.
//
// Arithmetic:
//   fixed  : VAL_T = ap_ufixed<32,8,AP_RND,AP_SAT> (spmv/libfpga/common.h:38). product =
//            min((a*b + 2^23) >> 24, 2^32-1) (pe.h:64), accumulation saturating (pe.h:72). All terms
//            are >= 0 and the clamp is at a constant, so y = min(sum of products, 2^32-1) in ANY
//            order: products are summed exactly (64-bit) and clamped once in the drain.
//   float  : fp32 multiply then fp32 add, not fused (pe-pob.h:64-66, pe-stall.h:53,138).
#ifndef HISPARSE_B200_SPMV_KERNELS_CUH_
#define HISPARSE_B200_SPMV_KERNELS_CUH_

#include <cuda_runtime.h>
#include <cstdint>
#include "tile_format.h"

namespace hsb {

constexpr int kWarps = kWarpsPerCta;                 // one CTA per SM
constexpr int kThreads = kWarps * 32;
constexpr uint32_t kXTileBytes = kMaxTileCols * 4;   // 224 KB
constexpr uint32_t kXTileOffset = kColBias * 4;      // xs[0..7] = 0: what padding slots (column id 0) gather
constexpr uint32_t kSmemBytes = 232448 - 256;        // 227 KB (the most a CTA can opt in to) less the static variables; a launch asks for what its tiles need
                                                     // (x tile + the zero words + the combining tables when they fit)
#ifndef HSB_BULK_PIECE
#define HSB_BULK_PIECE 16384
#endif
constexpr uint32_t kBulkPiece = HSB_BULK_PIECE;      // bytes per cp.async.bulk
#ifndef HSB_PREFETCH
#define HSB_PREFETCH 4
#endif
constexpr int kPrefetch = HSB_PREFETCH;              // slice steps in flight per warp
#ifndef HSB_ROW_AHEAD
#define HSB_ROW_AHEAD 1
#endif
constexpr int kRowAhead = HSB_ROW_AHEAD;             // slices whose row ids are loaded ahead of their use
#ifndef HSB_NARROW_RING
#define HSB_NARROW_RING 12
#endif
constexpr int kNarrowRing = HSB_NARROW_RING;         // narrow layout: units (192 B per warp) in flight per warp

struct SpmvParams {
    const uint32_t *vals;
    const uint16_t *cols;
    const uint32_t *slice_rows;
    const uint32_t *cta_seg;      // gridDim.x + 1 : CTA b runs segs[cta_seg[b] .. cta_seg[b+1])
    const Segment *segs;          // (tile, [t_lo, t_hi) tile-relative steps, tile geometry): one x staging each
    const uint32_t *x;            // packed dense vector, raw 32-bit words
    void *acc;                    // row accumulators of THIS launch (uint64 fixed / fp32 float), rows + 1 entries,
                                  // all zero once guard_val has been seen
    void *drain_acc;              // accumulators of the PREVIOUS launch still to be drained into y, or null
    uint32_t *y;                  // packed result, raw 32-bit words
    uint32_t drain_begin, drain_end;  // rows of drain_acc to drain
    uint32_t trash_row;           // accumulator index of unused lanes (== rows)
    unsigned long long *trace;    // optional [gridDim.x][kWarps + 2] SM-clock stamps (profiling aid), or null
    // Flag pipeline (host <-> device overlap without stream events between launches, which would undo the
    // programmatic-dependent-launch overlap of consecutive SpMVs). All flags are 32-bit sequence
    // numbers in device memory, compared cyclically; null pointers switch the feature off.
    const uint32_t *wait_x_flag;  // the copy stream writes wait_x_val here once this launch's x has landed; in a multi-GPU
    uint32_t wait_x_val;          // iteration: wait_x_count consecutive flags, one per rank whose slice of x must have arrived
    uint32_t wait_x_count;
    const uint32_t *wait_y_flag;  // ... and wait_y_val here once the last download of `y` has been read out
    uint32_t wait_y_val;
    uint32_t *done_dev;           // launch `seq` publishes seq - 1 here (device memory) once its predecessor has completed ...
    uint32_t *done_seq;           // ... and here (mapped host memory, read by the host and the copy streams), or null
    uint32_t seq;
    uint32_t sync_start;          // 1: wait for the predecessor before the first row update (two accumulator buffers in
                                  // rotation instead of four: long launches with accumulators too big for L2)
    const uint32_t *guard_flag;   // == done_dev when this launch must see guard_val there before its first row update
    uint32_t guard_val;           // (launch seq - 3 has re-zeroed the accumulator buffer), else null
    uint32_t *error_flag;         // set to 1 if a flag wait timed out (no hang: the launch then makes NO row update; the word
                                  // is in mapped host memory when the flag pipeline is on, else in device memory)
    unsigned long long *timeline; // optional [256][8] %globaltimer stamps indexed by seq % 256 (profiling aid), or null:
                                  // 0 first CTA start, 1 CTA0 x flag seen, 2 CTA0 predecessor complete, 3 CTA0 drain done,
                                  // 4 CTA0 x tile staged, 5 last CTA end, 6 CTA0 matrix work done
    uint32_t *y_host;             // device alias of a mapped page-locked host buffer the drain ALSO writes (rows
    uint32_t y_host_rows;         // < y_host_rows), or null: a download without the copy engine
    // Row-block shards on several GPUs (hsb_gather_connect): the drain ALSO stores every result word into the
    // gathered-y buffer of the target ranks at this rank's row offset (peer pointers over NVLink) and, when the
    // whole grid is through, raises this rank's arrival flag there -- the gather of y is the drain's epilogue.
    const struct GatherTargets *gather;   // device-resident table, or null
    uint32_t gather_seq;          // value the arrival flags receive: gathered drains so far, this one included
    uint32_t acquire;             // 1: flag waits are acquire loads (+ fence.proxy.async before the x TMA)
    uint32_t x_after_grid;        // 1: x was written by the kernel in front of this launch on the stream (axpb step): thread 0
                                  // waits for that grid (griddepcontrol.wait) before it stages the first x tile
    uint32_t comb_offset;         // byte offset (from the start of dynamic shared memory) of the two row-update combining
                                  // tables, kCombineSlots x 32 accumulators each (Segment::comb_n), or 0: no room, no combining
    uint32_t narrow;              // 1: the matrix is in the narrow layout (tile_format.h): units of 32 elements, a row
                                  // unit in front of every slice, shares cut at slice boundaries; slice_rows unused
};

constexpr int kMaxPeers = 16;
struct GatherTargets {
    uint32_t *y[kMaxPeers];       // target g: its gathered-y buffer + this rank's first row
    uint32_t *flag[kMaxPeers];    // target g: &arrival[this rank] in its flag array
    uint32_t *ticket;             // this rank's completion counter (CTAs of one drain)
    int n;                        // targets (1: gather to a root, world: all-gather)
};

enum { kArithFixed = 0, kArithFloat = 1 };

cudaError_t launch_spmv(int arith, const SpmvParams &p, int grid, uint32_t smem_bytes, cudaStream_t stream);
}  // namespace hsb
#include "spmv_iterate.h"     // hsb_iterate as one cooperative launch: IterateParams, IteratePeers, launch_iterate
namespace hsb {
// drain only: y[r] = clamp(acc[r]), acc[r] = 0 for r in [row_begin, row_end) and the trash slot; y_host (device alias of a
// mapped page-locked host buffer, or null): rows < y_host_rows are ALSO written there (a download without the copy engine)
cudaError_t launch_drain(int arith, void *acc, uint32_t *y, uint32_t row_begin, uint32_t row_end,
                         uint32_t trash_row, const GatherTargets *gather, uint32_t gather_seq, uint32_t *y_host,
                         uint32_t y_host_rows, cudaStream_t stream);
// stream-ordered wait (one thread, acquire) until `count` consecutive arrival flags have reached `val`; a
// timed-out wait raises *error_flag
cudaError_t launch_wait_flags(const uint32_t *flags, uint32_t count, uint32_t val, uint32_t *error_flag, cudaStream_t stream);
// y = final result of the last launch (drained from acc when acc != null); x_next[col_offset + r] = alpha (*) y[r] (+) beta
cudaError_t launch_axpb(int arith, void *acc, uint32_t *y, uint32_t *x_next, uint32_t rows, uint32_t x_limit,
                        uint32_t alpha, uint32_t beta, uint32_t col_offset, uint32_t trash_row, cudaStream_t stream);
// Multi-GPU form of launch_axpb: the slice alpha (*) y (+) beta is stored into the next x buffer of EVERY rank
// (peer pointers over NVLink, this rank included) at col_offset, and when the whole grid has finished, `seq`
// is written into this rank's slot of every rank's arrival-flag array -- compute and all-gather in one kernel.
struct PeerTargets {
    uint32_t *x_next[kMaxPeers];   // next x buffer of rank g
    uint32_t *flag[kMaxPeers];     // &arrival[my rank] in rank g's flag array
    int world;
};
cudaError_t launch_axpb_peers(int arith, void *acc, uint32_t *y, const PeerTargets &t, uint32_t rows, uint32_t x_limit,
                              uint32_t alpha, uint32_t beta, uint32_t col_offset, uint32_t trash_row, uint32_t seq,
                              uint32_t *ticket, cudaStream_t stream);
// sm_count: SMs of the device (grid sizing of the element-wise kernels)
cudaError_t configure_kernels(int sm_count);

}  // namespace hsb
#endif
