// sm_100a SpMV kernels over the tile-stream format (tile_format.h).
//
// One persistent CTA per SM replaces one HiSparse "cluster" pipeline
// (spmv/libfpga/spmv_cluster.h:196-373):
//
//   reference unit (file:line)                          here
//   ------------------------------------------------    --------------------------------------------
//   spmv_vector_loader + axis_duplicate + vecbuf_writer x tile -> shared memory with cp.async.bulk
//     (spmv_vector_loader.cpp:7-79, stream_utils.h:8,     (TMA bulk copy, mbarrier complete_tx);
//      vecbuf_access_unit.h:92-136)                       one 128 KB tile = LOGICAL_VB_SIZE words
//   CPSR_matrix_loader (spmv_cluster.h:34-107)          128-bit ld.global.nc streaming loads, each
//                                                         warp-wide load one contiguous 512 B
//   shuffler<EDGE> + vecbuf_reader                      per-lane shared-memory gather xs[col]
//     (shuffle.h:380-468, vecbuf_access_unit.h:18-84)     (bank = col % 32, conflicts replayed by HW)
//   shuffler<UPDATE> + pe (shuffle.h, pe.h:22-90)       register accumulation per lane + warp
//                                                         segmented scan keyed by end-of-segment flags
//   pe dump + result_packer + axis_merge + result_drain one red.global.add per (row, tile) segment
//     (pe.h:95-116, spmv_cluster.h:133-193,                into the row accumulator, then a clamp pass
//      stream_utils.h:36-75, spmv_result_drain.cpp)
//
// Arithmetic:
//   fixed  : VAL_T = ap_ufixed<32,8,AP_RND,AP_SAT> (spmv/libfpga/common.h:38). product =
//            min((a*b + 2^23) >> 24, 2^32-1) (pe.h:64), accumulation saturating (pe.h:72). All terms
//            are >= 0 and the clamp is at a constant, so y = min(sum of products, 2^32-1) in ANY
//            order: we add the clamped products exactly in 64 bits and clamp once at the end.
//   float  : fp32 multiply then fp32 add, not fused (pe-pob.h:64-66, pe-stall.h:53,138).
#ifndef HISPARSE_B200_SPMV_KERNELS_CUH_
#define HISPARSE_B200_SPMV_KERNELS_CUH_

#include <cuda_runtime.h>
#include <cstdint>
#include "tile_format.h"

namespace hsb {

constexpr int kThreads = 1024;                       // 32 warps: one CTA per SM
constexpr int kWarps = kThreads / 32;
constexpr uint32_t kXTileBytes = kMaxTileCols * 4;   // 128 KB
constexpr uint32_t kBulkPiece = 16384;               // bytes per cp.async.bulk

struct SpmvParams {
    const uint32_t *vals;
    const uint16_t *cidx;
    const ChunkDesc *chunks;
    const TileDesc *tiles;
    const uint32_t *seg_row;
    const uint32_t *x;            // packed dense vector, raw 32-bit words
    void *acc;                    // fixed: uint64 per row; float: the fp32 result itself
    uint32_t chunk_begin, chunk_end;
};

enum { kArithFixed = 0, kArithFloat = 1 };

void launch_spmv_tiles(int arith, const SpmvParams &p, int grid, cudaStream_t stream);
void launch_finalize_fixed(const unsigned long long *acc, uint32_t *y, uint32_t row_begin, uint32_t row_end,
                           cudaStream_t stream);
cudaError_t configure_kernels();

}  // namespace hsb
#endif
