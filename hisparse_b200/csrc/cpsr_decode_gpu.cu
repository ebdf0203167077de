// GPU-side ingestion of the reference's per-HBM-channel CPSR packet images:
// the 16 buffers that sw/host.cpp:263-299 migrates to the FPGA are copied to HBM unchanged and decoded
// THERE into a device-resident COO list, which gpu_format.cu turns into tile streams -- no host-side
// re-formatting of buffers produced by the reference's own host code.
//
// One WARP walks one piece of a lane stream -- (physical channel, row partition, interleave slot, lane) x column
// partition -- 32 entries at a time, exactly as CPSR_matrix_loader does (spmv/libfpga/spmv_cluster.h:34-107;
// float: spmv-fp/libfpga/spmv_cluster.h:39-129): an entry with index 0xFFFFFFFF advances the lane's row by
// 8 * IF * marker, every other entry belongs to the current row, and the row counter restarts with every column
// partition. The row of an entry is therefore a prefix sum over the markers in front of it: a warp-wide scan per
// 32 entries plus a carry. The eight lanes of a packet sit in the eight warps of one CTA, so the 64-byte packets
// are fetched from HBM once. Two passes: count the entries every piece keeps, scan, write (row, column, value)
// triples -- a COO list in image order, which the GPU formatter sorts into tile streams directly (no CSR is built).
// Result rows follow the drain order y[(r/8)*128 + pc*8 + r%8] (spmv_result_drain.cpp:36-113). The one-thread-per-
// lane-stream version this replaces had 128 threads of parallelism for a fixed-point image with one row partition.
#include <cub/cub.cuh>

#include <algorithm>
#include <string>

#include "cpsr_decode.h"

namespace hsb {
namespace {

struct DecodeArgs {
    const uint32_t *img[16];     // device copies of the channel images (16 words per packet)
    unsigned long long n_pkts[16];
    uint32_t IF, ob_size, vb_size, marker_q824;
    uint32_t n_row_parts, n_col_parts, num_rows, num_cols;
};

// piece p = ((j * n_col_parts + i) * 16 + pc) * IF + f, lane k = warp of the CTA.
// PASS 0: kept[p * 8 + k] = entries of the piece that survive.  PASS 1: write them at offset[p * 8 + k] + rank.
template <int PASS>
__global__ void __launch_bounds__(256) k_decode(const DecodeArgs a, uint32_t *__restrict__ kept_or_offset,
                                                uint32_t *__restrict__ coo_rows, uint32_t *__restrict__ coo_cols,
                                                uint32_t *__restrict__ coo_vals, int *__restrict__ bad) {
    const uint32_t piece = blockIdx.x, k = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t f = piece % a.IF, pc = (piece / a.IF) % 16u, ij = piece / (a.IF * 16u);
    const uint32_t j = ij / a.n_col_parts, i = ij % a.n_col_parts;
    const uint32_t *img = a.img[pc];
    const unsigned long long limit = a.n_pkts[pc];
    const unsigned long long base = (unsigned long long)a.n_row_parts * a.n_col_parts * (1 + a.IF);
    if (limit < base) { if (threadIdx.x == 0) *bad = 1; return; }
    const uint32_t rows_here = min(a.ob_size, a.num_rows - j * a.ob_size);
    const uint32_t part_len = rows_here / 16u, row0 = j * a.ob_size;
    const unsigned long long start = img[(unsigned long long)ij * (1 + a.IF) * 16];
    const uint32_t len = img[((unsigned long long)ij * (1 + a.IF) + 1 + f) * 16 + k];
    uint32_t row_carry = f * 8u + k;                          // row of the next entry if no marker intervenes
    uint32_t out = PASS == 1 ? kept_or_offset[piece * 8u + k] : 0u;
    for (uint32_t n0 = 0; n0 < len; n0 += 32u) {
        const uint32_t n = n0 + lane;
        uint32_t idx = 0, v = 0, adv = 0;
        bool valid = false;
        if (n < len) {
            const unsigned long long pi = base + start + (unsigned long long)n * a.IF + f;
            if (pi >= limit) {
                *bad = 2;                                      // (no early exit: the warp-wide scans below need every lane)
            } else {
                idx = img[pi * 16 + k];
                v = img[pi * 16 + 8 + k];
                if (idx == 0xFFFFFFFFu) adv = 8u * a.IF * (a.marker_q824 ? (v >> 24) : v);
                else valid = true;
            }
        }
        // inclusive scan of the row advances: an entry's row counts the markers in FRONT of it
        uint32_t incl = adv;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += t;
        }
        const uint32_t row_local = row_carry + incl - adv;
        row_carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        valid = valid && row_local < part_len;                 // rows beyond the partition are never dumped by the PE (pe.h:95-116)
        unsigned long long col = 0;
        if (valid) {
            col = (unsigned long long)i * a.vb_size + idx;
            if (idx >= a.vb_size) { *bad = 3; valid = false; }
            else if (col >= a.num_cols) { *bad = 4; valid = false; }
        }
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, valid);
        if (PASS == 1 && valid) {
            const uint32_t at = out + __popc(mask & ((1u << lane) - 1u));
            coo_rows[at] = row0 + (row_local / 8u) * 128u + pc * 8u + (row_local % 8u);
            coo_cols[at] = (uint32_t)col;
            coo_vals[at] = v;
        }
        out += __popc(mask);
    }
    if (PASS == 0 && lane == 0) kept_or_offset[piece * 8u + k] = out;
}

}  // namespace

// Decode host channel images on the device into a device COO list in image order (cudaMalloc'ed; the caller frees).
cudaError_t cpsr_decode_gpu(const ImplConfig &cfg, const uint32_t *const images[16], const size_t n_packets[16],
                            uint32_t num_row_partitions, uint32_t num_col_partitions, uint32_t num_rows,
                            uint32_t num_cols, cudaStream_t stream, uint32_t **d_rows, uint32_t **d_cols,
                            uint32_t **d_vals, uint64_t *nnz, std::string *err) {
#define CD_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); return e__; } } while (0)
    DecodeArgs a;
    uint32_t *d_img = nullptr, *kept = nullptr, *offset = nullptr, *rows = nullptr, *cols = nullptr, *vals = nullptr;
    int *bad = nullptr;
    void *tmp = nullptr;
    auto cleanup = [&]() { cudaFree(d_img); cudaFree(kept); cudaFree(offset); cudaFree(bad); cudaFree(tmp); };
    auto fail = [&](const char *m) { if (err) *err = m; cleanup(); cudaFree(rows); cudaFree(cols); cudaFree(vals); return cudaErrorInvalidValue; };
    *d_rows = *d_cols = *d_vals = nullptr;
    *nnz = 0;
    size_t total = 0;
    for (int c = 0; c < 16; c++) total += n_packets[c];
    CD_TRY(cudaMalloc(&d_img, std::max<size_t>(total, 1) * 64));
    size_t off = 0;
    for (int c = 0; c < 16; c++) {
        a.img[c] = d_img + off * 16;
        a.n_pkts[c] = n_packets[c];
        if (n_packets[c]) CD_TRY(cudaMemcpyAsync(d_img + off * 16, images[c], n_packets[c] * 64, cudaMemcpyHostToDevice, stream));
        off += n_packets[c];
    }
    a.IF = cfg.interleave; a.ob_size = cfg.ob_size; a.vb_size = cfg.vb_size; a.marker_q824 = cfg.marker_is_q824 ? 1u : 0u;
    a.n_row_parts = num_row_partitions; a.n_col_parts = num_col_partitions; a.num_rows = num_rows; a.num_cols = num_cols;
    const uint64_t pieces64 = (uint64_t)num_row_partitions * num_col_partitions * 16u * cfg.interleave;
    if (pieces64 >= (1ull << 28)) return fail("too many partitions");
    const uint32_t pieces = (uint32_t)pieces64, n_streams = pieces * 8u;
    CD_TRY(cudaMalloc(&kept, ((size_t)n_streams + 1) * 4));
    CD_TRY(cudaMalloc(&offset, ((size_t)n_streams + 1) * 4));
    CD_TRY(cudaMalloc(&bad, sizeof(int)));
    CD_TRY(cudaMemsetAsync(kept, 0, ((size_t)n_streams + 1) * 4, stream));
    CD_TRY(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    if (pieces) k_decode<0><<<pieces, 256, 0, stream>>>(a, kept, nullptr, nullptr, nullptr, bad);
    size_t need = 0;
    CD_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, kept, offset, (int)n_streams + 1, stream));
    CD_TRY(cudaMalloc(&tmp, std::max<size_t>(need, 1)));
    CD_TRY(cub::DeviceScan::ExclusiveSum(tmp, need, kept, offset, (int)n_streams + 1, stream));
    uint32_t h_nnz = 0;
    int h_bad = 0;
    CD_TRY(cudaMemcpyAsync(&h_nnz, offset + n_streams, 4, cudaMemcpyDeviceToHost, stream));
    CD_TRY(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CD_TRY(cudaStreamSynchronize(stream));
    if (h_bad) return fail(h_bad == 1 ? "channel image shorter than its header" : h_bad == 2 ? "channel image truncated"
                           : h_bad == 3 ? "column index beyond the vector buffer" : "column index out of range");
    CD_TRY(cudaMalloc(&rows, std::max<size_t>(h_nnz, 1) * 4));
    CD_TRY(cudaMalloc(&cols, std::max<size_t>(h_nnz, 1) * 4));
    CD_TRY(cudaMalloc(&vals, std::max<size_t>(h_nnz, 1) * 4));
    if (pieces) k_decode<1><<<pieces, 256, 0, stream>>>(a, offset, rows, cols, vals, bad);
    CD_TRY(cudaGetLastError());
    CD_TRY(cudaStreamSynchronize(stream));
    cleanup();
    *d_rows = rows; *d_cols = cols; *d_vals = vals; *nnz = h_nnz;
    return cudaSuccess;
#undef CD_TRY
}

}  // namespace hsb
