// GPU-side ingestion of the reference's per-HBM-channel CPSR packet images:
// the 16 buffers that sw/host.cpp:263-299 migrates to the FPGA are copied to HBM unchanged and decoded
// THERE into a device-resident CSR, which gpu_format.cu turns into tile streams -- no host-side
// re-formatting of buffers produced by the reference's own host code.
//
// One thread walks one lane stream, i.e. one (physical channel, row partition, interleave slot, lane)
// across all column partitions, exactly as CPSR_matrix_loader does (spmv/libfpga/spmv_cluster.h:34-107;
// float: spmv-fp/libfpga/spmv_cluster.h:39-129): an entry with index 0xFFFFFFFF advances the lane's
// row by 8 * IF * marker, every other entry belongs to the current row. A matrix row is owned by
// exactly one lane stream, so counting and placing need no atomics and the entries of a row keep
// the reference's order (column partition ascending, stream order inside). Result rows follow the
// drain order y[(r/8)*128 + pc*8 + r%8] (spmv_result_drain.cpp:36-113).
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

// PASS 0: counts[row] = entries of the row.  PASS 1: place them at cursor[row]++ (cursor starts as indptr).
template <int PASS>
__global__ void k_decode(const DecodeArgs a, uint32_t *__restrict__ counts_or_cursor, uint32_t *__restrict__ indices,
                         uint32_t *__restrict__ vals, int *__restrict__ bad) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t per_part = 16u * a.IF * 8u;
    if (tid >= a.n_row_parts * per_part) return;
    const uint32_t j = tid / per_part, rem = tid % per_part;
    const uint32_t pc = rem / (a.IF * 8u), f = (rem / 8u) % a.IF, k = rem % 8u;
    const uint32_t *img = a.img[pc];
    const unsigned long long limit = a.n_pkts[pc];
    const uint32_t np = a.n_row_parts * a.n_col_parts;
    const unsigned long long base = (unsigned long long)np * (1 + a.IF);
    const uint32_t rows_here = min(a.ob_size, a.num_rows - j * a.ob_size);
    const uint32_t part_len = rows_here / 16u;
    const uint32_t row0 = j * a.ob_size;
    if (limit < base) { *bad = 1; return; }
    for (uint32_t i = 0; i < a.n_col_parts; i++) {
        const unsigned long long ij = (unsigned long long)j * a.n_col_parts + i;
        const unsigned long long start = img[ij * (1 + a.IF) * 16];
        const uint32_t len = img[(ij * (1 + a.IF) + 1 + f) * 16 + k];
        uint32_t row_local = f * 8u + k;
        uint32_t run_row = 0xFFFFFFFFu, run = 0;          // entries of run_row seen in this column partition
        for (uint32_t n = 0; n < len; n++) {
            const unsigned long long pi = base + start + (unsigned long long)n * a.IF + f;
            if (pi >= limit) { *bad = 2; return; }
            const uint32_t idx = img[pi * 16 + k], v = img[pi * 16 + 8 + k];
            if (idx == 0xFFFFFFFFu) {
                row_local += 8u * a.IF * (a.marker_q824 ? (v >> 24) : v);
                continue;
            }
            if (row_local >= part_len) continue;             // never dumped by the PE (pe.h:95-116)
            if (idx >= a.vb_size) { *bad = 3; return; }
            const unsigned long long col = (unsigned long long)i * a.vb_size + idx;
            if (col >= a.num_cols) { *bad = 4; return; }
            const uint32_t row = row0 + (row_local / 8u) * 128u + pc * 8u + (row_local % 8u);
            if (row != run_row) {
                if (run_row != 0xFFFFFFFFu) counts_or_cursor[run_row] += run;
                run_row = row;
                run = 0;
            }
            if (PASS == 1) {
                const uint32_t at = counts_or_cursor[row] + run;
                indices[at] = (uint32_t)col;
                vals[at] = v;
            }
            run++;
        }
        if (run_row != 0xFFFFFFFFu) counts_or_cursor[run_row] += run;
    }
}

}  // namespace

// Decode host channel images on the device into a device CSR (cudaMalloc'ed; the caller frees).
cudaError_t cpsr_decode_gpu(const ImplConfig &cfg, const uint32_t *const images[16], const size_t n_packets[16],
                            uint32_t num_row_partitions, uint32_t num_col_partitions, uint32_t num_rows,
                            uint32_t num_cols, cudaStream_t stream, uint32_t **d_indptr, uint32_t **d_indices,
                            uint32_t **d_vals, uint64_t *nnz, std::string *err) {
#define CD_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); return e__; } } while (0)
    DecodeArgs a;
    uint32_t *d_img = nullptr, *counts = nullptr, *indptr = nullptr, *indices = nullptr, *vals = nullptr;
    int *bad = nullptr;
    void *tmp = nullptr;
    auto cleanup = [&]() { cudaFree(d_img); cudaFree(counts); cudaFree(bad); cudaFree(tmp); };
    auto fail = [&](const char *m) { if (err) *err = m; cleanup(); cudaFree(indptr); cudaFree(indices); cudaFree(vals); return cudaErrorInvalidValue; };
    *d_indptr = *d_indices = *d_vals = nullptr;
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
    CD_TRY(cudaMalloc(&counts, ((size_t)num_rows + 1) * 4));
    CD_TRY(cudaMalloc(&indptr, ((size_t)num_rows + 1) * 4));
    CD_TRY(cudaMalloc(&bad, sizeof(int)));
    CD_TRY(cudaMemsetAsync(counts, 0, ((size_t)num_rows + 1) * 4, stream));
    CD_TRY(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    const uint32_t threads = num_row_partitions * 16u * cfg.interleave * 8u;
    const int TB = 128;
    if (threads) k_decode<0><<<(threads + TB - 1) / TB, TB, 0, stream>>>(a, counts, nullptr, nullptr, bad);
    size_t need = 0;
    CD_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, counts, indptr, (int)num_rows + 1, stream));
    CD_TRY(cudaMalloc(&tmp, std::max<size_t>(need, 1)));
    CD_TRY(cub::DeviceScan::ExclusiveSum(tmp, need, counts, indptr, (int)num_rows + 1, stream));
    uint32_t h_nnz = 0;
    int h_bad = 0;
    CD_TRY(cudaMemcpyAsync(&h_nnz, indptr + num_rows, 4, cudaMemcpyDeviceToHost, stream));
    CD_TRY(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CD_TRY(cudaStreamSynchronize(stream));
    if (h_bad) return fail(h_bad == 1 ? "channel image shorter than its header" : h_bad == 2 ? "channel image truncated"
                           : h_bad == 3 ? "column index beyond the vector buffer" : "column index out of range");
    CD_TRY(cudaMalloc(&indices, std::max<size_t>(h_nnz, 1) * 4));
    CD_TRY(cudaMalloc(&vals, std::max<size_t>(h_nnz, 1) * 4));
    // cursor array for the placing pass = a copy of indptr (counts is reused)
    CD_TRY(cudaMemcpyAsync(counts, indptr, ((size_t)num_rows + 1) * 4, cudaMemcpyDeviceToDevice, stream));
    if (threads) k_decode<1><<<(threads + TB - 1) / TB, TB, 0, stream>>>(a, counts, indices, vals, bad);
    CD_TRY(cudaGetLastError());
    CD_TRY(cudaStreamSynchronize(stream));
    cleanup();
    *d_indptr = indptr; *d_indices = indices; *d_vals = vals; *nnz = h_nnz;
    return cudaSuccess;
#undef CD_TRY
}

}  // namespace hsb
