// GPU-side preprocessing: CSR -> tile-stream format built on the device.
// The reference formats on one host thread (csr2cpsr, sw/data_formatter.h:468-544; 0.02 - 10.6 s per
// dataset, paper Table 8); tile_format.cpp does it on all host cores; this does the same three
// stages with device-wide sorts and scans, so that a matrix that already lives in HBM (generated on
// the device, or the previous stage of a pipeline) never visits the host:
//
//   stage A  key every non-zero by (row partition, column tile, row) and radix-sort   == column
//            partitioning (util_convert_csr_to_dds, :256-313) for all tiles at once
//   stage B  run-length encode the keys -> row segments; cut them into lane streams of <= 128;
//            radix-sort the streams by (tile, length descending); 32 streams = 1 slice
//   stage C  one warp per slice copies the 32 streams into the warp-coalesced slot order
//
// The result is the layout tile_format.h describes, except that the order of the non-zeros inside a
// lane stream is a per-lane rotation instead of the host builder's greedy bank scheduling.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "gpu_format.h"

namespace hsb {
namespace {

#define GF_TRY(expr)                                  \
    do {                                              \
        cudaError_t e__ = (expr);                     \
        if (e__ != cudaSuccess) return e__;           \
    } while (0)

struct StreamRec {
    uint32_t row;
    uint32_t len;
    uint32_t src;     // position of its first non-zero in the tile-sorted order
    uint32_t pad_;
};

__device__ __forceinline__ uint32_t n_pieces(uint32_t n, uint32_t max_len) { return (n + max_len - 1) / max_len; }

// one thread per non-zero: key = ((part * T + tile) << 32) | row; the row comes from the CSR row pointers, or
// (coo_rows != null) from a COO list
__global__ void k_keys(uint64_t nnz, uint32_t rows, const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ coo_rows,
                       const uint32_t *__restrict__ indices, uint32_t rows_per_part, uint32_t tile_cols,
                       uint32_t T, uint32_t cols, unsigned long long *__restrict__ keys, uint32_t *__restrict__ ids,
                       int *__restrict__ bad) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    uint32_t lo = 0, hi = rows;                       // last row r with indptr[r] <= e
    if (coo_rows) {
        lo = coo_rows[e];
        if (lo >= rows) { *bad = 1; keys[e] = 0; ids[e] = (uint32_t)e; return; }
    } else {
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (indptr[mid] <= e) lo = mid; else hi = mid;
        }
    }
    const uint32_t col = indices[e];
    if (col >= cols) { *bad = 1; keys[e] = 0; ids[e] = (uint32_t)e; return; }
    const uint32_t tp = (lo / rows_per_part) * T + col / tile_cols;
    keys[e] = ((unsigned long long)tp << 32) | lo;
    ids[e] = (uint32_t)e;
}

__global__ void k_piece_counts(uint32_t n_segs, const uint32_t *__restrict__ seg_len, uint32_t max_len, uint32_t *__restrict__ pieces) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_segs) pieces[s] = n_pieces(seg_len[s], max_len);
}

// one thread per segment: emit its lane streams (balanced piece lengths) and their sort keys
__global__ void k_streams(uint32_t n_segs, const unsigned long long *__restrict__ seg_key,
                          const uint32_t *__restrict__ seg_len, const uint32_t *__restrict__ seg_start,
                          const uint32_t *__restrict__ stream_off, uint32_t max_len, StreamRec *__restrict__ recs,
                          unsigned long long *__restrict__ skeys, uint32_t *__restrict__ sids) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segs) return;
    const uint32_t n = seg_len[s], p = n_pieces(n, max_len), base = n / p, extra = n % p;
    const unsigned long long tp = seg_key[s] >> 32;
    uint32_t src = seg_start[s], o = stream_off[s];
    for (uint32_t q = 0; q < p; q++, o++) {
        const uint32_t l = base + (q < extra ? 1u : 0u);
        recs[o] = StreamRec{(uint32_t)seg_key[s], l, src, 0u};
        skeys[o] = (tp << 8) | (kMaxStreamLen - l);           // tile major, longer streams first
        sids[o] = o;
        src += l;
    }
}

// tile_stream_begin[tp] = first sorted stream whose tile >= tp   (NT + 1 entries)
__global__ void k_tile_bounds(uint32_t NT, uint32_t n_streams, const unsigned long long *__restrict__ skeys_sorted,
                              uint32_t *__restrict__ tile_stream_begin) {
    uint32_t tp = blockIdx.x * blockDim.x + threadIdx.x;
    if (tp > NT) return;
    uint32_t lo = 0, hi = n_streams;                  // first index with (key >> 8) >= tp
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if ((skeys_sorted[mid] >> 8) >= tp) hi = mid; else lo = mid + 1;
    }
    tile_stream_begin[tp] = lo;
}

__global__ void k_tile_slices(uint32_t NT, const uint32_t *__restrict__ tile_stream_begin, uint32_t *__restrict__ tile_slices) {
    uint32_t tp = blockIdx.x * blockDim.x + threadIdx.x;
    if (tp < NT) tile_slices[tp] = (tile_stream_begin[tp + 1] - tile_stream_begin[tp] + kLanes - 1) / kLanes;
}

// one thread per slice: its tile (binary search) and its step count
__global__ void k_slice_steps(uint32_t n_slices, uint32_t NT, const uint32_t *__restrict__ tile_slice_begin,
                              const uint32_t *__restrict__ tile_stream_begin, const uint32_t *__restrict__ sids_sorted,
                              const StreamRec *__restrict__ recs, uint32_t slot_block, uint32_t row_units,
                              uint32_t *__restrict__ slice_tile, uint32_t *__restrict__ slice_steps) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slices) return;
    uint32_t lo = 0, hi = NT;                         // last tile with tile_slice_begin[tile] <= s
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (tile_slice_begin[mid] <= s) lo = mid; else hi = mid;
    }
    // skip empty tiles that share the same slice_begin: take the LAST tile with begin <= s that is non-empty
    while (tile_slice_begin[lo + 1] <= s) lo++;
    const uint32_t first = tile_stream_begin[lo] + (s - tile_slice_begin[lo]) * kLanes;
    slice_tile[s] = lo;
    slice_steps[s] = (recs[sids_sorted[first]].len + slot_block - 1) / slot_block + row_units;   // narrow: + the row unit
}

// one thread per (tile, c): cnt_ge[c] = slices of the tile with more than c steps (slices are sorted)
__global__ void k_cnt_ge(uint32_t NT, const uint32_t *__restrict__ tile_slice_begin,
                         const uint32_t *__restrict__ slice_steps, uint32_t row_units, uint32_t *__restrict__ cnt_ge) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NT * 32u) return;
    const uint32_t tp = i >> 5, c = i & 31u;
    uint32_t lo = tile_slice_begin[tp], hi = tile_slice_begin[tp + 1];
    const uint32_t b = lo;
    while (lo < hi) {                                  // first slice with steps <= c
        uint32_t mid = (lo + hi) >> 1;
        if (slice_steps[mid] - row_units > c) lo = mid + 1; else hi = mid;
    }
    cnt_ge[i] = lo - b;
}

// one warp per slice: lane l copies stream l into the slot order of the kernel's vector loads
__global__ void k_fill(uint32_t n_slices, uint32_t rows, uint32_t tile_cols, uint32_t T,
                       const uint32_t *__restrict__ slice_tile, const uint32_t *__restrict__ slice_off,
                       const uint32_t *__restrict__ tile_slice_begin, const uint32_t *__restrict__ tile_stream_begin,
                       const uint32_t *__restrict__ sids_sorted, const StreamRec *__restrict__ recs,
                       const uint32_t *__restrict__ ids_sorted, const uint32_t *__restrict__ indices,
                       const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ vals, uint16_t *__restrict__ cols16,
                       uint32_t *__restrict__ slice_rows, int narrow) {
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (s >= n_slices) return;
    const uint32_t tp = slice_tile[s];
    const uint32_t i = tile_stream_begin[tp] + (s - tile_slice_begin[tp]) * kLanes + lane;
    uint32_t row = rows, len = 0, src = 0;
    if (i < tile_stream_begin[tp + 1]) {
        const StreamRec r = recs[sids_sorted[i]];
        row = r.row; len = r.len; src = r.src;
    }
    // row ids: wide layout in slice_rows, narrow layout in the slice's leading row unit (column ids stay 0)
    const size_t base = (size_t)slice_off[s] * (narrow ? kUnitElems : kStepElems);
    if (narrow) vals[base + lane] = row; else slice_rows[(size_t)s * kLanes + lane] = row;
    const uint32_t col_base = (tp % T) * tile_cols;
    for (uint32_t k = 0; k < len; k++) {
        // per-lane rotation of the stream: neighbouring lanes that walk the same dense row start
        // at different columns, i.e. different shared-memory banks
        uint32_t from = k + lane;
        from = from >= len ? from % len : from;
        const uint32_t e = ids_sorted[src + from];
        const size_t at = narrow ? base + (size_t)(1 + k) * kUnitElems + lane
                                 : base + (size_t)(k / kSlotBlock) * kStepElems + (size_t)lane * kSlotBlock + (k % kSlotBlock);
        vals[at] = vals_in[e];
        cols16[at] = (uint16_t)(indices[e] - col_base + kColBias);
    }
}

template <class T> cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)); }

struct Scratch {
    std::vector<void *> ptrs;
    template <class T> cudaError_t get(T **p, size_t n) {
        cudaError_t e = dalloc(p, n);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    ~Scratch() { for (void *p : ptrs) cudaFree(p); }
};

inline int bits_for(uint64_t v) { int b = 0; while ((1ull << b) <= v && b < 63) b++; return std::max(b, 1); }

}  // namespace

cudaError_t build_tiled_gpu(uint32_t rows, uint32_t cols, uint64_t nnz, const uint32_t *d_indptr, const uint32_t *d_coo_rows,
                            const uint32_t *d_indices, const uint32_t *d_vals, uint32_t rows_per_part,
                            uint32_t tile_cols, cudaStream_t stream, TiledMatrix *meta, DeviceFormat *out,
                            std::string *err) {
    auto fail = [&](const char *m) { if (err) *err = m; return cudaErrorInvalidValue; };
    if (tile_cols == 0 || tile_cols > kMaxTileCols || (tile_cols & 7u)) return fail("bad tile_cols");
    // the device-wide sorts and scans take 32-bit signed item counts, and step / unit offsets (at most 2 per
    // non-zero) are 32-bit exclusive sums
    if (nnz > 0x7FFFFFFFull) return fail("more than 2^31 - 1 non-zeros in one matrix (shard it by row blocks)");
    TiledMatrix &M = *meta;
    M = TiledMatrix();
    M.rows = rows; M.cols = cols; M.nnz = nnz;
    M.rows_per_part = rows_per_part ? rows_per_part : std::max(rows, 1u);
    M.n_row_parts = rows ? (rows + M.rows_per_part - 1) / M.rows_per_part : 0;
    M.tile_cols = tile_cols;
    M.n_col_tiles = std::max(1u, (cols + tile_cols - 1) / tile_cols);
    const uint32_t T = M.n_col_tiles, NT = M.n_row_parts * T;
    *out = DeviceFormat();
    M.tiles.assign(NT, TileDesc());
    M.part_slice_begin.assign(M.n_row_parts + 1, 0);
    for (uint32_t tp = 0; tp < NT; tp++) {
        TileDesc &td = M.tiles[tp];
        std::memset(&td, 0, sizeof td);
        td.col_base = (tp % T) * tile_cols;
        uint32_t width = std::min(tile_cols, cols > td.col_base ? cols - td.col_base : 0u);
        td.col_count = (width + 7u) & ~7u;
        td.row_part = tp / T;
    }
    if (nnz == 0 || NT == 0) return cudaSuccess;
    if (NT >= (1u << 24)) return fail("too many tiles");

    Scratch sc;
    const int TB = 256;
    auto blocks = [&](uint64_t n) { return (unsigned)((n + TB - 1) / TB); };
    size_t tmp_bytes = 0;
    void *tmp = nullptr;
    auto ensure_tmp = [&](size_t need) -> cudaError_t {
        if (need <= tmp_bytes) return cudaSuccess;
        if (tmp) cudaFree(tmp);
        tmp_bytes = need;
        return cudaMalloc(&tmp, need);
    };
    struct TmpGuard { void **p; ~TmpGuard() { if (*p) cudaFree(*p); } } guard{&tmp};

    // ---- stage A: sort the non-zeros by (tile, row), stable ---------------------------------
    unsigned long long *keys = nullptr, *keys_s = nullptr;
    uint32_t *ids = nullptr, *ids_s = nullptr;
    int *bad = nullptr;
    GF_TRY(sc.get(&keys, nnz)); GF_TRY(sc.get(&keys_s, nnz));
    GF_TRY(sc.get(&ids, nnz)); GF_TRY(sc.get(&ids_s, nnz));
    GF_TRY(sc.get(&bad, 1));
    GF_TRY(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    k_keys<<<blocks(nnz), TB, 0, stream>>>(nnz, rows, d_indptr, d_coo_rows, d_indices, M.rows_per_part, tile_cols, T, cols, keys, ids, bad);
    const int key_bits = 32 + bits_for(NT);
    size_t need = 0;
    GF_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, keys, keys_s, ids, ids_s, (int)nnz, 0, key_bits, stream));
    GF_TRY(ensure_tmp(need));
    GF_TRY(cub::DeviceRadixSort::SortPairs(tmp, need, keys, keys_s, ids, ids_s, (int)nnz, 0, key_bits, stream));
    int h_bad = 0;
    GF_TRY(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, stream));

    // ---- stage B: segments -> lane streams -> slices -----------------------------------------
    unsigned long long *seg_key = nullptr;
    uint32_t *seg_len = nullptr, *d_nsegs = nullptr;
    GF_TRY(sc.get(&seg_key, nnz)); GF_TRY(sc.get(&seg_len, nnz)); GF_TRY(sc.get(&d_nsegs, 1));
    GF_TRY(cub::DeviceRunLengthEncode::Encode(nullptr, need, keys_s, seg_key, seg_len, d_nsegs, (int)nnz, stream));
    GF_TRY(ensure_tmp(need));
    GF_TRY(cub::DeviceRunLengthEncode::Encode(tmp, need, keys_s, seg_key, seg_len, d_nsegs, (int)nnz, stream));
    uint32_t n_segs = 0;
    GF_TRY(cudaMemcpyAsync(&n_segs, d_nsegs, 4, cudaMemcpyDeviceToHost, stream));
    GF_TRY(cudaStreamSynchronize(stream));
    if (h_bad) return fail("row or column index out of range");

    // layout (tile_format.h): narrow when the segments are nearly all one or two entries long
    M.narrow = choose_narrow(nnz, n_segs);
    const uint32_t max_len = M.narrow ? kNarrowMaxLen : kMaxStreamLen;
    const uint32_t slot_block = M.narrow ? 1u : (uint32_t)kSlotBlock, row_units = M.narrow ? 1u : 0u;
    uint32_t *seg_start = nullptr, *pieces = nullptr, *stream_off = nullptr;
    GF_TRY(sc.get(&seg_start, n_segs + 1)); GF_TRY(sc.get(&pieces, n_segs + 1)); GF_TRY(sc.get(&stream_off, n_segs + 1));
    GF_TRY(cudaMemsetAsync(pieces + n_segs, 0, 4, stream));
    k_piece_counts<<<blocks(n_segs), TB, 0, stream>>>(n_segs, seg_len, max_len, pieces);
    GF_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, seg_len, seg_start, (int)n_segs, stream));
    GF_TRY(ensure_tmp(need));
    GF_TRY(cub::DeviceScan::ExclusiveSum(tmp, need, seg_len, seg_start, (int)n_segs, stream));
    GF_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, pieces, stream_off, (int)n_segs + 1, stream));
    GF_TRY(ensure_tmp(need));
    GF_TRY(cub::DeviceScan::ExclusiveSum(tmp, need, pieces, stream_off, (int)n_segs + 1, stream));
    uint32_t n_streams = 0;
    GF_TRY(cudaMemcpyAsync(&n_streams, stream_off + n_segs, 4, cudaMemcpyDeviceToHost, stream));
    GF_TRY(cudaStreamSynchronize(stream));

    StreamRec *recs = nullptr;
    unsigned long long *skeys = nullptr, *skeys_s = nullptr;
    uint32_t *sids = nullptr, *sids_s = nullptr;
    GF_TRY(sc.get(&recs, n_streams)); GF_TRY(sc.get(&skeys, n_streams)); GF_TRY(sc.get(&skeys_s, n_streams));
    GF_TRY(sc.get(&sids, n_streams)); GF_TRY(sc.get(&sids_s, n_streams));
    k_streams<<<blocks(n_segs), TB, 0, stream>>>(n_segs, seg_key, seg_len, seg_start, stream_off, max_len, recs, skeys, sids);
    const int skey_bits = 8 + bits_for(NT);
    GF_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, skeys, skeys_s, sids, sids_s, (int)n_streams, 0, skey_bits, stream));
    GF_TRY(ensure_tmp(need));
    GF_TRY(cub::DeviceRadixSort::SortPairs(tmp, need, skeys, skeys_s, sids, sids_s, (int)n_streams, 0, skey_bits, stream));

    uint32_t *tile_stream_begin = nullptr, *tile_slices = nullptr, *tile_slice_begin = nullptr;
    GF_TRY(sc.get(&tile_stream_begin, NT + 1)); GF_TRY(sc.get(&tile_slices, NT + 1)); GF_TRY(sc.get(&tile_slice_begin, NT + 1));
    k_tile_bounds<<<blocks(NT + 1), TB, 0, stream>>>(NT, n_streams, skeys_s, tile_stream_begin);
    GF_TRY(cudaMemsetAsync(tile_slices + NT, 0, 4, stream));
    k_tile_slices<<<blocks(NT), TB, 0, stream>>>(NT, tile_stream_begin, tile_slices);
    GF_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, tile_slices, tile_slice_begin, (int)NT + 1, stream));
    GF_TRY(ensure_tmp(need));
    GF_TRY(cub::DeviceScan::ExclusiveSum(tmp, need, tile_slices, tile_slice_begin, (int)NT + 1, stream));
    uint32_t n_slices = 0;
    GF_TRY(cudaMemcpyAsync(&n_slices, tile_slice_begin + NT, 4, cudaMemcpyDeviceToHost, stream));
    GF_TRY(cudaStreamSynchronize(stream));

    uint32_t *slice_tile = nullptr, *slice_steps = nullptr, *slice_off = nullptr, *cnt_ge = nullptr;
    GF_TRY(sc.get(&slice_tile, n_slices)); GF_TRY(sc.get(&slice_steps, n_slices + 1)); GF_TRY(sc.get(&slice_off, n_slices + 1));
    GF_TRY(sc.get(&cnt_ge, (size_t)NT * 32));
    GF_TRY(cudaMemsetAsync(slice_steps + n_slices, 0, 4, stream));
    k_slice_steps<<<blocks(n_slices), TB, 0, stream>>>(n_slices, NT, tile_slice_begin, tile_stream_begin, sids_s, recs, slot_block, row_units, slice_tile, slice_steps);
    GF_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, slice_steps, slice_off, (int)n_slices + 1, stream));
    GF_TRY(ensure_tmp(need));
    GF_TRY(cub::DeviceScan::ExclusiveSum(tmp, need, slice_steps, slice_off, (int)n_slices + 1, stream));
    k_cnt_ge<<<blocks((uint64_t)NT * 32), TB, 0, stream>>>(NT, tile_slice_begin, slice_steps, row_units, cnt_ge);
    uint32_t n_steps = 0;
    GF_TRY(cudaMemcpyAsync(&n_steps, slice_off + n_slices, 4, cudaMemcpyDeviceToHost, stream));

    // host copies of the (small) geometry tables: the launch planner runs on the host
    std::vector<uint32_t> h_tsb(NT + 1), h_steps(n_slices), h_off(n_slices + 1), h_tile(n_slices), h_cnt((size_t)NT * 32);
    GF_TRY(cudaMemcpyAsync(h_tsb.data(), tile_slice_begin, (NT + 1) * 4, cudaMemcpyDeviceToHost, stream));
    GF_TRY(cudaMemcpyAsync(h_steps.data(), slice_steps, (size_t)n_slices * 4, cudaMemcpyDeviceToHost, stream));
    GF_TRY(cudaMemcpyAsync(h_off.data(), slice_off, ((size_t)n_slices + 1) * 4, cudaMemcpyDeviceToHost, stream));
    GF_TRY(cudaMemcpyAsync(h_tile.data(), slice_tile, (size_t)n_slices * 4, cudaMemcpyDeviceToHost, stream));
    GF_TRY(cudaMemcpyAsync(h_cnt.data(), cnt_ge, (size_t)NT * 32 * 4, cudaMemcpyDeviceToHost, stream));
    GF_TRY(cudaStreamSynchronize(stream));

    // ---- stage C: fill the slices -------------------------------------------------------------
    const size_t n_elems = (size_t)n_steps * M.step_elems();
    GF_TRY(dalloc(&out->vals, n_elems + 4));
    GF_TRY(dalloc(&out->cols, n_elems + 8));
    GF_TRY(dalloc(&out->slice_rows, (M.narrow ? 0 : (size_t)n_slices * kLanes) + 4));
    GF_TRY(cudaMemsetAsync(out->vals, 0, n_elems * 4, stream));
    GF_TRY(cudaMemsetAsync(out->cols, 0, n_elems * 2, stream));                 // kPadCol == 0
    k_fill<<<blocks((uint64_t)n_slices * 32), TB, 0, stream>>>(n_slices, rows, tile_cols, T, slice_tile, slice_off,
                                                             tile_slice_begin, tile_stream_begin, sids_s, recs, ids_s,
                                                             d_indices, d_vals, out->vals, out->cols, out->slice_rows, M.narrow ? 1 : 0);
    GF_TRY(cudaGetLastError());
    GF_TRY(cudaStreamSynchronize(stream));
    out->n_elems = n_elems; out->n_slices = n_slices; out->n_streams = n_streams;

    M.n_streams = n_streams;
    M.slices.resize(n_slices);
    for (uint32_t s = 0; s < n_slices; s++) M.slices[s] = SliceDesc{h_off[s], (h_tile[s] << 8) | h_steps[s]};
    for (uint32_t tp = 0; tp < NT; tp++) {
        TileDesc &td = M.tiles[tp];
        td.slice_begin = h_tsb[tp];
        td.slice_end = h_tsb[tp + 1];
        td.step_begin = h_tsb[tp] < n_slices ? h_off[h_tsb[tp]] : n_steps;
        std::memcpy(td.cnt_ge, &h_cnt[(size_t)tp * 32], 32 * 4);
    }
    for (uint32_t j = 0; j <= M.n_row_parts; j++) M.part_slice_begin[j] = h_tsb[(size_t)j * T];
    return cudaSuccess;
}

}  // namespace hsb
