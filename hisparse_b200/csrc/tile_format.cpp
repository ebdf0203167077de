// Host-side builder for the tile-stream format (see tile_format.h). This is the B200
// counterpart of the reference's CSR -> CPSR preprocessing (sw/data_formatter.h:468-544):
// same inputs (a CSR with 32-bit indices and 32-bit value words, a row-partition length, a
// column-partition length), different output layout. Two passes (count, then place), both
// parallel over nnz-balanced row slices, so preprocessing is not the single-threaded
// bottleneck it is in the reference (paper Table 8: 0.02 - 10.6 s).
#include "tile_format.h"

#include <algorithm>
#include <cstring>
#include <thread>

namespace hsb {

uint32_t choose_tile_cols(uint32_t cols) {
    if (cols == 0) return 8;
    uint32_t n_tiles = (cols + kMaxTileCols - 1) / kMaxTileCols;
    uint32_t w = (cols + n_tiles - 1) / n_tiles;
    w = (w + 7u) & ~7u;
    return std::min(w, kMaxTileCols);
}

namespace {

struct Slice {
    uint32_t part, r0, r1;       // rows [r0, r1) of row partition `part`
};

template <class F> void parallel_for(size_t n, int n_threads, F f) {
    if (n_threads <= 1 || n <= 1) {
        for (size_t i = 0; i < n; i++) f(i);
        return;
    }
    std::vector<std::thread> th;
    size_t nt = std::min<size_t>(n_threads, n);
    for (size_t t = 0; t < nt; t++)
        th.emplace_back([&, t]() { for (size_t i = t; i < n; i += nt) f(i); });
    for (auto &x : th) x.join();
}

}  // namespace

bool build_tiled(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                 const uint32_t *vals, uint32_t rows_per_part, uint32_t tile_cols, int n_threads,
                 TiledMatrix *out, std::string *err) {
    auto fail = [&](const char *m) { if (err) *err = m; return false; };
    if (tile_cols == 0 || tile_cols > kMaxTileCols || (tile_cols & 7u)) return fail("tile_cols must be a multiple of 8 in [8, 32768]");
    if (rows && indptr[0] != 0) return fail("indptr[0] must be 0");
    for (uint32_t r = 0; r < rows; r++)
        if (indptr[r + 1] < indptr[r]) return fail("indptr is not monotone");
    const uint64_t nnz = rows ? indptr[rows] : 0;
    if (n_threads <= 0) n_threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (nnz < (1u << 16)) n_threads = 1;

    TiledMatrix &M = *out;
    M = TiledMatrix();
    M.rows = rows; M.cols = cols; M.nnz = nnz;
    M.rows_per_part = rows_per_part ? rows_per_part : std::max(rows, 1u);
    M.n_row_parts = rows ? (rows + M.rows_per_part - 1) / M.rows_per_part : 0;
    M.tile_cols = tile_cols;
    M.n_col_tiles = std::max(1u, (cols + tile_cols - 1) / tile_cols);
    const uint32_t T = M.n_col_tiles;

    // nnz-balanced row slices inside every row partition
    std::vector<Slice> slices;
    std::vector<uint32_t> part_slice_begin(M.n_row_parts + 1, 0);
    for (uint32_t j = 0; j < M.n_row_parts; j++) {
        uint32_t r0 = j * M.rows_per_part, r1 = (uint32_t)std::min<uint64_t>(rows, (uint64_t)r0 + M.rows_per_part);
        uint64_t e0 = indptr[r0], e1 = indptr[r1];
        int ns = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_threads, (e1 - e0) >> 14));
        uint32_t prev = r0;
        for (int s = 1; s <= ns; s++) {
            uint32_t cut = r1;
            if (s < ns) {
                uint64_t target = e0 + (e1 - e0) * s / ns;
                cut = (uint32_t)(std::lower_bound(indptr + prev, indptr + r1, (uint32_t)target) - indptr);
                cut = std::min(std::max(cut, prev), r1);
            }
            if (cut > prev || s == ns) slices.push_back(Slice{j, prev, cut});
            prev = cut;
        }
        part_slice_begin[j + 1] = (uint32_t)slices.size();
    }
    const size_t NS = slices.size();

    // pass 1: per (slice, tile) non-zero and segment counts
    std::vector<uint64_t> cnt(NS * T, 0), seg(NS * T, 0);
    bool bad_col = false;
    parallel_for(NS, n_threads, [&](size_t s) {
        const Slice &sl = slices[s];
        std::vector<uint32_t> stamp(T, 0xFFFFFFFFu);
        uint64_t *c = &cnt[s * T], *g = &seg[s * T];
        for (uint32_t r = sl.r0; r < sl.r1; r++)
            for (uint32_t e = indptr[r]; e < indptr[r + 1]; e++) {
                uint32_t col = indices[e];
                if (col >= cols) { bad_col = true; continue; }
                uint32_t t = col / tile_cols;
                c[t]++;
                if (stamp[t] != r) { stamp[t] = r; g[t]++; }
            }
    });
    if (bad_col) return fail("column index out of range");

    // layout: tiles in (row partition, column tile) order; chunks and segments numbered globally
    M.tiles.resize((size_t)M.n_row_parts * T);
    M.part_chunk_begin.assign(M.n_row_parts + 1, 0);
    std::vector<uint64_t> pos0(NS * T), seg0(NS * T);   // first stream position / segment of a slice in a tile
    uint64_t chunk_cursor = 0, seg_cursor = 0;
    for (uint32_t j = 0; j < M.n_row_parts; j++) {
        M.part_chunk_begin[j] = (uint32_t)chunk_cursor;
        for (uint32_t t = 0; t < T; t++) {
            uint64_t n = 0;
            for (uint32_t s = part_slice_begin[j]; s < part_slice_begin[j + 1]; s++) {
                pos0[(size_t)s * T + t] = chunk_cursor * kChunkNnz + n;
                seg0[(size_t)s * T + t] = seg_cursor;
                n += cnt[(size_t)s * T + t];
                seg_cursor += seg[(size_t)s * T + t];
            }
            TileDesc &td = M.tiles[(size_t)j * T + t];
            std::memset(&td, 0, sizeof(td));
            td.col_base = t * tile_cols;
            uint32_t width = std::min(tile_cols, cols > td.col_base ? cols - td.col_base : 0u);
            td.col_count = (width + 7u) & ~7u;
            td.row_part = j;
            td.chunk_begin = (uint32_t)chunk_cursor;
            chunk_cursor += (n + kChunkNnz - 1) / kChunkNnz;
            td.chunk_end = (uint32_t)chunk_cursor;
        }
    }
    M.part_chunk_begin[M.n_row_parts] = (uint32_t)chunk_cursor;
    if (chunk_cursor >= (1ull << 31) || seg_cursor >= (1ull << 32)) return fail("matrix too large for 32-bit chunk/segment ids");
    const size_t NC = (size_t)chunk_cursor;
    M.vals.assign(NC * kChunkNnz, 0u);
    M.cidx.assign(NC * kChunkNnz, (uint16_t)0);
    M.seg_row.assign((size_t)seg_cursor, 0u);
    M.chunks.assign(NC, ChunkDesc{0, 0});

    // pass 2: place every non-zero; flag the last one of each (row, tile) segment
    parallel_for(NS, n_threads, [&](size_t s) {
        const Slice &sl = slices[s];
        std::vector<uint64_t> cur(pos0.begin() + s * T, pos0.begin() + (s + 1) * T);
        std::vector<uint64_t> scur(seg0.begin() + s * T, seg0.begin() + (s + 1) * T);
        std::vector<uint64_t> last(T, ~0ull);
        std::vector<uint32_t> touched;
        for (uint32_t r = sl.r0; r < sl.r1; r++) {
            for (uint32_t e = indptr[r]; e < indptr[r + 1]; e++) {
                uint32_t col = indices[e], t = col / tile_cols;
                uint64_t p = cur[t]++;
                size_t chunk = (size_t)(p / kChunkNnz);
                int j = (int)(p % kChunkNnz);
                M.vals[chunk * kChunkNnz + val_slot(j / kNnzPerLane, j % kNnzPerLane)] = vals[e];
                M.cidx[p] = (uint16_t)(col - t * tile_cols);
                if (last[t] == ~0ull) touched.push_back(t);
                last[t] = p;
            }
            for (uint32_t t : touched) {
                M.cidx[last[t]] |= kSegEndFlag;
                M.seg_row[scur[t]++] = r;
                last[t] = ~0ull;
            }
            touched.clear();
        }
    });

    // chunk descriptors: running segment count + "stream continues past the last flag"
    std::vector<uint32_t> flags_in(NC, 0);
    parallel_for((size_t)M.tiles.size(), n_threads, [&](size_t ti) {
        const TileDesc &td = M.tiles[ti];
        for (uint32_t c = td.chunk_begin; c < td.chunk_end; c++) {
            const uint16_t *w = &M.cidx[(size_t)c * kChunkNnz];
            uint32_t n = 0;
            int last_flag = -1;
            for (int j = 0; j < kChunkNnz; j++)
                if (w[j] & kSegEndFlag) { n++; last_flag = j; }
            flags_in[c] = n;
            // every tile stream ends with a flagged non-zero, so real entries after the last flag
            // exist exactly when the chunk is not the tile's last one or ... it simply has them:
            bool cont = (c + 1 < td.chunk_end) && last_flag != kChunkNnz - 1;
            M.chunks[c].tile = (uint32_t)ti | (cont ? kChunkContinues : 0u);
        }
    });
    uint32_t run = 0;
    for (size_t c = 0; c < NC; c++) { M.chunks[c].seg_base = run; run += flags_in[c]; }
    return true;
}

}  // namespace hsb
