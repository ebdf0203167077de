// Host-only part of the C ABI: format inspection and CPSR decoding, usable without a GPU.
#include "../../include/hisparse_b200.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "cpsr_decode.h"
#include "format_handle.h"
#include "tile_format.h"

extern "C" {

hsb_format *hsb_format_build(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                             const void *vals, uint32_t rows_per_partition, uint32_t tile_cols) {
    hsb_format *f = new hsb_format;
    std::string err;
    if (!hsb::build_tiled(rows, cols, indptr, indices, (const uint32_t *)vals, rows_per_partition,
                          tile_cols ? tile_cols : hsb::choose_tile_cols(cols, rows, rows ? indptr[rows] : 0), 0, &f->M, &err)) {
        delete f;
        return nullptr;
    }
    return f;
}

int hsb_format_stats(const hsb_format *f, hsb_stats *out) {
    if (!f || !out) return HSB_EINVAL;
    std::memset(out, 0, sizeof *out);
    const hsb::TiledMatrix &M = f->M;
    out->nnz = M.nnz; out->rows = M.rows; out->cols = M.cols; out->n_row_parts = M.n_row_parts;
    out->n_col_tiles = M.n_col_tiles; out->tile_cols = M.tile_cols; out->n_slices = M.n_slices();
    out->n_streams = M.n_streams; out->n_elems = M.n_elems(); out->format_bytes = M.format_bytes();
    out->layout = M.narrow ? 1u : 0u;
    out->algorithmic_bytes = 8ull * M.nnz + 4ull * ((uint64_t)M.rows + 1) + 4ull * M.rows + 4ull * M.cols;
    return HSB_OK;
}

// The launch plan hsb_spmv would use on `ctas` CTAs, flattened for inspection: one record of
// 4 words per (segment, warp) = {cta, tile, first step, end step} (tile-relative steps).
long long hsb_format_plan(const hsb_format *f, uint32_t ctas, uint32_t *records, size_t capacity_records) {
    if (!f || ctas == 0) return HSB_EINVAL;
    const hsb::TiledMatrix &M = f->M;
    std::vector<uint32_t> cta_seg;
    std::vector<hsb::Segment> segs;
    hsb::plan_launch(M, 0, (uint32_t)M.tiles.size(), ctas, &cta_seg, &segs);
    size_t n = 0;
    for (uint32_t b = 0; b < ctas; b++)
        for (uint32_t g = cta_seg[b]; g < cta_seg[b + 1]; g++)
            for (int w = 0; w < hsb::kWarpsPerCta; w++) {
                if (segs[g].warp_t[w] == segs[g].warp_t[w + 1]) continue;
                if (records && n < capacity_records) {
                    records[4 * n + 0] = b; records[4 * n + 1] = segs[g].tile;
                    records[4 * n + 2] = segs[g].warp_t[w]; records[4 * n + 3] = segs[g].warp_t[w + 1];
                }
                n++;
            }
    return (long long)n;
}

int hsb_format_expand(const hsb_format *f, uint32_t *indptr, uint32_t *indices, uint32_t *vals) {
    if (!f || !indptr) return HSB_EINVAL;
    const hsb::TiledMatrix &M = f->M;
    // pass 0 counts per row, pass 1 places; both walk slice by slice, lane by lane, with the very
    // address arithmetic the kernel uses (slice_elem == the warp's vector-load pattern)
    std::vector<uint32_t> cursor;
    std::fill(indptr, indptr + M.rows + 1, 0u);
    uint64_t streams_seen = 0;
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            for (uint32_t r = 0; r < M.rows; r++) indptr[r + 1] += indptr[r];
            if (indptr[M.rows] != M.nnz) return HSB_EINVAL;
            cursor.assign(indptr, indptr + M.rows);
        }
        for (size_t ti = 0; ti < M.tiles.size(); ti++) {
            const hsb::TileDesc &td = M.tiles[ti];
            const uint32_t row_lo = td.row_part * M.rows_per_part;
            uint32_t prev_steps = 0xFFFFFFFFu;
            for (uint32_t s = td.slice_begin; s < td.slice_end; s++) {
                const hsb::SliceDesc &sd = M.slices[s];
                if ((sd.tile_steps >> 8) != ti) return HSB_EINVAL;
                const uint32_t steps = sd.tile_steps & 0xFFu;        // narrow layout: 1 row unit + L step units
                const uint32_t max_steps = M.narrow ? 1 + hsb::kNarrowMaxLen : hsb::kMaxStreamLen / hsb::kSlotBlock;
                if (steps <= (M.narrow ? 1u : 0u) || steps > max_steps || steps > prev_steps) return HSB_EINVAL;
                prev_steps = steps;                                  // sorted by length inside a tile
                const size_t base = (size_t)sd.off * M.step_elems();
                if (base + (size_t)steps * M.step_elems() > M.vals.size()) return HSB_EINVAL;
                const uint32_t slots = M.narrow ? steps - 1 : steps * hsb::kSlotBlock;
                for (int lane = 0; lane < hsb::kLanes; lane++) {
                    if (M.narrow && M.cols16[base + lane] != hsb::kPadCol) return HSB_EINVAL;       // row unit: no column ids
                    const uint32_t row = M.narrow ? M.vals[base + lane] : M.slice_rows[(size_t)s * hsb::kLanes + lane];
                    bool ended = false;
                    uint32_t len = 0;
                    for (uint32_t k = 0; k < slots; k++) {
                        const size_t e = M.narrow ? base + (size_t)(1 + k) * hsb::kUnitElems + lane : hsb::slice_elem(base, lane, k);
                        const uint16_t c16 = M.cols16[e];
                        if (c16 == hsb::kPadCol) {                   // padding: zero value, and only at the tail
                            if (M.vals[e]) return HSB_EINVAL;
                            ended = true;
                            continue;
                        }
                        if (ended || row >= M.rows || row < row_lo) return HSB_EINVAL;
                        if (c16 < hsb::kColBias) return HSB_EINVAL;
                        const uint32_t lc = c16 - hsb::kColBias;
                        if (lc >= td.col_count || td.col_base + lc >= M.cols) return HSB_EINVAL;
                        len++;
                        if (pass == 0) {
                            indptr[row + 1]++;
                        } else {
                            uint32_t at = cursor[row]++;
                            indices[at] = td.col_base + lc;
                            vals[at] = M.vals[e];
                        }
                    }
                    if (row == M.rows && len) return HSB_EINVAL;     // unused lanes hold nothing
                    if (pass == 0 && len) streams_seen++;
                }
            }
        }
    }
    return streams_seen == M.n_streams ? HSB_OK : HSB_EINVAL;
}

void hsb_format_free(hsb_format *f) { delete f; }

int hsb_cpsr_to_csr(int impl, const void *const ch[HSB_NUM_HBM_CHANNELS],
                    const size_t ch_packets[HSB_NUM_HBM_CHANNELS], unsigned num_row_partitions,
                    unsigned num_col_partitions, unsigned num_rows, unsigned num_cols, uint32_t *indptr,
                    uint32_t *indices, uint32_t *vals, size_t capacity, size_t *nnz) {
    hsb::ImplConfig cfg;
    if (!hsb::impl_config(impl, &cfg) || !ch || !nnz) return HSB_EINVAL;
    const uint32_t *imgs[16];
    for (int i = 0; i < 16; i++) imgs[i] = (const uint32_t *)ch[i];
    std::vector<uint32_t> rows_in(num_row_partitions);
    for (unsigned j = 0; j < num_row_partitions; j++)
        rows_in[j] = (uint32_t)std::min<uint64_t>(cfg.ob_size, (uint64_t)num_rows - (uint64_t)j * cfg.ob_size);
    hsb::HostCsr csr;
    std::string err;
    if (!hsb::cpsr_decode(cfg, imgs, ch_packets, num_row_partitions, num_col_partitions, 0, num_row_partitions,
                          rows_in.data(), num_cols, &csr, &err))
        return HSB_EINVAL;
    *nnz = csr.indices.size();
    if (capacity < csr.indices.size()) return capacity == 0 ? HSB_OK : HSB_ENOMEM;
    if (indptr) std::memcpy(indptr, csr.indptr.data(), csr.indptr.size() * 4);
    if (indices && !csr.indices.empty()) std::memcpy(indices, csr.indices.data(), csr.indices.size() * 4);
    if (vals && !csr.vals.empty()) std::memcpy(vals, csr.vals.data(), csr.vals.size() * 4);
    return HSB_OK;
}

}  // extern "C"
