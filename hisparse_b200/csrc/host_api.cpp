// Host-only part of the C ABI: format inspection and CPSR decoding, usable without a GPU.
#include "../../include/hisparse_b200.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "cpsr_decode.h"
#include "tile_format.h"

struct hsb_format {
    hsb::TiledMatrix M;
};

extern "C" {

hsb_format *hsb_format_build(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                             const void *vals, uint32_t rows_per_partition, uint32_t tile_cols) {
    hsb_format *f = new hsb_format;
    std::string err;
    if (!hsb::build_tiled(rows, cols, indptr, indices, (const uint32_t *)vals, rows_per_partition,
                          tile_cols ? tile_cols : hsb::choose_tile_cols(cols), 0, &f->M, &err)) {
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
    out->n_col_tiles = M.n_col_tiles; out->tile_cols = M.tile_cols; out->n_chunks = M.n_chunks();
    out->n_segments = M.seg_row.size(); out->format_bytes = M.format_bytes();
    out->algorithmic_bytes = 8ull * M.nnz + 4ull * ((uint64_t)M.rows + 1) + 4ull * M.rows + 4ull * M.cols;
    return HSB_OK;
}

int hsb_format_expand(const hsb_format *f, uint32_t *indptr, uint32_t *indices, uint32_t *vals) {
    if (!f || !indptr) return HSB_EINVAL;
    const hsb::TiledMatrix &M = f->M;
    // pass 0 counts per row, pass 1 places; both walk chunk by chunk, lane by lane, like the kernel
    std::vector<uint32_t> cursor;
    std::fill(indptr, indptr + M.rows + 1, 0u);
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            for (uint32_t r = 0; r < M.rows; r++) indptr[r + 1] += indptr[r];
            if (indptr[M.rows] != M.nnz) return HSB_EINVAL;
            cursor.assign(indptr, indptr + M.rows);
        }
        for (size_t ti = 0; ti < M.tiles.size(); ti++) {
            const hsb::TileDesc &td = M.tiles[ti];
            uint32_t row_lo = td.row_part * M.rows_per_part;
            for (uint32_t c = td.chunk_begin; c < td.chunk_end; c++) {
                const hsb::ChunkDesc &cd = M.chunks[c];
                if ((cd.tile & ~hsb::kChunkContinues) != ti) return HSB_EINVAL;
                uint32_t seg = cd.seg_base;
                int last_flag = -1, last_real = -1;
                for (int lane = 0; lane < hsb::kLanes; lane++)
                    for (int k = 0; k < hsb::kNnzPerLane; k++) {
                        int j = lane * hsb::kNnzPerLane + k;
                        uint16_t w = M.cidx[(size_t)c * hsb::kChunkNnz + j];
                        uint32_t v = M.vals[(size_t)c * hsb::kChunkNnz + hsb::val_slot(lane, k)];
                        bool flag = w & hsb::kSegEndFlag;
                        // padding: after the tile's last flag (zero value, zero column, no flag)
                        bool is_pad = !(cd.tile & hsb::kChunkContinues) && !flag && [&] {
                            for (int q = j; q < hsb::kChunkNnz; q++)
                                if (M.cidx[(size_t)c * hsb::kChunkNnz + q] & hsb::kSegEndFlag) return false;
                            return true;
                        }();
                        if (is_pad) {
                            if (w || v) return HSB_EINVAL;
                            continue;
                        }
                        last_real = j;
                        if (seg >= M.seg_row.size()) return HSB_EINVAL;
                        uint32_t row = M.seg_row[seg];
                        if (row < row_lo || row >= M.rows) return HSB_EINVAL;
                        uint32_t col = td.col_base + (w & 0x7FFFu);
                        if ((w & 0x7FFFu) >= td.col_count || col >= M.cols) return HSB_EINVAL;
                        if (pass == 0) {
                            indptr[row + 1]++;
                        } else {
                            uint32_t at = cursor[row]++;
                            indices[at] = col;
                            vals[at] = v;
                        }
                        if (flag) { seg++; last_flag = j; }
                    }
                bool cont = last_real > last_flag;
                if (cont != bool(cd.tile & hsb::kChunkContinues)) return HSB_EINVAL;
                if (c + 1 < td.chunk_end && M.chunks[c + 1].seg_base != seg) return HSB_EINVAL;
            }
        }
    }
    return HSB_OK;
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
