#include "cpsr_decode.h"

#include <algorithm>

namespace hsb {

bool impl_config(int impl, ImplConfig *out) {
    switch (impl) {
    case 0: *out = ImplConfig{1, 16u * 8u * 8192u, 8u * 4096u, true}; return true;    // spmv/libfpga/common.h:164-179
    case 1: *out = ImplConfig{1, 16u * 8u * 1024u, 8u * 4096u, false}; return true;   // spmv-fp/libfpga/common.h:178-182
    case 2: *out = ImplConfig{8, 16u * 8u * 8192u, 8u * 4096u, false}; return true;   // spmv-fp/libfpga/common.h:185-188
    default: return false;
    }
}

namespace {
// 16 words per 64-byte packet: [0..7] column indices, [8..15] value words (common.h:44-50)
inline const uint32_t *pkt(const uint32_t *img, size_t i) { return img + i * 16; }
}  // namespace

size_t cpsr_image_packets(const ImplConfig &cfg, const uint32_t *image, uint32_t num_partitions) {
    const uint32_t IF = cfg.interleave;
    size_t data_end = 0;
    for (uint32_t ij = 0; ij < num_partitions; ij++) {
        size_t start = pkt(image, (size_t)ij * (1 + IF))[0];
        uint32_t longest = 0;
        for (uint32_t f = 0; f < IF; f++)
            for (int k = 0; k < 8; k++) longest = std::max(longest, pkt(image, (size_t)ij * (1 + IF) + 1 + f)[k]);
        data_end = std::max(data_end, start + (size_t)longest * IF);
    }
    return (size_t)num_partitions * (1 + IF) + data_end;
}

bool cpsr_decode(const ImplConfig &cfg, const uint32_t *const images[16], const size_t *n_packets,
                 uint32_t num_row_partitions, uint32_t num_col_partitions, uint32_t part_begin,
                 uint32_t part_end, const uint32_t *rows_in_part, uint32_t num_cols, HostCsr *out,
                 std::string *err) {
    auto fail = [&](const char *m) { if (err) *err = m; return false; };
    const uint32_t IF = cfg.interleave;
    const uint32_t np = num_row_partitions * num_col_partitions;
    if (part_end > num_row_partitions || part_begin > part_end) return fail("row partition range out of bounds");
    const size_t base = (size_t)np * (1 + IF);

    std::vector<uint32_t> part_row0(part_end - part_begin + 1, 0);
    for (uint32_t j = part_begin; j < part_end; j++) {
        uint32_t n = rows_in_part[j - part_begin];
        if (n % 128u || n > cfg.ob_size) return fail("rows of a partition must be a multiple of 128 and <= LOGICAL_OB_SIZE");
        part_row0[j - part_begin + 1] = part_row0[j - part_begin] + n;
    }
    const uint32_t rows = part_row0.back();
    out->rows = rows;
    out->cols = num_cols;
    out->indptr.assign((size_t)rows + 1, 0);

    // The same walk is done twice: count per row, then place.
    for (int pass = 0; pass < 2; pass++) {
        std::vector<uint32_t> cursor;
        if (pass == 1) {
            for (uint32_t r = 0; r < rows; r++) out->indptr[r + 1] += out->indptr[r];
            out->indices.assign(out->indptr[rows], 0);
            out->vals.assign(out->indptr[rows], 0);
            cursor.assign(out->indptr.begin(), out->indptr.end() - 1);
        }
        for (uint32_t pc = 0; pc < 16; pc++) {
            const uint32_t *img = images[pc];
            const size_t limit = n_packets ? n_packets[pc] : ~(size_t)0;
            if (limit < base) return fail("channel image shorter than its header");
            for (uint32_t j = part_begin; j < part_end; j++) {
                const uint32_t part_len = rows_in_part[j - part_begin] / 16u;   // rows of this cluster
                for (uint32_t i = 0; i < num_col_partitions; i++) {
                    const size_t ij = (size_t)j * num_col_partitions + i;
                    const size_t start = pkt(img, ij * (1 + IF))[0];
                    for (uint32_t f = 0; f < IF; f++) {
                        const uint32_t *lens = pkt(img, ij * (1 + IF) + 1 + f);
                        for (uint32_t k = 0; k < 8; k++) {
                            uint32_t row_local = f * 8 + k;            // spmv_cluster.h:59-64 / fp :78-86
                            for (uint32_t n = 0; n < lens[k]; n++) {
                                size_t pi = base + start + (size_t)n * IF + f;
                                if (pi >= limit) return fail("channel image truncated");
                                const uint32_t *p = pkt(img, pi);
                                uint32_t idx = p[k], v = p[8 + k];
                                if (idx == 0xFFFFFFFFu) {              // end-of-row marker
                                    uint32_t adv = cfg.marker_is_q824 ? (v >> 24) : v;
                                    row_local += 8u * adv * IF;
                                    continue;
                                }
                                if (row_local >= part_len) continue;   // never dumped by the PE (pe.h:95-116)
                                if (idx >= cfg.vb_size) return fail("column index beyond the vector buffer");
                                uint64_t col = (uint64_t)i * cfg.vb_size + idx;
                                if (col >= num_cols) return fail("column index out of range");
                                // drain order: y[(r/8)*128 + pc*8 + r%8]  (spmv_result_drain.cpp:36-113)
                                uint32_t row = part_row0[j - part_begin] + (row_local / 8u) * 128u + pc * 8u + (row_local % 8u);
                                if (pass == 0) {
                                    out->indptr[row + 1]++;
                                } else {
                                    uint32_t at = cursor[row]++;
                                    out->indices[at] = (uint32_t)col;
                                    out->vals[at] = v;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    return true;
}

}  // namespace hsb
