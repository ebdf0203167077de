// spmv::io -- CSR -> CPSR ("cyclic packed streams of rows") formatting with the names, argument
// meaning and error behaviour of the reference's sw/data_formatter.h, plus the per-HBM-channel
// packet image layout that the reference repeats inline in sw/host.cpp:163-231,
// sw/benchmark.cpp:127-195 and spmv_csim/csim.cpp:229-297 (here: build_channel_images).
//
// Own implementation. Where the reference builds intermediate copies per stage (DDS partitions,
// marker-padded CSR, then packing), this one counts first and writes every packet once.
// Checked against the reference's golden vectors (unit_tests/test_io.cpp) and against the
// reference formatter itself by tests/test_host_cpp.py.
#ifndef HISPARSE_B200_HOST_DATA_FORMATTER_H_
#define HISPARSE_B200_HOST_DATA_FORMATTER_H_

#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <numeric>
#include <type_traits>
#include <vector>

#include "data_loader.h"

namespace spmv {
namespace io {

// sw/data_formatter.h:15-29
template <typename DataT>
void util_round_csr_matrix_dim(CSRMatrix<DataT> &m, uint32_t row_divisor, uint32_t col_divisor) {
    if (m.num_rows % row_divisor) {
        uint32_t pad = row_divisor - m.num_rows % row_divisor;
        m.adj_indptr.resize(m.adj_indptr.size() + pad, m.adj_indptr[m.num_rows]);
        m.num_rows += pad;
    }
    if (m.num_cols % col_divisor) m.num_cols += col_divisor - m.num_cols % col_divisor;
}

// sw/data_formatter.h:32-48
template <typename DataT> void util_normalize_csr_matrix_by_outdegree(CSRMatrix<DataT> &m) {
    std::vector<uint32_t> deg(m.num_cols, 0);
    for (uint32_t c : m.adj_indices) deg[c]++;
    for (size_t e = 0; e < m.adj_indices.size(); e++) m.adj_data[e] = DataT(1.0 / deg[m.adj_indices[e]]);
}

namespace detail {
// value stored in an end-of-row marker slot: the count itself for numeric types, its raw bit
// pattern when DataT is float (sw/data_formatter.h:69-74,154-158; decoded with val2bit on the
// device side, spmv-fp/libfpga/spmv_cluster.h:104)
template <typename DataT> inline DataT marker_value(uint32_t k) {
    if (std::is_same<DataT, float>::value) {
        DataT v = DataT();
        std::memcpy(static_cast<void *>(&v), &k, std::min(sizeof(DataT), sizeof(k)));
        return v;
    }
    return DataT(k);
}
// per-row marker counts: 0 = the row gets no marker
inline std::vector<uint32_t> marker_counts(const std::vector<uint32_t> &indptr, uint32_t stride, bool skip_empty_rows) {
    const uint32_t n = (uint32_t)indptr.size() - 1;
    std::vector<uint32_t> mk(n, 1);
    if (!skip_empty_rows) return mk;
    assert(n % stride == 0);
    auto keeps = [&](uint32_t r) { return r < stride || indptr[r + 1] != indptr[r]; };
    for (uint32_t r = 0; r < n; r++) mk[r] = keeps(r) ? 1 : 0;
    for (uint32_t r = 0; r < n; r++)
        if (mk[r])
            for (uint32_t q = r + stride; q < n && !keeps(q); q += stride) mk[r]++;
    return mk;
}
template <typename DataT>
void apply_markers(std::vector<DataT> &data, std::vector<uint32_t> &indices, std::vector<uint32_t> &indptr,
                   uint32_t idx_marker, const std::vector<uint32_t> &mk) {
    const uint32_t n = (uint32_t)indptr.size() - 1;
    size_t extra = 0;
    for (uint32_t v : mk) extra += v ? 1 : 0;
    std::vector<DataT> d2;
    std::vector<uint32_t> i2, p2(n + 1, 0);
    d2.reserve(data.size() + extra);
    i2.reserve(data.size() + extra);
    for (uint32_t r = 0; r < n; r++) {
        for (uint32_t e = indptr[r]; e < indptr[r + 1]; e++) { d2.push_back(data[e]); i2.push_back(indices[e]); }
        if (mk[r]) { d2.push_back(marker_value<DataT>(mk[r])); i2.push_back(idx_marker); }
        p2[r + 1] = (uint32_t)d2.size();
    }
    data.swap(d2); indices.swap(i2); indptr.swap(p2);
}
}  // namespace detail

// sw/data_formatter.h:51-83
template <typename DataT>
void util_pad_marker_end_of_row_no_skip_empty_rows(std::vector<DataT> &adj_data, std::vector<uint32_t> &adj_indices,
                                                   std::vector<uint32_t> &adj_indptr, uint32_t idx_marker) {
    detail::apply_markers(adj_data, adj_indices, adj_indptr, idx_marker, detail::marker_counts(adj_indptr, 1, false));
}
// sw/data_formatter.h:87-171
template <typename DataT>
void util_pad_marker_end_of_row_skip_empty_rows(std::vector<DataT> &adj_data, std::vector<uint32_t> &adj_indices,
                                                std::vector<uint32_t> &adj_indptr, uint32_t idx_marker,
                                                uint32_t interleave_stride) {
    detail::apply_markers(adj_data, adj_indices, adj_indptr, idx_marker,
                          detail::marker_counts(adj_indptr, interleave_stride, true));
}
// sw/data_formatter.h:175-187
template <typename DataT>
void util_pad_marker_end_of_row(std::vector<DataT> &adj_data, std::vector<uint32_t> &adj_indices,
                                std::vector<uint32_t> &adj_indptr, uint32_t idx_marker, uint32_t interleave_stride,
                                bool skip_empty_rows = false) {
    if (skip_empty_rows)
        util_pad_marker_end_of_row_skip_empty_rows(adj_data, adj_indices, adj_indptr, idx_marker, interleave_stride);
    else
        util_pad_marker_end_of_row_no_skip_empty_rows(adj_data, adj_indices, adj_indptr, idx_marker);
}

// sw/data_formatter.h:196-238
template <typename packed_val_t, typename packed_idx_t, uint32_t pack_size> struct CPSRMatrix {
    uint32_t num_row_partitions = 0;
    uint32_t num_col_partitions = 0;
    uint32_t num_hbm_channels = 0;
    bool skip_empty_rows = false;
    std::vector<std::vector<packed_val_t> > formatted_adj_data;
    std::vector<std::vector<packed_idx_t> > formatted_adj_indices;
    std::vector<std::vector<packed_idx_t> > formatted_adj_indptr;

    size_t flat(uint32_t j, uint32_t i, uint32_t c) const {
        return ((size_t)j * num_col_partitions + i) * num_hbm_channels + c;
    }
    std::vector<packed_val_t> get_packed_data(uint32_t j, uint32_t i, uint32_t c) { return formatted_adj_data[flat(j, i, c)]; }
    std::vector<packed_idx_t> get_packed_indices(uint32_t j, uint32_t i, uint32_t c) { return formatted_adj_indices[flat(j, i, c)]; }
    std::vector<packed_idx_t> get_packed_indptr(uint32_t j, uint32_t i, uint32_t c) { return formatted_adj_indptr[flat(j, i, c)]; }
};

// sw/data_formatter.h:256-313 : column partitioning, column ids rebased to the partition
template <typename DataT>
void util_convert_csr_to_dds(uint32_t num_rows, uint32_t num_cols, const DataT *adj_data, const uint32_t *adj_indices,
                             const uint32_t *adj_indptr, uint32_t num_cols_per_partition,
                             std::vector<DataT> partitioned_adj_data[], std::vector<uint32_t> partitioned_adj_indices[],
                             std::vector<uint32_t> partitioned_adj_indptr[]) {
    const uint32_t parts = (num_cols + num_cols_per_partition - 1) / num_cols_per_partition;
    for (uint32_t p = 0; p < parts; p++) {
        partitioned_adj_indptr[p].assign(num_rows + 1, 0);
        partitioned_adj_data[p].clear();
        partitioned_adj_indices[p].clear();
    }
    for (uint32_t r = 0; r < num_rows; r++) {
        for (uint32_t e = adj_indptr[r]; e < adj_indptr[r + 1]; e++) {
            uint32_t p = adj_indices[e] / num_cols_per_partition;
            partitioned_adj_data[p].push_back(adj_data[e]);
            partitioned_adj_indices[p].push_back(adj_indices[e] - p * num_cols_per_partition);
        }
        for (uint32_t p = 0; p < parts; p++) partitioned_adj_indptr[p][r + 1] = (uint32_t)partitioned_adj_data[p].size();
    }
}

// sw/data_formatter.h:336-365
template <typename DataT>
void util_reorder_rows_ascending_nnz(std::vector<DataT> const &adj_data, std::vector<uint32_t> const &adj_indices,
                                     std::vector<uint32_t> const &adj_indptr, std::vector<DataT> &reordered_adj_data,
                                     std::vector<uint32_t> &reordered_adj_indices,
                                     std::vector<uint32_t> &reordered_adj_indptr) {
    const uint32_t n = (uint32_t)adj_indptr.size() - 1;
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return adj_indptr[a + 1] - adj_indptr[a] < adj_indptr[b + 1] - adj_indptr[b];
    });
    reordered_adj_indptr.push_back(0);
    for (uint32_t r : order) {
        for (uint32_t e = adj_indptr[r]; e < adj_indptr[r + 1]; e++) {
            reordered_adj_data.push_back(adj_data[e]);
            reordered_adj_indices.push_back(adj_indices[e]);
        }
        reordered_adj_indptr.push_back((uint32_t)reordered_adj_data.size());
    }
}

// sw/data_formatter.h:384-446 : row r -> pack r/(C*P), channel (r/P)%C, lane r%P; packet n of a channel
// holds element n of each of its P lane streams; shorter lanes are zero padded
template <typename DataT, typename packed_val_t, typename packed_idx_t>
void util_pack_rows(std::vector<DataT> const &adj_data, std::vector<uint32_t> const &adj_indices,
                    std::vector<uint32_t> const &adj_indptr, uint32_t num_hbm_channels, uint32_t pack_size,
                    std::vector<packed_val_t> packed_adj_data[], std::vector<packed_idx_t> packed_adj_indices[],
                    std::vector<packed_idx_t> packed_adj_indptr[]) {
    const uint32_t n = (uint32_t)adj_indptr.size() - 1, stride = num_hbm_channels * pack_size;
    const uint32_t packs = (n + stride - 1) / stride;
    for (uint32_t c = 0; c < num_hbm_channels; c++) {
        packed_idx_t run;
        for (uint32_t l = 0; l < pack_size; l++) run.data[l] = 0;
        packed_adj_indptr[c].assign(1, run);
        for (uint32_t s = 0; s < packs; s++) {
            for (uint32_t l = 0; l < pack_size; l++) {
                uint32_t r = s * stride + c * pack_size + l;
                if (r < n) run.data[l] += adj_indptr[r + 1] - adj_indptr[r];
            }
            packed_adj_indptr[c].push_back(run);
        }
        uint32_t longest = 0;
        for (uint32_t l = 0; l < pack_size; l++) longest = std::max<uint32_t>(longest, run.data[l]);
        packed_val_t zero_v;
        packed_idx_t zero_i;
        for (uint32_t l = 0; l < pack_size; l++) { zero_v.data[l] = DataT(); zero_i.data[l] = 0; }
        packed_adj_data[c].assign(longest, zero_v);
        packed_adj_indices[c].assign(longest, zero_i);
        for (uint32_t l = 0; l < pack_size; l++) {
            uint32_t at = 0;
            for (uint32_t r = c * pack_size + l; r < n; r += stride)
                for (uint32_t e = adj_indptr[r]; e < adj_indptr[r + 1]; e++, at++) {
                    packed_adj_data[c][at].data[l] = adj_data[e];
                    packed_adj_indices[c][at].data[l] = adj_indices[e];
                }
        }
    }
}

// sw/data_formatter.h:468-544
template <typename packed_val_t, typename packed_idx_t, typename DataT, typename IndexT, uint32_t pack_size>
CPSRMatrix<packed_val_t, packed_idx_t, pack_size> csr2cpsr(CSRMatrix<DataT> const &csr_matrix, uint32_t idx_marker,
                                                          uint32_t out_buf_len, uint32_t vec_buf_len,
                                                          uint32_t num_hbm_channels, bool skip_empty_rows) {
    const uint32_t stride = pack_size * num_hbm_channels;
    if (csr_matrix.num_rows % stride != 0) {
        std::cout << "The number of rows of the sparse matrix should divide " << stride << ". "
                  << "Please use spmv::io::util_round_csr_matrix_dim. "
                  << "Exit!" << std::endl;
        exit(EXIT_FAILURE);
    }
    if (csr_matrix.num_cols % pack_size != 0) {
        std::cout << "The number of columns of the sparse matrix should divide " << pack_size << ". "
                  << "Please use spmv::io::util_round_csr_matrix_dim. "
                  << "Exit!" << std::endl;
        exit(EXIT_FAILURE);
    }
    assert(out_buf_len % stride == 0);
    assert(vec_buf_len % pack_size == 0);
    CPSRMatrix<packed_val_t, packed_idx_t, pack_size> out;
    out.skip_empty_rows = skip_empty_rows;
    out.num_hbm_channels = num_hbm_channels;
    out.num_row_partitions = (csr_matrix.num_rows + out_buf_len - 1) / out_buf_len;
    out.num_col_partitions = (csr_matrix.num_cols + vec_buf_len - 1) / vec_buf_len;
    const size_t blocks = (size_t)out.num_row_partitions * out.num_col_partitions * num_hbm_channels;
    out.formatted_adj_data.resize(blocks);
    out.formatted_adj_indices.resize(blocks);
    out.formatted_adj_indptr.resize(blocks);
    std::vector<std::vector<DataT> > pd(out.num_col_partitions);
    std::vector<std::vector<IndexT> > pi(out.num_col_partitions), pp(out.num_col_partitions);
    for (uint32_t j = 0; j < out.num_row_partitions; j++) {
        const uint32_t r0 = j * out_buf_len;
        const uint32_t nr = std::min<uint32_t>(out_buf_len, csr_matrix.num_rows - r0);
        const IndexT base = csr_matrix.adj_indptr[r0];
        std::vector<IndexT> slice(csr_matrix.adj_indptr.begin() + r0, csr_matrix.adj_indptr.begin() + r0 + nr + 1);
        for (auto &v : slice) v -= base;
        util_convert_csr_to_dds<DataT>(nr, csr_matrix.num_cols, csr_matrix.adj_data.data() + base,
                                       csr_matrix.adj_indices.data() + base, slice.data(), vec_buf_len, pd.data(),
                                       pi.data(), pp.data());
        for (uint32_t i = 0; i < out.num_col_partitions; i++) {
            util_pad_marker_end_of_row<DataT>(pd[i], pi[i], pp[i], idx_marker, stride, skip_empty_rows);
            const size_t at = out.flat(j, i, 0);
            util_pack_rows<DataT, packed_val_t, packed_idx_t>(pd[i], pi[i], pp[i], num_hbm_channels, pack_size,
                                                              &out.formatted_adj_data[at],
                                                              &out.formatted_adj_indices[at],
                                                              &out.formatted_adj_indptr[at]);
        }
    }
    return out;
}

// The per-HBM-channel packet images the kernels read (sw/host.cpp:163-231): for physical channel pc,
// `num_partitions * (1 + IF)` header packets (partition start * IF; lane lengths of each of the IF
// virtual channels pc + f*C), then the data packets of its IF virtual channels interleaved, every
// (partition, virtual channel) padded to the partition's longest.
template <typename mat_pkt_t, typename packed_val_t, typename packed_idx_t, uint32_t pack_size>
std::vector<std::vector<mat_pkt_t> > build_channel_images(CPSRMatrix<packed_val_t, packed_idx_t, pack_size> &cpsr,
                                                          uint32_t num_physical_channels, uint32_t interleave_factor) {
    const uint32_t IF = interleave_factor, C = num_physical_channels;
    const uint32_t np = cpsr.num_row_partitions * cpsr.num_col_partitions;
    assert(cpsr.num_hbm_channels == C * IF);
    std::vector<std::vector<mat_pkt_t> > images(C);
    for (uint32_t pc = 0; pc < C; pc++) {
        std::vector<uint32_t> longest(np, 0), start(np + 1, 0);
        for (uint32_t ij = 0; ij < np; ij++) {
            for (uint32_t f = 0; f < IF; f++)
                longest[ij] = std::max<uint32_t>(longest[ij], (uint32_t)cpsr.formatted_adj_indices[(size_t)ij * C * IF + pc + f * C].size());
            start[ij + 1] = start[ij] + longest[ij];
        }
        mat_pkt_t zero;
        std::memset(static_cast<void *>(&zero), 0, sizeof zero);
        auto &img = images[pc];
        img.assign((size_t)np * (1 + IF) + (size_t)start[np] * IF, zero);
        const size_t base = (size_t)np * (1 + IF);
        for (uint32_t ij = 0; ij < np; ij++) {
            img[(size_t)ij * (1 + IF)].indices.data[0] = start[ij] * IF;
            for (uint32_t f = 0; f < IF; f++) {
                const size_t blk = (size_t)ij * C * IF + pc + f * C;
                img[(size_t)ij * (1 + IF) + 1 + f].indices = cpsr.formatted_adj_indptr[blk].back();
                const auto &idx = cpsr.formatted_adj_indices[blk];
                const auto &val = cpsr.formatted_adj_data[blk];
                for (size_t n = 0; n < idx.size(); n++) {
                    mat_pkt_t &p = img[base + ((size_t)start[ij] + n) * IF + f];
                    p.indices = idx[n];
                    p.vals = val[n];
                }
            }
        }
    }
    return images;
}

}  // namespace io
}  // namespace spmv
#endif
