// TEST INFRASTRUCTURE ONLY (oracle/_ref). The reference's host-side formatter
// (/root/reference/sw/data_formatter.h + data_loader.h, unmodified) instantiated with the
// small shapes that the reference's own known-answer tests use
// (unit_tests/test_io.cpp:206-398: pack_size 2, 1-2 channels), so the golden vectors
// written there can be checked against (a) the reference code itself and (b) our restatement.
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <vector>
#include "data_loader.h"
#include "data_formatter.h"

namespace {
const uint32_t P = 2;
template <class T> struct pk { T data[P]; };

template <class DataT>
void *fmt(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices, const DataT *data,
          uint32_t out_buf_len, uint32_t vec_buf_len, uint32_t num_channels, int skip) {
    using namespace spmv::io;
    CSRMatrix<DataT> m;
    m.num_rows = rows; m.num_cols = cols;
    m.adj_indptr.assign(indptr, indptr + rows + 1);
    m.adj_indices.assign(indices, indices + indptr[rows]);
    m.adj_data.assign(data, data + indptr[rows]);
    auto *out = new CPSRMatrix<pk<DataT>, pk<uint32_t>, P>;
    *out = csr2cpsr<pk<DataT>, pk<uint32_t>, DataT, uint32_t, P>(m, 0xFFFFFFFFu, out_buf_len, vec_buf_len,
                                                                num_channels, skip != 0);
    return out;
}
template <class DataT>
size_t get(void *hp, uint32_t j, uint32_t i, uint32_t c, uint32_t *idx, DataT *val, uint32_t *indptr, size_t *n_indptr) {
    auto *h = (spmv::io::CPSRMatrix<pk<DataT>, pk<uint32_t>, P> *)hp;
    auto ind = h->get_packed_indices(j, i, c);
    auto dat = h->get_packed_data(j, i, c);
    auto ptr = h->get_packed_indptr(j, i, c);
    if (idx) for (size_t n = 0; n < ind.size(); n++) for (uint32_t k = 0; k < P; k++) {
        idx[n * P + k] = ind[n].data[k]; val[n * P + k] = dat[n].data[k];
    }
    if (indptr) for (size_t n = 0; n < ptr.size(); n++) for (uint32_t k = 0; k < P; k++) indptr[n * P + k] = ptr[n].data[k];
    *n_indptr = ptr.size();
    return ind.size();
}
}  // namespace

extern "C" {
void *ref2_csr2cpsr_i32(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices, const int32_t *data,
                        uint32_t ob, uint32_t vb, uint32_t nch, int skip) { return fmt<int32_t>(rows, cols, indptr, indices, data, ob, vb, nch, skip); }
void *ref2_csr2cpsr_f32(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices, const float *data,
                        uint32_t ob, uint32_t vb, uint32_t nch, int skip) { return fmt<float>(rows, cols, indptr, indices, data, ob, vb, nch, skip); }
size_t ref2_get_i32(void *h, uint32_t j, uint32_t i, uint32_t c, uint32_t *idx, int32_t *val, uint32_t *indptr, size_t *n_indptr) { return get<int32_t>(h, j, i, c, idx, val, indptr, n_indptr); }
size_t ref2_get_f32(void *h, uint32_t j, uint32_t i, uint32_t c, uint32_t *idx, float *val, uint32_t *indptr, size_t *n_indptr) { return get<float>(h, j, i, c, idx, val, indptr, n_indptr); }

// util_pack_rows (sw/data_formatter.h:384-446) on a plain CSR, pack_size 2 (test_io.cpp:206-245)
size_t ref2_pack_rows_f32(uint32_t rows, const uint32_t *indptr, const uint32_t *indices, const float *data,
                          uint32_t nch, uint32_t c, uint32_t *idx, float *val, uint32_t *iptr, size_t *n_indptr) {
    std::vector<float> d(data, data + indptr[rows]);
    std::vector<uint32_t> ii(indices, indices + indptr[rows]), ip(indptr, indptr + rows + 1);
    std::vector<std::vector<pk<float>>> pd(nch);
    std::vector<std::vector<pk<uint32_t>>> pi(nch), pp(nch);
    spmv::io::util_pack_rows<float, pk<float>, pk<uint32_t>>(d, ii, ip, nch, P, pd.data(), pi.data(), pp.data());
    for (size_t n = 0; n < pi[c].size(); n++) for (uint32_t k = 0; k < P; k++) { idx[n*P+k] = pi[c][n].data[k]; val[n*P+k] = pd[c][n].data[k]; }
    for (size_t n = 0; n < pp[c].size(); n++) for (uint32_t k = 0; k < P; k++) iptr[n*P+k] = pp[c][n].data[k];
    *n_indptr = pp[c].size();
    return pi[c].size();
}
// util_convert_csr_to_dds (sw/data_formatter.h:256-313), test_io.cpp:143-174
void ref2_csr_to_dds_f32(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices, const float *data,
                         uint32_t cols_per_part, uint32_t part, float *odata, uint32_t *oidx, uint32_t *optr, uint32_t *onnz) {
    uint32_t ncp = (cols + cols_per_part - 1) / cols_per_part;
    std::vector<std::vector<float>> pd(ncp);
    std::vector<std::vector<uint32_t>> pi(ncp), pp(ncp);
    spmv::io::util_convert_csr_to_dds<float>(rows, cols, data, indices, indptr, cols_per_part, pd.data(), pi.data(), pp.data());
    for (size_t n = 0; n < pd[part].size(); n++) { odata[n] = pd[part][n]; oidx[n] = pi[part][n]; }
    for (size_t n = 0; n < pp[part].size(); n++) optr[n] = pp[part][n];
    *onnz = pd[part].size();
}
void ref2_round_dims(uint32_t *rows, uint32_t *cols, uint32_t rd, uint32_t cd) {
    spmv::io::CSRMatrix<float> m; m.num_rows = *rows; m.num_cols = *cols; m.adj_indptr.assign(*rows + 1, 0);
    spmv::io::util_round_csr_matrix_dim(m, rd, cd);
    *rows = m.num_rows; *cols = m.num_cols;
}
void ref2_free_i32(void *h) { delete (spmv::io::CPSRMatrix<pk<int32_t>, pk<uint32_t>, P> *)h; }
void ref2_free_f32(void *h) { delete (spmv::io::CPSRMatrix<pk<float>, pk<uint32_t>, P> *)h; }
}
