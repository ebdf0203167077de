// CPU-only self-check of the host formatter (data_formatter.h) used by tests/test_host_cpp.py:
// prints, for a CSR read from stdin, every CPSR block and the 16 channel images as plain numbers so
// the Python test can compare them with the reference's golden vectors and with the oracle.
//   stdin : rows cols nnz P C IF OB VB skip kind(0 int,1 float,2 q8.24)  indptr...  indices...  data(words)...
#include <cstdio>
#include <iostream>

#include "data_formatter.h"
#include "fixed_point.h"

template <uint32_t P> struct pi { uint32_t data[P]; };
template <class T, uint32_t P> struct pv { T data[P]; };
template <class T, uint32_t P> struct pkt { pi<P> indices; pv<T, P> vals; };

template <class T> uint32_t bits(const T &v) { uint32_t b = 0; std::memcpy(&b, &v, 4); return b; }

template <class T, uint32_t P> int run(uint32_t rows, uint32_t cols, uint32_t C, uint32_t IF, uint32_t OB, uint32_t VB, bool skip,
                                       std::vector<uint32_t> &ip, std::vector<uint32_t> &ix, std::vector<uint32_t> &words) {
    spmv::io::CSRMatrix<T> m;
    m.num_rows = rows; m.num_cols = cols; m.adj_indptr = ip; m.adj_indices = ix;
    for (uint32_t w : words) { T v; std::memcpy(static_cast<void *>(&v), &w, 4); m.adj_data.push_back(v); }
    auto cp = spmv::io::csr2cpsr<pv<T, P>, pi<P>, T, uint32_t, P>(m, 0xFFFFFFFFu, OB, VB, C * IF, skip);
    std::printf("%u %u\n", cp.num_row_partitions, cp.num_col_partitions);
    for (uint32_t j = 0; j < cp.num_row_partitions; j++)
        for (uint32_t i = 0; i < cp.num_col_partitions; i++)
            for (uint32_t c = 0; c < C * IF; c++) {
                auto d = cp.get_packed_data(j, i, c);
                auto x = cp.get_packed_indices(j, i, c);
                auto p = cp.get_packed_indptr(j, i, c);
                std::printf("B %u %u %u %zu %zu\n", j, i, c, x.size(), p.size());
                for (size_t n = 0; n < x.size(); n++) for (uint32_t l = 0; l < P; l++) std::printf("%u %u ", x[n].data[l], bits(d[n].data[l]));
                std::printf("\n");
                for (size_t n = 0; n < p.size(); n++) for (uint32_t l = 0; l < P; l++) std::printf("%u ", p[n].data[l]);
                std::printf("\n");
            }
    if (P == 8) {
        auto img = spmv::io::build_channel_images<pkt<T, P> >(cp, C, IF);
        for (uint32_t c = 0; c < C; c++) {
            std::printf("I %u %zu\n", c, img[c].size());
            for (auto &q : img[c]) { for (uint32_t l = 0; l < P; l++) std::printf("%u ", q.indices.data[l]); for (uint32_t l = 0; l < P; l++) std::printf("%u ", bits(q.vals.data[l])); }
            std::printf("\n");
        }
    }
    return 0;
}

int main() {
    uint32_t rows, cols, nnz, P, C, IF, OB, VB, skip, kind;
    std::cin >> rows >> cols >> nnz >> P >> C >> IF >> OB >> VB >> skip >> kind;
    std::vector<uint32_t> ip(rows + 1), ix(nnz), w(nnz);
    for (auto &v : ip) std::cin >> v;
    for (auto &v : ix) std::cin >> v;
    for (auto &v : w) std::cin >> v;
    if (P == 2 && kind == 0) return run<int32_t, 2>(rows, cols, C, IF, OB, VB, skip, ip, ix, w);
    if (P == 2 && kind == 1) return run<float, 2>(rows, cols, C, IF, OB, VB, skip, ip, ix, w);
    if (P == 8 && kind == 1) return run<float, 8>(rows, cols, C, IF, OB, VB, skip, ip, ix, w);
    if (P == 8 && kind == 2) return run<spmv::ufixed_q8_24, 8>(rows, cols, C, IF, OB, VB, skip, ip, ix, w);
    return 2;
}
