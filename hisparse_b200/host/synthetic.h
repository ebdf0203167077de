// Synthetic CSR generators for the host drivers: the reference's two (sw/host.cpp:382-430) plus a
// seeded R-MAT / uniform-random generator standing in for the datasets the reference downloads
// (datasets/download.sh) -- counter-based splitmix64, so results do not depend on thread count.
#ifndef HISPARSE_B200_HOST_SYNTHETIC_H_
#define HISPARSE_B200_HOST_SYNTHETIC_H_

#include <algorithm>
#include <cstdint>
#include <vector>

#include "data_loader.h"

inline spmv::io::CSRMatrix<float> create_dense_CSR(unsigned num_rows, unsigned num_cols) {
    spmv::io::CSRMatrix<float> m;
    m.num_rows = num_rows; m.num_cols = num_cols;
    m.adj_data.assign((size_t)num_rows * num_cols, 1.0f);
    m.adj_indices.resize((size_t)num_rows * num_cols);
    m.adj_indptr.resize(num_rows + 1);
    for (size_t i = 0; i < num_rows; i++)
        for (size_t j = 0; j < num_cols; j++) m.adj_indices[i * num_cols + j] = (uint32_t)j;
    for (size_t i = 0; i <= num_rows; i++) m.adj_indptr[i] = (uint32_t)(num_cols * i);
    return m;
}

inline spmv::io::CSRMatrix<float> create_uniform_sparse_CSR(unsigned num_rows, unsigned num_cols, unsigned nnz_per_row) {
    spmv::io::CSRMatrix<float> m;
    m.num_rows = num_rows; m.num_cols = num_cols;
    m.adj_data.assign((size_t)num_rows * nnz_per_row, 1.0f);
    m.adj_indices.resize((size_t)num_rows * nnz_per_row);
    m.adj_indptr.resize(num_rows + 1);
    const unsigned step = num_cols / nnz_per_row;
    for (size_t i = 0; i < num_rows; i++)
        for (size_t j = 0; j < nnz_per_row; j++) m.adj_indices[i * nnz_per_row + j] = (uint32_t)((step * j + i) % num_cols);
    for (size_t i = 0; i <= num_rows; i++) m.adj_indptr[i] = (uint32_t)(nnz_per_row * i);
    return m;
}

namespace synth {
inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline float u01(uint64_t h) { return (float)((h >> 40) * (1.0 / 16777216.0)); }

// edges -> CSR with sorted, unique columns per row; values U(0,1) * scale
inline spmv::io::CSRMatrix<float> from_edges(uint32_t rows, uint32_t cols, std::vector<uint64_t> &keys, uint64_t seed, float scale) {
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    spmv::io::CSRMatrix<float> m;
    m.num_rows = rows; m.num_cols = cols;
    m.adj_indptr.assign((size_t)rows + 1, 0);
    m.adj_indices.resize(keys.size());
    m.adj_data.resize(keys.size());
    for (size_t e = 0; e < keys.size(); e++) {
        uint32_t r = (uint32_t)(keys[e] / cols), c = (uint32_t)(keys[e] % cols);
        m.adj_indptr[r + 1]++;
        m.adj_indices[e] = c;
        m.adj_data[e] = scale * u01(splitmix64(seed ^ (keys[e] * 0x2545F4914F6CDD1Dull)));
    }
    for (uint32_t r = 0; r < rows; r++) m.adj_indptr[r + 1] += m.adj_indptr[r];
    return m;
}

inline spmv::io::CSRMatrix<float> random_CSR(uint32_t rows, uint32_t cols, uint64_t nnz, uint64_t seed, float scale = 1.0f) {
    std::vector<uint64_t> keys(nnz);
    for (uint64_t e = 0; e < nnz; e++) {
        uint64_t h = splitmix64(seed + e);
        keys[e] = (uint64_t)((h >> 32) % rows) * cols + (uint32_t)h % cols;
    }
    return from_edges(rows, cols, keys, seed, scale);
}

// R-MAT (a, b, c, d) = (.57, .19, .19, .05); ids scrambled by a multiplicative permutation
inline spmv::io::CSRMatrix<float> rmat_CSR(uint32_t n, uint64_t edges, uint64_t seed, float scale = 1.0f) {
    int levels = 0;
    while ((1ull << levels) < n) levels++;
    std::vector<uint64_t> keys;
    keys.reserve(edges);
    uint64_t ctr = 0;
    while (keys.size() < edges) {
        uint64_t r = 0, c = 0;
        for (int l = 0; l < levels; l++) {
            float u = u01(splitmix64(seed * 0x100000001B3ull + (ctr++)));
            r = (r << 1) | (u >= 0.76f);
            c = (c << 1) | ((u >= 0.57f && u < 0.76f) || u >= 0.95f);
        }
        if (r >= n || c >= n) continue;
        r = (r * 0x9E3779B1ull + 12345u) % n;            // scramble hubs (bijective when gcd(mult, n) == 1 is not required for a stand-in)
        c = (c * 0x85EBCA77ull + 6789u) % n;
        keys.push_back(r * n + c);
    }
    return from_edges(n, n, keys, seed, scale);
}
}  // namespace synth
#endif
