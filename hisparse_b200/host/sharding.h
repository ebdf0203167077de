// Row-block sharding of a CSR across GPUs (the C++ twin of hisparse_b200/sharding.py):
// rank g holds rows [bounds[g], bounds[g+1]) -- contiguous blocks balanced by non-zeros, boundaries on the
// reference's row granularity (PACK_SIZE * NUM_HBM_CHANNELS = 128 rows, sw/data_formatter.h:475) -- plus a
// full replica of x, and produces y for its block.
#ifndef HISPARSE_B200_HOST_SHARDING_H_
#define HISPARSE_B200_HOST_SHARDING_H_

#include <algorithm>
#include <cstdint>
#include <vector>

#include "data_loader.h"

namespace spmv {
namespace shard {

// nnz-balanced row boundaries, multiples of `granularity` (the last one = rows): world + 1 entries
inline std::vector<uint32_t> shard_bounds(const std::vector<uint32_t> &indptr, int world, uint32_t granularity = 128) {
    const uint32_t rows = (uint32_t)indptr.size() - 1;
    const uint64_t nnz = indptr[rows];
    std::vector<uint32_t> bounds(1, 0u);
    for (int g = 1; g < world; g++) {
        const uint64_t target = nnz * (uint64_t)g / (uint64_t)world;
        uint32_t r = (uint32_t)(std::lower_bound(indptr.begin(), indptr.end(), (uint32_t)target) - indptr.begin());
        r = std::min<uint64_t>(rows, ((uint64_t)r + granularity / 2) / granularity * granularity);
        bounds.push_back(std::max(bounds.back(), r));
    }
    bounds.push_back(rows);
    return bounds;
}

// CSR of rows [r0, r1) with indptr rebased to 0 (all columns kept)
template <typename T>
io::CSRMatrix<T> extract_shard(const io::CSRMatrix<T> &m, uint32_t r0, uint32_t r1) {
    io::CSRMatrix<T> s;
    s.num_rows = r1 - r0;
    s.num_cols = m.num_cols;
    const uint32_t e0 = m.adj_indptr[r0], e1 = m.adj_indptr[r1];
    s.adj_indptr.resize(s.num_rows + 1);
    for (uint32_t r = r0; r <= r1; r++) s.adj_indptr[r - r0] = m.adj_indptr[r] - e0;
    s.adj_indices.assign(m.adj_indices.begin() + e0, m.adj_indices.begin() + e1);
    s.adj_data.assign(m.adj_data.begin() + e0, m.adj_data.begin() + e1);
    return s;
}

}  // namespace shard
}  // namespace spmv
#endif
