// CPU-only check of the .npz reader (data_loader.h): prints dims, nnz and simple checksums.
#include <cstdio>
#include "data_loader.h"
int main(int argc, char **argv) {
    if (argc < 2) return 2;
    try {
        auto m = spmv::io::load_csr_matrix_from_float_npz(argv[1]);
        double sd = 0; unsigned long long si = 0, sp = 0;
        for (float v : m.adj_data) sd += v;
        for (uint32_t v : m.adj_indices) si += v;
        for (uint32_t v : m.adj_indptr) sp += v;
        std::printf("%u %u %zu %.6f %llu %llu\n", m.num_rows, m.num_cols, m.adj_data.size(), sd, si, sp);
    } catch (const std::exception &e) {
        std::printf("ERROR %s\n", e.what());
        return 1;
    }
    return 0;
}
