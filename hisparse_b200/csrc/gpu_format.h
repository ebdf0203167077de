// GPU-side builder of the tile-stream format (see gpu_format.cu, tile_format.h).
#ifndef HISPARSE_B200_GPU_FORMAT_H_
#define HISPARSE_B200_GPU_FORMAT_H_

#include <cuda_runtime.h>

#include <string>

#include "tile_format.h"

namespace hsb {

struct DeviceFormat {
    uint32_t *vals = nullptr;        // n_elems words
    uint16_t *cols = nullptr;        // n_elems stored column ids
    uint32_t *slice_rows = nullptr;  // n_slices * 32
    uint64_t n_elems = 0, n_slices = 0, n_streams = 0;
};

// CSR (d_indptr; d_coo_rows null) or COO (d_coo_rows: the row of every non-zero, any order; d_indptr null) in device
// memory -> tile streams in device memory. `meta` receives everything the host-side
// launch planner needs (tiles, slice list, partition table, counts); its vals / cols16 / slice_rows
// vectors stay empty. Synchronises `stream`.
cudaError_t build_tiled_gpu(uint32_t rows, uint32_t cols, uint64_t nnz, const uint32_t *d_indptr, const uint32_t *d_coo_rows,
                            const uint32_t *d_indices, const uint32_t *d_vals, uint32_t rows_per_part,
                            uint32_t tile_cols, cudaStream_t stream, TiledMatrix *meta, DeviceFormat *out,
                            std::string *err);

}  // namespace hsb
#endif
