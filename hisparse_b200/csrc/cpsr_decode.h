// Decoder for the reference's per-HBM-channel CPSR packet images (sw/host.cpp:163-231): the
// inverse of csr2cpsr + channel layout, i.e. what CPSR_matrix_loader does on the FPGA
// (spmv/libfpga/spmv_cluster.h:34-107; float: spmv-fp/libfpga/spmv_cluster.h:39-129) followed by
// the row numbering of the result path (pe.h:95-116, stream_utils.h:36-75,
// spmv_result_drain.cpp:36-113). Used by hsb_upload_matrix_cpsr so that buffers produced by the
// reference's own host code can be handed to this engine unchanged.
#ifndef HISPARSE_B200_CPSR_DECODE_H_
#define HISPARSE_B200_CPSR_DECODE_H_

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace hsb {

struct ImplConfig {
    uint32_t interleave;     // INTERLEAVE_FACTOR
    uint32_t ob_size;        // LOGICAL_OB_SIZE
    uint32_t vb_size;        // LOGICAL_VB_SIZE
    bool marker_is_q824;     // marker value in bits 31..24 (fixed) or raw integer bits (float)
};
bool impl_config(int impl, ImplConfig *out);

struct HostCsr {
    uint32_t rows = 0, cols = 0;
    std::vector<uint32_t> indptr, indices, vals;
};

// Number of packets a channel image must hold, recovered from its header packets.
size_t cpsr_image_packets(const ImplConfig &cfg, const uint32_t *image, uint32_t num_partitions);

// Decode row partitions [part_begin, part_end) of the 16 images into a CSR whose row 0 is the
// first row of partition part_begin. rows_in_part[j - part_begin] = rows of partition j (a multiple
// of 128). Column ids are global. n_packets may be null (no bounds checking of the images).
bool cpsr_decode(const ImplConfig &cfg, const uint32_t *const images[16], const size_t *n_packets,
                 uint32_t num_row_partitions, uint32_t num_col_partitions, uint32_t part_begin,
                 uint32_t part_end, const uint32_t *rows_in_part, uint32_t num_cols, HostCsr *out,
                 std::string *err);

}  // namespace hsb

#ifdef __CUDACC__
#include <cuda_runtime.h>
namespace hsb {
// The same decoding done on the GPU (cpsr_decode_gpu.cu): host images in, device COO list (row, column, value per
// non-zero, image order) out (caller frees).
cudaError_t cpsr_decode_gpu(const ImplConfig &cfg, const uint32_t *const images[16], const size_t n_packets[16],
                            uint32_t num_row_partitions, uint32_t num_col_partitions, uint32_t num_rows,
                            uint32_t num_cols, cudaStream_t stream, uint32_t **d_rows, uint32_t **d_cols,
                            uint32_t **d_vals, uint64_t *nnz, std::string *err);
}  // namespace hsb
#endif
#endif
