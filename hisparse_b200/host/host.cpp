// Test driver for the B200 SpMV engine, mirroring the reference's sw/host.cpp: same
// preprocessing calls (util_round_csr_matrix_dim, csr_matrix_convert_from_float, csr2cpsr, per-
// channel packet images), same kernel-invocation sequence per row partition, same compute_ref and
// verify (|y - ref| < 1e-4), same test names. The OpenCL calls are replaced by the C ABI of
// include/hisparse_b200.h. Build with -DFP_POB / -DFP_STALL for the float variants.
//
//   usage: host [device-index]
#include <assert.h>

#include <cmath>
#include <iomanip>
#include <iostream>

#include "common.h"
#include "data_formatter.h"
#include "data_loader.h"
#include "runtime.h"
#include "synthetic.h"

//--------------------------------------------------------------------------------------------------
// reference and verify utils (sw/host.cpp:33-86)
//--------------------------------------------------------------------------------------------------
void compute_ref(spmv::io::CSRMatrix<float> &mat, std::vector<float> &vector, std::vector<float> &ref_result) {
    ref_result.assign(mat.num_rows, 0.0f);
    for (size_t r = 0; r < mat.num_rows; r++)
        for (size_t i = mat.adj_indptr[r]; i < mat.adj_indptr[r + 1]; i++)
            ref_result[r] += mat.adj_data[i] * vector[mat.adj_indices[i]];
}

// The acceptance check of the reference's harness (sw/host.cpp:50-74): every result within 1e-4 (absolute) of the
// fp32 reference. Own body; the messages keep the reference's wording because scripts grep for them.
bool verify(const std::vector<float> &reference_results, const std::vector<VAL_T> &kernel_results) {
    const float epsilon = 0.0001f;
    const size_t n = reference_results.size();
    if (kernel_results.size() != n) {
        std::cout << "Error: Size mismatch" << std::endl
                  << "  Reference result size: " << n << "  Kernel result size: " << kernel_results.size() << std::endl;
        return false;
    }
    size_t first_bad = n;
    for (size_t i = 0; i < n && first_bad == n; i++)
        if (!(std::fabs(float(kernel_results[i]) - reference_results[i]) < epsilon)) first_bad = i;     // (NaN fails too)
    if (first_bad == n) return true;
    std::cout << "Error: Result mismatch" << std::endl
              << "  i = " << first_bad << "  Reference result = " << reference_results[first_bad]
              << "  Kernel result = " << kernel_results[first_bad] << std::endl;
    return false;
}

// packed result words (8 per packet, natural row order: sw/host.cpp:76-86) -> one value per row
void unpack_vector(const aligned_vector<PACKED_VAL_T> &packed, std::vector<VAL_T> &flat) {
    flat.clear();
    flat.reserve(packed.size() * PACK_SIZE);
    for (const PACKED_VAL_T &pkt : packed) flat.insert(flat.end(), pkt.data, pkt.data + PACK_SIZE);
}

//---------------------------------------------------------------
// test harness (sw/host.cpp:136-377)
//---------------------------------------------------------------
bool spmv_test_harness(hsb_runtime &runtime, spmv::io::CSRMatrix<float> &ext_matrix, bool skip_empty_rows) {
    using namespace spmv::io;
    std::cout << "INFO : Test started" << std::endl;
    util_round_csr_matrix_dim<float>(ext_matrix, PACK_SIZE * NUM_HBM_CHANNELS * INTERLEAVE_FACTOR, PACK_SIZE);
    CSRMatrix<VAL_T> mat = csr_matrix_convert_from_float<VAL_T>(ext_matrix);

    size_t num_row_partitions = (mat.num_rows + LOGICAL_OB_SIZE - 1) / LOGICAL_OB_SIZE;
    size_t num_col_partitions = (mat.num_cols + LOGICAL_VB_SIZE - 1) / LOGICAL_VB_SIZE;
    size_t num_partitions = num_row_partitions * num_col_partitions;
    size_t num_virtual_hbm_channels = NUM_HBM_CHANNELS * INTERLEAVE_FACTOR;
    CPSRMatrix<PACKED_VAL_T, PACKED_IDX_T, PACK_SIZE> cpsr_matrix =
        csr2cpsr<PACKED_VAL_T, PACKED_IDX_T, VAL_T, IDX_T, PACK_SIZE>(mat, IDX_MARKER, LOGICAL_OB_SIZE, LOGICAL_VB_SIZE,
                                                                      num_virtual_hbm_channels, skip_empty_rows);
    std::vector<std::vector<SPMV_MAT_PKT_T> > channel_packets =
        build_channel_images<SPMV_MAT_PKT_T>(cpsr_matrix, NUM_HBM_CHANNELS, INTERLEAVE_FACTOR);
    std::cout << "INFO : Matrix loading/preprocessing complete!" << std::endl;

    // input vector: rand() % 2 as in the reference
    std::vector<float> vector_f(ext_matrix.num_cols);
    std::generate(vector_f.begin(), vector_f.end(), [&]() { return float(rand() % 2); });
    aligned_vector<PACKED_VAL_T> vector(mat.num_cols / PACK_SIZE);
    for (size_t i = 0; i < vector.size(); i++)
        for (size_t k = 0; k < PACK_SIZE; k++) vector[i].data[k] = VAL_T(vector_f[i * PACK_SIZE + k]);
    aligned_vector<PACKED_VAL_T> result(mat.num_rows / PACK_SIZE);
    std::cout << "INFO : Input/result initialization complete!" << std::endl;

    // device buffers + transfers (sw/host.cpp:263-299)
    const void *ch[NUM_HBM_CHANNELS];
    size_t ch_packets[NUM_HBM_CHANNELS];
    for (size_t c = 0; c < NUM_HBM_CHANNELS; c++) {
        ch[c] = channel_packets[c].data();
        ch_packets[c] = channel_packets[c].size();
    }
    HSB_CHECK(hsb_upload_matrix_cpsr(runtime.ctx, ch, ch_packets, (unsigned)num_row_partitions,
                                     (unsigned)num_col_partitions, mat.num_rows, mat.num_cols));
    HSB_CHECK(hsb_upload_vector(runtime.ctx, vector.data(), mat.num_cols));
    std::cout << "INFO : Host -> Device data transfer complete!" << std::endl;

    // invoke kernel, one call per row partition (sw/host.cpp:329-358)
    std::cout << "INFO : Invoking kernel:" << std::endl;
    std::cout << "  row_partitions: " << num_row_partitions << std::endl;
    std::cout << "  col_partitions: " << num_col_partitions << std::endl;
    size_t rows_per_ch_in_last_row_part = (mat.num_rows % LOGICAL_OB_SIZE == 0)
                                              ? LOGICAL_OB_SIZE / NUM_HBM_CHANNELS
                                              : mat.num_rows % LOGICAL_OB_SIZE / NUM_HBM_CHANNELS;
    for (size_t row_part_id = 0; row_part_id < num_row_partitions; row_part_id++) {
        unsigned part_len = LOGICAL_OB_SIZE / NUM_HBM_CHANNELS;
        if (row_part_id == num_row_partitions - 1) part_len = rows_per_ch_in_last_row_part;
        HSB_CHECK(hsb_spmv_row_partition(runtime.ctx, (unsigned)row_part_id, part_len, (unsigned)num_col_partitions,
                                         (unsigned)num_partitions, mat.num_cols));
        HSB_CHECK(hsb_sync(runtime.ctx));
    }
    std::cout << "INFO : SpMV kernel complete!" << std::endl;

    std::vector<float> ref_result;
    compute_ref(ext_matrix, vector_f, ref_result);
    std::cout << "INFO : Compute reference complete!" << std::endl;

    HSB_CHECK(hsb_download_result(runtime.ctx, result.data(), mat.num_rows));
    std::cout << "INFO : Device -> Host data transfer complete!" << std::endl;
    std::vector<VAL_T> upk_result;
    unpack_vector(result, upk_result);
    return verify(ref_result, upk_result);
}

//---------------------------------------------------------------
// test cases (sw/host.cpp:438-530); datasets are optional (not shipped by the reference)
//---------------------------------------------------------------
std::string GRAPH_DATASET_DIR = "../datasets/graph/";
std::string NN_DATASET_DIR = "../datasets/pruned_nn/";

static bool run_case(hsb_runtime &runtime, const char *title, spmv::io::CSRMatrix<float> mat_f, bool skip) {
    std::cout << "------ Running test: " << title << std::endl;
    bool ok = spmv_test_harness(runtime, mat_f, skip);
    std::cout << (ok ? "INFO : Testcase passed." : "INFO : Testcase failed.") << std::endl;
    return ok;
}
bool test_basic(hsb_runtime &rt) { return run_case(rt, "on basic dense matrix", create_dense_CSR(128, 128), false); }
bool test_basic_sparse(hsb_runtime &rt) {
    return run_case(rt, "on basic sparse matrix", create_uniform_sparse_CSR(1000, 1024, 10), false);
}
bool test_medium_sparse(hsb_runtime &rt) {
    return run_case(rt, "on uniform 10K 10", create_uniform_sparse_CSR(10000, 10000, 10), false);
}
bool test_large_sparse(hsb_runtime &rt) {
    return run_case(rt, "on uniform 100K 10", create_uniform_sparse_CSR(100000, 100000, 10), false);
}
bool test_random_small_values(hsb_runtime &rt) {     // exercises product rounding, unlike the {0,1} reference inputs
    return run_case(rt, "on random 4096 x 4096, 1%", synth::random_CSR(4096, 4096, 167772, 0xC0FFEE01, 0.01f), true);
}
bool test_rmat(hsb_runtime &rt) {
    return run_case(rt, "on R-MAT 50K (power law, skip empty rows)", synth::rmat_CSR(50000, 1500000, 7, 0.001f), true);
}
static bool test_dataset(hsb_runtime &rt, const std::string &path, const char *title, bool skip) {
    std::ifstream probe(path);
    if (!probe) {
        std::cout << "------ Skipping test: " << title << " (" << path << " not present)" << std::endl;
        return true;
    }
    spmv::io::CSRMatrix<float> mat_f = spmv::io::load_csr_matrix_from_float_npz(path);
    for (auto &x : mat_f.adj_data) x = 1.0f / mat_f.num_cols;    // the reference's `1 / num_cols` is integer 0
    return run_case(rt, title, mat_f, skip);
}

int main(int argc, char **argv) {
    int device = argc > 1 ? atoi(argv[1]) : 0;
    hsb_runtime runtime(device, HSB_IMPL);
    std::cout << "INFO : Using " << hsb_version() << " on device " << device << std::endl;
    bool passed = true;
    passed = passed && test_basic(runtime);
    passed = passed && test_basic_sparse(runtime);
    passed = passed && test_medium_sparse(runtime);
    passed = passed && test_large_sparse(runtime);
    passed = passed && test_random_small_values(runtime);
    passed = passed && test_rmat(runtime);
    passed = passed && test_dataset(runtime, GRAPH_DATASET_DIR + "gplus_108K_13M_csr_float32.npz", "on google_plus", false);
    passed = passed && test_dataset(runtime, GRAPH_DATASET_DIR + "ogbl_ppa_576K_42M_csr_float32.npz", "on ogbl_ppa", false);
    passed = passed && test_dataset(runtime, NN_DATASET_DIR + "transformer_50_512_33288_csr_float32.npz", "on transformer-50-t", true);
    passed = passed && test_dataset(runtime, NN_DATASET_DIR + "transformer_95_512_33288_csr_float32.npz", "on transformer-95-t", true);
    std::cout << (passed ? "===== All Test Passed! =====" : "===== Test FAILED! =====") << std::endl;
    return passed ? 0 : 1;
}
