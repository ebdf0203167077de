// TEST INFRASTRUCTURE ONLY (oracle/_ref). Compiles the UNMODIFIED reference C simulation
//   /root/reference/spmv_csim/csim.cpp  (+ spmv/ or spmv-fp/ sources and sw/*.h it includes)
// against the clean-room HLS shim in oracle/shim/ and exposes a flat C ABI over it so the
// parity tests can call the reference's own code: its CSR->CPSR formatter, its dataflow
// `top_wrapper`, its CPU SpMV `compute_ref` and its end-to-end `spmv_test_harness`.
// Nothing here is product code; nothing in the product links against it.
//
// Build: see oracle/Makefile (one shared object per IMPL: fixed / float_pob / float_stall,
// mirroring spmv_csim/Makefile:27-38). The reference sources are #included from where they
// lie under /root/reference through forwarding headers generated at build time.
#define main reference_csim_main
#include "csim.cpp"   // resolved with -I/root/reference/spmv_csim
#undef main

#include <atomic>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <thread>

namespace {

inline uint32_t val_bits(const VAL_T &v) {
#if defined(FP_POB) || defined(FP_STALL)
    uint32_t b; std::memcpy(&b, &v, 4); return b;
#else
    return (uint32_t)(v(31, 0));
#endif
}

struct cpsr_handle {
    spmv::io::CPSRMatrix<PACKED_VAL_T, PACKED_IDX_T, PACK_SIZE> m;
    uint32_t rows, cols;
};

spmv::io::CSRMatrix<float> make_csr(uint32_t rows, uint32_t cols, const uint32_t *indptr,
                                    const uint32_t *indices, const float *data) {
    spmv::io::CSRMatrix<float> m;
    m.num_rows = rows;
    m.num_cols = cols;
    m.adj_indptr.assign(indptr, indptr + rows + 1);
    uint32_t nnz = indptr[rows];
    m.adj_indices.assign(indices, indices + nnz);
    m.adj_data.assign(data, data + nnz);
    return m;
}

struct cout_silencer {
    std::streambuf *old;
    std::ostringstream sink;
    cout_silencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~cout_silencer() { std::cout.rdbuf(old); }
};

}  // namespace

extern "C" {

// [0]=PACK_SIZE [1]=NUM_HBM_CHANNELS [2]=INTERLEAVE_FACTOR [3]=LOGICAL_OB_SIZE
// [4]=LOGICAL_VB_SIZE [5]=sizeof(SPMV_MAT_PKT_T) [6]=sizeof(VAL_T) [7]=impl (0 fixed,1 pob,2 stall)
void ref_config(unsigned out[8]) {
    out[0] = PACK_SIZE; out[1] = NUM_HBM_CHANNELS; out[2] = INTERLEAVE_FACTOR;
    out[3] = LOGICAL_OB_SIZE; out[4] = LOGICAL_VB_SIZE;
    out[5] = sizeof(SPMV_MAT_PKT_T); out[6] = sizeof(VAL_T);
#if defined(FP_POB)
    out[7] = 1;
#elif defined(FP_STALL)
    out[7] = 2;
#else
    out[7] = 0;
#endif
}

// float -> VAL_T, element-wise, exactly as sw/data_loader.h:76-84 does it (std::copy).
void ref_val_from_float(const float *in, size_t n, uint32_t *out_bits) {
    for (size_t i = 0; i < n; i++) { VAL_T v = in[i]; out_bits[i] = val_bits(v); }
}

// VAL_T product-accumulate chain exactly as spmv/libfpga/pe.h:64,72 (fixed) or
// pe-pob.h:64-66 (float): acc = acc + a*b, for i in order. Used to pin the closed form.
uint32_t ref_mac_chain(const uint32_t *a_bits, const uint32_t *b_bits, size_t n) {
    VAL_T acc = 0;
    for (size_t i = 0; i < n; i++) {
        VAL_T a, b;
#if defined(FP_POB) || defined(FP_STALL)
        std::memcpy(&a, &a_bits[i], 4); std::memcpy(&b, &b_bits[i], 4);
#else
        a(31, 0) = a_bits[i]; b(31, 0) = b_bits[i];
#endif
        VAL_T incr = a * b;
        VAL_T nq = acc + incr;
        acc = nq;
    }
    return val_bits(acc);
}

// Reference formatter: util_round_csr_matrix_dim + csr_matrix_convert_from_float + csr2cpsr
// with the compiled-in PACK_SIZE / VAL_T (spmv_csim/csim.cpp:213-227).
void *ref_csr2cpsr(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                   const float *data, uint32_t row_divisor, uint32_t col_divisor,
                   uint32_t out_buf_len, uint32_t vec_buf_len, uint32_t num_channels, int skip_empty_rows) {
    using namespace spmv::io;
    CSRMatrix<float> mf = make_csr(rows, cols, indptr, indices, data);
    util_round_csr_matrix_dim<float>(mf, row_divisor, col_divisor);
    CSRMatrix<VAL_T> mat = csr_matrix_convert_from_float<VAL_T>(mf);
    cpsr_handle *h = new cpsr_handle;
    h->rows = mat.num_rows; h->cols = mat.num_cols;
    h->m = csr2cpsr<PACKED_VAL_T, PACKED_IDX_T, VAL_T, IDX_T, PACK_SIZE>(
        mat, IDX_MARKER, out_buf_len, vec_buf_len, num_channels, skip_empty_rows != 0);
    return h;
}
void ref_cpsr_dims(void *hp, uint32_t out[4]) {
    cpsr_handle *h = (cpsr_handle *)hp;
    out[0] = h->rows; out[1] = h->cols; out[2] = h->m.num_row_partitions; out[3] = h->m.num_col_partitions;
}
size_t ref_cpsr_len(void *hp, uint32_t j, uint32_t i, uint32_t c) {
    return ((cpsr_handle *)hp)->m.get_packed_indices(j, i, c).size();
}
// idx/val: len*PACK_SIZE words (packet-major); lens: PACK_SIZE words = final per-lane stream lengths
void ref_cpsr_get(void *hp, uint32_t j, uint32_t i, uint32_t c, uint32_t *idx, uint32_t *val, uint32_t *lens) {
    cpsr_handle *h = (cpsr_handle *)hp;
    auto ind = h->m.get_packed_indices(j, i, c);
    auto dat = h->m.get_packed_data(j, i, c);
    auto ptr = h->m.get_packed_indptr(j, i, c);
    for (size_t n = 0; n < ind.size(); n++)
        for (unsigned k = 0; k < PACK_SIZE; k++) {
            idx[n * PACK_SIZE + k] = ind[n].data[k];
            val[n * PACK_SIZE + k] = val_bits(dat[n].data[k]);
        }
    for (unsigned k = 0; k < PACK_SIZE; k++) lens[k] = ptr.back().data[k];
}
void ref_cpsr_free(void *hp) { delete (cpsr_handle *)hp; }

// The reference dataflow pipeline, one row partition (spmv_csim/csim.cpp:22-136).
// ch[c] points at NUM_HBM_CHANNELS channel images of 64-byte packets; x / y are packed
// raw 32-bit words (Q8.24 bits or IEEE bits). sizeof(VAL_T)==4 in the shim, so the images
// are bit-identical to what sw/host.cpp:163-231 hands to the FPGA.
int ref_top_wrapper(const void *const *ch, const void *x, void *y, unsigned row_part_id,
                    unsigned part_len, unsigned num_col_partitions, unsigned num_partitions,
                    unsigned num_cols) {
    static_assert(sizeof(SPMV_MAT_PKT_T) == 64, "packet must be 64 bytes");
    static_assert(NUM_HBM_CHANNELS == 16, "16 channels");
    const SPMV_MAT_PKT_T *c[16];
    for (int i = 0; i < 16; i++) c[i] = (const SPMV_MAT_PKT_T *)ch[i];
    top_wrapper(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[9], c[10], c[11], c[12],
                c[13], c[14], c[15], (const PACKED_VAL_T *)x, (PACKED_VAL_T *)y, row_part_id,
                part_len, num_col_partitions, num_partitions, num_cols);
    return 0;
}

// The reference CPU SpMV (spmv_csim/csim.cpp:143-158 == sw/host.cpp:33-48).
void ref_compute_ref(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                     const float *data, const float *x, float *y) {
    spmv::io::CSRMatrix<float> m = make_csr(rows, cols, indptr, indices, data);
    std::vector<float> xv(x, x + cols), yv;
    compute_ref(m, xv, yv);
    std::memcpy(y, yv.data(), sizeof(float) * rows);
}
// Same loop timed in place (no marshalling inside the timed region): returns seconds per run.
double ref_time_compute_ref(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                            const float *data, const float *x, float *y, int runs) {
    spmv::io::CSRMatrix<float> m = make_csr(rows, cols, indptr, indices, data);
    std::vector<float> xv(x, x + cols), yv;
    compute_ref(m, xv, yv);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < runs; r++) compute_ref(m, xv, yv);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    std::memcpy(y, yv.data(), sizeof(float) * rows);
    return ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)) / runs;
}

// The same reference loop on all host threads: the rows are cut into `threads` contiguous blocks of
// about equal non-zeros and every thread runs the UNMODIFIED compute_ref on its block's CSR (x shared
// by value per thread, as compute_ref takes it). Wall time from a common start to the last finisher.
double ref_time_compute_ref_mt(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                               const float *data, const float *x, float *y, int runs, int threads) {
    if (threads < 1) threads = 1;
    std::vector<uint32_t> cut(threads + 1, rows);
    cut[0] = 0;
    for (int t = 1; t < threads; t++) {
        const uint64_t target = (uint64_t)indptr[rows] * t / threads;
        uint32_t lo = cut[t - 1], hi = rows;
        while (lo < hi) { uint32_t mid = lo + (hi - lo) / 2; if (indptr[mid] < target) lo = mid + 1; else hi = mid; }
        cut[t] = lo;
    }
    std::vector<spmv::io::CSRMatrix<float>> blocks(threads);
    std::vector<std::vector<float>> xs(threads, std::vector<float>(x, x + cols)), ys(threads);
    for (int t = 0; t < threads; t++) {
        const uint32_t r0 = cut[t], r1 = cut[t + 1], e0 = indptr[r0];
        std::vector<uint32_t> ip(r1 - r0 + 1);
        for (uint32_t r = r0; r <= r1; r++) ip[r - r0] = indptr[r] - e0;
        blocks[t] = make_csr(r1 - r0, cols, ip.data(), indices + e0, data + e0);
        compute_ref(blocks[t], xs[t], ys[t]);                       // warm
    }
    std::atomic<int> ready(0);
    std::atomic<bool> go(false);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t] {
            ready++;
            while (!go.load(std::memory_order_acquire)) {}
            for (int r = 0; r < runs; r++) compute_ref(blocks[t], xs[t], ys[t]);
        });
    while (ready.load() < threads) {}
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    go.store(true, std::memory_order_release);
    for (auto &th : pool) th.join();
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (int t = 0; t < threads; t++)
        if (cut[t + 1] > cut[t]) std::memcpy(y + cut[t], ys[t].data(), sizeof(float) * (cut[t + 1] - cut[t]));
    return ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)) / runs;
}

// The reference's own end-to-end test harness (format -> channel images -> top_wrapper ->
// compute_ref -> verify 1e-4), spmv_csim/csim.cpp:203-381. x comes from rand()%2 inside.
int ref_test_harness(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                     const float *data, int skip_empty_rows, unsigned seed) {
    spmv::io::CSRMatrix<float> m = make_csr(rows, cols, indptr, indices, data);
    srand(seed);
    cout_silencer quiet;
    return spmv_test_harness(m, skip_empty_rows != 0) ? 1 : 0;
}

// The reference's three synthetic self-tests (spmv_csim/csim.cpp:443-479).
int ref_selftest(void) {
    cout_silencer quiet;
    bool ok = test_basic() && test_basic_sparse() && test_large_sparse();
    return ok ? 1 : 0;
}

}  // extern "C"
