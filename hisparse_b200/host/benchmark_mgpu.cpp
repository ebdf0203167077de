// Multi-GPU timing driver: the shape of the reference's sw/benchmark.cpp:355-416 (load the dataset, preprocess,
// upload once, NUM_RUNS back-to-back SpMVs, print `{Preprocessing | SpMV | GBPS | GOPS}`) with the matrix cut
// into nnz-balanced row blocks, ONE PROCESS PER GPU:
//
//   usage: benchmark_mgpu <dataset.npz | dense:R:C | uniform:R:C:K | random:R:C:NNZ:SEED | rmat:N:EDGES:SEED> <n_gpus>
//
// The parent draws an NCCL unique id and forks n_gpus - 1 children BEFORE anything touches CUDA; rank r drives
// device r. Exchange steps:
//   x   ncclBroadcast from rank 0 straight into every engine's device vector (hsb_device_x), on the engine's stream
//   y   no collective call: hsb_gather_export / ncclAllGather of the blobs / hsb_gather_connect, after which every
//       result drain ALSO stores the rank's block into rank 0's gathered vector over NVLink (peer pointers) and
//       raises an arrival flag -- the role axis_merge + spmv_result_drain play for the 16 clusters
//       (spmv/libfpga/stream_utils.h:36-75, spmv/spmv_result_drain.cpp:36-113)
// Rank 0 checks the gathered y against the host reference of the WHOLE matrix (float: |y - ref| within 1e-5 of
// sum|a x|; fixed point: bit for bit against the closed form of pe.h:64,72) and prints the reference's result
// line for the sharded run, the time of the broadcast and what the gather adds.
#include <nccl.h>
#include <cuda_runtime.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstring>
#include <iostream>
#include <sstream>

#include "common.h"
#include "data_formatter.h"
#include "data_loader.h"
#include "runtime.h"
#include "sharding.h"
#include "synthetic.h"

const unsigned NUM_RUNS = 50;

#define NCCL_CHECK(call)                                                                                   \
    do {                                                                                                   \
        ncclResult_t r__ = (call);                                                                         \
        if (r__ != ncclSuccess) {                                                                          \
            printf("%s:%d Error calling " #call ": %s\n", __FILE__, __LINE__, ncclGetErrorString(r__));   \
            exit(EXIT_FAILURE);                                                                            \
        }                                                                                                  \
    } while (0)
#define CUDA_CHECK(call)                                                                                   \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) {                                                                          \
            printf("%s:%d Error calling " #call ": %s\n", __FILE__, __LINE__, cudaGetErrorString(e__));   \
            exit(EXIT_FAILURE);                                                                            \
        }                                                                                                  \
    } while (0)

static spmv::io::CSRMatrix<float> load(const std::string &spec) {
    std::vector<std::string> f;
    std::stringstream ss(spec);
    for (std::string t; std::getline(ss, t, ':');) f.push_back(t);
    auto num = [&](size_t i) { return i < f.size() ? std::strtoull(f[i].c_str(), nullptr, 0) : 0ull; };
    if (f[0] == "dense") return create_dense_CSR((unsigned)num(1), (unsigned)num(2));
    if (f[0] == "uniform") return create_uniform_sparse_CSR((unsigned)num(1), (unsigned)num(2), (unsigned)num(3));
    if (f[0] == "random") return synth::random_CSR((uint32_t)num(1), (uint32_t)num(2), num(3), num(4), 0.05f);
    if (f[0] == "rmat") return synth::rmat_CSR((uint32_t)num(1), num(2), num(3), 0.05f);
    spmv::io::CSRMatrix<float> m = spmv::io::load_csr_matrix_from_float_npz(spec);
    for (auto &x : m.adj_data) x = 1.0f / m.num_cols;
    return m;
}

static uint32_t word_of(const VAL_T &v) { uint32_t w; std::memcpy(&w, &v, 4); return w; }

// max over ranks of a host double, through a device scalar on the engine's stream
static double max_over_ranks(double v, double *d_scalar, ncclComm_t comm, cudaStream_t s) {
    CUDA_CHECK(cudaMemcpyAsync(d_scalar, &v, 8, cudaMemcpyHostToDevice, s));
    NCCL_CHECK(ncclAllReduce(d_scalar, d_scalar, 1, ncclDouble, ncclMax, comm, s));
    CUDA_CHECK(cudaMemcpyAsync(&v, d_scalar, 8, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    return v;
}

static int run_rank(int rank, int world, const ncclUniqueId &id, const std::string &dataset) {
    using namespace spmv::io;
    using namespace std::chrono;
    CUDA_CHECK(cudaSetDevice(rank));
    ncclComm_t comm;
    NCCL_CHECK(ncclCommInitRank(&comm, world, id, rank));
    hsb_runtime runtime(rank, HSB_IMPL);
    cudaStream_t stream = (cudaStream_t)hsb_stream(runtime.ctx);

    // every rank builds the same matrix (or reads the same file) and keeps its row block
    auto t0 = high_resolution_clock::now();
    CSRMatrix<float> ext_matrix = load(dataset);
    util_round_csr_matrix_dim<float>(ext_matrix, PACK_SIZE * NUM_HBM_CHANNELS * INTERLEAVE_FACTOR, PACK_SIZE);
    CSRMatrix<VAL_T> mat = csr_matrix_convert_from_float<VAL_T>(ext_matrix);
    const std::vector<uint32_t> bounds = spmv::shard::shard_bounds(mat.adj_indptr, world, PACK_SIZE * NUM_HBM_CHANNELS * INTERLEAVE_FACTOR);
    CSRMatrix<VAL_T> mine = spmv::shard::extract_shard(mat, bounds[rank], bounds[rank + 1]);
    auto t1 = high_resolution_clock::now();
    const uint32_t rows_per_part = mine.num_rows > LOGICAL_OB_SIZE ? LOGICAL_OB_SIZE : 0;
    HSB_CHECK(hsb_upload_matrix_csr(runtime.ctx, mine.num_rows, mine.num_cols, mine.adj_indptr.data(),
                                    mine.adj_indices.data(), mine.adj_data.data(), rows_per_part));
    const double preprocess_s = duration<double>(high_resolution_clock::now() - t1).count();
    const double load_s = duration<double>(t1 - t0).count();
    hsb_stats st;
    HSB_CHECK(hsb_get_stats(runtime.ctx, &st));
    const size_t l2 = hsb_device_l2_bytes(rank);
    const int replicas = (int)std::max<uint64_t>(2, (uint64_t)(2.5 * (double)l2 / (double)std::max<uint64_t>(st.format_bytes, 1)) + 1);
    HSB_CHECK(hsb_set_replicas(runtime.ctx, std::min(replicas, 64)));

    // x: made on rank 0, broadcast into the engines' device vectors
    aligned_vector<VAL_T> x(mat.num_cols);
    srand(12345);
    for (size_t i = 0; i < x.size(); i++) x[i] = rank == 0 ? VAL_T(float(rand() % 2)) : VAL_T(0.0f);
    HSB_CHECK(hsb_upload_vector(runtime.ctx, x.data(), mat.num_cols));
    HSB_CHECK(hsb_sync(runtime.ctx));
    void *d_x = hsb_device_x(runtime.ctx);
    double *d_scalar = nullptr;
    CUDA_CHECK(cudaMalloc(&d_scalar, 8));
    NCCL_CHECK(ncclBroadcast(d_x, d_x, mat.num_cols, ncclUint32, 0, comm, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));

    // (a) resident shards, no exchange: the slowest rank's time counts
    float step_ms = 0;
    HSB_CHECK(hsb_time_spmv(runtime.ctx, 5, 1, &step_ms, nullptr));
    max_over_ranks(0.0, d_scalar, comm, stream);                              // line the ranks up
    HSB_CHECK(hsb_time_spmv(runtime.ctx, 0, NUM_RUNS, &step_ms, nullptr));
    const double spmv_ms = max_over_ranks(step_ms, d_scalar, comm, stream);

    // (b) one broadcast of x, alone
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) NCCL_CHECK(ncclBroadcast(d_x, d_x, mat.num_cols, ncclUint32, 0, comm, stream));
    CUDA_CHECK(cudaEventRecord(e0, stream));
    for (int i = 0; i < 20; i++) NCCL_CHECK(ncclBroadcast(d_x, d_x, mat.num_cols, ncclUint32, 0, comm, stream));
    CUDA_CHECK(cudaEventRecord(e1, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    float bc = 0;
    CUDA_CHECK(cudaEventElapsedTime(&bc, e0, e1));
    const double bcast_ms = max_over_ranks(bc / 20.0, d_scalar, comm, stream);

    // gather of y into rank 0, fused into the drains from here on
    unsigned char blob[HSB_PEER_BLOB_BYTES], *d_blobs = nullptr;
    std::vector<unsigned char> blobs((size_t)world * HSB_PEER_BLOB_BYTES);
    HSB_CHECK(hsb_gather_export(runtime.ctx, mat.num_rows, rank == 0, blob));
    CUDA_CHECK(cudaMalloc(&d_blobs, blobs.size()));
    CUDA_CHECK(cudaMemcpyAsync(d_blobs + (size_t)rank * HSB_PEER_BLOB_BYTES, blob, HSB_PEER_BLOB_BYTES, cudaMemcpyHostToDevice, stream));
    NCCL_CHECK(ncclAllGather(d_blobs + (size_t)rank * HSB_PEER_BLOB_BYTES, d_blobs, HSB_PEER_BLOB_BYTES, ncclUint8, comm, stream));
    CUDA_CHECK(cudaMemcpyAsync(blobs.data(), d_blobs, blobs.size(), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    HSB_CHECK(hsb_gather_connect(runtime.ctx, world, rank, bounds[rank], blobs.data()));
    max_over_ranks(0.0, d_scalar, comm, stream);

    // parity of the gathered vector
    HSB_CHECK(hsb_spmv(runtime.ctx));
    HSB_CHECK(hsb_sync(runtime.ctx));
    max_over_ranks(0.0, d_scalar, comm, stream);                              // every rank's drain has landed
    bool ok = true;
    if (rank == 0) {
        std::vector<uint32_t> y(mat.num_rows);
        HSB_CHECK(hsb_download_gathered(runtime.ctx, y.data(), mat.num_rows));
        srand(12345);
        for (size_t i = 0; i < x.size(); i++) x[i] = VAL_T(float(rand() % 2));   // (rank 0 still has it; kept explicit)
        size_t bad = 0;
        for (uint32_t r = 0; r < mat.num_rows; r++) {
#if defined(FP_POB) || defined(FP_STALL)
            double ref = 0, scale = 0;
            for (uint32_t i = mat.adj_indptr[r]; i < mat.adj_indptr[r + 1]; i++) {
                const double t = (double)mat.adj_data[i] * (double)x[mat.adj_indices[i]];
                ref += t;
                scale += std::fabs(t);
            }
            float got;
            std::memcpy(&got, &y[r], 4);
            if (std::fabs((double)got - ref) > 1e-5 * scale + 1e-30) bad++;
#else
            uint64_t sum = 0;                                                 // closed form of pe.h:64,72 for non-negative terms
            for (uint32_t i = mat.adj_indptr[r]; i < mat.adj_indptr[r + 1]; i++)
                sum += spmv::ufixed_q8_24::mul(mat.adj_data[i], x[mat.adj_indices[i]]).raw;
            if (y[r] != (sum > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)sum)) bad++;
#endif
        }
        ok = bad == 0;
        std::cout << "INFO : gathered y on rank 0 (" << mat.num_rows << " rows from " << world << " row blocks): "
                  << (ok ? "matches the host reference of the whole matrix" : "MISMATCH") << " (" << bad << " bad rows)" << std::endl;
    }
    max_over_ranks(0.0, d_scalar, comm, stream);

    // (c) the same loop with the gather fused into the drains
    HSB_CHECK(hsb_time_spmv(runtime.ctx, 5, 1, &step_ms, nullptr));
    max_over_ranks(0.0, d_scalar, comm, stream);
    HSB_CHECK(hsb_time_spmv(runtime.ctx, 0, NUM_RUNS, &step_ms, nullptr));
    const double gather_ms = max_over_ranks(step_ms, d_scalar, comm, stream);

    if (rank == 0) {
        const double nnz = (double)mat.adj_indptr[mat.num_rows];
        std::cout << "INFO : " << world << " GPUs, row blocks at";
        for (uint32_t b : bounds) std::cout << ' ' << b;
        std::cout << "; rank 0: " << st.nnz << " nnz, " << st.n_col_tiles << " column tiles of " << st.tile_cols
                  << (st.layout ? ", narrow layout" : ", wide layout") << ", format " << (double)st.format_bytes / (double)st.nnz << " B/nnz" << std::endl;
        std::cout << "INFO : dataset load + shard " << load_s << " s (host), ncclBroadcast(x) " << bcast_ms << " ms ("
                  << mat.num_cols * 4.0 / 1e6 / bcast_ms << " GB/s), SpMV with y gathered on rank 0 " << gather_ms << " ms ("
                  << 2.0 * nnz / 1e6 / gather_ms << " GOPS): the gather adds " << gather_ms - spmv_ms << " ms" << std::endl;
        // the reference's result line (sw/benchmark.cpp:80-87), for the sharded SpMV with resident x
        std::cout << "{Preprocessing: " << preprocess_s << " s | SpMV: " << spmv_ms << " ms | "
                  << nnz * 8.0 / 1024.0 / 1024.0 / 1024.0 / (spmv_ms / 1000.0) << " GBPS | " << 2.0 * nnz / 1e6 / spmv_ms << " GOPS }" << std::endl;
    }
    cudaFree(d_blobs);
    cudaFree(d_scalar);
    hsb_destroy(runtime.ctx);
    runtime.ctx = nullptr;
    ncclCommDestroy(comm);
    return ok ? 0 : 1;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        std::cout << "Usage: " << argv[0] << " <dataset.npz | dense:R:C | uniform:R:C:K | random:R:C:NNZ:SEED | rmat:N:EDGES:SEED> <n_gpus>" << std::endl;
        return 0;
    }
    const std::string dataset = argv[1];
    const int world = atoi(argv[2]);
    if (world < 1 || world > 16) { std::cout << "ERROR : n_gpus must be in [1, 16]" << std::endl; return 1; }
    std::cout << "------ Running multi-GPU benchmark on " << dataset << " with " << world << " GPU(s), one process each" << std::endl;
    std::cout.flush();
    ncclUniqueId id;
    NCCL_CHECK(ncclGetUniqueId(&id));                      // before the forks: every rank inherits it; no CUDA call yet
    std::vector<pid_t> kids;
    int rank = 0;
    for (int r = 1; r < world; r++) {
        pid_t p = fork();
        if (p < 0) { perror("fork"); return 1; }
        if (p == 0) { rank = r; kids.clear(); break; }
        kids.push_back(p);
    }
    int rc = run_rank(rank, world, id, dataset);
    if (rank != 0) _exit(rc);
    for (pid_t p : kids) {
        int status = 0;
        waitpid(p, &status, 0);
        if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) rc = 1;
    }
    std::cout << (rc == 0 ? "===== Benchmark Finished =====" : "===== Benchmark FAILED =====") << std::endl;
    return rc;
}
