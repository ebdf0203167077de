/*
 * hisparse_b200.h -- C ABI of the B200-native SpMV engine that stands in for HiSparse's
 * FPGA kernel pipeline (spmv_vector_loader -> spmv_sk0/1/2 -> spmv_result_drain).
 *
 * The reference has no plugin/FFI registry; its host <-> accelerator boundary is the kernel
 * argument contract, which exists in two equivalent forms:
 *   (1) the OpenCL form   sw/host.cpp:263-371   (16 CL_BUFFER_RDONLY channel images + x + y,
 *       enqueueMigrateMemObjects, setArg, 5 x enqueueTask + finish per row partition) and
 *   (2) the C-simulation form   spmv_csim/csim.cpp:22-46   `top_wrapper(...)`.
 * Every entry point below names the reference interface it replaces. Plain pointers and sizes
 * only; all buffers are HOST buffers unless the name says `device`. All functions returning
 * `int` return 0 on success and a negative HSB_E* code on failure; hsb_last_error() gives the
 * message (the reference prints file:line and exits, xrt/includes/xcl2/xcl2.hpp:40-46 -- the C++
 * wrapper in hisparse_b200/host/ reproduces that behaviour on top of these codes).
 *
 * Value words are always 32 bits: raw Q8.24 (ap_ufixed<32,8,AP_RND,AP_SAT>,
 * spmv/libfpga/common.h:35-38) for HSB_IMPL_FIXED, IEEE fp32 bits for the float variants.
 */
#ifndef HISPARSE_B200_H_
#define HISPARSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSB_NUM_HBM_CHANNELS 16     /* spmv/libfpga/common.h:173-176 : 4 + 6 + 6 clusters */
#define HSB_PACK_SIZE 8             /* spmv/libfpga/common.h:30 */
#define HSB_IDX_MARKER 0xFFFFFFFFu  /* spmv/libfpga/common.h:8 */

/* IMPL= of sw/Makefile:2-12 and spmv_csim/Makefile:27-38 */
enum { HSB_IMPL_FIXED = 0, HSB_IMPL_FLOAT_POB = 1, HSB_IMPL_FLOAT_STALL = 2 };

enum {
    HSB_OK = 0,
    HSB_EINVAL = -1,     /* bad argument / malformed matrix or channel image */
    HSB_ECUDA = -2,      /* CUDA runtime error (message has the CUDA error string) */
    HSB_ESTATE = -3,     /* call order violated (e.g. spmv before a matrix upload) */
    HSB_ENOMEM = -4
};

typedef struct hsb_ctx hsb_ctx;

/* compile-time constants of the selected implementation, as the reference's common.h defines them
 * (spmv/libfpga/common.h:162-179, spmv-fp/libfpga/common.h:178-199) */
typedef struct hsb_config {
    uint32_t pack_size;            /* 8 */
    uint32_t num_hbm_channels;     /* 16 */
    uint32_t interleave_factor;    /* 1 (fixed, float_pob) or 8 (float_stall) */
    uint32_t logical_ob_size;      /* rows per row partition: 1048576 / 131072 / 1048576 */
    uint32_t logical_vb_size;      /* columns per column partition: 32768 */
} hsb_config;

typedef struct hsb_stats {
    uint64_t nnz;                  /* non-zeros of the resident matrix */
    uint32_t rows, cols;           /* padded dimensions */
    uint32_t n_row_parts, n_col_tiles, tile_cols;
    uint64_t n_slices;             /* warp work units: 32 lane streams each */
    uint64_t n_streams;            /* lane streams = row segments after splitting long rows */
    uint64_t n_elems;              /* stored non-zero slots including padding */
    uint64_t format_bytes;         /* bytes of the tile-stream format in HBM */
    uint64_t algorithmic_bytes;    /* 8*nnz + 4*(rows+1) + 4*rows + 4*cols   */
    uint64_t kernel_launches;      /* kernels of this library launched so far on this context */
    uint32_t sm_count, grid;
    uint32_t replicas;
    uint32_t layout;               /* 0: wide tile streams (4 slots per lane and step), 1: narrow (hypersparse matrices:
                                      1 slot per lane and unit, row ids in the stream) -- csrc/tile_format.h */
    double preprocess_seconds;     /* host formatting time of the last upload ("Preprocessing" in benchmark.cpp:80-87) */
} hsb_stats;

const char *hsb_version(void);
const char *hsb_last_error(void);
int hsb_device_count(void);
int hsb_get_config(int impl, hsb_config *out);

/* == cl::Context + cl::CommandQueue + cl::Kernel x5 (sw/host.cpp:556-590). One context per GPU. */
hsb_ctx *hsb_create(int device, int impl);
void hsb_destroy(hsb_ctx *ctx);

/* Page-locked host memory: the reference's aligned_allocator (xrt/includes/xcl2/xcl2.hpp:61-76)
 * gives page-aligned host vectors that the device buffers alias (CL_MEM_USE_HOST_PTR). */
void *hsb_host_alloc(size_t bytes);
void hsb_host_free(void *p);

/* == the 16 CL_BUFFER_RDONLY channel buffers + enqueueMigrateMemObjects (sw/host.cpp:263-299).
 * ch[c] points at ch_packets[c] 64-byte packets laid out exactly as sw/host.cpp:163-231 builds
 * them (header packets, then data packets, interleaved for float_stall). num_rows / num_cols are
 * the padded matrix dimensions (the reference passes them through part_len / num_cols later). */
int hsb_upload_matrix_cpsr(hsb_ctx *ctx, const void *const ch[HSB_NUM_HBM_CHANNELS],
                           const size_t ch_packets[HSB_NUM_HBM_CHANNELS], unsigned num_row_partitions,
                           unsigned num_col_partitions, unsigned num_rows, unsigned num_cols);

/* Fast path that skips CPSR: hand over the CSR the reference's formatter would consume
 * (spmv::io::CSRMatrix<VAL_T>, sw/data_loader.h:19-30) -- vals are 32-bit VAL_T words.
 * rows_per_partition: row-partition length (LOGICAL_OB_SIZE); 0 = a single partition. */
int hsb_upload_matrix_csr(hsb_ctx *ctx, uint32_t rows, uint32_t cols, const uint32_t *indptr,
                          const uint32_t *indices, const void *vals, uint32_t rows_per_partition);

/* The same with the formatting done ON THE GPU (device-wide sorts and scans instead of host threads):
 * from a host CSR, or from a CSR that already lives in device memory (32-bit device pointers to
 * indptr[rows+1], indices[nnz], vals[nnz]); the device CSR is not modified and may be freed afterwards. */
int hsb_upload_matrix_csr_gpu(hsb_ctx *ctx, uint32_t rows, uint32_t cols, const uint32_t *indptr,
                              const uint32_t *indices, const void *vals, uint32_t rows_per_partition);
int hsb_upload_matrix_csr_device(hsb_ctx *ctx, uint32_t rows, uint32_t cols, uint64_t nnz,
                                 const uint32_t *d_indptr, const uint32_t *d_indices, const void *d_vals,
                                 uint32_t rows_per_partition);

/* == vector_buf on HBM[20] + migrate (sw/host.cpp:282-298). x_packed: num_cols 32-bit words.
 * Asynchronous on a copy stream; the next hsb_spmv* waits for it. */
int hsb_upload_vector(hsb_ctx *ctx, const void *x_packed, unsigned num_cols);

/* == setArg(row_part_id, part_len) + enqueueTask x5 + finish for ONE row partition
 * (sw/host.cpp:335-357; kernel arguments spmv/spmv_sk0.cpp:13-24, spmv_vector_loader.cpp:95-101,
 * spmv_result_drain.cpp:11-18). part_len = rows per cluster in this partition (rows / 16). The
 * call is asynchronous on the context's stream; hsb_sync() is the finish(). */
int hsb_spmv_row_partition(hsb_ctx *ctx, unsigned row_part_id, unsigned part_len,
                           unsigned num_col_partitions, unsigned num_partitions, unsigned num_cols);

/* all row partitions back to back (the loop at sw/benchmark.cpp:317-340), asynchronous.
 * The packed result becomes final on the device at the next hsb_sync() / hsb_download_result()
 * (or at the end of the next hsb_spmv*, whose CTAs drain their predecessor's row accumulators last). */
int hsb_spmv(hsb_ctx *ctx);
int hsb_sync(hsb_ctx *ctx);

/* == enqueueMigrateMemObjects(result_buf -> host) + finish (sw/host.cpp:370-371).
 * y_packed: num_rows 32-bit words, natural row order. Synchronises. */
int hsb_download_result(hsb_ctx *ctx, void *y_packed, unsigned num_rows);
/* The same without the finish(): the copy runs on its own stream (the reference's queue is out of
 * order too, sw/host.cpp:586-590); y_packed is valid after the next hsb_sync() -- not earlier: when the
 * sums of the last SpMV are still in the row accumulators the copy is deferred and attached to the next
 * hsb_spmv (which drains them anyway, at its end) or issued by hsb_sync. Together with the multi-buffered
 * x of hsb_upload_vector and a double-buffered device y, upload(k+1), SpMV(k) and download(k-1)
 * overlap, and the compute stream carries nothing but back-to-back kernel launches. */
int hsb_download_result_async(hsb_ctx *ctx, void *y_packed, unsigned num_rows);

/* == spmv_csim/csim.cpp:22-46 `top_wrapper`: one row partition, host buffers in, host buffer out.
 * Channel image lengths are recovered from the image headers. Uploads, runs, downloads. */
int hsb_top_wrapper(int impl, const void *const matrix_hbm[HSB_NUM_HBM_CHANNELS],
                    const void *packed_dense_vector, void *packed_dense_result, unsigned row_part_id,
                    unsigned part_len, unsigned num_col_partitions, unsigned num_partitions,
                    unsigned num_cols);

/* ---- iterative callers (the step either side of SpMV in the reference's intended use: PageRank-style
 * x <- alpha (*) A x (+) beta, unit_tests/test_app.cpp:51-136) -- everything stays on the device ---------
 * alpha / beta are 32-bit value words (Q8.24 or fp32 bits); the arithmetic is the implementation's own:
 * fixed = rounded saturating product (spmv/libfpga/pe.h:64) + saturating add (pe.h:72); float = fp32
 * multiply then add (spmv-fp/libfpga/pe-pob.h:64-66). */
/* Finalise y of the last SpMV and write x_next[col_offset + r] = alpha (*) y[r] (+) beta for every row r into
 * the NEXT vector buffer (fused with the row drain). col_offset = this GPU's first row when the matrix is
 * row-block sharded: each rank fills its slice, then the ranks all-gather hsb_device_x_next() in place. */
int hsb_axpb_to_vector(hsb_ctx *ctx, uint32_t alpha_word, uint32_t beta_word, uint32_t col_offset);
void *hsb_device_x_next(hsb_ctx *ctx);
/* the next vector buffer becomes the one the following SpMVs read (after the all-gather, if any) */
int hsb_vector_commit(hsb_ctx *ctx);
/* Multi-GPU iteration without the host: one process per GPU; after every rank has uploaded its row-block
 * shard and the same x, the ranks exchange hsb_peer_export blobs (e.g. torch.distributed.all_gather) and
 * call hsb_peer_connect (CUDA IPC, peer access over NVLink). hsb_axpb_to_peers then stores this rank's
 * slice of the next vector into the next x buffer of EVERY rank and raises an arrival flag there when the
 * kernel has finished -- the compute step and the all-gather are one kernel -- and after
 * hsb_vector_commit the next hsb_spmv polls the arrival flags of all ranks before it stages x. No NCCL
 * call, no host synchronisation inside the iteration. */
#define HSB_PEER_BLOB_BYTES 192
int hsb_peer_export(hsb_ctx *ctx, void *blob);
int hsb_peer_connect(hsb_ctx *ctx, int world, int rank, const void *blobs /* world x HSB_PEER_BLOB_BYTES */);
int hsb_axpb_to_peers(hsb_ctx *ctx, uint32_t alpha_word, uint32_t beta_word, uint32_t col_offset);
/* iters x { hsb_spmv; hsb_axpb_to_peers(alpha, beta, col_offset); hsb_vector_commit } as ONE resident kernel per GPU
 * (a cooperative launch, as hsb_iterate): the SpMV and the update are separated by a grid barrier, the iterations by the
 * arrival flags of all ranks -- the multi-GPU iteration is a single kernel, no host and no NCCL inside. Every rank calls
 * it with the same iters; afterwards y = this rank's rows of A x of the last iteration and the current vector is the
 * updated one on every rank (the next SpMV polls the last arrival flags). Needs all of the device's SMs for one grid:
 * ranks that share a GPU cannot be resident together (use the launch-per-step calls there). */
int hsb_iterate_peers(hsb_ctx *ctx, int iters, uint32_t alpha_word, uint32_t beta_word, uint32_t col_offset);
/* Gather of y across row-block shards, fused into the result drain (the role of axis_merge +
 * spmv_result_drain, spmv/libfpga/stream_utils.h:36-75 and spmv/spmv_result_drain.cpp:36-113, which assemble
 * ONE y from the 16 clusters' streams): once connected, every drain of this context ALSO stores its result
 * words into the gathered vector of each target rank at `row_offset` (peer stores over NVLink) and, when the
 * drain's grid is through, raises this rank's arrival flag there. Every rank calls hsb_gather_export
 * (want_buffer != 0 on the ranks that receive the whole y: one root for a gather, everybody for an
 * all-gather), the blobs are exchanged (HSB_PEER_BLOB_BYTES each, e.g. torch.distributed.all_gather or
 * ncclAllGather), every rank calls hsb_gather_connect. All ranks must issue the same sequence of SpMVs.
 * hsb_gather_wait (targets only) makes later work on the context's stream wait for the blocks of ALL ranks
 * of the last SpMV; hsb_download_gathered = wait + copy to the host. The gathered vector is overwritten by
 * the next drain: consume it (or synchronise the ranks) before issuing further SpMVs. */
int hsb_gather_export(hsb_ctx *ctx, uint32_t total_rows, int want_buffer, void *blob);
int hsb_gather_connect(hsb_ctx *ctx, int world, int rank, uint32_t row_offset, const void *blobs);
int hsb_gather_wait(hsb_ctx *ctx);
void *hsb_device_y_gathered(hsb_ctx *ctx);
int hsb_download_gathered(hsb_ctx *ctx, void *y_packed, uint32_t total_rows);
/* NUMA node of the GPU (sysfs), or -1 when the platform does not say; hsb_host_alloc prefers that node */
int hsb_device_numa_node(int device);
/* L2 cache size of the device in bytes (0 on error): what a timed loop's working set has to exceed */
size_t hsb_device_l2_bytes(int device);
/* single GPU, rows <= cols: iters x { hsb_spmv; hsb_axpb_to_vector(alpha, beta, 0); hsb_vector_commit } -- the same
 * results, issued as ONE cooperative launch whose grid stays resident: the two dependencies of an iteration (all row
 * updates before the drain, the whole new vector before the next SpMV) are grid-wide barriers instead of kernel
 * boundaries. Afterwards y = A x of the last iteration (hsb_download_result) and the current vector is the updated one.
 * hsb_set_option "iterate_persistent" 0 (or a gather of y being connected) selects the launch-per-step form. */
int hsb_iterate(hsb_ctx *ctx, int iters, uint32_t alpha_word, uint32_t beta_word);

/* ---- measurement and multi-GPU plumbing (no reference counterpart) ------------------------ */
int hsb_get_stats(hsb_ctx *ctx, hsb_stats *out);
/* keep n copies of the matrix in HBM and rotate through them on successive hsb_spmv() calls so
 * that a timed loop never re-reads a matrix that is still in the L2 (hsb_device_l2_bytes: 126 MB on B200) */
int hsb_set_replicas(hsb_ctx *ctx, int n);
/* CUDA-event timing on the context's stream: `steps` full SpMVs after `warmup` untimed ones.
 * step_ms = (stop - start) / steps over the whole loop; kernel_ms = mean duration of the main
 * tile kernel alone, from per-launch event pairs in a second loop of the same length. */
int hsb_time_spmv(hsb_ctx *ctx, int warmup, int steps, float *step_ms, float *kernel_ms);
/* End-to-end timing with HOST buffers (what bench.py reports as e2e): `iters` times
 *   hsb_upload_vector(x_host[k & 1]); hsb_spmv(); hsb_download_result[_async](y_host[k & 1]);
 * then hsb_sync(), on the wall clock -- the loop of sw/benchmark.cpp:315-343 with the transfers
 * inside, issued from C the way the reference's C++ benchmark issues its OpenCL calls. */
int hsb_time_e2e(hsb_ctx *ctx, const void *const x_host[2], void *const y_host[2], unsigned num_cols,
                 unsigned num_rows, int iters, int async_download, double *seconds_per_spmv);
/* Tuning / A-B switches (synchronises first). "flags": 1 = flag pipeline (default when the driver offers
 * stream memory operations), 0 = stream events between launches; "xwait_once": 1 = launches stop polling
 * the x flag once one that polled it has completed (default); "host_drain": 1 = deferred downloads into
 * page-locked memory are written by the kernel's drain itself instead of the copy engine (default); "acquire":
 * 1 = the kernels' flag waits are acquire loads + fence.proxy.async (default), 0 = relaxed (A/B aid); "xflag_copy":
 * 1 = the "vector has landed" flag is written by a 4-byte copy behind the vector's copy (default), 0 = by a stream
 * memory operation; "iterate_persistent": 1 = hsb_iterate is one cooperative launch (default), 0 = a launch per step. */
int hsb_set_option(hsb_ctx *ctx, const char *name, int value);
/* Profiling aid: SM-clock stamps of the last launch, [sm_count][34] = per warp "my slices are done",
 * then CTA "finished" (wait for the predecessor and drain included); the last word is unused. out == NULL arms (capacity != 0) or
 * disarms (capacity == 0) the trace and returns the number of words. */
int hsb_debug_trace(hsb_ctx *ctx, unsigned long long *out, size_t capacity);
/* Profiling aid: %globaltimer stamps (ns) of the last 256 SpMV launches, [256][8] indexed by launch number
 * % 256: 0 first CTA started, 1 CTA 0 saw its x flag, 2 CTA 0 saw the predecessor complete, 3 CTA 0 finished
 * its drain share and published it, 4 CTA 0 has its x tile, 5 last CTA ended, 6 CTA 0 finished its matrix work. out == NULL arms (capacity != 0) / disarms.
 * Returns the number of the last launch. */
int hsb_debug_timeline(hsb_ctx *ctx, unsigned long long *out, size_t capacity);
/* Profiling aid: cudaProfilerStart (on != 0) / cudaProfilerStop around a range of calls, for
 * `ncu --replay-mode app-range` (DRAM counters over a whole loop of overlapping launches). Synchronises first. */
int hsb_debug_profiler(hsb_ctx *ctx, int on);
/* Profiling aid: steps and slices the whole-matrix plan gives to every CTA. */
int hsb_debug_plan(hsb_ctx *ctx, uint32_t *steps, uint32_t *slices, size_t capacity);
/* raw device pointers / stream for callers that move x or y with NCCL (torch.distributed);
 * call hsb_sync() before reading device y */
void *hsb_device_x(hsb_ctx *ctx);
void *hsb_device_y(hsb_ctx *ctx);   /* query after hsb_sync(): the result buffer alternates when downloads are deferred */
void *hsb_stream(hsb_ctx *ctx);

/* ---- synthetic matrices generated on the device (benchmark inputs; no reference counterpart: the
 * reference loads .npz datasets, sw/data_loader.h:51-70) ------------------------------------------ */
typedef struct hsb_device_csr {
    uint32_t rows, cols;
    uint64_t nnz;
    uint32_t *d_indptr, *d_indices, *d_vals;   /* device pointers: rows + 1, nnz, nnz words */
    int device;
} hsb_device_csr;
/* Rows [first_global_row, first_global_row + rows) of a power-law matrix with `cols` columns (BASELINE
 * config C5): row degree = truncated discrete Pareto(alpha) with mean `mean_degree`, clipped to
 * max_degree; a fraction band_fraction of a row's columns lies within +-band_half_width of the
 * diagonal, the rest is uniform; column ids sorted and unique per row. Values U[0,1) * value_scale as
 * fp32 bits (value_kind 0) or raw Q8.24 words (value_kind 1). A pure function of (seed, global row):
 * shards generated on different GPUs are pieces of the same matrix. The result feeds
 * hsb_upload_matrix_csr_device; release it with hsb_device_csr_free. */
int hsb_synth_powerlaw_csr_device(int device, uint32_t rows, uint32_t cols, uint64_t first_global_row,
                                  double mean_degree, double alpha, uint32_t max_degree,
                                  uint32_t band_half_width, double band_fraction, uint64_t seed,
                                  int value_kind, float value_scale, hsb_device_csr *out);
/* any of indptr / indices / vals may be NULL */
int hsb_device_csr_download(const hsb_device_csr *m, uint32_t *indptr, uint32_t *indices, uint32_t *vals);
void hsb_device_csr_free(hsb_device_csr *m);
const char *hsb_synth_last_error(void);

/* ---- host-side format inspection (no GPU needed; used by the CPU test-suite) --------------- */
typedef struct hsb_format hsb_format;
/* CSR -> tile-stream format on the host, exactly as hsb_upload_matrix_csr builds it.
 * tile_cols == 0 picks the width automatically. NULL on malformed input. */
hsb_format *hsb_format_build(uint32_t rows, uint32_t cols, const uint32_t *indptr, const uint32_t *indices,
                             const void *vals, uint32_t rows_per_partition, uint32_t tile_cols);
/* Copy of the matrix resident on the device (whichever way it was built) for inspection. */
hsb_format *hsb_format_from_context(hsb_ctx *ctx);
int hsb_format_stats(const hsb_format *f, hsb_stats *out);
/* Walk the chunk stream the way the kernel addresses it (value slots, end-of-segment flags,
 * seg_row, chunk descriptors) and rebuild the CSR. indices/vals hold nnz words each. Returns
 * HSB_EINVAL if the stream is internally inconsistent. Entries of a row come back grouped by
 * column tile, CSR order inside a tile. */
int hsb_format_expand(const hsb_format *f, uint32_t *indptr, uint32_t *indices, uint32_t *vals);
/* The launch plan the engine would cut for `ctas` CTAs (whole matrix): one record of 4 words per warp share,
 * {cta, tile, first step, end step} with tile-relative steps. Returns the number of records (call with
 * records == NULL to size the buffer). Every step of every tile is covered exactly once. */
long long hsb_format_plan(const hsb_format *f, uint32_t ctas, uint32_t *records, size_t capacity_records);
void hsb_format_free(hsb_format *f);
/* Decode reference channel images back to CSR (what hsb_upload_matrix_cpsr does first).
 * indptr: num_rows + 1 words; indices / vals: capacity words each; *nnz receives the count
 * (call with capacity 0 to size the buffers). */
int hsb_cpsr_to_csr(int impl, const void *const ch[HSB_NUM_HBM_CHANNELS],
                    const size_t ch_packets[HSB_NUM_HBM_CHANNELS], unsigned num_row_partitions,
                    unsigned num_col_partitions, unsigned num_rows, unsigned num_cols, uint32_t *indptr,
                    uint32_t *indices, uint32_t *vals, size_t capacity, size_t *nnz);

#ifdef __cplusplus
}
#endif
#endif /* HISPARSE_B200_H_ */
