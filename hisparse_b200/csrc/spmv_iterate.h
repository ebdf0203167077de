// hsb_iterate / hsb_iterate_peers as one cooperative launch (kernel: spmv_iterate.cuh). Included by spmv_kernels.cuh.
#ifndef HISPARSE_B200_SPMV_ITERATE_H_
#define HISPARSE_B200_SPMV_ITERATE_H_
namespace hsb {

// `iters` iterations x <- alpha (*) (A x) (+) beta in one cooperative launch (spmv_iterate_kernel): iteration k reads
// x0 (k even) / x1 (k odd) and writes the other; y holds A x of the last iteration; p.acc must be all zero and is left all zero.
// grid <= number of SMs (one CTA per SM: every CTA resident). p: vals / cols / slice_rows / cta_seg / segs / acc / y /
// trash_row / seq / done_dev / done_seq / error_flag / comb_offset / narrow as for launch_spmv, everything else zero.
struct IterateParams {
    uint32_t *x0, *x1;            // iteration k reads x0 (k even) or x1 (k odd) and writes the other
    uint32_t *barrier;            // grid-barrier counter, zero at launch; iters * 2 * grid must stay below 2^31
    uint32_t iters, alpha, beta;
    uint32_t rows, x_limit;       // rows of the matrix; elements of x that may be written (x_next[r] for r < x_limit)
};
// Multi-GPU form (peers != null): iteration k reads this rank's x buffer (buf0 + k) % 4 (it.x0 = buffer 0, buffers
// x_stride words apart) once the arrival flags of all ranks show seq0 + k (k == 0: only if wait_first), and stores its
// slice alpha (*) y (+) beta at col_offset of buffer (buf0 + k + 1) % 4 of EVERY rank, then raises arrival[rank] = seq0 + k + 1
// on every rank. All ranks launch the same number of iterations.
struct IteratePeers {
    uint32_t *x_base[kMaxPeers];  // rank g's x buffer 0
    uint32_t *flag[kMaxPeers];    // &arrival[this rank] in rank g's flag array
    const uint32_t *arrival;      // this rank's flag array: arrival[g] written by rank g
    unsigned long long x_stride;
    uint32_t world, buf0, seq0, col_offset, wait_first;
};
cudaError_t launch_iterate(int arith, const SpmvParams &p, const IterateParams &it, const IteratePeers *peers, int grid,
                           uint32_t smem_bytes, cudaStream_t stream);

}  // namespace hsb
#endif
