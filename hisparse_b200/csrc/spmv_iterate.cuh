// The iterative callers' kernel (hsb_iterate): included by spmv_kernels.cu after its device functions (work loop, row
// updates, arithmetic policies, drain helpers) -- not a stand-alone header.
namespace {
// ---------------------------------------------------------------------------------------
// x <- alpha (*) (A x) (+) beta, `iters` times, in ONE launch (hsb_iterate): the iterative callers of the reference
// (PageRank-style pull iterations). The launch-per-SpMV form pays two grid completions per iteration (SpMV -> update
// kernel -> SpMV: about 3.8 us each on B200, more than the work of a small matrix); here the grid stays resident
// (a cooperative launch: one CTA per SM, all co-resident) and the two dependencies of an iteration -- every row
// update before the drain, every element of the new vector before the next x tile is staged -- are two grid-wide
// barriers on a counter in device memory (arrive: fence + atomic add; wait: acquire loads), about 1 us each.
// Work loop, row updates and arithmetic are those of spmv_tiles_kernel; one accumulator buffer, re-zeroed by the drain.
// ---------------------------------------------------------------------------------------
// Grid barrier in two halves on a monotonic counter: every CTA arrives once per barrier (after a CTA-wide barrier: thread
// 0's fence is cumulative over what the other threads wrote), and barrier number b (1, 2, ...) is complete when the
// counter has reached b * gridDim.x. A CTA waits for barrier b before it arrives at b + 1, so the count cannot run ahead.
// The barriers' fences: __threadfence[_system]() emit the sequentially consistent form (MEMBAR.SC). Acquire-release is all
// the release / acquire patterns need (HSB_ITER_SC_FENCES=0 builds that form), but it measures the same on one and on
// two GPUs (tools/iterate_fence_ab.sh: 5.12 / 5.11, 19.95 / 19.98, 10.25 / 10.28, 18.66 / 18.71 us), so the stronger one stays.
#ifndef HSB_ITER_SC_FENCES
#define HSB_ITER_SC_FENCES 1
#endif
__device__ __forceinline__ void fence_gpu() {
#if HSB_ITER_SC_FENCES
    __threadfence();
#else
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}
__device__ __forceinline__ void fence_sys() {
#if HSB_ITER_SC_FENCES
    __threadfence_system();
#else
    asm volatile("fence.acq_rel.sys;" ::: "memory");
#endif
}
__device__ __forceinline__ void grid_arrive(uint32_t *word) {
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_gpu();
        atomicAdd(word, 1u);
    }
}
// one thread; false: timed out (a CTA of the grid never arrived)
__device__ __forceinline__ bool grid_wait(const uint32_t *word, uint32_t target) {
    uint32_t v = 0, polls = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(word) : "memory");
        if ((int32_t)(v - target) >= 0) return true;
        if (++polls > (1u << 14)) {
            if (polls - (1u << 14) > kFlagPolls) return false;
            __nanosleep(1000);
        }
    }
}

// y = final row sums, accumulators re-zeroed, x_next = alpha (*) y (+) beta: pe dump + result drain + the caller's update
// in one pass; four rows per 128-bit access when a thread has several rows
template <class A>
__device__ __forceinline__ void drain_axpb_rows(void *acc, uint32_t *y, uint32_t *x_next, uint32_t rows, uint32_t x_limit,
                                                uint32_t alpha, uint32_t beta, uint32_t t, uint32_t n_threads) {
    const bool vec = rows > n_threads && (reinterpret_cast<uintptr_t>(x_next) & 15u) == 0;
    const uint32_t e4 = vec ? (min(rows, x_limit) & ~3u) : 0u;
    for (uint32_t c = t; c < (e4 >> 2); c += n_threads) {
        const typename A::Raw4 raw = A::load4(acc, 4u * c);
        A::zero4(acc, 4u * c, false);
        const uint4 v = A::final4(raw);
        *reinterpret_cast<uint4 *>(y + 4u * c) = v;
        *reinterpret_cast<uint4 *>(x_next + 4u * c) = make_uint4(A::axpb(alpha, v.x, beta), A::axpb(alpha, v.y, beta),
                                                                A::axpb(alpha, v.z, beta), A::axpb(alpha, v.w, beta));
    }
    for (uint32_t r = e4 + t; r < rows; r += n_threads) {
        const uint32_t v = A::drain(acc, r);
        y[r] = v;
        if (r < x_limit) x_next[r] = A::axpb(alpha, v, beta);
    }
}

// kPeers: the matrix is a row-block shard and the vector lives on every GPU (hsb_peer_connect). The update stores this
// rank's slice of the next vector into the x buffer of EVERY rank over NVLink; the CTA that arrives last at the second
// barrier -- it knows every store of the grid has been issued and fenced -- raises this rank's arrival flag on every rank,
// and the next iteration starts when the flags of ALL ranks (this one's included, which stands for the second barrier)
// show the iteration's number: the whole multi-GPU iteration is one resident kernel per GPU, no host, no NCCL.
template <class A, bool kNarrow, bool kPeers>
__global__ void __launch_bounds__(kThreads, 1) spmv_iterate_kernel(const SpmvParams p, const IterateParams it, const IteratePeers pr) {
    unsigned char *smem_raw = reinterpret_cast<unsigned char *>(xs);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t abort_flag, dead;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t g0 = __ldg(p.cta_seg + blockIdx.x), g1 = __ldg(p.cta_seg + blockIdx.x + 1);
    typedef typename A::acc_t acc_t;
    if (tid == 0) {
        mbar_init(&bar, 1);
        abort_flag = 0u;
        dead = 0u;
    }
    if (tid < kColBias) xs[tid] = 0u;
    acc_t *comb_base = (!kNarrow && p.comb_offset) ? reinterpret_cast<acc_t *>(smem_raw + p.comb_offset) : nullptr;
    if (comb_base)
        for (uint32_t i = tid; i < 2u * kCombineSlots * kLanes; i += kThreads) comb_base[i] = acc_t(0);
    __syncthreads();
    uint32_t parity = 0;
    for (uint32_t k = 0; k < it.iters; k++) {
        const uint32_t *x = kPeers ? it.x0 + (size_t)((pr.buf0 + k) & 3u) * pr.x_stride : ((k & 1u) ? it.x1 : it.x0);
        // The whole new vector, and every re-zeroed accumulator, before this iteration's first x tile is staged (the
        // row updates come after the tile). Only thread 0 waits: the other warps go ahead and fill their prefetch
        // rings with matrix data, which does not depend on x. A timed-out wait still stages the tile -- the warps
        // must be released -- but nobody multiplies (abort_flag), and the CTA leaves at the next barrier.
        if (tid == 0 && (k > 0 || (kPeers && pr.wait_first))) {
            bool ok = true;
            if (kPeers) for (uint32_t g = 0; g < pr.world; g++) ok &= wait_flag_geq<kAcqSys>(pr.arrival + g, pr.seq0 + k);
            else ok = grid_wait(it.barrier, 2u * k * gridDim.x);
            if (!ok) {
                abort_flag = 1u;
                dead = 1u;
                raise_error(p.error_flag);
            }
        }
        for (uint32_t g = g0; g < g1; g++) {
            const Segment *sg = p.segs + g;
            const uint4 h0 = __ldg(reinterpret_cast<const uint4 *>(sg));        // tile, t_lo, t_hi, col_base
            const uint4 h1 = __ldg(reinterpret_cast<const uint4 *>(sg) + 1);    // col_count, slice_begin, n_slices, step_begin
            const uint32_t cnt = __ldg(&sg->cnt_ge[lane]);
            if (tid == 0) {
                // the vector was stored by other SMs through the generic proxy and ordered by the barrier's acquire:
                // the bulk copy below reads it through the async proxy
                asm volatile("fence.proxy.async.global;" ::: "memory");
                fence_proxy_async();
                const uint32_t bytes = h1.x * 4u;
                mbar_arrive_expect_tx(&bar, bytes);
                const unsigned char *src = reinterpret_cast<const unsigned char *>(x + h0.w);
                const uint32_t pieces = (bytes + kBulkPiece - 1) / kBulkPiece;
                uint32_t q = blockIdx.x % pieces;
                for (uint32_t i = 0; i < pieces; i++) {
                    const uint32_t off = q * kBulkPiece;
                    bulk_g2s(smem_raw + kXTileOffset + off, src + off, min(kBulkPiece, bytes - off), &bar);
                    q = q + 1 == pieces ? 0 : q + 1;
                }
            }
            const uint32_t ta = __ldg(&sg->warp_t[warp]), tb = __ldg(&sg->warp_t[warp + 1]);
            const uint32_t first_slice = __ldg(&sg->warp_slice[warp]);
            const uint32_t comb_first = __ldg(&sg->comb_first), comb_n = comb_base ? __ldg(&sg->comb_n) : 0u;
            acc_t *comb = comb_n ? comb_base + ((g - g0) & 1u) * (kCombineSlots * kLanes) : nullptr;
            if (kNarrow) stream_units_narrow<A>(p, &bar, parity, cnt, h1.w, ta, tb, first_slice, lane, false, &abort_flag);
            else (void)stream_steps<A>(p, &bar, parity, cnt, h1.y, h1.z, h1.w, ta, tb, first_slice, lane, false, &abort_flag,
                                       comb, comb_first, warp < comb_n);
            parity ^= 1u;
            __syncthreads();                                       // everyone is done with this x tile
            if (comb)
                for (uint32_t s = warp; s < comb_n; s += kWarps) {
                    const acc_t v = comb[s * kLanes + lane];
                    comb[s * kLanes + lane] = acc_t(0);
                    if (abort_flag) continue;
                    const uint32_t row = __ldg(p.slice_rows + (size_t)(h1.y + comb_first + s) * kLanes + lane);
                    if (__all_sync(0xFFFFFFFFu, row == __shfl_sync(0xFFFFFFFFu, row, 0))) {
                        const acc_t t = A::warp_sum(v);
                        if (lane == 0) A::emit(p.acc, row, t);
                    } else {
                        A::emit(p.acc, row, v);
                    }
                }
        }
        // every row update of this iteration, grid-wide, before anybody reads a row sum
        grid_arrive(it.barrier);
        if (tid == 0 && !dead && !grid_wait(it.barrier, (2u * k + 1u) * gridDim.x)) {
            dead = 1u;
            raise_error(p.error_flag);
        }
        __syncthreads();
        if (dead) break;
        if (kPeers) {
            const size_t wb = (size_t)((pr.buf0 + k + 1u) & 3u) * pr.x_stride + pr.col_offset;
            for (uint32_t r = blockIdx.x * kThreads + tid; r < it.rows; r += gridDim.x * kThreads) {
                const uint32_t v = A::drain(p.acc, r);
                p.y[r] = v;
                if (pr.col_offset + r < it.x_limit) {
                    const uint32_t w = A::axpb(it.alpha, v, it.beta);
                    for (uint32_t g = 0; g < pr.world; g++) pr.x_base[g][wb + r] = w;      // NVLink peer stores, coalesced
                }
            }
            if (blockIdx.x == 0 && tid == 0) (void)A::drain(p.acc, p.trash_row);
            // second barrier, arrive half, at system scope; the last CTA of the grid publishes the slice on every rank
            __syncthreads();
            if (tid == 0) {
                fence_sys();
                if (atomicAdd(it.barrier, 1u) + 1u == (2u * k + 2u) * gridDim.x) {
                    fence_sys();
                    for (uint32_t g = 0; g < pr.world; g++)
                        publish_flag(pr.flag[g], pr.seq0 + k + 1u);
                }
            }
        } else {
            drain_axpb_rows<A>(p.acc, p.y, (k & 1u) ? it.x0 : it.x1, it.rows, it.x_limit, it.alpha, it.beta,
                               blockIdx.x * kThreads + tid, gridDim.x * kThreads);
            if (blockIdx.x == 0 && tid == 0) (void)A::drain(p.acc, p.trash_row);
            if (k + 1u < it.iters) grid_arrive(it.barrier);        // waited for at the top of the next iteration
        }
    }
    // an ordinary launch: its predecessor was complete before it started
    if (blockIdx.x == 0 && tid == 0) {
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.done_dev), "r"(p.seq - 1u) : "memory");
        if (p.done_seq) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p.done_seq), "r"(p.seq - 1u) : "memory");
    }
}

}  // namespace

cudaError_t configure_iterate_kernels() {
    const void *kernels[] = {(const void *)spmv_iterate_kernel<FixedArith, false, false>, (const void *)spmv_iterate_kernel<FloatArith, false, false>,
                             (const void *)spmv_iterate_kernel<FixedArith, true, false>, (const void *)spmv_iterate_kernel<FloatArith, true, false>,
                             (const void *)spmv_iterate_kernel<FixedArith, false, true>, (const void *)spmv_iterate_kernel<FloatArith, false, true>,
                             (const void *)spmv_iterate_kernel<FixedArith, true, true>, (const void *)spmv_iterate_kernel<FloatArith, true, true>};
    for (const void *k : kernels) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_iterate(int arith, const SpmvParams &p, const IterateParams &it, const IteratePeers *peers, int grid,
                           uint32_t smem_bytes, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;           // the grid barriers need every CTA resident
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    IteratePeers none = {};
    const IteratePeers &pr = peers ? *peers : none;
    const bool fixed = arith == kArithFixed;
#define HSB_ITER_LAUNCH(N, P)                                                                                   \
    (fixed ? cudaLaunchKernelEx(&cfg, spmv_iterate_kernel<FixedArith, N, P>, p, it, pr)                         \
           : cudaLaunchKernelEx(&cfg, spmv_iterate_kernel<FloatArith, N, P>, p, it, pr))
    if (peers) return p.narrow ? HSB_ITER_LAUNCH(true, true) : HSB_ITER_LAUNCH(false, true);
    return p.narrow ? HSB_ITER_LAUNCH(true, false) : HSB_ITER_LAUNCH(false, false);
#undef HSB_ITER_LAUNCH
}

