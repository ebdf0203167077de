// sm_100a SpMV kernels. See spmv_kernels.cuh for the mapping to the reference's units.
#include "spmv_kernels.cuh"

namespace hsb {
namespace {

// ---------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (cp.async.bulk -> UBLKCP), streaming 128-bit loads
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// generic-proxy accesses to shared memory before this point are ordered before later
// async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_desc(const ChunkDesc *p) {
    uint2 r;
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// ---------------------------------------------------------------------------------------
// arithmetic policies
// ---------------------------------------------------------------------------------------
struct FixedArith {                                   // ap_ufixed<32,8,AP_RND,AP_SAT>
    typedef unsigned long long acc_t;
    static __device__ __forceinline__ acc_t zero() { return 0ull; }
    static __device__ __forceinline__ acc_t mac(acc_t acc, uint32_t a, uint32_t x) {
        unsigned long long p = (unsigned long long)a * (unsigned long long)x;      // exact, 48 fraction bits
        p = (p + (1ull << 23)) >> 24;                                              // AP_RND
        p = (p >> 32) ? 0xFFFFFFFFull : p;                                         // AP_SAT of the product
        return acc + p;                                                            // exact; clamp deferred
    }
    static __device__ __forceinline__ acc_t add(acc_t a, acc_t b) { return a + b; }
    static __device__ __forceinline__ void emit(void *acc, uint32_t row, acc_t v) {
        if (v) atomicAdd(reinterpret_cast<unsigned long long *>(acc) + row, v);    // RED.ADD.64
    }
};
struct FloatArith {                                   // fp32 multiply, then fp32 add (not fused)
    typedef float acc_t;
    static __device__ __forceinline__ acc_t zero() { return 0.0f; }
    static __device__ __forceinline__ acc_t mac(acc_t acc, uint32_t a, uint32_t x) {
        return __fadd_rn(acc, __fmul_rn(__uint_as_float(a), __uint_as_float(x)));
    }
    static __device__ __forceinline__ acc_t add(acc_t a, acc_t b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ void emit(void *acc, uint32_t row, acc_t v) {
        if (v != 0.0f) atomicAdd(reinterpret_cast<float *>(acc) + row, v);         // RED.ADD.F32
    }
};

struct ChunkRegs {
    uint4 v0, v1;      // 8 value words of this lane
    uint4 ci;          // 8 x (local column | end-of-segment flag)
    uint2 desc;        // ChunkDesc
};

__device__ __forceinline__ void load_chunk(const SpmvParams &p, uint32_t ch, uint32_t lane, ChunkRegs &r) {
    const uint4 *vp = reinterpret_cast<const uint4 *>(p.vals + (size_t)ch * kChunkNnz);
    r.v0 = ldg_stream(vp + lane);
    r.v1 = ldg_stream(vp + kLanes + lane);
    r.ci = ldg_stream(reinterpret_cast<const uint4 *>(p.cidx + (size_t)ch * kChunkNnz) + lane);
    r.desc = ldg_desc(p.chunks + ch);
}

template <class A>
__device__ __forceinline__ void process_chunk(const SpmvParams &p, const uint32_t *xs, const ChunkRegs &r,
                                              uint32_t lane) {
    typedef typename A::acc_t acc_t;
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t w[kNnzPerLane] = {r.ci.x & 0xFFFFu, r.ci.x >> 16, r.ci.y & 0xFFFFu, r.ci.y >> 16,
                                     r.ci.z & 0xFFFFu, r.ci.z >> 16, r.ci.w & 0xFFFFu, r.ci.w >> 16};
    const uint32_t v[kNnzPerLane] = {r.v0.x, r.v0.y, r.v0.z, r.v0.w, r.v1.x, r.v1.y, r.v1.z, r.v1.w};

    // gather x from the shared-memory tile (the vecbuf_reader step)
    uint32_t xv[kNnzPerLane];
#pragma unroll
    for (int k = 0; k < kNnzPerLane; k++) xv[k] = xs[w[k] & 0x7FFFu];

    // how many segments end in lanes before this one
    uint32_t nfl = 0;
#pragma unroll
    for (int k = 0; k < kNnzPerLane; k++) nfl += w[k] >> 15;
    uint32_t incl = nfl;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, incl, d);
        if (lane >= (uint32_t)d) incl += t;
    }
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    uint32_t seg = r.desc.x + incl - nfl;             // segment the lane's first flag closes
    const uint32_t first_seg = seg;

    // lane-local multiply-accumulate; segments that start AND end inside the lane are emitted at once
    acc_t run = A::zero(), head = A::zero();
    bool seen = false;
#pragma unroll
    for (int k = 0; k < kNnzPerLane; k++) {
        run = A::mac(run, v[k], xv[k]);
        if (w[k] & kSegEndFlag) {
            if (!seen) {
                head = run;
                seen = true;
            } else {
                A::emit(p.acc, __ldg(p.seg_row + seg), run);
            }
            seg++;
            run = A::zero();
        }
    }

    // warp segmented scan of the open tails: I(l) = sum of tails from the nearest flagged lane
    // at or below l (or lane 0) up to l
    const uint32_t fmask = __ballot_sync(FULL, seen);
    const uint32_t below = fmask & (FULL >> (31 - lane));
    const uint32_t dist = below ? lane - (31 - __clz(below)) : lane;
    acc_t I = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        acc_t t = __shfl_up_sync(FULL, I, d);
        if (dist >= (uint32_t)d) I = A::add(t, I);
    }
    acc_t carry = __shfl_up_sync(FULL, I, 1);
    if (lane == 0) carry = A::zero();
    if (seen) A::emit(p.acc, __ldg(p.seg_row + first_seg), A::add(carry, head));
    // what is left open at the end of the chunk belongs to the next chunk's first segment
    if (lane == 31 && (r.desc.y & kChunkContinues)) A::emit(p.acc, __ldg(p.seg_row + r.desc.x + total), I);
}

template <class A>
__global__ void __launch_bounds__(kThreads, 1) spmv_tiles_kernel(const SpmvParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *xs = reinterpret_cast<uint32_t *>(smem_raw);
    __shared__ __align__(8) uint64_t bar;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n = p.chunk_end - p.chunk_begin;
    const uint32_t c0 = p.chunk_begin + (uint32_t)(((unsigned long long)n * blockIdx.x) / gridDim.x);
    const uint32_t c1 = p.chunk_begin + (uint32_t)(((unsigned long long)n * (blockIdx.x + 1)) / gridDim.x);
    if (c0 >= c1) return;

    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();

    uint32_t parity = 0;
    uint32_t c = c0;
    uint32_t tile = __ldg(&p.chunks[c0].tile) & ~kChunkContinues;
    while (c < c1) {
        const TileDesc td = p.tiles[tile];
        if (td.chunk_end <= c) { tile++; continue; }          // empty tile
        const uint32_t sub_end = min(c1, td.chunk_end);

        // stage the x tile: the vector loader + vecbuf writer of the reference
        if (tid == 0) {
            fence_proxy_async();
            const uint32_t bytes = td.col_count * 4u;
            mbar_arrive_expect_tx(&bar, bytes);
            const unsigned char *src = reinterpret_cast<const unsigned char *>(p.x + td.col_base);
            for (uint32_t off = 0; off < bytes; off += kBulkPiece)
                bulk_g2s(smem_raw + off, src + off, min(kBulkPiece, bytes - off), &bar);
        }

        // the matrix stream does not depend on x: issue the first loads before waiting for the tile
        uint32_t ch = c + warp;
        ChunkRegs cur;
        if (ch < sub_end) load_chunk(p, ch, lane, cur);
        mbar_wait(&bar, parity);
        parity ^= 1u;
        while (ch < sub_end) {
            const uint32_t nx = ch + kWarps;
            ChunkRegs nxt;
            if (nx < sub_end) load_chunk(p, nx, lane, nxt);
            process_chunk<A>(p, xs, cur, lane);
            cur = nxt;
            ch = nx;
        }
        c = sub_end;
        tile++;
        __syncthreads();                                       // everyone is done with this x tile
    }
}

__global__ void finalize_fixed_kernel(const unsigned long long *__restrict__ acc, uint32_t *__restrict__ y,
                                      uint32_t row_begin, uint32_t row_end) {
    uint32_t r = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (r < row_end) {
        unsigned long long a = acc[r];
        y[r] = (a >> 32) ? 0xFFFFFFFFu : (uint32_t)a;          // AP_SAT of the accumulator (pe.h:72)
    }
}

}  // namespace

cudaError_t configure_kernels() {
    cudaError_t e = cudaFuncSetAttribute(spmv_tiles_kernel<FixedArith>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kXTileBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(spmv_tiles_kernel<FloatArith>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)kXTileBytes);
}

void launch_spmv_tiles(int arith, const SpmvParams &p, int grid, cudaStream_t stream) {
    if (p.chunk_end <= p.chunk_begin) return;
    if (arith == kArithFixed)
        spmv_tiles_kernel<FixedArith><<<grid, kThreads, kXTileBytes, stream>>>(p);
    else
        spmv_tiles_kernel<FloatArith><<<grid, kThreads, kXTileBytes, stream>>>(p);
}

void launch_finalize_fixed(const unsigned long long *acc, uint32_t *y, uint32_t row_begin, uint32_t row_end,
                           cudaStream_t stream) {
    if (row_end <= row_begin) return;
    uint32_t n = row_end - row_begin;
    finalize_fixed_kernel<<<(n + 255) / 256, 256, 0, stream>>>(acc, y, row_begin, row_end);
}

}  // namespace hsb
