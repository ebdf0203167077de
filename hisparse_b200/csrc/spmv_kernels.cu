// sm_100a SpMV kernel. See spmv_kernels.cuh for the mapping to the reference's units.
#include "spmv_kernels.cuh"

#include <algorithm>
#include <cstdlib>

namespace hsb {
namespace {

// ---------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (cp.async.bulk -> UBLKCP), streaming vector loads
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
// L2 residency: the matrix stream is read exactly once per launch, so it is marked evict-first and leaves
// the L2 to what IS reused -- the x tiles (read by every CTA of a tile) and above all the row accumulators,
// which the row updates hit again and again (hypersparse shards: 50-100 MB of accumulators against a 4 GB
// stream). HSB_L2_HINTS=0 builds without the hints.
#ifndef HSB_L2_HINTS
#define HSB_L2_HINTS 1
#endif
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg_stream128(const uint4 *p) {
    uint4 r;
#if HSB_L2_HINTS
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(policy_evict_first()));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#endif
    return r;
}
__device__ __forceinline__ uint2 ldg_stream64(const uint2 *p) {
    uint2 r;
#if HSB_L2_HINTS
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;"
                 : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(policy_evict_first()));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
#endif
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream32(const uint32_t *p) {
    uint32_t r;
#if HSB_L2_HINTS
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(policy_evict_first()));
#else
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
#endif
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream16(const uint16_t *p) {
    uint16_t r;
#if HSB_L2_HINTS
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(r) : "l"(p), "l"(policy_evict_first()));
#else
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
#endif
    return r;
}
// Bounded spin on a sequence flag another engine writes (a stream memory operation on a copy stream, a peer
// GPU's kernel): true once (int)(flag - val) >= 0. kFlagPolls polls of ~1 us, then give up rather than hang
// the GPU: the caller raises the error flag and SKIPS its work, so a stale vector is never multiplied.
#ifndef HSB_FLAG_POLLS
#define HSB_FLAG_POLLS (8u << 20)
#endif
constexpr uint32_t kFlagPolls = HSB_FLAG_POLLS;
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// A flag wait timed out. (An atomic exchange, although every writer stores the same 1 and the word may live in mapped
// page-locked host memory: a plain `st.relaxed.sys` in its place made the fixed-point kernel 2.6 us slower per SpMV
// on C2 -- 16.56 against 13.91 us, same box, the store never executed -- one of the code-generation cliffs of a
// kernel that sits exactly at its 64-register limit.)
__device__ __forceinline__ void raise_error(uint32_t *flag) { atomicExch(flag, 1u); }
// kAcqNone: relaxed polls. kAcqSys / kAcqGpu: every poll is an acquire load at that scope, so the successful one
// synchronises with the writer's release (st.release.sys of a peer GPU's kernel, the copy engine's flag write
// behind its copy, st.release.gpu of an earlier launch): everything the writer did before raising the flag is
// visible to what this thread -- and, through the CTA's mbarrier, every thread it releases -- does afterwards.
// (An acquire LOAD, not a relaxed load followed by fence.acq_rel.sys: the fence also has to wait until this SM's
// own outstanding system-scope writes -- the previous drain's posted PCIe writes among them -- have been
// acknowledged, which cost 1.7 us per launch device-resident and 9 us with host buffers; HSB_ACQ_FENCE=1 builds
// that variant for A/B measurements.)
enum { kAcqNone = 0, kAcqSys = 1, kAcqGpu = 2 };
#ifndef HSB_ACQ_FENCE
#define HSB_ACQ_FENCE 0
#endif
template <int kAcq>
__device__ __forceinline__ bool wait_flag_geq(const uint32_t *flag, uint32_t val) {
#pragma unroll 1
    for (uint32_t i = 0; i < kFlagPolls; i++) {
        uint32_t v;
        if (kAcq == kAcqNone || HSB_ACQ_FENCE) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        else if (kAcq == kAcqSys) asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        else asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - val) >= 0) {
#if HSB_ACQ_FENCE
            if (kAcq != kAcqNone) asm volatile("fence.acq_rel.sys;" ::: "memory");
#endif
            return true;
        }
        __nanosleep(i < 64 ? 100 : 1000);
    }
    return false;
}
__device__ __forceinline__ bool wait_flag_geq(const uint32_t *flag, uint32_t val, bool acquire) {
    return acquire ? wait_flag_geq<kAcqSys>(flag, val) : wait_flag_geq<kAcqNone>(flag, val);
}
// An arrival flag on a peer GPU (or this one), written AFTER a system-scope fence that ordered the data stores: fence +
// relaxed store is a release pattern. (`st.release.sys` per target instead makes every store wait for the previous one's
// NVLink round trip: the publication then costs one round trip per rank -- HSB_FLAG_RELEASE_STORES=1 builds that form.)
#ifndef HSB_FLAG_RELEASE_STORES
#define HSB_FLAG_RELEASE_STORES 0
#endif
__device__ __forceinline__ void publish_flag(uint32_t *flag, uint32_t seq) {
#if HSB_FLAG_RELEASE_STORES
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
#else
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
#endif
}
// Grid-wide completion of a drain with a gather epilogue: the CTA that takes the last ticket knows every
// peer store of this drain has been issued and fenced, and raises this rank's arrival flag on every target.
__device__ __forceinline__ void gather_publish(const GatherTargets *gt, uint32_t seq) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(gt->ticket, 1u) == gridDim.x - 1) {
            *gt->ticket = 0u;
            __threadfence_system();
            for (int g = 0; g < gt->n; g++)
                publish_flag(gt->flag[g], seq);
        }
    }
}
// ---------------------------------------------------------------------------------------
// arithmetic policies
// ---------------------------------------------------------------------------------------
struct FixedArith {                                   // ap_ufixed<32,8,AP_RND,AP_SAT>
    typedef unsigned long long acc_t;
    // a lane stream has <= 128 terms: the high part (<= 2^24 each) and the low part (< 2^8 each)
    // of the rounded products are summed separately in 32 bits and recombined exactly
    uint32_t hi, lo;
    __device__ __forceinline__ void clear() { hi = 0; lo = 0; }
    __device__ __forceinline__ void mac(uint32_t a, uint32_t x) {
        unsigned long long q = (unsigned long long)a * x + 0x800000ull;   // exact product + half LSB (AP_RND)
        // (q >> 24) = (q.hi << 8) + (q.lo >> 24). A product that saturates (AP_SAT: q.hi >= 2^24) is
        // replaced by a value >= 2^32, which forces the drain's clamp -- exactly what saturation means
        // for a sum of non-negative terms.
        hi += min((uint32_t)(q >> 32), 0x1000000u);
        lo += (uint32_t)q >> 24;
    }
    __device__ __forceinline__ acc_t total() const { return ((acc_t)hi << 8) + lo; }
    static __device__ __forceinline__ void emit(void *acc, uint32_t row, acc_t v) {
        if (!v) return;
#if HSB_L2_HINTS
        asm volatile("red.global.add.L2::cache_hint.u64 [%0], %1, %2;"
                     ::"l"(reinterpret_cast<unsigned long long *>(acc) + row), "l"(v), "l"(policy_evict_last()) : "memory");
#else
        atomicAdd(reinterpret_cast<unsigned long long *>(acc) + row, v);           // RED.ADD.64
#endif
    }
    static __device__ __forceinline__ acc_t warp_sum(acc_t v) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
        return v;
    }
    static __device__ __forceinline__ void add_shared(acc_t *slot, acc_t v) { if (v) atomicAdd(slot, v); }
    // alpha (*) y (+) beta with the PE's product rounding / saturation (pe.h:64) and saturating add (pe.h:72)
    static __device__ __forceinline__ uint32_t axpb(uint32_t alpha, uint32_t y, uint32_t beta) {
        unsigned long long q = ((unsigned long long)alpha * y + 0x800000ull) >> 24;
        q = min(q, 0xFFFFFFFFull) + beta;
        return q > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)q;
    }
    static __device__ __forceinline__ uint32_t drain(void *acc, uint32_t row) {
        unsigned long long *p = reinterpret_cast<unsigned long long *>(acc) + row;
        unsigned long long a = __ldcg(p);
        *p = 0ull;
        return (a >> 32) ? 0xFFFFFFFFu : (uint32_t)a;                              // AP_SAT (pe.h:72)
    }
    // four consecutive rows (row % 4 == 0) for drain_rows: raw accumulator words, their re-zeroing, the result words
    struct Raw4 { ulonglong2 a, b; };
    static __device__ __forceinline__ Raw4 load4(const void *acc, uint32_t row) {
        const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(reinterpret_cast<const unsigned long long *>(acc) + row);
        Raw4 r;
        r.a = __ldcg(p);
        r.b = __ldcg(p + 1);
        return r;
    }
    static __device__ __forceinline__ void zero4(void *acc, uint32_t row, bool stream) {
        ulonglong2 *p = reinterpret_cast<ulonglong2 *>(reinterpret_cast<unsigned long long *>(acc) + row);
        const ulonglong2 z = make_ulonglong2(0ull, 0ull);
        if (stream) { __stcs(p, z); __stcs(p + 1, z); } else { p[0] = z; p[1] = z; }
    }
    static __device__ __forceinline__ uint4 final4(const Raw4 &r) {
        auto sat = [](unsigned long long a) { return (a >> 32) ? 0xFFFFFFFFu : (uint32_t)a; };    // AP_SAT (pe.h:72)
        return make_uint4(sat(r.a.x), sat(r.a.y), sat(r.b.x), sat(r.b.y));
    }
};
struct FloatArith {                                   // fp32 multiply, then fp32 add (not fused)
    typedef float acc_t;
    float s;
    __device__ __forceinline__ void clear() { s = 0.0f; }
    __device__ __forceinline__ void mac(uint32_t a, uint32_t x) {
        s = __fadd_rn(s, __fmul_rn(__uint_as_float(a), __uint_as_float(x)));
    }
    __device__ __forceinline__ acc_t total() const { return s; }
    static __device__ __forceinline__ void emit(void *acc, uint32_t row, acc_t v) {
#if HSB_L2_HINTS
        asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;"
                     ::"l"(reinterpret_cast<float *>(acc) + row), "f"(v), "l"(policy_evict_last()) : "memory");
#else
        atomicAdd(reinterpret_cast<float *>(acc) + row, v);                        // RED.ADD.F32
#endif
    }
    static __device__ __forceinline__ acc_t warp_sum(acc_t v) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
        return v;
    }
    static __device__ __forceinline__ void add_shared(acc_t *slot, acc_t v) { atomicAdd(slot, v); }
    static __device__ __forceinline__ uint32_t axpb(uint32_t alpha, uint32_t y, uint32_t beta) {
        return __float_as_uint(__fadd_rn(__fmul_rn(__uint_as_float(alpha), __uint_as_float(y)), __uint_as_float(beta)));
    }
    static __device__ __forceinline__ uint32_t drain(void *acc, uint32_t row) {
        float *p = reinterpret_cast<float *>(acc) + row;
        float a = __ldcg(p);
        *p = 0.0f;
        return __float_as_uint(a);
    }
    typedef uint4 Raw4;
    static __device__ __forceinline__ Raw4 load4(const void *acc, uint32_t row) {
        return __ldcg(reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(acc) + row));
    }
    static __device__ __forceinline__ void zero4(void *acc, uint32_t row, bool stream) {
        uint4 *p = reinterpret_cast<uint4 *>(reinterpret_cast<uint32_t *>(acc) + row);
        if (stream) __stcs(p, make_uint4(0u, 0u, 0u, 0u)); else *p = make_uint4(0u, 0u, 0u, 0u);
    }
    static __device__ __forceinline__ uint4 final4(const Raw4 &r) { return r; }
};

// The result drain proper (pe dump + result_packer + axis_merge + spmv_result_drain of the reference, pe.h:95-116,
// spmv_cluster.h:133-193, stream_utils.h:36-75, spmv_result_drain.cpp:36-113, and the PE's reset loop pe.h:131-135):
// y[r] = clamp(acc[r]); acc[r] = 0 for r in [begin, end), by `n_threads` threads of which this is `t`. The result
// words also go to the mapped host buffer of a deferred download and to the gathered vectors of the target ranks.
// Four rows per 128-bit access and kDrainUnroll independent accesses in flight per thread: with one 4-byte load
// per thread and iteration the drain of a 12.5 M-row shard was latency bound at 0.4 TB/s and took 580 us of a
// 1.42 ms SpMV. `stream` (launches whose accumulator buffers exceed the L2, sync_start): the re-zeroing stores
// and y are written with the streaming (evict-first) policy, so that the drained buffer stops competing for the
// L2 with the live one.
constexpr int kDrainUnroll = 4;
template <class A>
__device__ __forceinline__ void drain_rows(void *acc, uint32_t *y, uint32_t begin, uint32_t end, uint32_t *y_host,
                                           uint32_t y_host_rows, const GatherTargets *gt, bool stream, uint32_t t,
                                           uint32_t n_threads) {
    const int n_targets = gt ? gt->n : 0;
    auto one = [&](uint32_t r) {
        const uint32_t v = A::drain(acc, r);
        y[r] = v;
        if (y_host && r < y_host_rows) __stcs(y_host + r, v);          // posted write over PCIe
        for (int g = 0; g < n_targets; g++) gt->y[g][r] = v;           // NVLink peer stores
    };
    // 16-byte accesses need 4-aligned rows in every destination: the buffers themselves are 256-byte aligned
    // (cudaMalloc), a gather target starts at this rank's row offset
    // ... and only pay off when a thread has several rows to drain: with at most one row per thread the scalar
    // loop keeps every SM busy and is the shorter path on the completion chain of small launches (C1, C3: +0.5 us
    // per SpMV with the vector loop; C2 with host buffers 18.3 -> 22.5 us, its posted writes coming from 26 SMs only)
    bool vec = end - begin > n_threads && (reinterpret_cast<uintptr_t>(y_host) & 15u) == 0;
    for (int g = 0; g < n_targets; g++) vec &= (reinterpret_cast<uintptr_t>(gt->y[g]) & 15u) == 0;
    const uint32_t b4 = vec ? min(end, (begin + 3u) & ~3u) : end;
    const uint32_t e4 = b4 + ((end - b4) & ~3u);
    for (uint32_t r = begin + t; r < b4; r += n_threads) one(r);
    for (uint32_t r = e4 + t; r < end; r += n_threads) one(r);
    const uint32_t n_chunks = (e4 - b4) >> 2;
    for (uint32_t c0 = t; c0 < n_chunks; c0 += n_threads * kDrainUnroll) {
        typename A::Raw4 raw[kDrainUnroll];
#pragma unroll
        for (int u = 0; u < kDrainUnroll; u++) {
            const uint32_t c = c0 + u * n_threads;
            if (c < n_chunks) raw[u] = A::load4(acc, b4 + 4u * c);
        }
#pragma unroll
        for (int u = 0; u < kDrainUnroll; u++) {
            const uint32_t c = c0 + u * n_threads;
            if (c >= n_chunks) break;
            const uint32_t r = b4 + 4u * c;
            A::zero4(acc, r, stream);
            const uint4 v = A::final4(raw[u]);
            if (stream) __stcs(reinterpret_cast<uint4 *>(y + r), v); else *reinterpret_cast<uint4 *>(y + r) = v;
            if (y_host) {
                if (r + 4u <= y_host_rows) __stcs(reinterpret_cast<uint4 *>(y_host + r), v);
                else {
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                    for (uint32_t k = 0; k < 4u && r + k < y_host_rows; k++) __stcs(y_host + r + k, w[k]);
                }
            }
            for (int g = 0; g < n_targets; g++) *reinterpret_cast<uint4 *>(gt->y[g] + r) = v;
        }
    }
}

// The CTA's dynamic shared memory: xs[0..7] = 0 (padding slots), xs[8 + c] = x[tile column c].
extern __shared__ __align__(128) uint32_t xs[];

// Gather of the vecbuf_reader step: x word of the low / high 16-bit column id packed in `c`.
// Written in PTX so that each address costs two ALU instructions (mask or shift, then one
// shift-add onto the shared base held in a register) instead of the three nvcc emits for xs[id].
__device__ __forceinline__ uint32_t gather_lo(uint32_t xs_base, uint32_t c) {
    uint32_t v;
    asm("{\n\t.reg .u32 t;\n\tand.b32 t, %1, 0xFFFF;\n\tmad.lo.u32 t, t, 4, %2;\n\tld.shared.u32 %0, [t];\n\t}"
        : "=r"(v) : "r"(c), "r"(xs_base));
    return v;
}
__device__ __forceinline__ uint32_t gather_hi(uint32_t xs_base, uint32_t c) {
    uint32_t v;
    asm("{\n\t.reg .u32 t;\n\tshr.u32 t, %1, 16;\n\tmad.lo.u32 t, t, 4, %2;\n\tld.shared.u32 %0, [t];\n\t}"
        : "=r"(v) : "r"(c), "r"(xs_base));
    return v;
}

template <class A>
__device__ __forceinline__ void mac4(A &acc, uint32_t xs_base, const uint4 &v, const uint2 &c) {
    acc.mac(v.x, gather_lo(xs_base, c.x));
    acc.mac(v.y, gather_hi(xs_base, c.x));
    acc.mac(v.z, gather_lo(xs_base, c.y));
    acc.mac(v.w, gather_hi(xs_base, c.y));
}

// The accumulator buffer of this launch was last used four launches ago and re-zeroed by the drain at the end of
// launch seq-3. In the steady state that launch is long gone; the guard only ever spins when several small
// launches are resident at once. Every warp checks for itself, before it waits for the x tile (so the poll
// overlaps the staging and nobody waits for anybody else's poll): lane 0 polls with acquire loads at GPU scope
// (the flag is a release store of an earlier launch on this GPU), the warp-wide shuffle orders the other lanes'
// row updates behind it. Launches with two buffers in rotation wait for their predecessor instead (sync_start).
// false: the wait timed out -- the caller makes no row update and the host hears of it.
__device__ __forceinline__ bool accumulators_ready(const SpmvParams &p, uint32_t lane) {
    if (p.sync_start) asm volatile("griddepcontrol.wait;" ::: "memory");
    uint32_t ok = 1u;
    if (p.guard_flag) {
        if (lane == 0) {
            ok = (p.acquire ? wait_flag_geq<kAcqGpu>(p.guard_flag, p.guard_val) : wait_flag_geq<kAcqNone>(p.guard_flag, p.guard_val)) ? 1u : 0u;
            if (!ok) raise_error(p.error_flag);
        }
        ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
    }
    return ok != 0u;
}

// Slice geometry of a tile from its 32-entry table (lane c holds cnt_ge[c], see TileDesc):
// first step of tile-relative slice i, and the balancing weight (steps + one unit per slice).
__device__ __forceinline__ uint32_t steps_before(uint32_t cnt, uint32_t i) {
    return __reduce_add_sync(0xFFFFFFFFu, min(i, cnt));
}
__device__ __forceinline__ uint32_t steps_of(uint32_t cnt, uint32_t i) {
    return __popc(__ballot_sync(0xFFFFFFFFu, cnt > i));
}
// One warp streams tile-relative steps [ta, tb): a flat, contiguous run of (512 B values + 256 B
// columns) steps that may start and end inside a slice. Lanes own lane streams; whenever a slice
// (or the run) ends, the lane's partial sum is added to its row and the accumulator restarts.
template <class A>
__device__ __forceinline__ bool stream_steps(const SpmvParams &p, uint64_t *bar, uint32_t parity,
                                             uint32_t cnt, uint32_t slice_begin, uint32_t n_slices,
                                             uint32_t step_begin, uint32_t ta, uint32_t tb, uint32_t first_slice,
                                             uint32_t lane, bool first_segment, const volatile uint32_t *abort_flag,
                                             typename A::acc_t *comb, uint32_t comb_first, bool comb_drainer) {
    uint32_t remaining = tb - ta;
    const size_t base = (size_t)(step_begin + ta) * kStepElems;
    const uint4 *vp = reinterpret_cast<const uint4 *>(p.vals + base) + lane;
    const uint2 *cp = reinterpret_cast<const uint2 *>(p.cols + base) + lane;

    // the matrix stream does not depend on x: fill the ring before waiting for the x tile
    uint4 vb[kPrefetch];
    uint2 cb[kPrefetch];
#pragma unroll
    for (int j = 0; j < kPrefetch; j++)
        if ((uint32_t)j < remaining) {
            vb[j] = ldg_stream128(vp + j * kLanes);
            cb[j] = ldg_stream64(cp + j * kLanes);
        }
    vp += kPrefetch * kLanes;
    cp += kPrefetch * kLanes;

    // row ids of the current slice and of the next kRowAhead slices (short slices -- hypersparse tiles --
    // finish every step or two, so one slice of look-ahead would expose the load latency)
    uint32_t sl = 0, left = 0, row = 0, row_next[kRowAhead] = {};
    const uint32_t *rp = p.slice_rows;
    if (remaining) {
        sl = first_slice;                                     // the slice that contains step ta (host plan)
        left = steps_before(cnt, sl + 1) - ta;                // steps of slice sl still ahead of us
        rp = p.slice_rows + (size_t)(slice_begin + sl) * kLanes + lane;
        row = __ldg(rp);
#pragma unroll
        for (int a = 0; a < kRowAhead; a++)
            if (sl + 1 + a < n_slices) row_next[a] = __ldg(rp + (1 + a) * kLanes);
    }

    // (first segment) the accumulator buffer must be ours before the first row update: every warp checks the
    // guard itself, while the x tile is still on its way (see the kernel)
    // (warps without work -- and without a share of the combining table to flush -- make no row update)
    const bool guard_ok = !first_segment || !(remaining || comb_drainer) || accumulators_ready(p, lane);
    mbar_wait(bar, parity);
    if (p.timeline && first_segment && blockIdx.x == 0 && threadIdx.x == 0) p.timeline[(size_t)(p.seq & 255u) * 8 + 4] = globaltimer();
    if (!guard_ok) return false;
    if (!remaining || *abort_flag) return true;             // a flag wait timed out: no row update from stale data

    uint32_t xs_base = smem_u32(xs);
    asm volatile("" : "+r"(xs_base));                        // keep it in a register: no per-step rematerialisation
    A acc;
    acc.clear();
    // Row update of a finished slice. Streams of one row sit in adjacent lanes (stable sort), and
    // the pieces of a long row fill whole slices: then one warp reduction + one atomic replaces 32
    // same-address atomics (which the L2 would serialise).
    auto flush = [&]() {
        typename A::acc_t v = acc.total();
        if (comb) {                                          // combined in shared memory, flushed once per segment (see the kernel)
            A::add_shared(comb + (sl - comb_first) * kLanes + lane, v);
            return;
        }
        if (__all_sync(0xFFFFFFFFu, row == __shfl_sync(0xFFFFFFFFu, row, 0))) {
            v = A::warp_sum(v);
            if (lane == 0) A::emit(p.acc, row, v);
        } else {
            A::emit(p.acc, row, v);
        }
    };
    auto step_done = [&]() {
        if (--left == 0) {                                   // warp-uniform: the slice is complete
            flush();
            acc.clear();
            sl++;
            rp += kLanes;
            row = row_next[0];
#pragma unroll
            for (int a = 0; a + 1 < kRowAhead; a++) row_next[a] = row_next[a + 1];
            if (sl + kRowAhead < n_slices) row_next[kRowAhead - 1] = __ldg(rp + kRowAhead * kLanes);
            if (sl + 8 + kRowAhead < n_slices && lane == 0)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (8 + kRowAhead) * kLanes));
            left = steps_of(cnt, sl);
        }
    };
    while (remaining >= (uint32_t)kPrefetch) {
#pragma unroll
        for (int j = 0; j < kPrefetch; j++) {
            mac4(acc, xs_base, vb[j], cb[j]);
            if (remaining > (uint32_t)(kPrefetch + j)) {     // step (consumed + kPrefetch + j) exists
                vb[j] = ldg_stream128(vp + j * kLanes);
                cb[j] = ldg_stream64(cp + j * kLanes);
            }
            step_done();
        }
        vp += kPrefetch * kLanes;
        cp += kPrefetch * kLanes;
        remaining -= kPrefetch;
    }
#pragma unroll
    for (int j = 0; j < kPrefetch - 1; j++)
        if ((uint32_t)j < remaining) {
            mac4(acc, xs_base, vb[j], cb[j]);
            step_done();
        }
    // the run ended inside a slice: hand over what has been accumulated so far
    if (left != steps_of(cnt, sl) && sl < n_slices) flush();
    return true;
}

// Narrow layout (hypersparse matrices, tile_format.h): one warp streams units [ta, tb) of a tile, 32 elements
// each (128 B of value words + 64 B of column ids). A slice is a ROW UNIT -- lane l's value word is the matrix
// row of its stream -- followed by L(slice) step units, one slot per lane; shares begin and end at slice
// boundaries. Row ids, values and column ids all arrive through the same kNarrowRing-deep register ring, so the
// latency of none of them is exposed, however short the slices are (the wide layout's separate row-id loads
// were 40 % of all stall samples on a C5 shard, whose slices are one or two steps long).
template <class A>
__device__ __forceinline__ void stream_units_narrow(const SpmvParams &p, uint64_t *bar, uint32_t parity, uint32_t cnt,
                                                    uint32_t unit_begin, uint32_t ta, uint32_t tb, uint32_t first_slice,
                                                    uint32_t lane, bool first_segment, const volatile uint32_t *abort_flag) {
    uint32_t remaining = tb - ta;
    const size_t base = (size_t)(unit_begin + ta) * kUnitElems + lane;
    const uint32_t *vp = p.vals + base;
    const uint16_t *cp = p.cols + base;
    uint32_t vb[kNarrowRing], cb[kNarrowRing];
#pragma unroll
    for (int j = 0; j < kNarrowRing; j++) {
        vb[j] = 0u;
        cb[j] = 0u;                                            // (slots beyond the share: a valid id for the look-ahead gather)
        if ((uint32_t)j < remaining) {
            vb[j] = ldg_stream32(vp + j * kUnitElems);
            cb[j] = ldg_stream16(cp + j * kUnitElems);
        }
    }
    vp += kNarrowRing * kUnitElems;
    cp += kNarrowRing * kUnitElems;

    const bool guard_ok = !first_segment || !remaining || accumulators_ready(p, lane);   // (warps without work make no row update)
    mbar_wait(bar, parity);
    if (p.timeline && first_segment && blockIdx.x == 0 && threadIdx.x == 0) p.timeline[(size_t)(p.seq & 255u) * 8 + 4] = globaltimer();
    if (!remaining || !guard_ok || *abort_flag) return;

    uint32_t xs_base = smem_u32(xs);
    asm volatile("" : "+r"(xs_base));
    A acc;
    acc.clear();
    uint32_t sl = first_slice, left = 0, row = 0;
    // x word of a column id; a row unit's ids are 0 and fetch the zero word in front of the tile, so the gather
    // can be issued for EVERY unit, one unit ahead of its use, before it is known what kind of unit it is
    auto gather = [&](uint32_t c) {
        uint32_t xv;
        asm volatile("{\n\t.reg .u32 t;\n\tmad.lo.u32 t, %1, 4, %2;\n\tld.shared.u32 %0, [t];\n\t}" : "=r"(xv) : "r"(c), "r"(xs_base));
        return xv;
    };
    bool full_len = false;                                     // the current slice has the maximum length: maybe pieces of ONE long row
    auto consume = [&](uint32_t v, uint32_t xv) {
        if (left == 0) {                                       // warp-uniform: a row unit opens slice sl
            row = v;
            left = steps_of(cnt, sl);
            full_len = left == kNarrowMaxLen;
            return;
        }
        acc.mac(v, xv);
        if (--left == 0) {                                     // the slice is complete: one row update per lane stream
            typename A::acc_t t = acc.total();
            // (a row cut into 32-entry streams fills whole slices of the maximum length: only those can be one row)
            if (full_len && __all_sync(0xFFFFFFFFu, row == __shfl_sync(0xFFFFFFFFu, row, 0))) {
                t = A::warp_sum(t);
                if (lane == 0) A::emit(p.acc, row, t);
            } else {
                A::emit(p.acc, row, t);
            }
            acc.clear();
            sl++;
        }
    };
    uint32_t xv_next = gather(cb[0]);
    // bulk of the share: every slot consumed is refilled, no per-unit bounds test
    while (remaining >= 2u * (uint32_t)kNarrowRing) {
#pragma unroll
        for (int j = 0; j < kNarrowRing; j++) {
            const uint32_t v = vb[j], xv = xv_next;
            vb[j] = ldg_stream32(vp + j * kUnitElems);
            cb[j] = ldg_stream16(cp + j * kUnitElems);
            xv_next = gather(cb[(j + 1) % kNarrowRing]);
            consume(v, xv);
        }
        vp += kNarrowRing * kUnitElems;
        cp += kNarrowRing * kUnitElems;
        remaining -= kNarrowRing;
    }
    while (remaining >= (uint32_t)kNarrowRing) {
#pragma unroll
        for (int j = 0; j < kNarrowRing; j++) {
            const uint32_t v = vb[j], xv = xv_next;
            if (remaining > (uint32_t)(kNarrowRing + j)) {
                vb[j] = ldg_stream32(vp + j * kUnitElems);
                cb[j] = ldg_stream16(cp + j * kUnitElems);
            }
            // the next unit sits in slot j + 1, or (j + 1 == ring size) in slot 0, reloaded at the top of this round;
            // when nothing follows the slot is stale and the word fetched is simply not used
            xv_next = gather(cb[(j + 1) % kNarrowRing]);
            consume(v, xv);
        }
        vp += kNarrowRing * kUnitElems;
        cp += kNarrowRing * kUnitElems;
        remaining -= kNarrowRing;
    }
#pragma unroll
    for (int j = 0; j < kNarrowRing - 1; j++)
        if ((uint32_t)j < remaining) {
            const uint32_t xv = xv_next;
            xv_next = gather(cb[j + 1]);
            consume(vb[j], xv);
        }
}

// kNarrow selects the layout at compile time: the two streaming loops live in separate kernels, so that each gets
// the whole 64-register budget of a 1024-thread CTA for its own prefetch ring
template <class A, bool kNarrow>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) spmv_tiles_kernel(const SpmvParams p) {
    unsigned char *smem_raw = reinterpret_cast<unsigned char *>(xs);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t abort_flag;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t g0 = __ldg(p.cta_seg + blockIdx.x), g1 = __ldg(p.cta_seg + blockIdx.x + 1);
    const long long t_start = clock64();
    unsigned long long *tl = p.timeline ? p.timeline + (size_t)(p.seq & 255u) * 8 : nullptr;
    if (tl && tid == 0) atomicMin(tl + 0, globaltimer());

    // Programmatic dependent launch, both ways round. (1) The successor may start as early as it finds a
    // free SM: nothing this launch still has to do can be disturbed by it (it accumulates into another
    // buffer, reads another x). (2) This launch does not wait for its predecessor before working either:
    // only the DRAIN of the predecessor's row sums needs the predecessor complete, so the
    // griddepcontrol.wait sits at the very end of the CTA, where it returns at once. (Waiting at the start
    // costs 3.6 us per launch in a device-resident loop and up to 10 us when PCIe copies are in flight:
    // the grid-completion flush the wait includes is slowed down by them -- measured with the
    // %globaltimer timeline.)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (g0 < g1) {
        if (tid == 0) {
            mbar_init(&bar, 1);                                   // arrive.expect_tx: the x tile's bytes
            abort_flag = 0u;
        }
        if (tid < kColBias) xs[tid] = 0u;                         // what padding slots multiply by
        typedef typename A::acc_t acc_t;
        acc_t *comb_base = (!kNarrow && p.comb_offset) ? reinterpret_cast<acc_t *>(smem_raw + p.comb_offset) : nullptr;
        if (comb_base)
            for (uint32_t i = tid; i < 2u * kCombineSlots * kLanes; i += kThreads) comb_base[i] = acc_t(0);
        __syncthreads();
        uint32_t parity = 0;
        for (uint32_t g = g0; g < g1; g++) {
            const Segment *sg = p.segs + g;
            const uint4 h0 = __ldg(reinterpret_cast<const uint4 *>(sg));        // tile, t_lo, t_hi, col_base
            const uint4 h1 = __ldg(reinterpret_cast<const uint4 *>(sg) + 1);    // col_count, slice_begin, n_slices, step_begin
            const uint32_t cnt = __ldg(&sg->cnt_ge[lane]);
            // stage the x tile: the vector loader + vecbuf writer of the reference
            if (tid == 0) {
                bool ok = true;
                if (g == g0 && p.wait_x_flag) {
                    for (uint32_t i = 0; i < p.wait_x_count; i++) ok &= wait_flag_geq(p.wait_x_flag + i, p.wait_x_val, p.acquire != 0);
                    // the vector was written through the generic proxy (peer SM stores) or by the copy engine;
                    // the bulk copy below reads it through the async proxy
                    if (p.acquire) asm volatile("fence.proxy.async.global;" ::: "memory");
                    if (!ok) {                                 // timed out: nobody multiplies the stale vector, the host hears of it
                        abort_flag = 1u;
                        raise_error(p.error_flag);
                    }
                }
                // x was written by the kernel in front of this one on the stream (the axpb step of an iterative caller,
                // itself a programmatic launch that lets this grid start early): wait for it here, ring already filling
                if (g == g0 && p.x_after_grid) {
                    asm volatile("griddepcontrol.wait;" ::: "memory");
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                }
                if (tl && blockIdx.x == 0 && g == g0) tl[1] = globaltimer();
                fence_proxy_async();
                const uint32_t bytes = h1.x * 4u;
                mbar_arrive_expect_tx(&bar, bytes);
                const unsigned char *src = reinterpret_cast<const unsigned char *>(p.x + h0.w);
                // Many CTAs stage the same tile at the same moment. Each starts at a different piece and wraps
                // around, so that at any instant they pull from different L2 slices instead of queueing on one.
                const uint32_t pieces = (bytes + kBulkPiece - 1) / kBulkPiece;
                uint32_t k = blockIdx.x % pieces;
                for (uint32_t i = 0; i < pieces; i++) {
                    const uint32_t off = k * kBulkPiece;
                    bulk_g2s(smem_raw + kXTileOffset + off, src + off, min(kBulkPiece, bytes - off), &bar);
                    k = k + 1 == pieces ? 0 : k + 1;
                }
            }
            // this warp's equal-cost share of the segment (host plan)
            const uint32_t ta = __ldg(&sg->warp_t[warp]), tb = __ldg(&sg->warp_t[warp + 1]);
            const uint32_t first_slice = __ldg(&sg->warp_slice[warp]);
            // Row updates of a segment that touches only a few slices are combined in shared memory first: the warps
            // that share a slice (the pruned transformer layers: one or two 32-step slices per CTA, split over 32
            // warps) add their partial sums into a table, and the table goes to the row accumulators once, after the
            // barrier -- 32 x fewer global row updates, all of which would hit the same few rows. Two tables alternate
            // by segment, so the flush of one overlaps the next segment's work.
            const uint32_t comb_first = __ldg(&sg->comb_first), comb_n = comb_base ? __ldg(&sg->comb_n) : 0u;
            acc_t *comb = comb_n ? comb_base + ((g - g0) & 1u) * (kCombineSlots * kLanes) : nullptr;
            bool ok = true;
            if (kNarrow) stream_units_narrow<A>(p, &bar, parity, cnt, h1.w, ta, tb, first_slice, lane, g == g0, &abort_flag);
            else ok = stream_steps<A>(p, &bar, parity, cnt, h1.y, h1.z, h1.w, ta, tb, first_slice, lane, g == g0, &abort_flag,
                                      comb, comb_first, warp < comb_n);
            parity ^= 1u;
            if (p.trace && lane == 0) p.trace[(size_t)blockIdx.x * (kWarps + 2) + warp] = clock64() - t_start;
            __syncthreads();                                       // everyone is done with this x tile
            if (comb)
                for (uint32_t s = warp; s < comb_n; s += kWarps) {
                    const acc_t v = comb[s * kLanes + lane];
                    comb[s * kLanes + lane] = acc_t(0);
                    if (!ok || abort_flag) continue;               // (a timed-out wait: no row update, the host hears of it)
                    const uint32_t row = __ldg(p.slice_rows + (size_t)(h1.y + comb_first + s) * kLanes + lane);
                    if (__all_sync(0xFFFFFFFFu, row == __shfl_sync(0xFFFFFFFFu, row, 0))) {
                        const acc_t t = A::warp_sum(v);
                        if (lane == 0) A::emit(p.acc, row, t);
                    } else {
                        A::emit(p.acc, row, v);
                    }
                }
        }
    }
    if (tl && blockIdx.x == 0 && tid == 0) tl[6] = globaltimer();

    // Drain the predecessor's accumulators (clamp / copy into y, re-zero): the pe dump + result drain of
    // the reference (pe.h:95-116, spmv_result_drain.cpp) and the PE's reset loop (pe.h:131-135).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tl && blockIdx.x == 0 && tid == 0) tl[2] = globaltimer();
    if (p.drain_acc) {
        if (p.wait_y_flag) {                                     // y is still being copied to the host
            if (lane == 0 && !wait_flag_geq(p.wait_y_flag, p.wait_y_val, false)) raise_error(p.error_flag);
            __syncwarp();
        }
        const GatherTargets *gt = p.gather;
        drain_rows<A>(p.drain_acc, p.y, p.drain_begin, p.drain_end, p.y_host, p.y_host_rows, gt, HSB_L2_HINTS && p.sync_start,
                      blockIdx.x * kThreads + tid, gridDim.x * kThreads);
        if (blockIdx.x == 0 && tid == 0) (void)A::drain(p.drain_acc, p.trash_row);
        if (gt) gather_publish(gt, p.gather_seq);
    }
    // The predecessor is complete and its writes are visible: announce it -- to later launches (guard), to
    // the host (upload throttling reads the mapped copy) and to the copy streams (cuStreamWaitValue32).
    // griddepcontrol.wait has already ordered the predecessor's writes; the device word is a release so that
    // the guard's acquire in a later launch pairs with it formally, the host word (read for write-after-read
    // throttling of uploads only) stays relaxed.
    if (blockIdx.x == 0 && tid == 0) {
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.done_dev), "r"(p.seq - 1u) : "memory");
        if (p.done_seq) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p.done_seq), "r"(p.seq - 1u) : "memory");
        if (tl) tl[3] = globaltimer();
    }
    if (p.trace && tid == 0) p.trace[(size_t)blockIdx.x * (kWarps + 2) + kWarps] = clock64() - t_start;
    if (tl && tid == 0) atomicMax(tl + 5, globaltimer());
}

template <class A>
__global__ void drain_kernel(void *acc, uint32_t *y, uint32_t row_begin, uint32_t row_end, uint32_t trash_row,
                             const GatherTargets *gt, uint32_t gather_seq, uint32_t *y_host, uint32_t y_host_rows) {
    drain_rows<A>(acc, y, row_begin, row_end, y_host, y_host_rows, gt, false, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    if (blockIdx.x == 0 && threadIdx.x == 0) (void)A::drain(acc, trash_row);
    if (gt) gather_publish(gt, gather_seq);
}

// root side of the gather: later work on the stream sees the blocks of all ranks
__global__ void wait_flags_kernel(const uint32_t *flags, uint32_t count, uint32_t val, uint32_t *error_flag) {
    for (uint32_t i = 0; i < count; i++)
        if (!wait_flag_geq(flags + i, val, true)) raise_error(error_flag);
}

// The step between two SpMVs of an iterative caller (PageRank-style x <- alpha (*) A x (+) beta): final y
// from the row accumulators (or from y itself when they are already drained), and the next vector written
// straight into an x buffer at this GPU's row offset -- no host round trip, no separate drain.
template <class A>
__global__ void axpb_kernel(void *acc, uint32_t *y, uint32_t *x_next, uint32_t rows, uint32_t x_limit, uint32_t alpha,
                            uint32_t beta, uint32_t col_offset, uint32_t trash_row) {
    // a programmatic dependent launch on both sides: the next SpMV may start (it fills its prefetch ring and then
    // waits for THIS grid before it stages x), and this grid was allowed to start before the SpMV whose sums it
    // finalises had retired
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
        uint32_t v;
        if (acc) { v = A::drain(acc, r); y[r] = v; } else { v = y[r]; }
        if (col_offset + r < x_limit) x_next[col_offset + r] = A::axpb(alpha, v, beta);
    }
    if (acc && blockIdx.x == 0 && threadIdx.x == 0) (void)A::drain(acc, trash_row);
}

template <class A>
__global__ void axpb_peers_kernel(void *acc, uint32_t *y, const PeerTargets t, uint32_t rows, uint32_t x_limit,
                                  uint32_t alpha, uint32_t beta, uint32_t col_offset, uint32_t trash_row, uint32_t seq,
                                  uint32_t *ticket) {
    // programmatic dependent launch on both sides, as axpb_kernel: the next SpMV may start at once -- what it needs
    // from this grid (and from the other ranks' grids) it learns from the arrival flags it polls before staging x
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
        uint32_t v;
        if (acc) { v = A::drain(acc, r); y[r] = v; } else { v = y[r]; }
        if (col_offset + r < x_limit) {
            const uint32_t w = A::axpb(alpha, v, beta);
            for (int g = 0; g < t.world; g++) t.x_next[g][col_offset + r] = w;      // NVLink peer stores, coalesced
        }
    }
    if (acc && blockIdx.x == 0 && threadIdx.x == 0) (void)A::drain(acc, trash_row);
    // grid-wide completion: the last CTA publishes the slice's arrival on every rank
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            *ticket = 0u;
            __threadfence_system();
            for (int g = 0; g < t.world; g++)
                publish_flag(t.flag[g], seq);
        }
    }
}

int g_sm_count = 148;                 // configure_kernels() replaces it with the device's count

}  // namespace

// hsb_iterate as one cooperative launch: kernel + launcher, built from the device functions above
#include "spmv_iterate.cuh"

cudaError_t launch_wait_flags(const uint32_t *flags, uint32_t count, uint32_t val, uint32_t *error_flag, cudaStream_t stream) {
    wait_flags_kernel<<<1, 1, 0, stream>>>(flags, count, val, error_flag);
    return cudaGetLastError();
}

cudaError_t launch_axpb_peers(int arith, void *acc, uint32_t *y, const PeerTargets &t, uint32_t rows, uint32_t x_limit,
                              uint32_t alpha, uint32_t beta, uint32_t col_offset, uint32_t trash_row, uint32_t seq,
                              uint32_t *ticket, cudaStream_t stream) {
    const int grid = (int)std::min<uint32_t>((std::max(rows, 1u) + 255) / 256, (uint32_t)g_sm_count * 4u);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    static const bool pdl = std::getenv("HSB_NO_PDL") == nullptr;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (arith == kArithFixed)
        return cudaLaunchKernelEx(&cfg, axpb_peers_kernel<FixedArith>, acc, y, t, rows, x_limit, alpha, beta, col_offset, trash_row, seq, ticket);
    return cudaLaunchKernelEx(&cfg, axpb_peers_kernel<FloatArith>, acc, y, t, rows, x_limit, alpha, beta, col_offset, trash_row, seq, ticket);
}

cudaError_t launch_axpb(int arith, void *acc, uint32_t *y, uint32_t *x_next, uint32_t rows, uint32_t x_limit,
                        uint32_t alpha, uint32_t beta, uint32_t col_offset, uint32_t trash_row, cudaStream_t stream) {
    const int grid = (int)std::min<uint32_t>((std::max(rows, 1u) + 255) / 256, (uint32_t)g_sm_count * 8u);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    static const bool pdl = std::getenv("HSB_NO_PDL") == nullptr;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (arith == kArithFixed) return cudaLaunchKernelEx(&cfg, axpb_kernel<FixedArith>, acc, y, x_next, rows, x_limit, alpha, beta, col_offset, trash_row);
    return cudaLaunchKernelEx(&cfg, axpb_kernel<FloatArith>, acc, y, x_next, rows, x_limit, alpha, beta, col_offset, trash_row);
}

cudaError_t configure_kernels(int sm_count) {
    if (sm_count > 0) g_sm_count = sm_count;
    const void *kernels[] = {(const void *)spmv_tiles_kernel<FixedArith, false>, (const void *)spmv_tiles_kernel<FloatArith, false>,
                             (const void *)spmv_tiles_kernel<FixedArith, true>, (const void *)spmv_tiles_kernel<FloatArith, true>};
    for (const void *k : kernels) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) return e;
    }
    return configure_iterate_kernels();
}

cudaError_t launch_spmv(int arith, const SpmvParams &p, int grid, uint32_t smem_bytes, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    static const bool pdl = std::getenv("HSB_NO_PDL") == nullptr;       // tracing aid: serialise launches
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: see griddepcontrol.wait in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (p.narrow) {
        if (arith == kArithFixed) return cudaLaunchKernelEx(&cfg, spmv_tiles_kernel<FixedArith, true>, p);
        return cudaLaunchKernelEx(&cfg, spmv_tiles_kernel<FloatArith, true>, p);
    }
    if (arith == kArithFixed) return cudaLaunchKernelEx(&cfg, spmv_tiles_kernel<FixedArith, false>, p);
    return cudaLaunchKernelEx(&cfg, spmv_tiles_kernel<FloatArith, false>, p);
}

cudaError_t launch_drain(int arith, void *acc, uint32_t *y, uint32_t row_begin, uint32_t row_end,
                         uint32_t trash_row, const GatherTargets *gather, uint32_t gather_seq, uint32_t *y_host,
                         uint32_t y_host_rows, cudaStream_t stream) {
    const uint32_t n = row_end > row_begin ? row_end - row_begin : 1;
    const int grid = (int)std::min<uint32_t>((n + 255) / 256, (uint32_t)g_sm_count * 8u);
    if (arith == kArithFixed)
        drain_kernel<FixedArith><<<grid, 256, 0, stream>>>(acc, y, row_begin, row_end, trash_row, gather, gather_seq, y_host, y_host_rows);
    else
        drain_kernel<FloatArith><<<grid, 256, 0, stream>>>(acc, y, row_begin, row_end, trash_row, gather, gather_seq, y_host, y_host_rows);
    return cudaGetLastError();
}

}  // namespace hsb
