// Device-side synthetic CSR generator for the large configurations of BASELINE.json (SURVEY.md
// section 8d, C5: 100 M x 100 M, 2 B non-zeros, power-law row degrees, 80 % of the columns inside a
// +-2^20 band around the diagonal, 20 % uniform -- "generated on device per shard").
//
// A shard is a contiguous block of rows [first_row, first_row + rows) of the global matrix. Everything
// is a pure function of (seed, global row, draw index), so shards generated on different GPUs / at
// different times are pieces of one and the same matrix:
//
//   degree(g)  = min(max_degree, floor(x_m * u^(-1/(alpha-1))))     u = U(0,1] from hash(seed, g)
//                (discrete truncated Pareto; x_m is calibrated on the host so that E[degree] = mean)
//   column k   = with probability band_fraction: (g + U{-W..W-1}) mod cols, else U{0..cols-1}
//   value      = U[0,1) * scale from hash(seed, g, column): fp32 bits, or raw Q8.24 words
//
// Draws are sorted per row and duplicates dropped (device radix sort + unique), so column ids are
// sorted and unique inside every row. No reference counterpart (the reference loads .npz files,
// sw/data_loader.h:51-70); measurement plumbing only.
#include "../../include/hisparse_b200.h"

#include <cub/cub.cuh>

#include <cmath>
#include <cstdio>
#include <string>

namespace hsb {
namespace {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {          // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct SynthParams {
    uint32_t rows, cols;
    uint64_t first_row, seed;
    double xm, inv_a;              // Pareto scale and 1 / (alpha - 1)
    uint32_t max_degree, band_half_width;
    uint32_t band_threshold;       // draw < threshold (of 2^32) -> band column
    int value_kind;                // 0: fp32 bits, 1: raw Q8.24 words
    float value_scale;
};

__device__ __forceinline__ uint32_t degree_of(const SynthParams &p, uint64_t g) {
    const uint64_t h = mix64(p.seed ^ mix64(g * 2 + 1));
    const double u = ((double)(h >> 11) + 1.0) * (1.0 / 9007199254740992.0);      // (0, 1]
    const double d = floor(p.xm * pow(u, -p.inv_a));
    return d >= (double)p.max_degree ? p.max_degree : (uint32_t)d;
}

__global__ void k_degrees(SynthParams p, uint32_t *__restrict__ deg) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < p.rows) deg[r] = degree_of(p, p.first_row + r);
    if (r == p.rows) deg[r] = 0;
}

// one thread per draw: key = (local row << 32) | column
__global__ void k_draws(SynthParams p, uint64_t n_draws, const uint32_t *__restrict__ off,
                        unsigned long long *__restrict__ keys) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_draws) return;
    uint32_t lo = 0, hi = p.rows;                     // last row with off[row] <= e
    while (hi - lo > 1) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (off[mid] <= e) lo = mid; else hi = mid;
    }
    const uint64_t g = p.first_row + lo, k = e - off[lo];
    const uint64_t h = mix64(mix64(p.seed + 0x51ED270B7F4A7C15ull * (g + 1)) ^ (k * 0xD1342543DE82EF95ull + 1));
    uint64_t col;
    if ((uint32_t)h < p.band_threshold) {
        const uint64_t span = 2ull * p.band_half_width;
        const uint64_t o = (h >> 32) % span;
        col = (g % p.cols + p.cols + o - p.band_half_width) % p.cols;
    } else {
        col = (h >> 32) % p.cols;
    }
    keys[e] = ((unsigned long long)lo << 32) | col;
}

__global__ void k_split(SynthParams p, uint64_t nnz, const unsigned long long *__restrict__ keys,
                        uint32_t *__restrict__ indices, uint32_t *__restrict__ vals) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const uint32_t col = (uint32_t)keys[e];
    const uint64_t g = p.first_row + (keys[e] >> 32);
    const uint64_t h = mix64(mix64(p.seed ^ 0xA5A5A5A5DEADBEEFull ^ (g * 0x9E3779B97F4A7C15ull)) + col);
    const uint32_t u24 = (uint32_t)(h >> 40);                                   // 24 random bits
    indices[e] = col;
    if (p.value_kind == 0)
        vals[e] = __float_as_uint(__fmul_rn((float)u24 * (1.0f / 16777216.0f), p.value_scale));
    else
        vals[e] = (uint32_t)fmin(floor((double)u24 * (double)p.value_scale + 0.5), 4294967295.0);   // u24 / 2^24 in Q8.24 == u24
}

// indptr[r] = first sorted key whose row >= r
__global__ void k_indptr(uint32_t rows, uint64_t nnz, const unsigned long long *__restrict__ keys,
                         uint32_t *__restrict__ indptr) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > rows) return;
    uint64_t lo = 0, hi = nnz;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if ((keys[mid] >> 32) >= r) hi = mid; else lo = mid + 1;
    }
    indptr[r] = (uint32_t)lo;
}

// E[min(M, floor(X))] for X ~ Pareto(xm, a): sum_{k=1..M} P(X >= k) = sum min(1, (xm/k)^a)
double expected_degree(double xm, double a, uint32_t M) {
    double s = 0;
    for (uint32_t k = 1; k <= M; k++) {
        const double q = xm / (double)k;
        s += q >= 1.0 ? 1.0 : std::pow(q, a);
    }
    return s;
}

inline int bits_for(uint64_t v) { int b = 1; while (b < 63 && (1ull << b) <= v) b++; return b; }

thread_local std::string g_synth_err;

}  // namespace
}  // namespace hsb

extern "C" {

const char *hsb_synth_last_error(void) { return hsb::g_synth_err.c_str(); }

int hsb_synth_powerlaw_csr_device(int device, uint32_t rows, uint32_t cols, uint64_t first_global_row,
                                  double mean_degree, double alpha, uint32_t max_degree,
                                  uint32_t band_half_width, double band_fraction, uint64_t seed,
                                  int value_kind, float value_scale, hsb_device_csr *out) {
    using namespace hsb;
    auto fail = [&](int code, const std::string &m) { g_synth_err = m; return code; };
    if (!out || rows == 0 || cols == 0 || alpha <= 1.0 || mean_degree <= 0 || max_degree == 0 ||
        band_fraction < 0 || band_fraction > 1 || (value_kind != 0 && value_kind != 1))
        return fail(HSB_EINVAL, "bad generator argument");
    if (band_half_width == 0 || 2ull * band_half_width > cols) band_half_width = cols / 2 ? cols / 2 : 1;
    *out = hsb_device_csr{};
#define SY_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            cudaFree(deg); cudaFree(off); cudaFree(keys); cudaFree(keys2); cudaFree(tmp); cudaFree(d_n); \
            cudaFree(ip); cudaFree(ix); cudaFree(vv);                                                  \
            cudaGetLastError();                                                                        \
            return fail(HSB_ECUDA, std::string("generator: ") + cudaGetErrorString(e__) + " (" #expr ")"); \
        }                                                                                              \
    } while (0)
    uint32_t *deg = nullptr, *off = nullptr, *ip = nullptr, *ix = nullptr, *vv = nullptr;
    unsigned long long *keys = nullptr, *keys2 = nullptr, *d_n = nullptr;
    void *tmp = nullptr;
    SY_TRY(cudaSetDevice(device));

    // calibrate the Pareto scale: expected_degree is increasing in xm
    const double a = alpha - 1.0;
    double lo = 1e-3, hi = std::max(4.0, mean_degree * 4.0);
    for (int it = 0; it < 48; it++) {
        const double mid = 0.5 * (lo + hi);
        if (expected_degree(mid, a, max_degree) < mean_degree) lo = mid; else hi = mid;
    }
    SynthParams p;
    p.rows = rows; p.cols = cols; p.first_row = first_global_row; p.seed = seed;
    p.xm = 0.5 * (lo + hi); p.inv_a = 1.0 / a;
    p.max_degree = max_degree; p.band_half_width = band_half_width;
    p.band_threshold = band_fraction >= 1.0 ? 0xFFFFFFFFu : (uint32_t)(band_fraction * 4294967296.0);
    p.value_kind = value_kind;
    p.value_scale = value_scale;

    const int TB = 256;
    auto blocks = [&](uint64_t n) { return (unsigned)((n + TB - 1) / TB); };
    SY_TRY(cudaMalloc(&deg, ((size_t)rows + 1) * 4));
    SY_TRY(cudaMalloc(&off, ((size_t)rows + 1) * 4));
    SY_TRY(cudaMalloc(&d_n, 8));
    SY_TRY(cudaMemset(d_n, 0, 8));
    k_degrees<<<blocks((uint64_t)rows + 1), TB>>>(p, deg);
    // the offsets are 32 bits: refuse shards whose expected size does not fit (the sampled total of a
    // shard that passes is within a few percent of the expectation)
    if ((double)rows * mean_degree > 2.0e9) { SY_TRY(cudaErrorInvalidValue); }
    size_t need = 0;
    SY_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, deg, off, (int)rows + 1));
    SY_TRY(cudaMalloc(&tmp, need));
    SY_TRY(cub::DeviceScan::ExclusiveSum(tmp, need, deg, off, (int)rows + 1));
    uint32_t n32 = 0;
    SY_TRY(cudaMemcpy(&n32, off + rows, 4, cudaMemcpyDeviceToHost));
    const uint64_t n_draws = n32;
    cudaFree(tmp); tmp = nullptr;
    if (n_draws >= (1ull << 31)) { SY_TRY(cudaErrorInvalidValue); }

    SY_TRY(cudaMalloc(&keys, std::max<uint64_t>(n_draws, 1) * 8));
    SY_TRY(cudaMalloc(&keys2, std::max<uint64_t>(n_draws, 1) * 8));
    uint64_t nnz = 0;
    if (n_draws) {
        k_draws<<<blocks(n_draws), TB>>>(p, n_draws, off, keys);
        const int key_bits = 32 + bits_for(rows);
        SY_TRY(cub::DeviceRadixSort::SortKeys(nullptr, need, keys, keys2, (int)n_draws, 0, key_bits));
        SY_TRY(cudaMalloc(&tmp, need));
        SY_TRY(cub::DeviceRadixSort::SortKeys(tmp, need, keys, keys2, (int)n_draws, 0, key_bits));
        cudaFree(tmp); tmp = nullptr;
        SY_TRY(cub::DeviceSelect::Unique(nullptr, need, keys2, keys, d_n, (int)n_draws));
        SY_TRY(cudaMalloc(&tmp, need));
        SY_TRY(cub::DeviceSelect::Unique(tmp, need, keys2, keys, d_n, (int)n_draws));
        unsigned long long h_n = 0;
        SY_TRY(cudaMemcpy(&h_n, d_n, 8, cudaMemcpyDeviceToHost));
        nnz = h_n;
        cudaFree(tmp); tmp = nullptr;
    }
    cudaFree(keys2); keys2 = nullptr;
    cudaFree(deg); deg = nullptr;
    cudaFree(off); off = nullptr;
    SY_TRY(cudaMalloc(&ip, ((size_t)rows + 1) * 4));
    SY_TRY(cudaMalloc(&ix, std::max<uint64_t>(nnz, 1) * 4));
    SY_TRY(cudaMalloc(&vv, std::max<uint64_t>(nnz, 1) * 4));
    if (nnz) k_split<<<blocks(nnz), TB>>>(p, nnz, keys, ix, vv);
    k_indptr<<<blocks((uint64_t)rows + 1), TB>>>(rows, nnz, keys, ip);
    SY_TRY(cudaGetLastError());
    SY_TRY(cudaDeviceSynchronize());
    cudaFree(keys); cudaFree(d_n);
    out->rows = rows; out->cols = cols; out->nnz = nnz;
    out->d_indptr = ip; out->d_indices = ix; out->d_vals = vv;
    out->device = device;
    return HSB_OK;
#undef SY_TRY
}

int hsb_device_csr_download(const hsb_device_csr *m, uint32_t *indptr, uint32_t *indices, uint32_t *vals) {
    if (!m || !m->d_indptr) { hsb::g_synth_err = "null device CSR"; return HSB_EINVAL; }
    cudaError_t e = cudaSetDevice(m->device);
    if (e == cudaSuccess && indptr) e = cudaMemcpy(indptr, m->d_indptr, ((size_t)m->rows + 1) * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && indices && m->nnz) e = cudaMemcpy(indices, m->d_indices, m->nnz * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && vals && m->nnz) e = cudaMemcpy(vals, m->d_vals, m->nnz * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { hsb::g_synth_err = cudaGetErrorString(e); cudaGetLastError(); return HSB_ECUDA; }
    return HSB_OK;
}

void hsb_device_csr_free(hsb_device_csr *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaFree(m->d_indptr); cudaFree(m->d_indices); cudaFree(m->d_vals);
    *m = hsb_device_csr{};
}

}  // extern "C"
