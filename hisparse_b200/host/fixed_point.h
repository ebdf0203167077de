// Q8.24 unsigned fixed point: the host-side stand-in for the reference's
//   typedef ap_ufixed<32, 8, AP_RND, AP_SAT> VAL_T;          (spmv/libfpga/common.h:35-38)
// Only what host code needs: conversion from/to float (round half up at 2^-24, clamp to
// [0, 2^32-1]), raw-bit access, and the two PE operations for host-side checks. The device
// arithmetic lives in csrc/spmv_kernels.cu.
#ifndef HISPARSE_B200_HOST_FIXED_POINT_H_
#define HISPARSE_B200_HOST_FIXED_POINT_H_

#include <cmath>
#include <cstdint>
#include <ostream>

namespace spmv {

struct ufixed_q8_24 {
    uint32_t raw;

    ufixed_q8_24() : raw(0) {}
    ufixed_q8_24(float v) : raw(quantize(v)) {}
    ufixed_q8_24(double v) : raw(quantize(v)) {}
    ufixed_q8_24(int v) : raw(v <= 0 ? 0u : (v >= 256 ? 0xFFFFFFFFu : (uint32_t)v << 24)) {}
    ufixed_q8_24(unsigned v) : raw(v >= 256u ? 0xFFFFFFFFu : v << 24) {}
    static ufixed_q8_24 from_raw(uint32_t bits) { ufixed_q8_24 r; r.raw = bits; return r; }

    static uint32_t quantize(double v) {
        if (!(v > 0.0)) return 0u;
        double s = std::floor(std::ldexp(v, 24) + 0.5);
        return s >= 4294967296.0 ? 0xFFFFFFFFu : (uint32_t)s;
    }
    explicit operator float() const { return (float)std::ldexp((double)raw, -24); }
    explicit operator double() const { return std::ldexp((double)raw, -24); }

    // `mat_val * vec_val` then `q + incr` as the PE does them (spmv/libfpga/pe.h:64,72)
    static ufixed_q8_24 mul(ufixed_q8_24 a, ufixed_q8_24 b) {
        uint64_t p = ((uint64_t)a.raw * b.raw + (1ull << 23)) >> 24;
        return from_raw(p > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)p);
    }
    static ufixed_q8_24 add(ufixed_q8_24 a, ufixed_q8_24 b) {
        uint64_t s = (uint64_t)a.raw + b.raw;
        return from_raw(s > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)s);
    }
    bool operator==(const ufixed_q8_24 &o) const { return raw == o.raw; }
};
inline std::ostream &operator<<(std::ostream &os, const ufixed_q8_24 &v) { return os << (double)v; }

static_assert(sizeof(ufixed_q8_24) == 4, "VAL_T must be one 32-bit word");

}  // namespace spmv
#endif
