// TEST INFRASTRUCTURE ONLY (oracle). Clean-room stand-in for the one Xilinx fixed-point
// type the reference uses: ap_ufixed<32, 8, AP_RND, AP_SAT> (spmv/libfpga/common.h:35-38).
// No Xilinx header was available; the arithmetic below follows the DOCUMENTED semantics:
//   * value = raw * 2^-(W-I), raw is an unsigned W-bit integer;
//   * `a * b` and `a + b` are exact (full-precision temporaries);
//   * assignment to the declared type quantises with AP_RND  = round to plus infinity
//     (add half an LSB, then truncate) and overflows with AP_SAT = clamp to [0, 2^W - 1];
//   * conversion from float/double quantises the same way; negatives clamp to 0.
// PARITY NOTE: because the real header is absent, rounding/saturation corner cases are
// pinned to this statement of the documented behaviour, not to Xilinx's implementation.
// The reference's own tests only use values {0, 1} and cannot distinguish the two.
#ifndef HISPARSE_ORACLE_SHIM_AP_FIXED_H_
#define HISPARSE_ORACLE_SHIM_AP_FIXED_H_

#include <cmath>
#include <cstdint>
#include <iostream>
#include "ap_int.h"

enum ap_q_mode { AP_RND, AP_RND_ZERO, AP_RND_MIN_INF, AP_RND_INF, AP_RND_CONV, AP_TRN, AP_TRN_ZERO };
enum ap_o_mode { AP_SAT, AP_SAT_ZERO, AP_SAT_SYM, AP_WRAP, AP_WRAP_SM };

// exact unsigned intermediate: value = mant * 2^-frac
struct ap_ufixed_exact {
    unsigned __int128 mant;
    int frac;
};

// read/write view of bits [hi:lo] of a 32-bit raw fixed-point word
struct ap_ufixed_range_ref {
    uint32_t *raw;
    int hi, lo;
    uint32_t mask() const { int w = hi - lo + 1; return w >= 32 ? 0xFFFFFFFFu : ((1u << w) - 1u); }
    unsigned long long get() const { return (*raw >> lo) & mask(); }
    void set(unsigned long long v) { *raw = (*raw & ~(mask() << lo)) | (((uint32_t)v & mask()) << lo); }
    operator unsigned long long() const { return get(); }
    ap_ufixed_range_ref &operator=(unsigned long long v) { set(v); return *this; }
    ap_ufixed_range_ref &operator=(const ap_ufixed_range_ref &o) { set(o.get()); return *this; }
    template <class T> ap_ufixed_range_ref &operator=(const T &o) {
        set((unsigned long long)o); return *this;
    }
};
inline std::ostream &operator<<(std::ostream &os, const ap_ufixed_range_ref &r) { return os << r.get(); }

template <int W, int I, ap_q_mode Q = AP_TRN, ap_o_mode O = AP_WRAP, int N = 0>
struct ap_ufixed {
    static_assert(W == 32 && Q == AP_RND && O == AP_SAT,
                  "oracle shim only models ap_ufixed<32, I, AP_RND, AP_SAT>");
    static const int F = W - I;
    uint32_t raw;

    static uint32_t from_exact(unsigned __int128 mant, int frac) {
        unsigned __int128 r;
        if (frac > F) {
            int sh = frac - F;
            r = (mant + ((unsigned __int128)1 << (sh - 1))) >> sh;   // AP_RND
        } else {
            r = mant << (F - frac);
        }
        return r > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)r;           // AP_SAT
    }
    static uint32_t from_double(double v) {
        if (!(v > 0.0)) return 0;                                       // negatives / NaN -> 0
        double s = std::floor(std::ldexp(v, F) + 0.5);                  // AP_RND
        if (s >= 4294967296.0) return 0xFFFFFFFFu;                      // AP_SAT
        return (uint32_t)s;
    }

    ap_ufixed() : raw(0) {}
    ap_ufixed(int v) : raw(v <= 0 ? 0u : from_exact((unsigned __int128)v, 0)) {}
    ap_ufixed(unsigned v) : raw(from_exact(v, 0)) {}
    ap_ufixed(long v) : raw(v <= 0 ? 0u : from_exact((unsigned __int128)v, 0)) {}
    ap_ufixed(unsigned long v) : raw(from_exact(v, 0)) {}
    ap_ufixed(unsigned long long v) : raw(from_exact(v, 0)) {}
    ap_ufixed(float v) : raw(from_double((double)v)) {}
    ap_ufixed(double v) : raw(from_double(v)) {}
    ap_ufixed(const ap_ufixed_exact &e) : raw(from_exact(e.mant, e.frac)) {}

    ap_ufixed_exact exact() const { return ap_ufixed_exact{raw, F}; }

    explicit operator float() const { return (float)std::ldexp((double)raw, -F); }
    explicit operator double() const { return std::ldexp((double)raw, -F); }
    float to_float() const { return (float)*this; }
    double to_double() const { return (double)*this; }

    // bit-level access: v(31,0) = raw bits, v(31,24) = integer part
    ap_ufixed_range_ref operator()(int hi, int lo) { return ap_ufixed_range_ref{&raw, hi, lo}; }
    unsigned long long operator()(int hi, int lo) const {
        uint32_t r = raw;
        return ap_ufixed_range_ref{&r, hi, lo}.get();
    }
    bool operator==(const ap_ufixed &o) const { return raw == o.raw; }
    bool operator!=(const ap_ufixed &o) const { return raw != o.raw; }
    bool operator<(const ap_ufixed &o) const { return raw < o.raw; }
    bool operator>(const ap_ufixed &o) const { return raw > o.raw; }

};

template <int W, int I, ap_q_mode Q, ap_o_mode O, int N>
inline ap_ufixed_exact operator*(const ap_ufixed<W, I, Q, O, N> &a, const ap_ufixed<W, I, Q, O, N> &b) {
    return ap_ufixed_exact{(unsigned __int128)a.raw * b.raw, 2 * (W - I)};
}
template <int W, int I, ap_q_mode Q, ap_o_mode O, int N>
inline ap_ufixed_exact operator+(const ap_ufixed<W, I, Q, O, N> &a, const ap_ufixed<W, I, Q, O, N> &b) {
    return ap_ufixed_exact{(unsigned __int128)a.raw + b.raw, W - I};
}
template <int W, int I, ap_q_mode Q, ap_o_mode O, int N>
inline std::ostream &operator<<(std::ostream &os, const ap_ufixed<W, I, Q, O, N> &v) {
    return os << (double)v;
}

#endif
