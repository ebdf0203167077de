// TEST INFRASTRUCTURE ONLY (oracle). Clean-room stand-in for the subset of Xilinx
// Vitis-HLS `ap_int.h` that the reference's SpMV sources use, so that the UNMODIFIED
// reference C-simulation (/root/reference/spmv_csim/csim.cpp) compiles with plain g++.
// Written from the documented behaviour of ap_uint (arbitrary-width unsigned integer
// with bit / range selection); no Xilinx source was available or consulted.
//
// Surface needed by the reference (file:line of a representative use):
//   ap_uint<2>  INST_T, compared/switch'ed against ints      spmv/libfpga/common.h:60,99-106
//   ap_uint<8>  lane masks: operator[] r/w, and_reduce()     spmv/libfpga/spmv_cluster.h:142-186
//   rrotate(n)  in-place rotate right                        spmv/libfpga/shuffle.h:54
//   ap_uint<288> AXIS word with (hi,lo) range r/w            spmv/libfpga/common.h:140-146
//   ap_uint<32> val2bit/bit2val                              spmv-fp/libfpga/common.h:39-51
#ifndef HISPARSE_ORACLE_SHIM_AP_INT_H_
#define HISPARSE_ORACLE_SHIM_AP_INT_H_

#include <cstdint>
#include <cstring>
#include <iostream>

template <int W> struct ap_uint;

// read/write view of bits [hi:lo] (width <= 64) of some word array
struct ap_range_ref {
    uint64_t *words;
    int hi, lo;
    int width() const { return hi - lo + 1; }
    unsigned long long get() const {
        unsigned long long v = 0;
        for (int b = width() - 1; b >= 0; b--) {
            int pos = lo + b;
            v = (v << 1) | ((words[pos >> 6] >> (pos & 63)) & 1ull);
        }
        return v;
    }
    void set(unsigned long long v) {
        for (int b = 0; b < width(); b++) {
            int pos = lo + b;
            uint64_t m = 1ull << (pos & 63);
            if ((v >> b) & 1ull) words[pos >> 6] |= m; else words[pos >> 6] &= ~m;
        }
    }
    operator unsigned long long() const { return get(); }
    ap_range_ref &operator=(unsigned long long v) { set(v); return *this; }
    ap_range_ref &operator=(const ap_range_ref &o) { set(o.get()); return *this; }
    template <class T> ap_range_ref &operator=(const T &o) {
        set((unsigned long long)o); return *this;
    }
};
inline std::ostream &operator<<(std::ostream &os, const ap_range_ref &r) { return os << r.get(); }

struct ap_bit_ref {
    uint64_t *words;
    int pos;
    operator bool() const { return (words[pos >> 6] >> (pos & 63)) & 1ull; }
    ap_bit_ref &operator=(bool v) {
        uint64_t m = 1ull << (pos & 63);
        if (v) words[pos >> 6] |= m; else words[pos >> 6] &= ~m;
        return *this;
    }
    ap_bit_ref &operator=(const ap_bit_ref &o) { return *this = bool(o); }
    bool operator!() const { return !bool(*this); }
};

template <int W> struct ap_uint {
    static const int NW = (W + 63) / 64;
    uint64_t w[NW];

    void trim() {
        if (W % 64) w[NW - 1] &= (~0ull >> (64 - W % 64));
    }
    void assign(unsigned long long v) {
        w[0] = v;
        for (int i = 1; i < NW; i++) w[i] = 0;
        trim();
    }
    ap_uint() { for (int i = 0; i < NW; i++) w[i] = 0; }
    ap_uint(int v) { assign((unsigned long long)(long long)v); }
    ap_uint(unsigned v) { assign(v); }
    ap_uint(long v) { assign((unsigned long long)v); }
    ap_uint(unsigned long v) { assign(v); }
    ap_uint(unsigned long long v) { assign(v); }
    ap_uint(bool v) { assign(v ? 1 : 0); }
    ap_uint(const ap_range_ref &r) { assign(r.get()); }
    ap_uint(const ap_bit_ref &r) { assign(bool(r) ? 1 : 0); }

    // single implicit integral conversion (keeps `switch`, `==`, arithmetic unambiguous)
    operator unsigned long long() const { return w[0]; }

    ap_bit_ref operator[](int i) { return ap_bit_ref{w, i}; }
    bool operator[](int i) const { return (w[i >> 6] >> (i & 63)) & 1ull; }
    ap_range_ref operator()(int hi, int lo) { return ap_range_ref{w, hi, lo}; }
    unsigned long long operator()(int hi, int lo) const {
        return ap_range_ref{const_cast<uint64_t *>(w), hi, lo}.get();
    }
    ap_range_ref range(int hi, int lo) { return (*this)(hi, lo); }

    bool and_reduce() const {
        for (int i = 0; i < W; i++) if (!((w[i >> 6] >> (i & 63)) & 1ull)) return false;
        return true;
    }
    bool or_reduce() const {
        for (int i = 0; i < NW; i++) if (w[i]) return true;
        return false;
    }
    // rotate right by n, in place: new bit i = old bit (i + n) mod W
    ap_uint &rrotate(int n) {
        ap_uint old = *this;
        for (int i = 0; i < W; i++) {
            int src = (i + n) % W;
            bool b = (old.w[src >> 6] >> (src & 63)) & 1ull;
            (*this)[i] = b;
        }
        return *this;
    }
    ap_uint &lrotate(int n) { return rrotate(W - (n % W)); }
};

template <int W> inline std::ostream &operator<<(std::ostream &os, const ap_uint<W> &v) {
    return os << (unsigned long long)v;
}

#endif
