// TEST INFRASTRUCTURE ONLY (oracle). Clean-room stand-in for Vitis-HLS `hls::stream<T>`
// as used in C simulation: an unbounded FIFO (spmv_csim/csim.cpp:47-48 relies on the
// streams buffering an entire stage's output because stages run sequentially).
#ifndef HISPARSE_ORACLE_SHIM_HLS_STREAM_H_
#define HISPARSE_ORACLE_SHIM_HLS_STREAM_H_

#include <cstdio>
#include <cstdlib>
#include <deque>

namespace hls {
template <typename T> class stream {
    std::deque<T> q_;
public:
    stream() {}
    explicit stream(const char *) {}
    stream(const stream &) = delete;
    stream &operator=(const stream &) = delete;
    void write(const T &v) { q_.push_back(v); }
    T read() {
        if (q_.empty()) {
            std::fprintf(stderr, "hls::stream shim: blocking read on an empty stream (deadlock in csim)\n");
            std::abort();
        }
        T v = q_.front();
        q_.pop_front();
        return v;
    }
    void read(T &v) { v = read(); }
    bool read_nb(T &v) {
        if (q_.empty()) return false;
        v = q_.front();
        q_.pop_front();
        return true;
    }
    bool write_nb(const T &v) { q_.push_back(v); return true; }
    bool empty() const { return q_.empty(); }
    bool full() const { return false; }
    size_t size() const { return q_.size(); }
    stream &operator<<(const T &v) { write(v); return *this; }
    stream &operator>>(T &v) { v = read(); return *this; }
};
}  // namespace hls
#endif
