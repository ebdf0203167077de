// Thin C++ layer over the C ABI (include/hisparse_b200.h) for host drivers written in the style of
// the reference's sw/host.cpp: HSB_CHECK reproduces OCL_CHECK's print-file:line-and-exit behaviour
// (xrt/includes/xcl2/xcl2.hpp:40-46), `aligned_vector` stands in for the page-aligned host vectors
// (xcl2.hpp:61-84), and `hsb_runtime` for the cl_runtime struct (sw/host.cpp:120-128).
#ifndef HISPARSE_B200_HOST_RUNTIME_H_
#define HISPARSE_B200_HOST_RUNTIME_H_

#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>

#include "../../include/hisparse_b200.h"

#define HSB_CHECK(call)                                                                         \
    do {                                                                                        \
        int rc__ = (call);                                                                      \
        if (rc__ != HSB_OK) {                                                                   \
            printf("%s:%d Error calling " #call ", error code is: %d (%s)\n", __FILE__, __LINE__, \
                   rc__, hsb_last_error());                                                     \
            exit(EXIT_FAILURE);                                                                 \
        }                                                                                       \
    } while (0)

// page-locked allocator: device transfers from these vectors are true DMA
template <typename T> struct pinned_allocator {
    using value_type = T;
    pinned_allocator() {}
    template <class U> pinned_allocator(const pinned_allocator<U> &) {}
    T *allocate(std::size_t n) {
        void *p = hsb_host_alloc(n * sizeof(T));
        if (!p) throw std::bad_alloc();
        return reinterpret_cast<T *>(p);
    }
    void deallocate(T *p, std::size_t) { hsb_host_free(p); }
    template <class U> bool operator==(const pinned_allocator<U> &) const { return true; }
    template <class U> bool operator!=(const pinned_allocator<U> &) const { return false; }
};
template <typename T> using aligned_vector = std::vector<T, pinned_allocator<T> >;

struct hsb_runtime {
    hsb_ctx *ctx = nullptr;
    explicit hsb_runtime(int device, int impl) {
        ctx = hsb_create(device, impl);
        if (!ctx) {
            printf("ERROR : Failed to open B200 device %d: %s, exit!\n", device, hsb_last_error());
            exit(EXIT_FAILURE);
        }
    }
    ~hsb_runtime() { hsb_destroy(ctx); }
    hsb_runtime(const hsb_runtime &) = delete;
    hsb_runtime &operator=(const hsb_runtime &) = delete;
};

#endif
