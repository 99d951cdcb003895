// host_util.cuh -- error plumbing and a device buffer RAII shared by the host code.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace admmb {

struct CudaFail {
    cudaError_t e;
    const char *what;
};
#define CK(call)                                                      \
    do {                                                              \
        cudaError_t e_ = (call);                                      \
        if (e_ != cudaSuccess) throw admmb::CudaFail{e_, #call};      \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    void alloc(size_t count)
    {
        if (count <= n && p) return;
        release();
        if (count == 0) return;
        CK(cudaMalloc((void **)&p, count * sizeof(T)));
        n = count;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

inline size_t round_up(size_t a, size_t m) { return (a + m - 1) / m * m; }

}  // namespace admmb
