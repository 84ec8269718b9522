// host_util.h -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/b200ida.h"

namespace b200ida {

char* last_error_buf();  // thread-local, 512 bytes (defined in ida_engine.cu)

inline int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define B200_CUDA_TRY(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return ::b200ida::set_error(B200IDA_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,   \
                                        cudaGetErrorString(_e));                                         \
    } while (0)

// launches of the engine's own kernels since the library was loaded (b200ida_kernel_launches): bench.py's `gpu_launches`
// is the difference of this counter over the timed region, not a hand count
long long& launch_counter();  // defined in ida_engine.cu
inline void count_launch(int n = 1) { launch_counter() += n; }

// RAII device scratch used by the host-memory (end-to-end) entry points
struct DevBuf {
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
    }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

}  // namespace b200ida
