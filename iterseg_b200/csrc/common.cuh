// Shared helpers for the iterseg_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "iterseg_b200.h"

namespace isg {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
// SMs set aside for the post stage while a U-Net of another frame runs beside it (0 = none):
// the persistent conv kernels launch num_sms() - post_sms() CTAs, the shared-memory-hungry
// flood kernels at most post_sms() SMs' worth, so neither waits for the other's CTAs to retire
int post_sms();

#define ISG_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            isg::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
                           cudaGetErrorString(_e));                                      \
            return ISG_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

// after every kernel launch: count it and surface launch-configuration errors
#define ISG_LAUNCHED()                                                                   \
    do {                                                                                 \
        isg::count_launch();                                                             \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            isg::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,             \
                           cudaGetErrorString(_e));                                      \
            return ISG_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define ISG_REQUIRE(cond, code, ...)                                                     \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            isg::set_error(__VA_ARGS__);                                                 \
            return (code);                                                               \
        }                                                                                \
    } while (0)

// bump allocator over a caller-provided workspace (256-byte aligned pieces)
struct Carver {
    char *base;
    size_t cap;
    size_t off;
    bool ok;
    __host__ Carver(void *p, size_t bytes) : base((char *)p), cap(bytes), off(0), ok(true) {
        size_t mis = (size_t)((uintptr_t)base & 255u);
        if (mis) off = 256 - mis;
    }
    template <typename T>
    __host__ T *take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        size_t at = off;
        off += bytes;
        if (base == nullptr) return nullptr;   // sizing pass
        if (off > cap) {
            ok = false;
            return nullptr;
        }
        return (T *)(base + at);
    }
};

static inline int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// float <-> order-preserving uint32 (ascending)
__host__ __device__ __forceinline__ uint32_t f32_ord(float v) {
    uint32_t b;
#ifdef __CUDA_ARCH__
    b = __float_as_uint(v);
#else
    union { float f; uint32_t u; } c; c.f = v; b = c.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord_f32(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}

}  // namespace isg
