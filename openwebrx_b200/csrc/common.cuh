// common.cuh — shared host/device helpers for libowrx_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "owrx_b200.h"

namespace owrx {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define OWRX_CUDA(expr)                                                                         \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return ::owrx::fail(OWRX_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                __FILE__, __LINE__);                                            \
    } while (0)

#define OWRX_LAUNCH_CHECK()                                                                     \
    do {                                                                                        \
        ::owrx::g_launches.fetch_add(1, std::memory_order_relaxed);                             \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess)                                                                  \
            return ::owrx::fail(OWRX_E_CUDA, "kernel launch failed: %s (%s:%d)",                \
                                cudaGetErrorString(_e), __FILE__, __LINE__);                    \
    } while (0)

// Select the device and verify it is a Blackwell-class part; there is no CPU fallback.
int select_device(int device, int* sm_count);
// true if `p` lies in page-locked host memory known to CUDA (cudaHostAlloc / cudaHostRegister)
bool host_is_pinned(const void* p);

// ---- small complex helpers -------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

}  // namespace owrx
