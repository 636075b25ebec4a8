// Shared device/host helpers for libeventful_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/eventful_b200.h"

// ---------------------------------------------------------------- error plumbing
extern thread_local char g_et_error[512];

inline int et_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_et_error, sizeof(g_et_error), fmt, ap);
    va_end(ap);
    return code;
}

#define ET_CHECK_ARG(cond, ...)                                  \
    do {                                                         \
        if (!(cond)) return et_fail(ET_ERR_ARG, __VA_ARGS__);    \
    } while (0)

#define ET_CHECK_LAUNCH(name)                                                                   \
    do {                                                                                        \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess)                                                                 \
            return et_fail(ET_ERR_CUDA, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

static inline bool et_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------- dtype traits
template <typename T>
struct ElemTraits;
template <>
struct ElemTraits<float> {
    static constexpr int VEC = 4;  // elements per 16-byte vector
    static __device__ __forceinline__ float to_float(float v) { return v; }
    static __device__ __forceinline__ float from_float(float v) { return v; }
};
template <>
struct ElemTraits<__nv_bfloat16> {
    static constexpr int VEC = 8;
    static __device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct ElemTraits<__half> {
    static constexpr int VEC = 8;
    static __device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_float(float v) { return __float2half_rn(v); }
};

// value of v after a round trip through T (round-to-nearest-even), i.e. what a
// torch elementwise op on dtype T would have stored
template <typename T>
__device__ __forceinline__ float round_to(float v) {
    return ElemTraits<T>::to_float(ElemTraits<T>::from_float(v));
}

template <typename T>
__device__ __forceinline__ void unpack16(const uint4& u, float* f) {
    if constexpr (sizeof(T) == 4) {
        f[0] = __uint_as_float(u.x);
        f[1] = __uint_as_float(u.y);
        f[2] = __uint_as_float(u.z);
        f[3] = __uint_as_float(u.w);
    } else {
        const T* h = reinterpret_cast<const T*>(&u);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = ElemTraits<T>::to_float(h[i]);
    }
}

template <typename T>
__device__ __forceinline__ uint4 pack16(const float* f) {
    uint4 u;
    if constexpr (sizeof(T) == 4) {
        u.x = __float_as_uint(f[0]);
        u.y = __float_as_uint(f[1]);
        u.z = __float_as_uint(f[2]);
        u.w = __float_as_uint(f[3]);
    } else {
        T* h = reinterpret_cast<T*>(&u);
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = ElemTraits<T>::from_float(f[i]);
    }
    return u;
}

// 128-bit streaming global load / store (read-only path, no L1 allocation)
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// plain 128-bit load (data that another kernel stage of the same launch may have written)
__device__ __forceinline__ uint4 ld16(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st16(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#define ET_DISPATCH_DTYPE(dtype, T, ...)                                       \
    switch (dtype) {                                                           \
        case ET_F32: { using T = float; __VA_ARGS__; break; }                  \
        case ET_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }         \
        case ET_F16: { using T = __half; __VA_ARGS__; break; }                 \
        default: return et_fail(ET_ERR_ARG, "unknown dtype %d", (int)(dtype)); \
    }

static inline cudaStream_t et_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every kernel begins with et_pdl_prologue(): it lets the NEXT kernel in the stream start launching right away and
// then waits until the PREVIOUS kernel has completed and flushed its writes.  Every launch goes through et_launch(),
// which sets cudaLaunchAttributeProgrammaticStreamSerialization, so launch latency and CTA scheduling of kernel n+1
// overlap the tail of kernel n.  Opt-in with EVENTFUL_B200_PDL=1: measured on B200 it gains nothing under CUDA-graph
// replay (265 vs 273 frames/s), where launches are already back to back.
__device__ __forceinline__ void et_pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
extern int g_et_pdl;
template <typename... KArgs, typename... Args>
inline void et_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_et_pdl;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface through ET_CHECK_LAUNCH
}

// Number of kernels this library has enqueued (bench.py reports it as gpu_launches).
extern long long g_et_launches;
#define ET_COUNT_LAUNCH(n) (g_et_launches += (n))
