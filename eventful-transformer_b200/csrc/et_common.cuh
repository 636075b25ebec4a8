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
    static constexpr bool IS_BF16 = false;
    static __device__ __forceinline__ float to_float(float v) { return v; }
    static __device__ __forceinline__ float from_float(float v) { return v; }
};
template <>
struct ElemTraits<__nv_bfloat16> {
    static constexpr int VEC = 8;
    static constexpr bool IS_BF16 = true;
    static __device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct ElemTraits<__half> {
    static constexpr int VEC = 8;
    static constexpr bool IS_BF16 = false;
    static __device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_float(float v) { return __float2half_rn(v); }
};

// value of v after a round trip through T (round-to-nearest-even), i.e. what a
// torch elementwise op on dtype T would have stored
template <typename T>
__device__ __forceinline__ float round_to(float v) {
    return ElemTraits<T>::to_float(ElemTraits<T>::from_float(v));
}

// Two fp32 values -> one packed pair of 16-bit floats (round-to-nearest-even, same results as the scalar conversions):
// a single F2FP instruction on the ALU pipe instead of two F2F on the quarter-rate conversion pipe.  `lo` -> bits 0-15.
template <typename T>
__device__ __forceinline__ uint32_t pack2_rn(float lo, float hi) {
    static_assert(sizeof(T) == 2, "pack2_rn: 16-bit element types only");
    uint32_t r;
    if constexpr (sizeof(T) == 2 && ElemTraits<T>::IS_BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// a, b <- their values after a round trip through T (pairwise form of round_to)
template <typename T>
__device__ __forceinline__ void round2_to(float& a, float& b) {
    if constexpr (sizeof(T) == 2) {
        const uint32_t r = pack2_rn<T>(a, b);
        if constexpr (ElemTraits<T>::IS_BF16) {
            a = __uint_as_float(r << 16);
            b = __uint_as_float(r & 0xffff0000u);
        } else {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r));
            a = f.x;
            b = f.y;
        }
    }
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2: two fp32 lanes per issue slot) -------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
template <typename T>
__device__ __forceinline__ void unpack16(const uint4& u, float* f) {
    if constexpr (sizeof(T) == 4) {
        f[0] = __uint_as_float(u.x);
        f[1] = __uint_as_float(u.y);
        f[2] = __uint_as_float(u.z);
        f[3] = __uint_as_float(u.w);
    } else {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (ElemTraits<T>::IS_BF16) {
                f[2 * i] = __uint_as_float(w[i] << 16);
                f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
            } else {
                const float2 h2 = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                f[2 * i] = h2.x;
                f[2 * i + 1] = h2.y;
            }
        }
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
        u.x = pack2_rn<T>(f[0], f[1]);
        u.y = pack2_rn<T>(f[2], f[3]);
        u.z = pack2_rn<T>(f[4], f[5]);
        u.w = pack2_rn<T>(f[6], f[7]);
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
__device__ __forceinline__ void et_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void et_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void et_pdl_prologue() {
    et_pdl_trigger();
    et_pdl_wait();
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

// ---------------------------------------------------------------- per-device host state
// SM count of the CURRENT device and "this kernel may use `bytes` of dynamic shared memory on the current device":
// both cached per device (a process may drive several GPUs) behind a mutex (calls may come from several host threads).
int et_sm_count();
int et_raise_smem_impl(const void* kernel, int bytes);
template <typename K>
inline int et_raise_smem(K kernel, int bytes) { return et_raise_smem_impl(reinterpret_cast<const void*>(kernel), bytes); }

// As et_launch, with the CTAs grouped into thread-block clusters of `cluster_x` along grid.x (co-scheduled on one GPC).
template <typename... KArgs, typename... Args>
inline void et_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_et_pdl;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cluster_x > 1 ? 2 : 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Number of kernels this library has enqueued (bench.py reports it as gpu_launches).
extern long long g_et_launches;
#define ET_COUNT_LAUNCH(n) (g_et_launches += (n))
