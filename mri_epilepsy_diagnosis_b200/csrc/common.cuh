// Common device/host helpers for libb200nn (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/b200nn.h"

namespace b200 {

// ---------------------------------------------------------------- error / launch bookkeeping
inline std::string& err_slot() {
    static thread_local std::string s;
    return s;
}
inline std::atomic<uint64_t>& launch_counter() {
    static std::atomic<uint64_t> c{0};
    return c;
}
inline int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err_slot() = buf;
    return 1;
}
#define B200_REQUIRE(cond, ...)                   \
    do {                                          \
        if (!(cond)) return ::b200::fail(__VA_ARGS__); \
    } while (0)

// Launch + count + check (no synchronisation).
#define B200_LAUNCH(kernel, grid, block, smem, stream, ...)                                        \
    do {                                                                                           \
        kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);                  \
        ::b200::launch_counter().fetch_add(1, std::memory_order_relaxed);                          \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ != cudaSuccess) return ::b200::fail("%s launch failed: %s", #kernel, cudaGetErrorString(e__)); \
    } while (0)

constexpr int kNumSMs = 148;   // B200

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- scalar conversions
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------- 16-byte vectors
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T, int V> struct Pack;   // V elements of T moved as one transaction when V == 16/sizeof(T)

// `raw` / ldraw / unpack: the packed form of one transaction -- streaming kernels issue the loads of several rows first and unpack
// row by row, so the loads in flight cost 4 registers each instead of V
template <> struct Pack<float, 4> {
    using raw = float4;
    static __device__ __forceinline__ raw ldraw(const float* p) { return *reinterpret_cast<const float4*>(p); }
    static __device__ __forceinline__ void unpack(const raw& v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
    static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
        float4 v = *reinterpret_cast<const float4*>(p);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&o)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    }
};
template <> struct Pack<float, 2> {
    using raw = float2;
    static __device__ __forceinline__ raw ldraw(const float* p) { return *reinterpret_cast<const float2*>(p); }
    static __device__ __forceinline__ void unpack(const raw& v, float (&o)[2]) { o[0] = v.x; o[1] = v.y; }
    static __device__ __forceinline__ void load(const float* p, float (&o)[2]) {
        float2 v = *reinterpret_cast<const float2*>(p);
        o[0] = v.x; o[1] = v.y;
    }
    static __device__ __forceinline__ void store(float* p, const float (&o)[2]) { *reinterpret_cast<float2*>(p) = make_float2(o[0], o[1]); }
};
template <> struct Pack<float, 1> {
    using raw = float;
    static __device__ __forceinline__ raw ldraw(const float* p) { return *p; }
    static __device__ __forceinline__ void unpack(const raw& v, float (&o)[1]) { o[0] = v; }
    static __device__ __forceinline__ void load(const float* p, float (&o)[1]) { o[0] = *p; }
    static __device__ __forceinline__ void store(float* p, const float (&o)[1]) { *p = o[0]; }
};
template <> struct Pack<__nv_bfloat16, 8> {
    using raw = uint4;
    static __device__ __forceinline__ raw ldraw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
    static __device__ __forceinline__ void unpack(const raw& v, float (&o)[8]) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[8]) {
        uint4 v = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            o[2 * i] = f.x; o[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&o)[8]) {
        uint4 v;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = v;
    }
};
template <> struct Pack<__nv_bfloat16, 1> {
    using raw = __nv_bfloat16;
    static __device__ __forceinline__ raw ldraw(const __nv_bfloat16* p) { return *p; }
    static __device__ __forceinline__ void unpack(const raw& v, float (&o)[1]) { o[0] = __bfloat162float(v); }
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[1]) { o[0] = __bfloat162float(*p); }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&o)[1]) { *p = __float2bfloat16_rn(o[0]); }
};

// V per-channel coefficients from shared memory, re-read at every use (volatile: the compiler must not hoist them into registers
// for the whole loop -- with 3..5 coefficient vectors of V = 8 channels that cost 24..40 registers and held the streaming kernels
// at 2-3 blocks per SM, too few loads in flight to cover the HBM latency)
template <int V> __device__ __forceinline__ void lds_coef(const float* p, float (&o)[V]) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    if constexpr (V % 4 == 0) {
#pragma unroll
        for (int q = 0; q < V / 4; ++q)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[4 * q]), "=f"(o[4 * q + 1]), "=f"(o[4 * q + 2]), "=f"(o[4 * q + 3]) : "r"(a + 16u * q));
    } else {
#pragma unroll
        for (int k = 0; k < V; ++k) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o[k]) : "r"(a + 4u * k));
    }
}

__device__ __forceinline__ float act_apply(float v, int act, float slope) {
    if (act == B200_ACT_RELU) return v < 0.f ? 0.f : v;           // NaN stays NaN, like at::relu (clamp_min)
    if (act == B200_ACT_LEAKY) return v > 0.f ? v : v * slope;
    return v;
}
// derivative gate from the OUTPUT of the activation (slope > 0 so sign(out) == sign(in))
__device__ __forceinline__ float act_gate(float out, int act, float slope) {
    if (act == B200_ACT_RELU) return out > 0.f ? 1.f : 0.f;
    if (act == B200_ACT_LEAKY) return out > 0.f ? 1.f : slope;
    return 1.f;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// grid size for grid-stride streaming kernels: a multiple of the SM count
inline int stream_grid(int64_t work_items, int block, int per_sm = 8) {
    int64_t need = ceil_div(work_items, block);
    int64_t cap = (int64_t)kNumSMs * per_sm;
    if (need >= cap) return (int)cap;
    return (int)(need < 1 ? 1 : need);
}

// dispatch on dtype and on whether 16-byte vectors can be used along the channel dimension
#define B200_DISPATCH_T(dtype, T, ...)                                  \
    do {                                                                 \
        if ((dtype) == B200_F32) { using T = float; __VA_ARGS__; }       \
        else if ((dtype) == B200_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
        else return ::b200::fail("unsupported dtype %d", (int)(dtype));  \
    } while (0)

}  // namespace b200
