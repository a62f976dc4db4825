// HBM-bound streaming kernels: activations, residual add, channel copies, layout transposes.
// All are grid-stride loops over 16-byte vectors with grids sized as a multiple of 148 SMs.
#pragma once
#include "common.cuh"

namespace b200 {

// ------------------------------------------------------------------ activations
template <typename T, int V>
__global__ void __launch_bounds__(256) act_fwd_kernel(int act, float slope, int64_t nvec, const T* __restrict__ x, T* __restrict__ y) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float v[V];
        Pack<T, V>::load(x + i * V, v);
#pragma unroll
        for (int k = 0; k < V; ++k) v[k] = act_apply(v[k], act, slope);
        Pack<T, V>::store(y + i * V, v);
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) act_bwd_kernel(int act, float slope, int64_t nvec, const T* __restrict__ y,
                                                      const T* __restrict__ dy, T* __restrict__ dx) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float o[V], g[V];
        Pack<T, V>::load(y + i * V, o);
        Pack<T, V>::load(dy + i * V, g);
#pragma unroll
        for (int k = 0; k < V; ++k) g[k] *= act_gate(o[k], act, slope);
        Pack<T, V>::store(dx + i * V, g);
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) prelu_fwd_kernel(int64_t nvec, const T* __restrict__ x, const float* __restrict__ a, T* __restrict__ y) {
    const float s = a[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float v[V];
        Pack<T, V>::load(x + i * V, v);
#pragma unroll
        for (int k = 0; k < V; ++k) v[k] = v[k] > 0.f ? v[k] : v[k] * s;
        Pack<T, V>::store(y + i * V, v);
    }
}

// dx = dy * (x > 0 ? 1 : a);  partial[block] = sum dy * x * (x <= 0)
template <typename T, int V>
__global__ void __launch_bounds__(256) prelu_bwd_kernel(int64_t nvec, const T* __restrict__ x, const float* __restrict__ a,
                                                        const T* __restrict__ dy, T* __restrict__ dx, float* __restrict__ partial) {
    const float s = a[0];
    float acc = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float v[V], g[V];
        Pack<T, V>::load(x + i * V, v);
        Pack<T, V>::load(dy + i * V, g);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            bool pos = v[k] > 0.f;
            acc += pos ? 0.f : g[k] * v[k];
            g[k] = pos ? g[k] : g[k] * s;
        }
        Pack<T, V>::store(dx + i * V, g);
    }
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void sum_partials_kernel(int n, const float* __restrict__ partial, float* __restrict__ out) {
    // one warp, fixed order -> deterministic
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) acc += (double)partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) out[0] = (float)acc;
}

template <typename T, int V>
__global__ void __launch_bounds__(256) add_act_kernel(int act, float slope, int64_t nvec, const T* __restrict__ a,
                                                      const T* __restrict__ b, T* __restrict__ y) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float u[V], v[V];
        Pack<T, V>::load(a + i * V, u);
        Pack<T, V>::load(b + i * V, v);
#pragma unroll
        for (int k = 0; k < V; ++k) u[k] = act_apply(u[k] + v[k], act, slope);
        Pack<T, V>::store(y + i * V, u);
    }
}

// ------------------------------------------------------------------ channel-slice copy (concat/split)
template <typename T, int V>
__global__ void __launch_bounds__(256) copy_channels_kernel(int64_t V_vox, int cvec, const T* __restrict__ src, int src_ctot, int src_off,
                                                            T* __restrict__ dst, int dst_ctot, int dst_off) {
    const int64_t total = V_vox * cvec;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = i / cvec;
        int c = (int)(i - v * cvec) * V;
        float t[V];
        Pack<T, V>::load(src + v * src_ctot + src_off + c, t);
        Pack<T, V>::store(dst + v * dst_ctot + dst_off + c, t);
    }
}

// ------------------------------------------------------------------ NCDHW <-> NDHWC (+dtype)
// 32x32 smem tile transpose of the (C, S) matrix of each batch item.
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) transpose_cs_kernel(int C, int64_t S, const TS* __restrict__ src, TD* __restrict__ dst, bool to_cl) {
    // to_cl: src is [n][C][S], dst is [n][S][C];  else the inverse
    __shared__ float tile[32][33];
    const int64_t n = blockIdx.z;
    const int64_t rows = to_cl ? C : S, cols = to_cl ? S : C;     // src matrix is rows x cols
    // blockIdx.x always walks the (large) S dimension, blockIdx.y the channels
    const int64_t r0 = (int64_t)(to_cl ? blockIdx.y : blockIdx.x) * 32, c0 = (int64_t)(to_cl ? blockIdx.x : blockIdx.y) * 32;
    const TS* s = src + n * rows * cols;
    TD* d = dst + n * rows * cols;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        int64_t r = r0 + k, c = c0 + tx;
        tile[k][tx] = (r < rows && c < cols) ? to_f<TS>(s[r * cols + c]) : 0.f;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        int64_t c = c0 + k, r = r0 + tx;                          // dst is cols x rows
        if (r < rows && c < cols) d[c * rows + r] = from_f<TD>(tile[tx][k]);
    }
}

}  // namespace b200
