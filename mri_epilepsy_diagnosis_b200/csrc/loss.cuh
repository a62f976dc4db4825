// Fused softmax(dim=1) + soft-Dice loss of the segmentation loop (segmentation/routine.py:272-274 with get_dice_score/
// get_dice_loss :239-253):
//     p = softmax(logits, dim=1);  tp = sum_v p*t;  fp = sum_v p*(1-t);  fn = sum_v (1-p)*t      per (n, c), t = targets (N,1,...)
//     loss = mean_{n,c} ( 1 - 2*tp / (2*tp + fp + fn + eps) )
// (the reference scores BOTH channels against the same foreground mask -- SURVEY a-8 -- and so does this).
// 2*tp + fp + fn == sum_v p + sum_v t, so the forward pass is ONE read of logits and targets producing 2C+1 sums per sample;
// the backward pass is one more read and one write of dlogits.  fp32 throughout; block partials are combined in a fixed order.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kDiceMaxC = 8;

template <typename T> __device__ __forceinline__ void load_logits(const T* p, int C, float (&z)[kDiceMaxC]) {
#pragma unroll
    for (int c = 0; c < kDiceMaxC; ++c) z[c] = c < C ? to_f<T>(p[c]) : -INFINITY;
}
template <> __device__ __forceinline__ void load_logits<float>(const float* p, int C, float (&z)[kDiceMaxC]) {
    if (C == 2) {
        const float2 v = *reinterpret_cast<const float2*>(p);
        z[0] = v.x; z[1] = v.y;
#pragma unroll
        for (int c = 2; c < kDiceMaxC; ++c) z[c] = -INFINITY;
        return;
    }
#pragma unroll
    for (int c = 0; c < kDiceMaxC; ++c) z[c] = c < C ? p[c] : -INFINITY;
}

__device__ __forceinline__ void softmax_small(float (&z)[kDiceMaxC], int C) {
    float m = z[0];
#pragma unroll
    for (int c = 1; c < kDiceMaxC; ++c) m = fmaxf(m, z[c]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kDiceMaxC; ++c) { z[c] = c < C ? __expf(z[c] - m) : 0.f; s += z[c]; }
    const float inv = 1.f / s;
#pragma unroll
    for (int c = 0; c < kDiceMaxC; ++c) z[c] *= inv;
}

// partial[(n*chunks + chunk)*(2C+1) + j]: j < C: sum p_c;  C <= j < 2C: sum p_c * t;  j == 2C: sum t
template <typename T>
__global__ void __launch_bounds__(256) dice_fwd_partial_kernel(const T* __restrict__ logits, const float* __restrict__ targets, int C, int64_t S,
                                                               float* __restrict__ partial) {
    __shared__ float red[8][2 * kDiceMaxC + 1];
    const int n = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float sp[kDiceMaxC], stp[kDiceMaxC], st = 0.f;
#pragma unroll
    for (int c = 0; c < kDiceMaxC; ++c) sp[c] = stp[c] = 0.f;
    const T* lg = logits + (int64_t)n * S * C;
    const float* tg = targets + (int64_t)n * S;
    for (int64_t v = (int64_t)chunk * 256 + threadIdx.x; v < S; v += (int64_t)chunks * 256) {
        float z[kDiceMaxC];
        load_logits<T>(lg + v * C, C, z);
        const float t = __ldg(tg + v);
        softmax_small(z, C);
#pragma unroll
        for (int c = 0; c < kDiceMaxC; ++c) { sp[c] += z[c]; stp[c] = fmaf(z[c], t, stp[c]); }
        st += t;
    }
    const int nj = 2 * C + 1;
#pragma unroll
    for (int j = 0; j < 2 * kDiceMaxC + 1; ++j) {
        float v = j < kDiceMaxC ? sp[j] : (j < 2 * kDiceMaxC ? stp[j - kDiceMaxC] : st);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < nj) {
        const int j = threadIdx.x;
        const int src = j < C ? j : (j < 2 * C ? kDiceMaxC + (j - C) : 2 * kDiceMaxC);
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][src];
        partial[((int64_t)n * chunks + chunk) * nj + j] = s;
    }
}

// sums[n*(2C+1) + j] (double-accumulated, fixed order) and the scalar loss; one block
__global__ void __launch_bounds__(256) dice_fwd_final_kernel(const float* __restrict__ partial, int N, int C, int chunks, float eps,
                                                             float* __restrict__ sums, float* __restrict__ loss) {
    __shared__ double acc[256];
    const int nj = 2 * C + 1, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = warp; e < N * nj; e += 8) {
        const int n = e / nj, j = e - n * nj;
        double s = 0.0;
        for (int k0 = lane; k0 < chunks; k0 += 32 * 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int k = k0 + 32 * u; v[u] = k < chunks ? __ldg(partial + ((int64_t)n * chunks + k) * nj + j) : 0.f; }
#pragma unroll
            for (int u = 0; u < 8; ++u) s += (double)v[u];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sums[e] = (float)s;
    }
    __syncthreads();
    double part = 0.0;
    for (int e = threadIdx.x; e < N * C; e += 256) {
        const int n = e / C, c = e - n * C;
        const double Sp = sums[n * nj + c], tp = sums[n * nj + C + c], T = sums[n * nj + 2 * C];
        part += 1.0 - 2.0 * tp / (Sp + T + (double)eps);
    }
    acc[threadIdx.x] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 256; ++i) s += acc[i];
        loss[0] = (float)(s / (double)(N * C));
    }
}

// dlogits = softmax'(.) applied to dL/dp = -dloss/(N*C) * 2*(t*den - tp)/den^2,  den = sum p + sum t + eps
template <typename T>
__global__ void __launch_bounds__(256) dice_bwd_kernel(const T* __restrict__ logits, const float* __restrict__ targets, const float* __restrict__ sums,
                                                       const float* __restrict__ dloss, int N, int C, int64_t S, float eps, T* __restrict__ dlogits) {
    const int n = blockIdx.y, nj = 2 * C + 1;
    float A[kDiceMaxC], B[kDiceMaxC];
    const float scale = -__ldg(dloss) / (float)(N * C);
#pragma unroll
    for (int c = 0; c < kDiceMaxC; ++c) {
        A[c] = B[c] = 0.f;
        if (c < C) {
            const float den = sums[n * nj + c] + sums[n * nj + 2 * C] + eps, tp = sums[n * nj + C + c];
            A[c] = scale * 2.f / den;
            B[c] = -scale * 2.f * tp / (den * den);
        }
    }
    const T* lg = logits + (int64_t)n * S * C;
    const float* tg = targets + (int64_t)n * S;
    T* dl = dlogits + (int64_t)n * S * C;
    for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < S; v += (int64_t)gridDim.x * 256) {
        float z[kDiceMaxC];
        load_logits<T>(lg + v * C, C, z);
        const float t = __ldg(tg + v);
        softmax_small(z, C);
        float g[kDiceMaxC], dot = 0.f;
#pragma unroll
        for (int c = 0; c < kDiceMaxC; ++c) { g[c] = fmaf(A[c], t, B[c]); dot = fmaf(g[c], z[c], dot); }
        if (C == 2 && sizeof(T) == 4) {
            *reinterpret_cast<float2*>(reinterpret_cast<float*>(dl) + v * 2) = make_float2(z[0] * (g[0] - dot), z[1] * (g[1] - dot));
        } else {
#pragma unroll
            for (int c = 0; c < kDiceMaxC; ++c) if (c < C) dl[v * C + c] = from_f<T>(z[c] * (g[c] - dot));
        }
    }
}

inline int dice_chunks(int N, int64_t S) {
    int64_t want = ceil_div(S, 256 * 8);
    int64_t cap = (kNumSMs * 8) / (N > 0 ? N : 1);
    if (cap < 1) cap = 1;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

}  // namespace b200
