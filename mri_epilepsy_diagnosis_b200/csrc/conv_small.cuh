// HBM-bound convolutions that are not GEMM-shaped enough for tcgen05: the Cin = 1 stems (3x3x3, pad 1) and the
// 1x1x1 heads with a handful of output channels (segmentation logits).  Direct CUDA-core kernels, fp32 accumulation:
//   stem fwd    thread = 2 output voxels x all CO channels; 27 coalesced input reads per voxel (L1-resident halo), weights
//               broadcast from shared memory as float4, 32-byte bf16 stores            (algorithmic bytes: 4|2 + 2*CO per voxel)
//   stem wgrad  lane = (output channel, row); each half-warp walks one x-row with a 3x3x3 sliding window in registers
//               (9 new broadcast loads + 27 FMAs per voxel and lane), block partials, fixed-order final sum
//   head fwd    Cin/8 lanes per voxel, one 16-byte load each, shuffle reduction         (2*Cin + 4*CO bytes per voxel)
//   head dgrad  the same lane mapping, 16-byte stores
//   head wgrad  per-lane CO x 8 accumulators over a grid-stride voxel loop, shuffle + shared-memory + fixed-order final sum
// Weight operands use the SIMT packing of conv_simt.cuh (fp32 [tap*IC + ic][OCp]), so callers are unaffected.
#pragma once
#include "common.cuh"

namespace b200 {

template <typename T> __device__ __forceinline__ float ldg_f(const T* p);
template <> __device__ __forceinline__ float ldg_f<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldg_f<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(__ushort_as_bfloat16(__ldg(reinterpret_cast<const unsigned short*>(p))));
}

// ------------------------------------------------------------------------------------------------ stem forward
// x [N][D][H][W] (one channel), w fp32 [27][CO], y bf16 [N][D][H][W][CO]
// One thread = output voxels (z, y0, x) and (z, y0+1, x): the 3 x 4 x 3 input neighbourhood (36 values) is loaded once into
// registers through 12 row pointers, every filter tap's weights are read once from shared memory (float4 broadcast) and used
// for both rows.  Consecutive lanes = consecutive x: coalesced loads and 32-byte stores.
template <typename TX, int CO>
__global__ void __launch_bounds__(256) stem3_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                        __nv_bfloat16* __restrict__ y, int N, int D, int H, int W) {
    __shared__ __align__(16) float ws[27 * CO];
    for (int i = threadIdx.x; i < 27 * CO; i += 256) ws[i] = w[i];
    __syncthreads();
    const int H2 = (H + 1) >> 1;
    const int64_t total = (int64_t)N * D * H2 * W;
    for (int64_t id = (int64_t)blockIdx.x * 256 + threadIdx.x; id < total; id += (int64_t)gridDim.x * 256) {
        int64_t r = id;
        const int xq = (int)(r % W); r /= W;
        const int y0 = (int)(r % H2) * 2; r /= H2;
        const int zq = (int)(r % D);
        const int64_t n = r / D;
        float in[3][4][3];
#pragma unroll
        for (int kz = 0; kz < 3; ++kz) {
            const int z = zq + kz - 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int yy = y0 + j - 1;
                const bool rok = (unsigned)z < (unsigned)D && (unsigned)yy < (unsigned)H;
                const TX* rp = x + ((n * D + (rok ? z : 0)) * H + (rok ? yy : 0)) * (int64_t)W + xq;
                in[kz][j][0] = (rok && xq > 0) ? ldg_f<TX>(rp - 1) : 0.f;
                in[kz][j][1] = rok ? ldg_f<TX>(rp) : 0.f;
                in[kz][j][2] = (rok && xq + 1 < W) ? ldg_f<TX>(rp + 1) : 0.f;
            }
        }
        float acc[2][CO];
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[0][c] = acc[1][c] = bias != nullptr ? __ldg(bias + c) : 0.f;
#pragma unroll
        for (int kz = 0; kz < 3; ++kz)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float a0 = in[kz][ky][kx], a1 = in[kz][ky + 1][kx];
                    const float4* wt = reinterpret_cast<const float4*>(ws + ((kz * 3 + ky) * 3 + kx) * CO);
#pragma unroll
                    for (int q = 0; q < CO / 4; ++q) {
                        const float4 f = wt[q];
                        acc[0][4 * q + 0] = fmaf(a0, f.x, acc[0][4 * q + 0]); acc[1][4 * q + 0] = fmaf(a1, f.x, acc[1][4 * q + 0]);
                        acc[0][4 * q + 1] = fmaf(a0, f.y, acc[0][4 * q + 1]); acc[1][4 * q + 1] = fmaf(a1, f.y, acc[1][4 * q + 1]);
                        acc[0][4 * q + 2] = fmaf(a0, f.z, acc[0][4 * q + 2]); acc[1][4 * q + 2] = fmaf(a1, f.z, acc[1][4 * q + 2]);
                        acc[0][4 * q + 3] = fmaf(a0, f.w, acc[0][4 * q + 3]); acc[1][4 * q + 3] = fmaf(a1, f.w, acc[1][4 * q + 3]);
                    }
                }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (y0 + u >= H) continue;
            __nv_bfloat16* dst = y + ((((n * D + zq) * H + y0 + u) * (int64_t)W) + xq) * CO;
#pragma unroll
            for (int q = 0; q < CO / 8; ++q) {
                uint4 pk;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(acc[u][8 * q + 2 * i], acc[u][8 * q + 2 * i + 1]);
                *reinterpret_cast<uint4*>(dst + 8 * q) = pk;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ stem wgrad
// partial[block][CO][28]: 27 taps + the bias gradient; dy bf16 [N][D][H][W][CO]
template <typename TX, int CO>
__global__ void __launch_bounds__(256) stem3_wgrad_kernel(const TX* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ partial,
                                                          int N, int D, int H, int W) {
    constexpr int RPW = 32 / CO;                 // rows walked concurrently by one warp
    __shared__ float red[8][CO][28];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int co = lane % CO, sub = lane / CO;
    float acc[28];
#pragma unroll
    for (int i = 0; i < 28; ++i) acc[i] = 0.f;
    const int64_t rows = (int64_t)N * D * H;
    const int64_t walkers = (int64_t)gridDim.x * 8 * RPW;
    for (int64_t row = ((int64_t)blockIdx.x * 8 + warp) * RPW + sub; row < rows; row += walkers) {
        const int yq = (int)(row % H);
        const int zq = (int)((row / H) % D);
        const int64_t n = row / ((int64_t)H * D);
        const TX* rp[9];
        bool rok[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const int z = zq + r / 3 - 1, yy = yq + r % 3 - 1;
            rok[r] = (unsigned)z < (unsigned)D && (unsigned)yy < (unsigned)H;
            rp[r] = x + ((n * D + (rok[r] ? z : 0)) * H + (rok[r] ? yy : 0)) * (int64_t)W;
        }
        const __nv_bfloat16* g = dy + row * (int64_t)W * CO + co;
        // sliding window per input row: win = (x-1, x, x+1).  (Measured alternatives that ran SLOWER on B200: a 4-voxel float4
        // variant -- 149 registers, one block per SM, 1.47 ms -- and a one-step-ahead prefetch of column x+2 -- 0.94 ms; this
        // plain form runs the 4 x 128^3 stem in 0.65 ms.)
        float win[9][3];
#pragma unroll
        for (int r = 0; r < 9; ++r) { win[r][0] = 0.f; win[r][1] = 0.f; win[r][2] = rok[r] ? ldg_f<TX>(rp[r]) : 0.f; }
        for (int xq = 0; xq < W; ++xq) {
#pragma unroll
            for (int r = 0; r < 9; ++r) {
                win[r][0] = win[r][1]; win[r][1] = win[r][2];
                win[r][2] = (rok[r] && xq + 1 < W) ? ldg_f<TX>(rp[r] + xq + 1) : 0.f;
            }
            const float gv = ldg_f<__nv_bfloat16>(g + (int64_t)xq * CO);
#pragma unroll
            for (int r = 0; r < 9; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k) acc[r * 3 + k] = fmaf(gv, win[r][k], acc[r * 3 + k]);
            acc[27] += gv;
        }
    }
    // lanes with the same co (different sub-rows) -> one value per warp, then across the 8 warps in a fixed order
#pragma unroll
    for (int i = 0; i < 28; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = CO; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (sub == 0) red[warp][co][i] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < CO * 28; e += 256) {
        float s = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) s += (&red[wq][0][0])[e];
        partial[(int64_t)blockIdx.x * CO * 28 + e] = s;
    }
}

// dw[co*27 + tap] / dbias[co] = sum over blocks (one warp per element, fixed order)
__global__ void __launch_bounds__(256) stem3_wgrad_reduce_kernel(const float* __restrict__ partial, int blocks, int CO, float* __restrict__ dw,
                                                                 float* __restrict__ dbias) {
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (e >= CO * 28) return;
    double s = 0.0;
    for (int b = lane; b < blocks; b += 32) s += (double)partial[(int64_t)b * CO * 28 + e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const int co = e / 28, i = e % 28;
        if (i < 27) dw[co * 27 + i] = (float)s;
        else if (dbias != nullptr) dbias[co] = (float)s;
    }
}

// ------------------------------------------------------------------------------------------------ 1x1x1 heads (few outputs)
// x bf16 [V][CI]; w fp32 [CI][OCp] (OCp = CO rounded up to 4); y [V][CO] (fp32 or bf16)
template <typename TY, int CO>
__global__ void __launch_bounds__(256) head_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                       TY* __restrict__ y, int64_t V, int CI) {
    const int LPV = CI / 8, OCp = (CO + 3) & ~3;
    const int lane = threadIdx.x & 31, j = lane % LPV;
    float wr[CO][8], bv[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) {
        bv[c] = bias != nullptr ? __ldg(bias + c) : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) wr[c][k] = __ldg(w + (int64_t)(j * 8 + k) * OCp + c);
    }
    const int64_t total = V * LPV;               // one 16-byte chunk per thread and step
    const int sh = 31 - __clz(LPV);              // LPV is a power of two
    const int64_t stride = (int64_t)gridDim.x * 256;
    constexpr int U = 4;                         // independent 16-byte loads in flight per thread (the kernel is a pure stream)
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i - lane < total; i += U * stride) {
        float xv[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t idx = i + u * stride;
            if (idx < total) Pack<__nv_bfloat16, 8>::load(x + idx * 8, xv[u]);
            else {
#pragma unroll
                for (int k = 0; k < 8; ++k) xv[u][k] = 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t idx = i + u * stride;
            if (idx - lane >= total) break;                                      // warp-uniform
            float s[CO];
#pragma unroll
            for (int c = 0; c < CO; ++c) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) a = fmaf(xv[u][k], wr[c][k], a);
                for (int o = 1; o < LPV; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                s[c] = a + bv[c];
            }
            if (idx < total && j == 0) {
                TY* dst = y + (idx >> sh) * CO;
#pragma unroll
                for (int c = 0; c < CO; ++c) dst[c] = from_f<TY>(s[c]);
            }
        }
    }
}

// dx[v][ci] = sum_co dy[v][co] * w[co][ci];  w fp32 [CO(ic)][OCp = CI]  (the SIMT dgrad packing)
template <typename TY, int CO>
__global__ void __launch_bounds__(256) head_dgrad_kernel(const TY* __restrict__ dy, const float* __restrict__ w, __nv_bfloat16* __restrict__ dx,
                                                         int64_t V, int CI, int OCp) {
    const int LPV = CI / 8;
    const int j = (int)(((int64_t)blockIdx.x * 256 + threadIdx.x) % LPV);       // gridDim.x*256 is a multiple of LPV
    float wr[CO][8];
#pragma unroll
    for (int c = 0; c < CO; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) wr[c][k] = __ldg(w + (int64_t)c * OCp + j * 8 + k);
    const int64_t total = V * LPV;
    const int sh = 31 - __clz(LPV);
    const int64_t stride = (int64_t)gridDim.x * 256;
    constexpr int U = 4;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += U * stride) {
        float g[U][CO];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t idx = i + u * stride;
#pragma unroll
            for (int c = 0; c < CO; ++c) g[u][c] = idx < total ? to_f<TY>(dy[(idx >> sh) * CO + c]) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t idx = i + u * stride;
            if (idx >= total) break;
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = 0.f;
#pragma unroll
            for (int c = 0; c < CO; ++c)
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = fmaf(g[u][c], wr[c][k], o[k]);
            Pack<__nv_bfloat16, 8>::store(dx + idx * 8, o);
        }
    }
}

// partial[block][CO][CI + 1]: dw[co][ci] and (last column) dbias[co]
template <typename TY, int CO>
__global__ void __launch_bounds__(256) head_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const TY* __restrict__ dy, float* __restrict__ partial,
                                                         int64_t V, int CI) {
    extern __shared__ float red[];               // [8 warps][CO][CI + 1]
    const int LPV = CI / 8, stride_c = CI + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = (int)(((int64_t)blockIdx.x * 256 + threadIdx.x) % LPV);
    float acc[CO][8], accb[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) {
        accb[c] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[c][k] = 0.f;
    }
    const int64_t total = V * LPV;
    const int sh = 31 - __clz(LPV);
    const int64_t stride = (int64_t)gridDim.x * 256;
    constexpr int U = 4;                         // (the order of the additions per thread is unchanged: u ascending = i ascending)
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += U * stride) {
        float xv[U][8], g[U][CO];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t idx = i + u * stride;
            if (idx < total) {
                Pack<__nv_bfloat16, 8>::load(x + idx * 8, xv[u]);
#pragma unroll
                for (int c = 0; c < CO; ++c) g[u][c] = to_f<TY>(dy[(idx >> sh) * CO + c]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) xv[u][k] = 0.f;
#pragma unroll
                for (int c = 0; c < CO; ++c) g[u][c] = 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int c = 0; c < CO; ++c) {
                accb[c] += g[u][c];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[c][k] = fmaf(g[u][c], xv[u][k], acc[c][k]);
            }
    }
    // lanes j, j+LPV, ... hold the same channels
#pragma unroll
    for (int c = 0; c < CO; ++c) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float a = acc[c][k];
            for (int o = LPV; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane < LPV) red[(warp * CO + c) * stride_c + j * 8 + k] = a;
        }
        float b = accb[c];
        for (int o = LPV; o < 32; o <<= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
        if (lane == 0) red[(warp * CO + c) * stride_c + CI] = b;       // every chunk lane saw every voxel once: take lane 0's
    }
    __syncthreads();
    for (int e = threadIdx.x; e < CO * stride_c; e += 256) {
        float s = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) s += red[wq * CO * stride_c + e];
        partial[(int64_t)blockIdx.x * CO * stride_c + e] = s;
    }
}

__global__ void __launch_bounds__(256) head_wgrad_reduce_kernel(const float* __restrict__ partial, int blocks, int CO, int CI, float* __restrict__ dw,
                                                                float* __restrict__ dbias) {
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int stride_c = CI + 1;
    if (e >= CO * stride_c) return;
    double s = 0.0;
    for (int b = lane; b < blocks; b += 32) s += (double)partial[(int64_t)b * CO * stride_c + e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const int co = e / stride_c, ci = e % stride_c;
        if (ci < CI) dw[co * CI + ci] = (float)s;
        else if (dbias != nullptr) dbias[co] = (float)s;
    }
}

// ------------------------------------------------------------------------------------------------ host side
// ------------------------------------------------------------------------------------------------ stem wgrad on tensor cores
// The FFMA kernel above is bound by its 432 multiply-adds per voxel (0.64 ms on 4 x 128^3, 16 % of the fp32 FMA rate, ~7 % of the
// HBM peak).  The same gradient is a (1,3,3) weight gradient over a 16-channel tensor whose channels are the three z-shifted copies
// of x -- each split into a bf16 "hi" and a bf16 "lo" part so that the fp32 volume loses nothing that matters (|x - hi - lo| <=
// 2^-17 |x|) -- which is exactly the shape the tcgen05 row_wgrad_kernel runs at its best:
//     X16[v][kz] = hi(x[z + kz - 1]),  X16[v][3 + kz] = lo(x[z + kz - 1]),  channels 6..15 = 0
//     dW[co][kz][ky][kx] = dW16[co][kz][ky][kx] + dW16[co][3 + kz][ky][kx]
// Cost: one expansion pass (read 4 B, write 32 B per voxel) + a 9-tap 16 -> Co tcgen05 wgrad + a 432-element combine.
template <typename TX>
__global__ void __launch_bounds__(256) stem3_expand_kernel(const TX* __restrict__ x, int N, int D, int H, int W, __nv_bfloat16* __restrict__ x16) {
    const int64_t plane = (int64_t)H * W, total = (int64_t)N * D * plane;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)((v / plane) % D);
        float val[3];
        val[0] = z > 0 ? ldg_f<TX>(x + v - plane) : 0.f;
        val[1] = ldg_f<TX>(x + v);
        val[2] = z + 1 < D ? ldg_f<TX>(x + v + plane) : 0.f;
        __nv_bfloat16 hi[3], lo[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            hi[k] = __float2bfloat16_rn(val[k]);
            lo[k] = __float2bfloat16_rn(val[k] - __bfloat162float(hi[k]));
        }
        uint4 a, b = make_uint4(0u, 0u, 0u, 0u);
        a.x = (uint32_t)__bfloat16_as_ushort(hi[0]) | ((uint32_t)__bfloat16_as_ushort(hi[1]) << 16);
        a.y = (uint32_t)__bfloat16_as_ushort(hi[2]) | ((uint32_t)__bfloat16_as_ushort(lo[0]) << 16);
        a.z = (uint32_t)__bfloat16_as_ushort(lo[1]) | ((uint32_t)__bfloat16_as_ushort(lo[2]) << 16);
        a.w = 0u;
        uint4* dst = reinterpret_cast<uint4*>(x16 + v * 16);
        dst[0] = a;
        dst[1] = b;
    }
}

__global__ void stem3_combine_kernel(const float* __restrict__ dw16 /* [Co][16][9] */, int Co, float* __restrict__ dw /* [Co][1][27] */) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Co * 27) return;
    const int co = e / 27, t = e - co * 27, kz = t / 9, r = t - kz * 9;
    dw[e] = dw16[(co * 16 + kz) * 9 + r] + dw16[(co * 16 + 3 + kz) * 9 + r];
}

// ------------------------------------------------------------------------------------------------ 1 -> 1 channel 3x3x3 convolution
// The autoencoder's final `vox` layer (AE_model.py:160-164: Conv3d(1, 1, 3, padding=1) on the full-resolution volume).  27 MACs per
// voxel: the generic implicit-GEMM kernel spent > 1 ms per pass on it (128 x 16 tiles with one useful column); these are plain
// streaming kernels.  `flip` = 1 evaluates the transposed geometry (dgrad): src = v + 1 - k instead of v - 1 + k.
// w: the SIMT packing [27][4] fp32 (column 0).
template <typename TX, typename TY>
__global__ void __launch_bounds__(256) c1k3_gather_kernel(const TX* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                          TY* __restrict__ y, int N, int D, int H, int W, int flip) {
    __shared__ float ws[27];
    if (threadIdx.x < 27) ws[threadIdx.x] = w[threadIdx.x * 4];
    __syncthreads();
    const float b = bias != nullptr ? bias[0] : 0.f;
    const int64_t total = (int64_t)N * D * H * W;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = v;
        const int xq = (int)(r % W); r /= W;
        const int yq = (int)(r % H); r /= H;
        const int zq = (int)(r % D);
        float acc = b;
#pragma unroll
        for (int kz = 0; kz < 3; ++kz) {
            const int z = zq + (flip ? 1 - kz : kz - 1);
            if ((unsigned)z >= (unsigned)D) continue;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = yq + (flip ? 1 - ky : ky - 1);
                if ((unsigned)yy >= (unsigned)H) continue;
                const TX* row = x + v + ((int64_t)(z - zq) * H + (yy - yq)) * W;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = xq + (flip ? 1 - kx : kx - 1);
                    if ((unsigned)xx < (unsigned)W) acc = fmaf(ldg_f<TX>(row + (xx - xq)), ws[(kz * 3 + ky) * 3 + kx], acc);
                }
            }
        }
        y[v] = from_f<TY>(acc);
    }
}

// partial[block][1][28]: dw[tap] = sum_v dy[v] * x[v + tap - 1], [27] = sum_v dy[v]   (reduced by stem3_wgrad_reduce_kernel, CO = 1)
template <typename TX, typename TG>
__global__ void __launch_bounds__(256) c1k3_wgrad_kernel(const TX* __restrict__ x, const TG* __restrict__ dy, float* __restrict__ partial, int N, int D,
                                                         int H, int W) {
    float acc[28];
#pragma unroll
    for (int i = 0; i < 28; ++i) acc[i] = 0.f;
    const int64_t total = (int64_t)N * D * H * W;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = v;
        const int xq = (int)(r % W); r /= W;
        const int yq = (int)(r % H); r /= H;
        const int zq = (int)(r % D);
        const float g = to_f<TG>(dy[v]);
        acc[27] += g;
#pragma unroll
        for (int kz = 0; kz < 3; ++kz) {
            const int z = zq + kz - 1;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = yq + ky - 1;
                const bool rok = (unsigned)z < (unsigned)D && (unsigned)yy < (unsigned)H;
                const TX* row = x + v + ((int64_t)(kz - 1) * H + (ky - 1)) * W;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = xq + kx - 1;
                    const float xv = (rok && (unsigned)xx < (unsigned)W) ? ldg_f<TX>(row + (kx - 1)) : 0.f;
                    acc[(kz * 3 + ky) * 3 + kx] = fmaf(g, xv, acc[(kz * 3 + ky) * 3 + kx]);
                }
            }
        }
    }
    __shared__ float red[8][28];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 28; ++i) {
        float a = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) red[warp][i] = a;
    }
    __syncthreads();
    if (threadIdx.x < 28) {
        float a = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) a += red[wq][threadIdx.x];
        partial[(int64_t)blockIdx.x * 28 + threadIdx.x] = a;
    }
}

inline bool c1k3_supported(const b200_conv_desc* d) {
    return !d->transposed && d->Ci == 1 && d->Co == 1 && d->kd == 3 && d->kh == 3 && d->kw == 3 && d->sd == 1 && d->sh == 1 && d->sw == 1 && d->pd == 1 &&
           d->ph == 1 && d->pw == 1 && d->dd == 1 && d->dh == 1 && d->dw == 1;
}

inline bool stem3_supported(const b200_conv_desc* d) {
    return !d->transposed && d->Ci == 1 && (d->Co == 8 || d->Co == 16 || d->Co == 32) && d->kd == 3 && d->kh == 3 && d->kw == 3 && d->sd == 1 &&
           d->sh == 1 && d->sw == 1 && d->pd == 1 && d->ph == 1 && d->pw == 1 && d->dd == 1 && d->dh == 1 && d->dw == 1 && d->y_dtype == B200_BF16;
}
inline bool head_supported(const b200_conv_desc* d) {
    if (d->transposed || d->kd != 1 || d->kh != 1 || d->kw != 1 || d->sd != 1 || d->sh != 1 || d->sw != 1 || d->pd || d->ph || d->pw) return false;
    if (d->x_dtype != B200_BF16 || d->Co < 1 || d->Co > 5) return false;
    const int lpv = d->Ci / 8;
    return d->Ci % 8 == 0 && lpv >= 1 && lpv <= 32 && (lpv & (lpv - 1)) == 0;
}
constexpr int kSmallBlocks = kNumSMs * 4;
inline size_t stem3_wgrad_ws_bytes(const b200_conv_desc* d) { return (size_t)kSmallBlocks * d->Co * 28 * 4; }
inline size_t head_wgrad_ws_bytes(const b200_conv_desc* d) { return (size_t)kSmallBlocks * d->Co * (d->Ci + 1) * 4; }

#define B200_STEM_CO(CO_, ...)                                 \
    do {                                                       \
        if ((CO_) == 8) { constexpr int CO = 8; __VA_ARGS__; }  \
        else if ((CO_) == 16) { constexpr int CO = 16; __VA_ARGS__; } \
        else { constexpr int CO = 32; __VA_ARGS__; }            \
    } while (0)
#define B200_HEAD_CO(CO_, ...)                                 \
    do {                                                       \
        switch (CO_) {                                         \
            case 1: { constexpr int CO = 1; __VA_ARGS__; } break; \
            case 2: { constexpr int CO = 2; __VA_ARGS__; } break; \
            case 3: { constexpr int CO = 3; __VA_ARGS__; } break; \
            case 4: { constexpr int CO = 4; __VA_ARGS__; } break; \
            default: { constexpr int CO = 5; __VA_ARGS__; } break; \
        }                                                      \
    } while (0)

inline int stem3_fwd_run(const b200_conv_desc* d, const void* x, const float* w, const float* bias, void* y, void* stream) {
    const int64_t work = (int64_t)d->N * d->Di * ((d->Hi + 1) / 2) * d->Wi;
    const int grid = (int)(ceil_div(work, 256) < kNumSMs * 8 ? ceil_div(work, 256) : kNumSMs * 8);
    B200_STEM_CO(d->Co, {
        if (d->x_dtype == B200_F32)
            B200_LAUNCH((stem3_fwd_kernel<float, CO>), grid, 256, 0, stream, (const float*)x, w, bias, (__nv_bfloat16*)y, d->N, d->Di, d->Hi, d->Wi);
        else
            B200_LAUNCH((stem3_fwd_kernel<__nv_bfloat16, CO>), grid, 256, 0, stream, (const __nv_bfloat16*)x, w, bias, (__nv_bfloat16*)y, d->N, d->Di,
                        d->Hi, d->Wi);
    });
    return 0;
}

inline int stem3_wgrad_run(const b200_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, void* workspace, void* stream) {
    float* partial = (float*)workspace;
    const int64_t rows = (int64_t)d->N * d->Di * d->Hi;
    B200_STEM_CO(d->Co, {
        const int64_t need = ceil_div(rows, 8 * (32 / CO));
        const int grid = (int)(need < kSmallBlocks ? need : kSmallBlocks);
        if (d->x_dtype == B200_F32)
            B200_LAUNCH((stem3_wgrad_kernel<float, CO>), grid, 256, 0, stream, (const float*)x, (const __nv_bfloat16*)dy, partial, d->N, d->Di, d->Hi,
                        d->Wi);
        else
            B200_LAUNCH((stem3_wgrad_kernel<__nv_bfloat16, CO>), grid, 256, 0, stream, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, partial, d->N,
                        d->Di, d->Hi, d->Wi);
        B200_LAUNCH(stem3_wgrad_reduce_kernel, (int)ceil_div(CO * 28, 8), 256, 0, stream, partial, grid, CO, dw, dbias);
    });
    return 0;
}

#define B200_C1K3_DT(in_dt, out_dt, TI, TO, ...)                                                     \
    do {                                                                                             \
        if ((in_dt) == B200_F32 && (out_dt) == B200_BF16) { using TI = float; using TO = __nv_bfloat16; __VA_ARGS__; }           \
        else if ((in_dt) == B200_BF16 && (out_dt) == B200_BF16) { using TI = __nv_bfloat16; using TO = __nv_bfloat16; __VA_ARGS__; } \
        else if ((in_dt) == B200_BF16 && (out_dt) == B200_F32) { using TI = __nv_bfloat16; using TO = float; __VA_ARGS__; }       \
        else { using TI = float; using TO = float; __VA_ARGS__; }                                    \
    } while (0)

// pass: B200_PASS_FWD (x -> y) or B200_PASS_DGRAD (dy -> dx, flipped taps)
inline int c1k3_gather_run(const b200_conv_desc* d, int pass, const void* in, const float* w, const float* bias, void* out, void* stream) {
    const int64_t V = (int64_t)d->N * d->Di * d->Hi * d->Wi;
    const int in_dt = pass == B200_PASS_DGRAD ? d->y_dtype : d->x_dtype, out_dt = pass == B200_PASS_DGRAD ? d->x_dtype : d->y_dtype;
    B200_C1K3_DT(in_dt, out_dt, TI, TO, {
        B200_LAUNCH((c1k3_gather_kernel<TI, TO>), stream_grid(V, 256, 16), 256, 0, stream, (const TI*)in, w, bias, (TO*)out, d->N, d->Di, d->Hi, d->Wi,
                    pass == B200_PASS_DGRAD ? 1 : 0);
    });
    return 0;
}
inline size_t c1k3_wgrad_ws_bytes() { return (size_t)kSmallBlocks * 28 * 4; }
inline int c1k3_wgrad_run(const b200_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, void* workspace, void* stream) {
    const int64_t V = (int64_t)d->N * d->Di * d->Hi * d->Wi;
    const int64_t need = ceil_div(V, 256 * 8);
    const int grid = (int)(need < kSmallBlocks ? (need < 1 ? 1 : need) : kSmallBlocks);
    float* partial = (float*)workspace;
    B200_C1K3_DT(d->x_dtype, d->y_dtype, TX, TG, {
        B200_LAUNCH((c1k3_wgrad_kernel<TX, TG>), grid, 256, 0, stream, (const TX*)x, (const TG*)dy, partial, d->N, d->Di, d->Hi, d->Wi);
    });
    B200_LAUNCH(stem3_wgrad_reduce_kernel, (int)ceil_div(28, 8), 256, 0, stream, partial, grid, 1, dw, dbias);
    return 0;
}

inline int head_grid(int64_t V, int lpv, int cap = kSmallBlocks) {
    int64_t need = ceil_div(V * lpv, 256 * 4);
    if (need > cap) need = cap;
    if (need < 1) need = 1;
    return (int)need;                            // 256 threads per block is a multiple of every lpv (power of two <= 32)
}
// fwd / dgrad keep no per-block partials: a fine grid (one 4-chunk pass per thread, up to 32 blocks per SM in flight over the run)
// instead of a persistent one whose block count (4 per SM) did not match the 3 blocks per SM the registers allow (1.33 waves)
constexpr int kHeadStreamBlocks = kNumSMs * 32;

inline int head_fwd_run(const b200_conv_desc* d, const void* x, const float* w, const float* bias, void* y, void* stream) {
    const int64_t V = (int64_t)d->N * d->Di * d->Hi * d->Wi;
    const int grid = head_grid(V, d->Ci / 8, kHeadStreamBlocks);
    B200_HEAD_CO(d->Co, {
        if (d->y_dtype == B200_F32) B200_LAUNCH((head_fwd_kernel<float, CO>), grid, 256, 0, stream, (const __nv_bfloat16*)x, w, bias, (float*)y, V, d->Ci);
        else B200_LAUNCH((head_fwd_kernel<__nv_bfloat16, CO>), grid, 256, 0, stream, (const __nv_bfloat16*)x, w, bias, (__nv_bfloat16*)y, V, d->Ci);
    });
    return 0;
}

inline int head_dgrad_run(const b200_conv_desc* d, const void* dy, const float* w, void* dx, void* stream) {
    const int64_t V = (int64_t)d->N * d->Di * d->Hi * d->Wi;
    const int grid = head_grid(V, d->Ci / 8, kHeadStreamBlocks);
    const int OCp = (d->Ci + 3) & ~3;
    B200_HEAD_CO(d->Co, {
        if (d->y_dtype == B200_F32) B200_LAUNCH((head_dgrad_kernel<float, CO>), grid, 256, 0, stream, (const float*)dy, w, (__nv_bfloat16*)dx, V, d->Ci, OCp);
        else B200_LAUNCH((head_dgrad_kernel<__nv_bfloat16, CO>), grid, 256, 0, stream, (const __nv_bfloat16*)dy, w, (__nv_bfloat16*)dx, V, d->Ci, OCp);
    });
    return 0;
}

inline int head_wgrad_run(const b200_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, void* workspace, void* stream) {
    const int64_t V = (int64_t)d->N * d->Di * d->Hi * d->Wi;
    const int grid = head_grid(V, d->Ci / 8);
    float* partial = (float*)workspace;
    const size_t smem = (size_t)8 * d->Co * (d->Ci + 1) * sizeof(float);
    B200_HEAD_CO(d->Co, {
        if (d->y_dtype == B200_F32)
            B200_LAUNCH((head_wgrad_kernel<float, CO>), grid, 256, smem, stream, (const __nv_bfloat16*)x, (const float*)dy, partial, V, d->Ci);
        else
            B200_LAUNCH((head_wgrad_kernel<__nv_bfloat16, CO>), grid, 256, smem, stream, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, partial, V, d->Ci);
        B200_LAUNCH(head_wgrad_reduce_kernel, (int)ceil_div(CO * (d->Ci + 1), 8), 256, 0, stream, partial, grid, CO, d->Ci, dw, dbias);
    });
    return 0;
}

}  // namespace b200
