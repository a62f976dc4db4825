// Convolutions on TINY volumes (< 16384 output voxels in the whole batch; wgrad: <= 4096): the deep levels of the autoencoder of BASELINE config 1
// (AE_model.py:4-120, channels up to 512 on 2^3 / 4^3 / 8^3 volumes, separable (3,1,1) kernels) and the fader heads.
// They are too small for a tensor-core tile grid, and the generic implicit-GEMM kernel (conv_simt.cuh) ran them as a handful of CTAs
// walking K = taps * Cin in 16-wide chunks with two barriers each: latency-bound, ~0.3 ms for 25 MFLOP (ncu, round 2).
// Here every output element gets its own thread, K is walked without barriers, and the only shared data is tiny:
//   fwd / dgrad  block = 128 consecutive output channels x VT output voxels; per tap the VT source voxels' channels are staged in
//                shared memory (broadcast reads), weights are read coalesced from the SIMT packing [tap*IC + ic][OCp] (L2-resident)
//   wgrad        block = 128 consecutive output channels x one (tap, ic); loops over all voxels (dy coalesced, x broadcast from a
//                per-block table of source offsets); writes dw in the parameter's layout directly, and the bias gradient
// fp32 accumulation in a fixed order: deterministic.
#pragma once
#include "common.cuh"
#include "conv_simt.cuh"

namespace b200 {

constexpr int kTinyMaxVox = 16384;        // VT = 4 voxels per block below 1024 voxels, 16 above (fewer re-reads of the weights through L2)

template <typename TI, typename TO, int kTinyVT>
__global__ void __launch_bounds__(128) conv_tiny_kernel(GatherGeom g, const TI* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                        TO* __restrict__ out) {
    extern __shared__ float xs[];                              // [VT][IC]
    __shared__ int64_t src[kTinyVT];                            // element offset of the source voxel of the current tap, or -1
    const int V = g.N * g.OD * g.OH * g.OW;
    const int v0 = blockIdx.x * kTinyVT;
    const int co = blockIdx.y * 128 + threadIdx.x;
    const int taps_hw = g.kh * g.kw, taps = g.kd * taps_hw;
    float acc[kTinyVT];
#pragma unroll
    for (int q = 0; q < kTinyVT; ++q) acc[q] = 0.f;
    for (int tap = 0; tap < taps; ++tap) {
        __syncthreads();                                        // xs / src of the previous tap are no longer read
        if (threadIdx.x < kTinyVT) {
            int64_t off = -1;
            int v = v0 + threadIdx.x;
            if (v < V) {
                const int x = v % g.OW; v /= g.OW;
                const int y = v % g.OH; v /= g.OH;
                const int z = v % g.OD, n = v / g.OD;
                const int kz = tap / taps_hw, kr = tap - kz * taps_hw, ky = kr / g.kw, kx = kr - ky * g.kw;
                int sz, sy, sx;
                if (src_coord(z, kz, g.sd, g.pd, g.dd, g.ID, g.transposed, &sz) && src_coord(y, ky, g.sh, g.ph, g.dh, g.IH, g.transposed, &sy) &&
                    src_coord(x, kx, g.sw, g.pw, g.dw, g.IW, g.transposed, &sx))
                    off = ((((int64_t)n * g.ID + sz) * g.IH + sy) * g.IW + sx) * g.IC;
            }
            src[threadIdx.x] = off;
        }
        __syncthreads();
        bool any = false;
#pragma unroll
        for (int q = 0; q < kTinyVT; ++q) any |= src[q] >= 0;
        if (!any) continue;                                     // block-uniform
        for (int e = threadIdx.x; e < kTinyVT * g.IC; e += 128) {
            const int q = e / g.IC, ic = e - q * g.IC;
            xs[e] = src[q] >= 0 ? to_f<TI>(in[src[q] + ic]) : 0.f;
        }
        __syncthreads();
        if (co < g.OC) {
            const float* wk = w + (int64_t)tap * g.IC * g.OCp + co;
#pragma unroll 4
            for (int ic = 0; ic < g.IC; ++ic) {
                const float wv = __ldg(wk + (int64_t)ic * g.OCp);
#pragma unroll
                for (int q = 0; q < kTinyVT; ++q) acc[q] = fmaf(xs[q * g.IC + ic], wv, acc[q]);
            }
        }
    }
    if (co >= g.OC) return;
    const float b = bias != nullptr ? bias[co] : 0.f;
#pragma unroll
    for (int q = 0; q < kTinyVT; ++q)
        if (v0 + q < V) out[(int64_t)(v0 + q) * g.OC + co] = from_f<TO>(acc[q] + b);
}

// dw[co][ci][tap] (param_is_ci_major = 0) = sum_v x[src(v, tap)][ci] * dy[v][co];  dbias[co] = sum_v dy[v][co]
// grid (taps * IC, ceil(OC / 128)); v walks the (OD,OH,OW) grid of dy (g is the WGRAD plan: gathered = x, second = dy)
constexpr int kTinyWgSlices = 4;         // the voxel loop of the wgrad kernel is split over threadIdx.y (fixed-order combine: deterministic)

template <typename TX, typename TG>
__global__ void __launch_bounds__(128 * kTinyWgSlices) conv_tiny_wgrad_kernel(GatherGeom g, const TX* __restrict__ x, const TG* __restrict__ gy,
                                                                              float* __restrict__ dw, float* __restrict__ dbias) {
    extern __shared__ float xv[];                               // [V] x value of this block's (tap, ic) at every output voxel (0 where padded)
    __shared__ float red[2][kTinyWgSlices][128];
    const int V = g.N * g.OD * g.OH * g.OW;
    const int tap = blockIdx.x / g.IC, ic = blockIdx.x - tap * g.IC;
    const int taps_hw = g.kh * g.kw, taps = g.kd * taps_hw;
    const int kz = tap / taps_hw, kr = tap - kz * taps_hw, ky = kr / g.kw, kx = kr - ky * g.kw;
    const int tid = threadIdx.y * 128 + threadIdx.x;
    for (int v = tid; v < V; v += 128 * kTinyWgSlices) {
        int r = v;
        const int xo = r % g.OW; r /= g.OW;
        const int yo = r % g.OH; r /= g.OH;
        const int zo = r % g.OD, n = r / g.OD;
        int sz, sy, sx;
        float val = 0.f;
        if (src_coord(zo, kz, g.sd, g.pd, g.dd, g.ID, 0, &sz) && src_coord(yo, ky, g.sh, g.ph, g.dh, g.IH, 0, &sy) &&
            src_coord(xo, kx, g.sw, g.pw, g.dw, g.IW, 0, &sx))
            val = to_f<TX>(x[((((int64_t)n * g.ID + sz) * g.IH + sy) * g.IW + sx) * g.IC + ic]);
        xv[v] = val;
    }
    __syncthreads();
    const int co = blockIdx.y * 128 + threadIdx.x;
    float acc = 0.f, accb = 0.f;
    if (co < g.OC) {
        const TG* gp = gy + co;
#pragma unroll 4
        for (int v = threadIdx.y; v < V; v += kTinyWgSlices) {
            const float gv = to_f<TG>(gp[(int64_t)v * g.OC]);
            acc = fmaf(xv[v], gv, acc);
            accb += gv;
        }
    }
    red[0][threadIdx.y][threadIdx.x] = acc;
    red[1][threadIdx.y][threadIdx.x] = accb;
    __syncthreads();
    if (threadIdx.y != 0 || co >= g.OC) return;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int q = 0; q < kTinyWgSlices; ++q) { a += red[0][q][threadIdx.x]; b += red[1][q][threadIdx.x]; }
    dw[((int64_t)co * g.IC + ic) * taps + tap] = a;
    if (dbias != nullptr && blockIdx.x == 0) dbias[co] = b;
}

inline bool conv_tiny_supported(const b200_conv_desc* d, int pass) {
    if (d->transposed) return false;
    const int64_t Vy = (int64_t)d->N * d->Do * d->Ho * d->Wo, Vx = (int64_t)d->N * d->Di * d->Hi * d->Wi;
    const int64_t V = pass == B200_PASS_DGRAD ? Vx : Vy;        // voxels of the produced tensor (fwd: y, dgrad: x) / of dy (wgrad)
    const int OC = pass == B200_PASS_DGRAD ? d->Ci : d->Co, IC = pass == B200_PASS_DGRAD ? d->Co : d->Ci;
    if (V >= kTinyMaxVox || Vx >= 8 * kTinyMaxVox || Vy >= 8 * kTinyMaxVox) return false;
    if (pass == B200_PASS_WGRAD && V > 4096) return false;      // the per-block x table lives in shared memory
    if (OC < 32 || IC > 2048 || (V >= 1024 && IC > 512)) return false;                     // coalescing needs a few warps of output channels; xs fits shared memory
    return true;
}

}  // namespace b200
