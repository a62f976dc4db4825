// Generic implicit-GEMM convolution on the fp32 FMA pipe (no tensor cores): every kernel shape, stride,
// dilation and padding the reference uses, fp32 or bf16 activations, fp32 accumulation.
// This is (a) the fp32 "tf32-off" product path (tolerance 1e-4), (b) the path for shapes that are not
// GEMM-shaped enough for tcgen05 (Cin=1 stems, Cout=2 heads, the HBM-bound separable (k,1,1) convs).
//
// GEMM view: OUT[v][co] = sum_K IN[src(v, tap)][ci] * W[K][co] with the flattened K = tap*IC + ci, so
// small-IC layers still fill 16-wide K chunks.  W is packed fp32 [K][OCp] (OCp = OC rounded up to 4).
//   forward geometry   : src = v*stride - pad + tap*dil
//   transposed geometry: src = (v + pad - tap*dil)/stride when divisible (dgrad of a conv, fwd of a ConvTranspose)
#pragma once
#include "common.cuh"

namespace b200 {

struct GatherGeom {
    int N;
    int IC, ID, IH, IW;      // tensor that is gathered from
    int OC, OD, OH, OW;      // tensor that is produced (fwd/dgrad) or the second operand (wgrad)
    int kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw;
    int transposed;
    int OCp;                 // padded OC of the packed weight
};

__device__ __forceinline__ bool src_coord(int o, int k, int s, int p, int dil, int in_size, bool transposed, int* out) {
    if (!transposed) {
        const int i = o * s - p + k * dil;
        *out = i;
        return i >= 0 && i < in_size;
    }
    const int t = o + p - k * dil;
    if (t < 0 || t % s != 0) return false;
    const int i = t / s;
    *out = i;
    return i < in_size;
}

template <typename TI>
__device__ __forceinline__ void load4(const TI* p, bool vec, int valid, float (&o)[4]) {
    // valid: number of in-range channels (0..4)
    if (vec && valid == 4) {
        if constexpr (sizeof(TI) == 4) {
            const float4 v = *reinterpret_cast<const float4*>(p);
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        } else {
            const uint2 v = *reinterpret_cast<const uint2*>(p);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
            const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
            o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = j < valid ? to_f<TI>(p[j]) : 0.f;
    }
}

// --------------------------------------------------------------------------- fwd / dgrad
// block 256 threads; tile TM output voxels x TN output channels; K chunks of 16.
template <typename TI, typename TO, int TM, int TN>
__global__ void __launch_bounds__(256) conv_gather_kernel(GatherGeom g, const TI* __restrict__ in, const float* __restrict__ w,
                                                          const float* __restrict__ bias, TO* __restrict__ out) {
    constexpr int TK = 16;
    constexpr int CG = TN / 4;                 // column groups (4 channels each)
    constexpr int RG = 256 / CG;               // row groups
    constexpr int RM = TM / RG;                // rows per thread
    static_assert(TM % RG == 0, "tile");
    __shared__ float As[TK][TM + 4];
    __shared__ __align__(16) float Bs[TK][TN];
    __shared__ int4 meta[TM];

    const int t = threadIdx.x;
    const int64_t V = (int64_t)g.N * g.OD * g.OH * g.OW;
    const int64_t v0 = (int64_t)blockIdx.x * TM;
    const int co0 = blockIdx.y * TN;
    for (int r = t; r < TM; r += 256) {
        int64_t v = v0 + r;
        int4 m = make_int4(-1, 0, 0, 0);
        if (v < V) {
            m.w = (int)(v % g.OW); v /= g.OW;
            m.z = (int)(v % g.OH); v /= g.OH;
            m.y = (int)(v % g.OD);
            m.x = (int)(v / g.OD);
        }
        meta[r] = m;
    }
    const int taps_hw = g.kh * g.kw;
    const int K = g.kd * taps_hw * g.IC;
    const bool vec_in = (g.IC % 4) == 0;
    const int tx = t % CG, ty = t / CG;
    float acc[RM][4];
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
    __syncthreads();

    for (int k0 = 0; k0 < K; k0 += TK) {
        // ---- A tile: TM rows x 16 k (4 quads)
        for (int e = t; e < TM * 4; e += 256) {
            const int row = e >> 2, kq = (e & 3) * 4;
            const int4 m = meta[row];
            float vals[4] = {0.f, 0.f, 0.f, 0.f};
            if (m.x >= 0) {
                if (vec_in) {
                    const int k = k0 + kq;
                    if (k < K) {
                        const int tap = k / g.IC, ci = k - tap * g.IC;
                        const int kz = tap / taps_hw, kr = tap - kz * taps_hw, ky = kr / g.kw, kx = kr - ky * g.kw;
                        int z, y, x;
                        if (src_coord(m.y, kz, g.sd, g.pd, g.dd, g.ID, g.transposed, &z) &&
                            src_coord(m.z, ky, g.sh, g.ph, g.dh, g.IH, g.transposed, &y) &&
                            src_coord(m.w, kx, g.sw, g.pw, g.dw, g.IW, g.transposed, &x))
                            load4<TI>(in + ((((int64_t)m.x * g.ID + z) * g.IH + y) * g.IW + x) * g.IC + ci, true, 4, vals);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = k0 + kq + j;
                        if (k < K) {
                            const int tap = k / g.IC, ci = k - tap * g.IC;
                            const int kz = tap / taps_hw, kr = tap - kz * taps_hw, ky = kr / g.kw, kx = kr - ky * g.kw;
                            int z, y, x;
                            if (src_coord(m.y, kz, g.sd, g.pd, g.dd, g.ID, g.transposed, &z) &&
                                src_coord(m.z, ky, g.sh, g.ph, g.dh, g.IH, g.transposed, &y) &&
                                src_coord(m.w, kx, g.sw, g.pw, g.dw, g.IW, g.transposed, &x))
                                vals[j] = to_f<TI>(in[((((int64_t)m.x * g.ID + z) * g.IH + y) * g.IW + x) * g.IC + ci]);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) As[kq + j][row] = vals[j];
        }
        // ---- B tile: 16 k x TN channels
        for (int e = t; e < TK * CG; e += 256) {
            const int kk = e / CG, cq = (e - kk * CG) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + kk < K && co0 + cq < g.OCp) v = *reinterpret_cast<const float4*>(w + (int64_t)(k0 + kk) * g.OCp + co0 + cq);
            *reinterpret_cast<float4*>(&Bs[kk][cq]) = v;
        }
        __syncthreads();
        const int kmax = min(TK, K - k0);
        for (int kk = 0; kk < kmax; ++kk) {
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                const float a = As[kk][ty + r * RG];
                acc[r][0] = fmaf(a, b.x, acc[r][0]);
                acc[r][1] = fmaf(a, b.y, acc[r][1]);
                acc[r][2] = fmaf(a, b.z, acc[r][2]);
                acc[r][3] = fmaf(a, b.w, acc[r][3]);
            }
        }
        __syncthreads();
    }
    // ---- epilogue
    const int co = co0 + tx * 4;
    if (co >= g.OC) return;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias != nullptr)
#pragma unroll
        for (int j = 0; j < 4; ++j) if (co + j < g.OC) bv[j] = bias[co + j];
    const bool vec_out = (g.OC % 4) == 0;
#pragma unroll
    for (int r = 0; r < RM; ++r) {
        const int64_t v = v0 + ty + r * RG;
        if (v >= V) continue;
        TO* o = out + v * g.OC + co;
        if (vec_out) {
            if constexpr (sizeof(TO) == 4) {
                *reinterpret_cast<float4*>(o) = make_float4(acc[r][0] + bv[0], acc[r][1] + bv[1], acc[r][2] + bv[2], acc[r][3] + bv[3]);
            } else {
                uint2 pk;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
                h[0] = __floats2bfloat162_rn(acc[r][0] + bv[0], acc[r][1] + bv[1]);
                h[1] = __floats2bfloat162_rn(acc[r][2] + bv[2], acc[r][3] + bv[3]);
                *reinterpret_cast<uint2*>(o) = pk;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (co + j < g.OC) o[j] = from_f<TO>(acc[r][j] + bv[j]);
        }
    }
}

// --------------------------------------------------------------------------- weight packing
// src: PyTorch parameter.  conv (transposed=0): (Co, Ci, taps);  convT (transposed=1): (Ci, Co, taps)
// dst: [K = tap*IC + ic][OCp] where (IC, OC) are the gathered/produced channel counts of the pass.
//   pass_swaps = 0: IC = channels of x (Ci), OC = channels of y (Co)        -- conv fwd, convT "fwd" (transposed gather)
//   pass_swaps = 1: IC = Co, OC = Ci                                        -- dgrad of either
__global__ void pack_weights_simt_kernel(int Ci, int Co, int taps, int param_is_ci_major, int pass_swaps, int OCp,
                                         const float* __restrict__ w, float* __restrict__ packed) {
    const int IC = pass_swaps ? Co : Ci, OC = pass_swaps ? Ci : Co;
    const int64_t total = (int64_t)taps * IC * OCp;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int oc = (int)(e % OCp);
        const int64_t k = e / OCp;
        const int ic = (int)(k % IC), tap = (int)(k / IC);
        float v = 0.f;
        if (oc < OC) {
            const int ci = pass_swaps ? oc : ic, co = pass_swaps ? ic : oc;
            v = param_is_ci_major ? w[((int64_t)ci * Co + co) * taps + tap] : w[((int64_t)co * Ci + ci) * taps + tap];
        }
        packed[e] = v;
    }
}

// --------------------------------------------------------------------------- wgrad
// partial[split][K][OCp] = sum over the split's voxels v of IN[src(v,tap)][ic] * G[v][oc]
// v ranges over the (OD,OH,OW) grid (the tensor NOT gathered from); geometry as in conv_gather_kernel.
template <typename TI, typename TG, int TN>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(GatherGeom g, const TI* __restrict__ in, const TG* __restrict__ grad,
                                                         int64_t vox_per_split, float* __restrict__ partial) {
    constexpr int TM = 64, TV = 16;
    constexpr int CG = TN / 4, RG = 256 / CG, RM = TM / RG;
    static_assert(TM % RG == 0 && RM >= 1, "tile");
    __shared__ float As[TV][TM + 4];           // [voxel][k]
    __shared__ __align__(16) float Bs[TV][TN]; // [voxel][oc]
    __shared__ int4 kmeta[TM];                 // (kz,ky,kx,ic) per k row, x=-1 when k >= K
    __shared__ int4 vmeta[TV];                 // (n,z,y,x) per voxel of the chunk, x=-1 past the end

    const int t = threadIdx.x;
    const int taps_hw = g.kh * g.kw;
    const int K = g.kd * taps_hw * g.IC;
    const int k0 = blockIdx.x * TM, co0 = blockIdx.y * TN, split = blockIdx.z;
    const int64_t V = (int64_t)g.N * g.OD * g.OH * g.OW;
    const int64_t v_begin = (int64_t)split * vox_per_split, v_end = min(V, v_begin + vox_per_split);
    for (int r = t; r < TM; r += 256) {
        const int k = k0 + r;
        int4 m = make_int4(-1, 0, 0, 0);
        if (k < K) {
            const int tap = k / g.IC;
            m.w = k - tap * g.IC;
            m.x = tap / taps_hw;
            const int kr = tap - m.x * taps_hw;
            m.y = kr / g.kw; m.z = kr - m.y * g.kw;
        }
        kmeta[r] = m;
    }
    const int tx = t % CG, ty = t / CG;
    float acc[RM][4];
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
    const bool vec_g = (g.OC % 4) == 0;
    __syncthreads();

    for (int64_t vb = v_begin; vb < v_end; vb += TV) {
        if (t < TV) {
            int64_t q = vb + t;
            int4 vm = make_int4(-1, 0, 0, 0);
            if (q < v_end) {
                vm.w = (int)(q % g.OW); q /= g.OW;
                vm.z = (int)(q % g.OH); q /= g.OH;
                vm.y = (int)(q % g.OD);
                vm.x = (int)(q / g.OD);
            }
            vmeta[t] = vm;
        }
        __syncthreads();
        // A: TV voxels x TM k; consecutive threads -> consecutive k (contiguous ic)
        for (int e = t; e < TV * TM; e += 256) {
            const int vv = e / TM, r = e - vv * TM;
            const int4 m = kmeta[r], vm = vmeta[vv];
            float val = 0.f;
            if (vm.x >= 0 && m.x >= 0) {
                int z, y, x;
                if (src_coord(vm.y, m.x, g.sd, g.pd, g.dd, g.ID, g.transposed, &z) &&
                    src_coord(vm.z, m.y, g.sh, g.ph, g.dh, g.IH, g.transposed, &y) &&
                    src_coord(vm.w, m.z, g.sw, g.pw, g.dw, g.IW, g.transposed, &x))
                    val = to_f<TI>(in[((((int64_t)vm.x * g.ID + z) * g.IH + y) * g.IW + x) * g.IC + m.w]);
            }
            As[vv][r] = val;
        }
        for (int e = t; e < TV * CG; e += 256) {
            const int vv = e / CG, cq = (e - vv * CG) * 4;
            const int64_t v = vb + vv;
            float vals[4] = {0.f, 0.f, 0.f, 0.f};
            if (v < v_end && co0 + cq < g.OC) load4<TG>(grad + v * g.OC + co0 + cq, vec_g, min(4, g.OC - co0 - cq), vals);
            *reinterpret_cast<float4*>(&Bs[vv][cq]) = make_float4(vals[0], vals[1], vals[2], vals[3]);
        }
        __syncthreads();
#pragma unroll
        for (int vv = 0; vv < TV; ++vv) {
            const float4 b = *reinterpret_cast<const float4*>(&Bs[vv][tx * 4]);
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                const float a = As[vv][ty + r * RG];
                acc[r][0] = fmaf(a, b.x, acc[r][0]);
                acc[r][1] = fmaf(a, b.y, acc[r][1]);
                acc[r][2] = fmaf(a, b.z, acc[r][2]);
                acc[r][3] = fmaf(a, b.w, acc[r][3]);
            }
        }
        __syncthreads();
    }
    const int co = co0 + tx * 4;
    if (co >= g.OCp) return;
#pragma unroll
    for (int r = 0; r < RM; ++r) {
        const int k = k0 + ty + r * RG;
        if (k < K) *reinterpret_cast<float4*>(partial + ((int64_t)split * K + k) * g.OCp + co) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
}

// dw(param layout) = sum over splits of partial[split][tap*IC+ic][oc]; one WARP per element, fixed order -> deterministic
//   gathered_is_ci: 1 when the gathered tensor carries the parameter's Ci channels (conv: x), 0 when Co (convT: dy)
__global__ void __launch_bounds__(256) conv_wgrad_reduce_kernel(int splits, int taps, int IC, int OC, int OCp, int Ci, int Co, int param_is_ci_major,
                                                                int gathered_is_ci, const float* __restrict__ partial, float* __restrict__ dw) {
    const int64_t K = (int64_t)taps * IC, total = K * OC;
    const int lane = threadIdx.x & 31;
    for (int64_t e = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); e < total; e += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int oc = (int)(e % OC);
        const int64_t k = e / OC;
        const int ic = (int)(k % IC), tap = (int)(k / IC);
        float acc = 0.f;
        for (int s = lane; s < splits; s += 32) acc += partial[((int64_t)s * K + k) * OCp + oc];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const int ci = gathered_is_ci ? ic : oc, co = gathered_is_ci ? oc : ic;
            const int64_t dst = param_is_ci_major ? ((int64_t)ci * Co + co) * taps + tap : ((int64_t)co * Ci + ci) * taps + tap;
            dw[dst] = acc;
        }
    }
}

// column sums (bias gradient): partial[chunk][C] then fixed-order sum
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ x, int C, int64_t R, int64_t rows_per_chunk, float* __restrict__ partial) {
    extern __shared__ float sm[];             // [rpi][C]
    const int rpi = C <= 256 ? 256 / C : 1;
    const int t = threadIdx.x;
    const int64_t r_end = min(R, (int64_t)(blockIdx.x + 1) * rows_per_chunk);
    for (int cbase = 0; cbase < C; cbase += 256) {
        const int c = cbase + (C <= 256 ? t % C : t), rr = C <= 256 ? t / C : 0;
        float acc = 0.f;
        if (c < C && rr < rpi)
            for (int64_t r = (int64_t)blockIdx.x * rows_per_chunk + rr; r < r_end; r += rpi) acc += to_f<T>(x[r * C + c]);
        if (C <= 256 && c < C && rr < rpi) sm[rr * C + c] = acc;      // wider tensors: one thread per channel, no shared staging
        __syncthreads();
        if (C <= 256) {
            if (t < C) { float s = 0.f; for (int j = 0; j < rpi; ++j) s += sm[j * C + t]; partial[(int64_t)blockIdx.x * C + t] = s; }
        } else if (c < C) partial[(int64_t)blockIdx.x * C + c] = acc;
        __syncthreads();
    }
}
// vectorised variant for channel counts that are multiples of the 16-byte vector: block = (256 / CV) rows x CV channel vectors,
// four rows in flight per thread, shared-memory reduction over the row groups; partial[chunk][C]
template <typename T, int V>
__global__ void __launch_bounds__(256) colsum_vec_partial_kernel(const T* __restrict__ x, int C, int64_t R, int64_t rows_per_chunk, float* __restrict__ partial) {
    extern __shared__ float sm[];             // [rpi][C]
    const int CV = C / V, rpi = 256 / CV;
    const int t = threadIdx.x, cv = t % CV, rr = t / CV;
    float s[V];
#pragma unroll
    for (int k = 0; k < V; ++k) s[k] = 0.f;
    if (rr < rpi) {
        const int64_t r_end = min(R, (int64_t)(blockIdx.x + 1) * rows_per_chunk);
        int64_t r = (int64_t)blockIdx.x * rows_per_chunk + rr;
        for (; r + 3 * rpi < r_end; r += 4 * rpi) {
            float a[4][V];
#pragma unroll
            for (int u = 0; u < 4; ++u) Pack<T, V>::load(x + (r + u * rpi) * C + cv * V, a[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < V; ++k) s[k] += a[u][k];
        }
        for (; r < r_end; r += rpi) {
            float a[V];
            Pack<T, V>::load(x + r * C + cv * V, a);
#pragma unroll
            for (int k = 0; k < V; ++k) s[k] += a[k];
        }
#pragma unroll
        for (int k = 0; k < V; ++k) sm[rr * C + cv * V + k] = s[k];
    }
    __syncthreads();
    for (int c = t; c < C; c += 256) {
        float acc = 0.f;
        for (int j = 0; j < rpi; ++j) acc += sm[j * C + c];
        partial[(int64_t)blockIdx.x * C + c] = acc;
    }
}

// dbias[c] = sum over rows of dy[r][c]: picks the vectorised kernel when it applies; bpart holds chunks*C floats
template <typename T>
inline int colsum_launch(const T* dy, int C, int64_t R, int chunks, int64_t rows_per_chunk, float* bpart, void* stream) {
    constexpr int V = 16 / sizeof(T);
    if (C % V == 0 && C / V <= 256 && aligned16(dy)) {
        const size_t smem = (size_t)(256 / (C / V)) * C * sizeof(float);
        if (smem <= 48 * 1024) {
            B200_LAUNCH((colsum_vec_partial_kernel<T, V>), chunks, 256, smem, stream, dy, C, R, rows_per_chunk, bpart);
            return 0;
        }
    }
    const size_t smem = (size_t)(C <= 256 ? (256 / C) * C : 1) * sizeof(float);
    B200_LAUNCH(colsum_partial_kernel<T>, chunks, 256, smem, stream, dy, C, R, rows_per_chunk, bpart);
    return 0;
}

__global__ void colsum_final_kernel(int chunks, int C, const float* __restrict__ partial, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0;
    for (int k = 0; k < chunks; ++k) s += (double)partial[(int64_t)k * C + c];
    out[c] = (float)s;
}

}  // namespace b200
