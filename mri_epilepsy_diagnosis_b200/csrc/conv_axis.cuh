// Direct kernels for the separable "1-D" convolutions of the autoencoder / fader family with FEW channels
// (classification/models/AE_model.py:9-26: kernel (k,1,1) / (1,k,1) / (1,1,k), k in {3,6}, stride 1 or 2 along that axis only):
// the Cin = 1 stems, the 8-channel layers of the shipped encoder_93_6_4 network, and the 16 -> 1 / 1 -> 1 reconstruction tails.
// These layers are ~370 FLOP per voxel and purely HBM-bound; the generic implicit-GEMM kernel spent its time on index
// arithmetic.  One thread = one produced voxel and all of its OC channels; the <= K taps along the axis are IC-wide vector loads.
//   gather  : forward (src = o*s - p + k) and dgrad / transposed (src = (o + p - k)/s when divisible) with the SIMT weight packing
//   wgrad   : one warp = 32 voxels x one tap k, IC x OC (+ bias) accumulators per lane, shuffle + shared + fixed-order final sums
#pragma once
#include "common.cuh"
#include "conv_simt.cuh"

namespace b200 {

struct AxisGeom {
    int64_t outer;        // product of the dims before the conv axis (n and the slower spatial dims)
    int in_len, out_len;  // extent of the conv axis in the gathered / produced tensor
    int64_t inner;        // product of the spatial dims after the conv axis (voxels, not elements)
    int K, s, p, transposed;
    int OCp;              // padded OC of the packed weight
};

template <typename T, int C> __device__ __forceinline__ void load_vec(const T* p, float (&o)[C]) {
    if constexpr (C * sizeof(T) % 16 == 0) {
        constexpr int V = 16 / sizeof(T);
#pragma unroll
        for (int q = 0; q < C / V; ++q) {
            float t[V];
            Pack<T, V>::load(p + q * V, t);
#pragma unroll
            for (int k = 0; k < V; ++k) o[q * V + k] = t[k];
        }
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = to_f<T>(p[c]);
    }
}
template <typename T, int C> __device__ __forceinline__ void store_vec(T* p, const float (&o)[C]) {
    if constexpr (C * sizeof(T) % 16 == 0) {
        constexpr int V = 16 / sizeof(T);
#pragma unroll
        for (int q = 0; q < C / V; ++q) {
            float t[V];
#pragma unroll
            for (int k = 0; k < V; ++k) t[k] = o[q * V + k];
            Pack<T, V>::store(p + q * V, t);
        }
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) p[c] = from_f<T>(o[c]);
    }
}

// out[outer][o][inner][OC] = bias + sum_k sum_ic in[outer][src(o,k)][inner][IC] * w[(k*IC + ic)*OCp + oc]
template <typename TI, typename TO, int IC, int OC>
__global__ void __launch_bounds__(256) axis_gather_kernel(AxisGeom g, const TI* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                          TO* __restrict__ out) {
    extern __shared__ float ws[];                         // [K][IC][OC]
    for (int i = threadIdx.x; i < g.K * IC * OC; i += 256) ws[i] = w[(int64_t)(i / OC) * g.OCp + i % OC];
    __syncthreads();
    // 32-bit index arithmetic (the host checks that the voxel count fits): the three 64-bit divisions per voxel this loop used to
    // do cost more instructions than the convolution itself (ncu, round 2: 1.18 TB/s on a pure streaming kernel)
    const unsigned total = (unsigned)(g.outer * g.out_len * g.inner), inner = (unsigned)g.inner, out_len = (unsigned)g.out_len;
    for (unsigned v = blockIdx.x * 256u + threadIdx.x; v < total; v += gridDim.x * 256u) {
        const unsigned in_ = v % inner;
        const unsigned r = v / inner;
        const int o = (int)(r % out_len);
        const int64_t ou = r / out_len;
        const TI* base = in + ((ou * g.in_len) * g.inner + in_) * IC;
        float acc[OC];
#pragma unroll
        for (int c = 0; c < OC; ++c) acc[c] = bias != nullptr ? __ldg(bias + c) : 0.f;
        for (int k = 0; k < g.K; ++k) {
            int i;
            if (!g.transposed) i = o * g.s - g.p + k;
            else {
                const int t = o + g.p - k;
                if (t < 0 || t % g.s != 0) continue;
                i = t / g.s;
            }
            if ((unsigned)i >= (unsigned)g.in_len) continue;
            float xv[IC];
            load_vec<TI, IC>(base + (int64_t)i * g.inner * IC, xv);
            const float* wk = ws + k * IC * OC;
#pragma unroll
            for (int ci = 0; ci < IC; ++ci) {
                if constexpr (OC % 4 == 0) {
#pragma unroll
                    for (int q = 0; q < OC / 4; ++q) {
                        const float4 f = *reinterpret_cast<const float4*>(wk + ci * OC + 4 * q);
                        acc[4 * q + 0] = fmaf(xv[ci], f.x, acc[4 * q + 0]);
                        acc[4 * q + 1] = fmaf(xv[ci], f.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(xv[ci], f.z, acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(xv[ci], f.w, acc[4 * q + 3]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < OC; ++c) acc[c] = fmaf(xv[ci], wk[ci * OC + c], acc[c]);
                }
            }
        }
        store_vec<TO, OC>(out + (int64_t)v * OC, acc);
    }
}

// partial[block][K][IC*OC + OC]: dw[k][ic][oc] = sum_v G[v][oc] * X[src(v,k)][ic]; the trailing OC entries of tap 0 hold sum_v G[v][oc]
template <typename TX, typename TG, int IC, int OC>
__global__ void __launch_bounds__(256) axis_wgrad_kernel(AxisGeom g, const TX* __restrict__ x, const TG* __restrict__ gy, float* __restrict__ partial) {
    constexpr int NA = IC * OC + OC;
    extern __shared__ float red[];                        // [8 warps][NA]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned gw = blockIdx.x * 8u + warp, nwarps = gridDim.x * 8u;                      // nwarps is a multiple of K
    const int k = (int)(gw % (unsigned)g.K);
    float acc[IC][OC], accb[OC];
#pragma unroll
    for (int c = 0; c < OC; ++c) {
        accb[c] = 0.f;
#pragma unroll
        for (int i = 0; i < IC; ++i) acc[i][c] = 0.f;
    }
    const unsigned total = (unsigned)(g.outer * g.out_len * g.inner);                          // voxels of gy (the produced tensor of the forward)
    const unsigned nchunks = (total + 31u) / 32u, inner = (unsigned)g.inner, out_len = (unsigned)g.out_len;
    const unsigned cstep = nwarps / (unsigned)g.K;
    for (unsigned chunk = gw / (unsigned)g.K; chunk < nchunks; chunk += cstep) {
        const unsigned v = chunk * 32u + lane;
        if (v >= total) continue;
        const unsigned in_ = v % inner;
        const unsigned r = v / inner;
        const int o = (int)(r % out_len);
        const int64_t ou = r / out_len;
        float gv[OC];
        load_vec<TG, OC>(gy + (int64_t)v * OC, gv);
        if (k == 0) {
#pragma unroll
            for (int c = 0; c < OC; ++c) accb[c] += gv[c];
        }
        const int i = o * g.s - g.p + k;
        if ((unsigned)i >= (unsigned)g.in_len) continue;
        float xv[IC];
        load_vec<TX, IC>(x + (((ou * g.in_len) + i) * g.inner + in_) * IC, xv);
#pragma unroll
        for (int ci = 0; ci < IC; ++ci)
#pragma unroll
            for (int c = 0; c < OC; ++c) acc[ci][c] = fmaf(xv[ci], gv[c], acc[ci][c]);
    }
#pragma unroll
    for (int e = 0; e < NA; ++e) {
        float v = e < IC * OC ? acc[e / OC][e % OC] : accb[e - IC * OC];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp * NA + e] = v;
    }
    __syncthreads();
    // warps of this block with the same tap (gw % K) are combined in warp order -> partial[block][k][NA]
    for (int e = threadIdx.x; e < g.K * NA; e += 256) {
        const int kk = e / NA, a = e - kk * NA;
        float s = 0.f;
        for (int wq = 0; wq < 8; ++wq)
            if ((int)(((int64_t)blockIdx.x * 8 + wq) % g.K) == kk) s += red[wq * NA + a];
        partial[((int64_t)blockIdx.x * g.K + kk) * NA + a] = s;
    }
}

// dw (PyTorch layout (Co, Ci, K)) and dbias from the block partials; one warp per element, fixed order
__global__ void __launch_bounds__(256) axis_wgrad_reduce_kernel(const float* __restrict__ partial, int blocks, int K, int IC, int OC, float* __restrict__ dw,
                                                                float* __restrict__ dbias) {
    const int NA = IC * OC + OC;
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (e >= K * NA) return;
    double s = 0.0;
    for (int b0 = lane; b0 < blocks; b0 += 32 * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int b = b0 + 32 * u; v[u] = b < blocks ? __ldg(partial + (int64_t)b * K * NA + e) : 0.f; }
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane != 0) return;
    const int k = e / NA, a = e - k * NA;
    if (a < IC * OC) {
        const int ci = a / OC, co = a % OC;
        dw[((int64_t)co * IC + ci) * K + k] = (float)s;
    } else if (k == 0 && dbias != nullptr) dbias[a - IC * OC] = (float)s;
}

// ------------------------------------------------------------------------------------------------ host side
// which spatial axis carries the kernel (0: D, 1: H, 2: W); -1 when the convolution is not a 1-D one
inline int axis_of(const b200_conv_desc* d) {
    const int k[3] = {d->kd, d->kh, d->kw}, s[3] = {d->sd, d->sh, d->sw}, p[3] = {d->pd, d->ph, d->pw}, dl[3] = {d->dd, d->dh, d->dw};
    int axis = -1;
    for (int a = 0; a < 3; ++a) {
        if (k[a] == 1) { if (s[a] != 1 || p[a] != 0) return -1; }
        else { if (axis >= 0 || dl[a] != 1) return -1; axis = a; }
    }
    return axis;
}
inline bool axis_ch_ok(int c) { return c == 1 || c == 8 || c == 16; }

inline bool axis_conv_supported(const b200_conv_desc* d, int pass) {
    if (d->transposed) return false;
    const int a = axis_of(d);
    if (a < 0) return false;
    const int K = a == 0 ? d->kd : (a == 1 ? d->kh : d->kw);
    if (K > 8 || !axis_ch_ok(d->Ci) || !axis_ch_ok(d->Co)) return false;
    if (pass == B200_PASS_WGRAD && d->Ci * d->Co > 128) return false;                // accumulators must stay in registers
    return true;
}

inline AxisGeom axis_geom(const b200_conv_desc* d, int pass) {
    AxisGeom g;
    const int a = axis_of(d);
    const int dims_in[3] = {d->Di, d->Hi, d->Wi}, dims_out[3] = {d->Do, d->Ho, d->Wo};
    const int K = a == 0 ? d->kd : (a == 1 ? d->kh : d->kw), s = a == 0 ? d->sd : (a == 1 ? d->sh : d->sw), p = a == 0 ? d->pd : (a == 1 ? d->ph : d->pw);
    // the non-conv dims are equal in x and y
    g.outer = d->N; g.inner = 1;
    for (int i = 0; i < a; ++i) g.outer *= dims_in[i];
    for (int i = a + 1; i < 3; ++i) g.inner *= dims_in[i];
    g.K = K; g.s = s; g.p = p;
    if (pass == B200_PASS_DGRAD) { g.in_len = dims_out[a]; g.out_len = dims_in[a]; g.transposed = 1; g.OCp = (d->Ci + 3) & ~3; }
    else { g.in_len = dims_in[a]; g.out_len = dims_out[a]; g.transposed = 0; g.OCp = (d->Co + 3) & ~3; }
    return g;
}

constexpr int kAxisBlocks = kNumSMs * 6;        // 8 warps per block: 8 * 148 * 6 warps, a multiple of every K <= 8 after rounding below
inline int axis_wgrad_blocks(int K) { return (kAxisBlocks / K) * K; }
inline size_t axis_wgrad_ws_bytes(const b200_conv_desc* d) {
    const int a = axis_of(d);
    const int K = a == 0 ? d->kd : (a == 1 ? d->kh : d->kw);
    return (size_t)axis_wgrad_blocks(K) * K * (d->Ci * d->Co + d->Co) * sizeof(float);
}

#define B200_AXIS_CH(ICV, OCV, ...)                                                                  \
    do {                                                                                             \
        const int key__ = (ICV) * 100 + (OCV);                                                       \
        switch (key__) {                                                                             \
            case 108: { constexpr int IC = 1, OC = 8; __VA_ARGS__; } break;                          \
            case 116: { constexpr int IC = 1, OC = 16; __VA_ARGS__; } break;                         \
            case 101: { constexpr int IC = 1, OC = 1; __VA_ARGS__; } break;                          \
            case 801: { constexpr int IC = 8, OC = 1; __VA_ARGS__; } break;                          \
            case 808: { constexpr int IC = 8, OC = 8; __VA_ARGS__; } break;                          \
            case 816: { constexpr int IC = 8, OC = 16; __VA_ARGS__; } break;                         \
            case 1601: { constexpr int IC = 16, OC = 1; __VA_ARGS__; } break;                        \
            case 1608: { constexpr int IC = 16, OC = 8; __VA_ARGS__; } break;                        \
            default: { constexpr int IC = 16, OC = 16; __VA_ARGS__; } break;                         \
        }                                                                                            \
    } while (0)

// in/out dtypes of the gather: (f32 -> bf16), (bf16 -> bf16), (bf16 -> f32), (f32 -> f32 only reached by dgrad of a mixed conv)
#define B200_AXIS_DT(in_dt, out_dt, TI, TO, ...)                                                     \
    do {                                                                                             \
        if ((in_dt) == B200_F32 && (out_dt) == B200_BF16) { using TI = float; using TO = __nv_bfloat16; __VA_ARGS__; }           \
        else if ((in_dt) == B200_BF16 && (out_dt) == B200_BF16) { using TI = __nv_bfloat16; using TO = __nv_bfloat16; __VA_ARGS__; } \
        else if ((in_dt) == B200_BF16 && (out_dt) == B200_F32) { using TI = __nv_bfloat16; using TO = float; __VA_ARGS__; }       \
        else { using TI = float; using TO = float; __VA_ARGS__; }                                    \
    } while (0)

inline int axis_gather_run(const b200_conv_desc* d, int pass, const void* in, const float* w, const float* bias, void* out, void* stream) {
    const AxisGeom g = axis_geom(d, pass);
    const int ICv = pass == B200_PASS_DGRAD ? d->Co : d->Ci, OCv = pass == B200_PASS_DGRAD ? d->Ci : d->Co;
    const int in_dt = pass == B200_PASS_DGRAD ? d->y_dtype : d->x_dtype, out_dt = pass == B200_PASS_DGRAD ? d->x_dtype : d->y_dtype;
    const int64_t total = g.outer * g.out_len * g.inner;
    B200_REQUIRE(total < ((int64_t)1 << 31) - 65536 * 256, "axis conv: more than 2^31 voxels");
    const int grid = (int)(ceil_div(total, 256) < (int64_t)kNumSMs * 16 ? ceil_div(total, 256) : (int64_t)kNumSMs * 16);
    const size_t smem = (size_t)g.K * ICv * OCv * sizeof(float);
    B200_REQUIRE(aligned16(in) && aligned16(out), "axis conv: pointers must be 16-byte aligned");
    B200_AXIS_CH(ICv, OCv, {
        B200_AXIS_DT(in_dt, out_dt, TI, TO, {
            B200_LAUNCH((axis_gather_kernel<TI, TO, IC, OC>), grid, 256, smem, stream, g, (const TI*)in, w, bias, (TO*)out);
        });
    });
    return 0;
}

inline int axis_wgrad_run(const b200_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, void* workspace, void* stream) {
    const AxisGeom g = axis_geom(d, B200_PASS_WGRAD);
    B200_REQUIRE(g.outer * g.out_len * g.inner < ((int64_t)1 << 31) - 64, "axis conv: more than 2^31 voxels");
    const int blocks = axis_wgrad_blocks(g.K);
    float* partial = (float*)workspace;
    const size_t smem = (size_t)8 * (d->Ci * d->Co + d->Co) * sizeof(float);
    B200_REQUIRE(aligned16(x) && aligned16(dy), "axis conv: pointers must be 16-byte aligned");
    B200_AXIS_CH(d->Ci, d->Co, {
        if constexpr (IC * OC <= 128) {
            B200_AXIS_DT(d->x_dtype, d->y_dtype, TX, TG, {
                B200_LAUNCH((axis_wgrad_kernel<TX, TG, IC, OC>), blocks, 256, smem, stream, g, (const TX*)x, (const TG*)dy, partial);
            });
        } else {
            return fail("axis wgrad: IC*OC too large");
        }
    });
    B200_LAUNCH(axis_wgrad_reduce_kernel, (int)ceil_div((int64_t)g.K * (d->Ci * d->Co + d->Co), 8), 256, 0, stream, partial, blocks, g.K, d->Ci, d->Co, dw,
                dbias);
    return 0;
}

}  // namespace b200
