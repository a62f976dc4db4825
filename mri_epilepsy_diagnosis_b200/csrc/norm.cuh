// BatchNorm / InstanceNorm / GroupNorm on channels-last tensors: statistics, apply(+act,+residual), backward.
// HBM-bound: every pass reads each element once with 16-byte vectors; per-channel partial sums are
// reduced in shared memory per block, then across blocks in a fixed order (deterministic), in double.
#pragma once
#include "common.cuh"

namespace b200 {

struct NormGeom {
    int N, C, kind, G;
    int64_t S;
    int NB;            // independent reduce-batches: 1 for BATCH, N otherwise
    int64_t R;         // rows (voxels) per reduce-batch
    int V;             // channel vector width actually used
    int CV;            // channel vectors per voxel = C / V
    int rpi;           // rows per block iteration = 256 / CV
    int chunks;        // blocks per reduce-batch
    int64_t rows_per_chunk;
    int groups;        // number of (mean, rstd) entries
};

inline bool norm_geom(const b200_norm_desc* d, bool vec_ok, NormGeom* g) {
    g->N = d->N; g->C = d->C; g->S = d->S; g->kind = d->kind; g->G = d->G;
    g->NB = d->kind == B200_NORM_BATCH ? 1 : d->N;
    g->R = d->kind == B200_NORM_BATCH ? (int64_t)d->N * d->S : d->S;
    int vfull = d->dtype == B200_F32 ? 4 : 8;
    g->V = (vec_ok && d->C % vfull == 0) ? vfull : 1;
    g->CV = d->C / g->V;
    if (g->CV > 256 || g->CV < 1) return false;
    g->rpi = 256 / g->CV;
    int64_t want = ceil_div(g->R, (int64_t)g->rpi * 16);
    int64_t cap = (kNumSMs * 4) / g->NB;
    if (cap < 1) cap = 1;
    g->chunks = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    g->rows_per_chunk = ceil_div(g->R, g->chunks);
    g->groups = d->kind == B200_NORM_BATCH ? d->C : d->kind == B200_NORM_INSTANCE ? d->N * d->C : d->N * d->G;
    return true;
}

__device__ __forceinline__ int norm_group_index(int kind, int n, int c, int C, int G) {
    return kind == B200_NORM_BATCH ? c : kind == B200_NORM_INSTANCE ? n * C + c : n * G + c / (C / G);
}

// partial[((nb*chunks + chunk)*2 + {0,1})*C + c] = sum(x-K), sum((x-K)^2) with K = first row of the reduce-batch
template <typename T, int V>
__global__ void __launch_bounds__(256) norm_stats_partial_kernel(const T* __restrict__ x, int C, int64_t R, int64_t rows_per_chunk,
                                                                 float* __restrict__ partial) {
    extern __shared__ float sm[];                 // [2][rpi][C]
    const int CV = C / V, rpi = 256 / CV;
    const int t = threadIdx.x, cv = t % CV, rr = t / CV;
    const bool active = rr < rpi;
    const int nb = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
    const T* base = x + (int64_t)nb * R * C;
    float K[V], s[V], q[V];
#pragma unroll
    for (int k = 0; k < V; ++k) s[k] = q[k] = 0.f;
    if (active) {
        Pack<T, V>::load(base + cv * V, K);
        const int64_t r_end = min(R, (int64_t)(chunk + 1) * rows_per_chunk);
        int64_t r = (int64_t)chunk * rows_per_chunk + rr;
        for (; r + 3 * rpi < r_end; r += 4 * rpi) {
            float a[4][V];
#pragma unroll
            for (int u = 0; u < 4; ++u) Pack<T, V>::load(base + (r + u * rpi) * C + cv * V, a[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < V; ++k) { float dlt = a[u][k] - K[k]; s[k] += dlt; q[k] += dlt * dlt; }
        }
        for (; r < r_end; r += rpi) {
            float a[V];
            Pack<T, V>::load(base + r * C + cv * V, a);
#pragma unroll
            for (int k = 0; k < V; ++k) { float dlt = a[k] - K[k]; s[k] += dlt; q[k] += dlt * dlt; }
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
            sm[rr * C + cv * V + k] = s[k];
            sm[(rpi + rr) * C + cv * V + k] = q[k];
        }
    }
    __syncthreads();
    for (int c = t; c < 2 * C; c += 256) {
        const int which = c / C, ch = c - which * C;
        float acc = 0.f;
        for (int j = 0; j < rpi; ++j) acc += sm[(which * rpi + j) * C + ch];
        partial[(((int64_t)nb * chunks + chunk) * 2 + which) * C + ch] = acc;
    }
}

// fixed-order (deterministic) warp reduction in double
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// This lane's share of the per-block partials p[(k*2 + {0,1})*C] (k = lane, lane+32, ...), loaded 8 at a time so that the loads
// overlap (these finalize kernels are pure latency: one round trip to L2 per batch instead of one per element); fixed order.
__device__ __forceinline__ void lane_partial_sums(const float* __restrict__ p, int chunks, int C, int lane, double& s0, double& s1) {
    for (int k0 = lane; k0 < chunks; k0 += 32 * 8) {
        float v0[8], v1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + 32 * j;
            const bool ok = k < chunks;
            v0[j] = ok ? __ldg(p + ((int64_t)k * 2 + 0) * C) : 0.f;
            v1[j] = ok ? __ldg(p + ((int64_t)k * 2 + 1) * C) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { s0 += (double)v0[j]; s1 += (double)v1[j]; }
    }
}

// one WARP per (mean, rstd) entry: the lanes split the per-block partials, the combination order is fixed
template <typename T>
__global__ void __launch_bounds__(256) norm_stats_finalize_kernel(const T* __restrict__ x, int N, int C, int64_t S, int kind, int G, int chunks, int64_t R,
                                                                  const float* __restrict__ partial, float eps, float momentum,
                                                                  float* __restrict__ mean, float* __restrict__ rstd,
                                                                  float* __restrict__ running_mean, float* __restrict__ running_var) {
    const int groups = kind == B200_NORM_BATCH ? C : kind == B200_NORM_INSTANCE ? N * C : N * G;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= groups) return;
    int nb, c0, c1;
    if (kind == B200_NORM_BATCH) { nb = 0; c0 = g; c1 = g + 1; }
    else if (kind == B200_NORM_INSTANCE) { nb = g / C; c0 = g % C; c1 = c0 + 1; }
    else { nb = g / G; c0 = (g % G) * (C / G); c1 = c0 + C / G; }
    double m_acc = 0.0, e2_acc = 0.0;
    for (int c = c0; c < c1; ++c) {
        double s = 0.0, q = 0.0;
        lane_partial_sums(partial + (int64_t)nb * chunks * 2 * C + c, chunks, C, lane, s, q);
        s = warp_sum_d(s); q = warp_sum_d(q);
        const double K = x != nullptr ? (double)to_f<T>(x[(int64_t)nb * R * C + c]) : 0.0;     // null: unshifted partials (conv epilogue)
        const double ms = s / (double)R;
        const double mc = K + ms;
        double vc = q / (double)R - ms * ms;
        if (vc < 0.0) vc = 0.0;
        m_acc += mc;
        e2_acc += vc + mc * mc;
    }
    if (lane != 0) return;
    const int nc = c1 - c0;
    const double m = m_acc / nc;
    double var = e2_acc / nc - m * m;
    if (var < 0.0) var = 0.0;
    mean[g] = (float)m;
    rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
    if (kind == B200_NORM_BATCH && running_mean != nullptr) {
        const double unbiased = R > 1 ? var * (double)R / (double)(R - 1) : var;
        running_mean[g] = (float)((1.0 - momentum) * (double)running_mean[g] + momentum * m);
        running_var[g] = (float)((1.0 - momentum) * (double)running_var[g] + momentum * unbiased);
    }
}

__global__ void norm_from_running_kernel(int C, float eps, const float* __restrict__ rm, const float* __restrict__ rv,
                                         float* __restrict__ mean, float* __restrict__ rstd) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) { mean[c] = rm[c]; rstd[c] = 1.0f / sqrtf(rv[c] + eps); }
}

// the affine form of the normalisation, y_pre = x * sc + sh: ONE definition, used by the forward apply pass and by the backward
// passes that recompute the activation gate from x (bit-identical pre-activation values, hence identical gates)
__device__ __forceinline__ void norm_scale_shift(float mean, float rstd, float ga, float be, float& sc, float& sh) {
    sc = rstd * ga;
    sh = be - mean * sc;
}

// y = act((x - mean) * rstd * gamma + beta [+ residual]); grid (gx, N); gx*256 is a multiple of CV so each thread owns fixed channels.
// Dynamic shared memory: 2*C floats (scale, shift of this sample's channels, see lds_coef).
template <typename T, int V, int U = 2>
__global__ void __launch_bounds__(256, U == 2 ? 4 : 3) norm_apply_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const T* __restrict__ res, T* __restrict__ y,
                                                         int C, int64_t S, int kind, int G, int act, float slope) {
    extern __shared__ __align__(16) float cf[];
    using P = Pack<T, V>;
    const int CV = C / V, n = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += 256) {
        const int g = norm_group_index(kind, n, c, C, G);
        norm_scale_shift(mean[g], rstd[g], gamma ? gamma[c] : 1.f, beta ? beta[c] : 0.f, cf[c], cf[C + c]);
    }
    __syncthreads();
    const int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int cv = (int)(i0 % CV);
    const float* csc = cf + cv * V;
    const float* csh = cf + C + cv * V;
    const int64_t total = S * CV, stride = (int64_t)gridDim.x * 256, base = (int64_t)n * total;
    int64_t i = i0;
    for (; i + (U - 1) * stride < total; i += U * stride) {            // U independent 16-byte streams per tensor in flight
        typename P::raw xr[U], rr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            xr[u] = P::ldraw(x + (base + i + u * stride) * V);
            if (res != nullptr) rr[u] = P::ldraw(res + (base + i + u * stride) * V);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float v[V], r[V], sc[V], sh[V];
            P::unpack(xr[u], v);
            if (res != nullptr) P::unpack(rr[u], r);
            lds_coef<V>(csc, sc);
            lds_coef<V>(csh, sh);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const float t = fmaf(v[k], sc[k], sh[k]);
                v[k] = act_apply(res != nullptr ? t + r[k] : t, act, slope);
            }
            P::store(y + (base + i + u * stride) * V, v);
        }
    }
    for (; i < total; i += stride) {
        float v[V], sc[V], sh[V];
        P::load(x + (base + i) * V, v);
        lds_coef<V>(csc, sc);
        lds_coef<V>(csh, sh);
        if (res != nullptr) {
            float r[V];
            P::load(res + (base + i) * V, r);
#pragma unroll
            for (int k = 0; k < V; ++k) v[k] = act_apply(fmaf(v[k], sc[k], sh[k]) + r[k], act, slope);
        } else {
#pragma unroll
            for (int k = 0; k < V; ++k) v[k] = act_apply(fmaf(v[k], sc[k], sh[k]), act, slope);
        }
        P::store(y + (base + i) * V, v);
    }
}

// partial[((nb*chunks+chunk)*2+{0,1})*C + c] = sum dy', sum dy' * xhat, with dy' = dy * act'(y).
// The gate act'(.) comes from the saved output y when given; with y == nullptr (fused activation WITHOUT residual) it is
// recomputed from x through the same affine form as the forward pass -- one tensor read less.  Two rows are in flight per thread.
// U rows in flight per thread, loaded in packed form.  The per-channel constants stay in REGISTERS here (measured: re-reading four
// coefficient vectors per row from shared memory made this kernel slower, 4.1 -> 3.3 TB/s, while it helped the two apply kernels).
// Dynamic shared memory: 2*rpi*C floats (block reduction).
template <typename T, int V, int U = 2>
__global__ void __launch_bounds__(256, 2) norm_bwd_partial_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ dy,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  int C, int64_t R, int64_t rows_per_chunk, int kind, int G, int act, float slope,
                                                                  float* __restrict__ partial) {
    extern __shared__ __align__(16) float sm[];
    using P = Pack<T, V>;
    const int CV = C / V, rpi = 256 / CV;
    const int t = threadIdx.x, cv = t % CV, rr = t / CV;
    const bool active = rr < rpi;
    const int nb = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
    const int64_t off = (int64_t)nb * R * C;
    const bool gate_y = act != B200_ACT_NONE && y != nullptr, gate_x = act != B200_ACT_NONE && y == nullptr;
    float a[V], b[V], mu[V], rs[V], sc[V], sh[V];
#pragma unroll
    for (int k = 0; k < V; ++k) a[k] = b[k] = 0.f;
    if (active) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int c = cv * V + k;
            const int g = norm_group_index(kind, nb, c, C, G);
            mu[k] = mean[g]; rs[k] = rstd[g];
            norm_scale_shift(mu[k], rs[k], gamma ? gamma[c] : 1.f, beta ? beta[c] : 0.f, sc[k], sh[k]);
        }
        auto row = [&](const float (&xv)[V], const float (&gv)[V], const float (&ov)[V]) {
#pragma unroll
            for (int k = 0; k < V; ++k) {
                float g = gv[k];
                if (gate_y) g *= act_gate(ov[k], act, slope);
                else if (gate_x) g *= act_gate(fmaf(xv[k], sc[k], sh[k]), act, slope);
                a[k] += g; b[k] += g * (xv[k] - mu[k]) * rs[k];
            }
        };
        const int64_t r_end = min(R, (int64_t)(chunk + 1) * rows_per_chunk);
        int64_t r = (int64_t)chunk * rows_per_chunk + rr;
        for (; r + (U - 1) * rpi < r_end; r += U * rpi) {
            typename P::raw xr[U], gr[U], orr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t p = off + (r + u * rpi) * C + cv * V;
                xr[u] = P::ldraw(x + p);
                gr[u] = P::ldraw(dy + p);
                if (gate_y) orr[u] = P::ldraw(y + p);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float xv[V], gv[V], ov[V];
                P::unpack(xr[u], xv);
                P::unpack(gr[u], gv);
                if (gate_y) P::unpack(orr[u], ov);
                row(xv, gv, ov);
            }
        }
        for (; r < r_end; r += rpi) {
            float xv[V], gv[V], ov[V];
            const int64_t p = off + r * C + cv * V;
            P::load(x + p, xv);
            P::load(dy + p, gv);
            if (gate_y) P::load(y + p, ov);
            row(xv, gv, ov);
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
            sm[rr * C + cv * V + k] = a[k];
            sm[(rpi + rr) * C + cv * V + k] = b[k];
        }
    }
    __syncthreads();
    for (int c = t; c < 2 * C; c += 256) {
        const int which = c / C, ch = c - which * C;
        float acc = 0.f;
        for (int j = 0; j < rpi; ++j) acc += sm[(which * rpi + j) * C + ch];
        partial[(((int64_t)nb * chunks + chunk) * 2 + which) * C + ch] = acc;
    }
}

// AB[(nb*C + c)*2 + {0,1}] = sum over chunks; one WARP per (nb, c)
__global__ void __launch_bounds__(256) norm_bwd_sum_kernel(int NB, int C, int chunks, const float* __restrict__ partial, float* __restrict__ AB) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= NB * C) return;
    const int nb = i / C, c = i % C;
    double a = 0.0, b = 0.0;
    lane_partial_sums(partial + (int64_t)nb * chunks * 2 * C + c, chunks, C, lane, a, b);
    a = warp_sum_d(a); b = warp_sum_d(b);
    if (lane == 0) {
        AB[(int64_t)i * 2] = (float)a;
        AB[(int64_t)i * 2 + 1] = (float)b;
    }
}

// BatchNorm (one reduce-batch, no cross-channel terms): the sum over chunks and the coefficients in ONE kernel, one warp per channel
__global__ void __launch_bounds__(256) norm_bwd_sum_coef_bn_kernel(int N, int C, int64_t S, int chunks, int training, const float* __restrict__ partial,
                                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    double P = 0.0, Q = 0.0;
    lane_partial_sums(partial + c, chunks, C, lane, P, Q);
    P = warp_sum_d(P); Q = warp_sum_d(Q);
    if (lane != 0) return;
    P = (double)(float)P; Q = (double)(float)Q;              // the two-kernel path rounds the sums to fp32 (AB buffer): keep results identical
    const double ga = gamma ? (double)gamma[c] : 1.0, rs = rstd[c], mu = mean[c];
    double k1 = ga * rs, k4 = 0.0, k5 = 0.0;
    if (training) {
        const double M = (double)N * (double)S;
        k4 = -ga * rs * rs * Q / M;
        k5 = -ga * rs * P / M - k4 * mu;
    }
    coef[(int64_t)c * 5] = (float)k1; coef[(int64_t)c * 5 + 1] = (float)k4; coef[(int64_t)c * 5 + 2] = (float)k5;
    float gsc, gsh;
    norm_scale_shift(mean[c], rstd[c], gamma ? gamma[c] : 1.f, beta ? beta[c] : 0.f, gsc, gsh);
    coef[(int64_t)c * 5 + 3] = gsc; coef[(int64_t)c * 5 + 4] = gsh;
    if (dgamma) dgamma[c] = (float)Q;
    if (dbeta) dbeta[c] = (float)P;
}

// coef[(nb*C + c)*5 + {0..4}] : dx = k1*dy' + k4*x + k5 ; (sc, sh) of the forward affine form; also dgamma/dbeta (thread nb==0 sums over nb)
__global__ void norm_bwd_coef_kernel(int N, int C, int64_t S, int kind, int G, int training, int world, const float* __restrict__ AB,
                                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int NB = kind == B200_NORM_BATCH ? 1 : N;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NB * C) return;
    const int nb = i / C, c = i % C;
    const int g = norm_group_index(kind, nb, c, C, G);
    const double ga = gamma ? (double)gamma[c] : 1.0, rs = rstd[g], mu = mean[g];
    double k1 = ga * rs, k4 = 0.0, k5 = 0.0;
    if (training) {
        double P, Q, M;
        if (kind == B200_NORM_GROUP) {
            const int cg = C / G, c0 = (c / cg) * cg;
            P = Q = 0.0;
            for (int cc = c0; cc < c0 + cg; ++cc) {
                const double gg = gamma ? (double)gamma[cc] : 1.0;
                P += gg * (double)AB[((int64_t)nb * C + cc) * 2];
                Q += gg * (double)AB[((int64_t)nb * C + cc) * 2 + 1];
            }
            M = (double)cg * (double)S;
            k4 = -rs * rs * Q / M;
            k5 = -rs * P / M - k4 * mu;
        } else {
            M = kind == B200_NORM_BATCH ? (double)N * (double)S * (double)world : (double)S;   // world>1: sums were all-reduced (SyncBN)
            P = AB[(int64_t)i * 2]; Q = AB[(int64_t)i * 2 + 1];
            k4 = -ga * rs * rs * Q / M;
            k5 = -ga * rs * P / M - k4 * mu;
        }
    }
    coef[(int64_t)i * 5] = (float)k1; coef[(int64_t)i * 5 + 1] = (float)k4; coef[(int64_t)i * 5 + 2] = (float)k5;
    float gsc, gsh;
    norm_scale_shift(mean[g], rstd[g], gamma ? gamma[c] : 1.f, beta ? beta[c] : 0.f, gsc, gsh);
    coef[(int64_t)i * 5 + 3] = gsc; coef[(int64_t)i * 5 + 4] = gsh;
    if (nb == 0 && (dgamma != nullptr || dbeta != nullptr)) {
        double dg = 0.0, db = 0.0;
        for (int b = 0; b < NB; ++b) { db += (double)AB[((int64_t)b * C + c) * 2]; dg += (double)AB[((int64_t)b * C + c) * 2 + 1]; }
        if (dgamma) dgamma[c] = (float)dg;
        if (dbeta) dbeta[c] = (float)db;
    }
}

// gsc/gsh (optional, [NBg*C] each, per_sample-indexed like coef): affine form for recomputing the gate from x when y == nullptr
// Dynamic shared memory: 5*C floats (k1, k4, k5, scale, shift per channel, see lds_coef).
template <typename T, int V, int U = 2>
__global__ void __launch_bounds__(256, U == 2 ? 3 : 2) norm_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ dy,
                                                             const float* __restrict__ coef, T* __restrict__ dx, T* __restrict__ dres,
                                                             int C, int64_t S, int per_sample, int act, float slope) {
    extern __shared__ __align__(16) float cf[];
    using P = Pack<T, V>;
    const int CV = C / V, n = blockIdx.y;
    for (int e = threadIdx.x; e < 5 * C; e += 256) {
        const int c = e / 5, j = e - c * 5;
        cf[j * C + c] = coef[((int64_t)(per_sample ? n : 0) * C) * 5 + e];
    }
    __syncthreads();
    const int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int cv = (int)(i0 % CV);
    const float* ck = cf + cv * V;
    const bool gate_y = act != B200_ACT_NONE && y != nullptr, gate_x = act != B200_ACT_NONE && y == nullptr;
    constexpr int W = V % 4 == 0 ? 4 : V;                     // channels per pass: bounds the live coefficient registers
    auto row = [&](float (&xv)[V], float (&gv)[V], const float (&ov)[V], int64_t p) {
#pragma unroll
        for (int h = 0; h < V; h += W) {
            if (gate_y) {
#pragma unroll
                for (int k = h; k < h + W; ++k) gv[k] *= act_gate(ov[k], act, slope);
            } else if (gate_x) {
                float sc[W], sh[W];
                lds_coef<W>(ck + 3 * C + h, sc);
                lds_coef<W>(ck + 4 * C + h, sh);
#pragma unroll
                for (int k = 0; k < W; ++k) gv[h + k] *= act_gate(fmaf(xv[h + k], sc[k], sh[k]), act, slope);
            }
            float k1[W], k4[W], k5[W];
            lds_coef<W>(ck + h, k1);
            lds_coef<W>(ck + C + h, k4);
            lds_coef<W>(ck + 2 * C + h, k5);
#pragma unroll
            for (int k = 0; k < W; ++k) xv[h + k] = fmaf(k1[k], gv[h + k], fmaf(k4[k], xv[h + k], k5[k]));
        }
        if (dres != nullptr) P::store(dres + p, gv);
        P::store(dx + p, xv);
    };
    const int64_t total = S * CV, stride = (int64_t)gridDim.x * 256, base = (int64_t)n * total;
    int64_t i = i0;
    for (; i + (U - 1) * stride < total; i += U * stride) {
        typename P::raw xr[U], gr[U], orr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t p = (base + i + u * stride) * V;
            xr[u] = P::ldraw(x + p);
            gr[u] = P::ldraw(dy + p);
            if (gate_y) orr[u] = P::ldraw(y + p);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float xv[V], gv[V], ov[V];
            P::unpack(xr[u], xv);
            P::unpack(gr[u], gv);
            if (gate_y) P::unpack(orr[u], ov);
            row(xv, gv, ov, (base + i + u * stride) * V);
        }
    }
    for (; i < total; i += stride) {
        float xv[V], gv[V], ov[V];
        const int64_t p = (base + i) * V;
        P::load(x + p, xv);
        P::load(dy + p, gv);
        if (gate_y) P::load(y + p, ov);
        row(xv, gv, ov, p);
    }
}

}  // namespace b200

// ------------------------------------------------------------------------------------------------ SyncBN glue (data parallel)
// N ranks x local batch == one device x global batch: every rank reduces its own (mean, rstd), packs
// (mean, E[x^2]) = (mean, var + mean^2), ONE all-reduce(AVG) over the ranks (equal per-rank counts), and the finalize kernel turns
// the averages back into (mean, rstd) and performs the running-statistics update with the GLOBAL element count.
namespace b200 {

__global__ void syncbn_pack_kernel(int C, float eps, const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ packed) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float m = mean[c], r = rstd[c];
    const float var = 1.f / (r * r) - eps;
    packed[c] = m;
    packed[C + c] = var + m * m;
}

__global__ void syncbn_finalize_kernel(int C, float eps, float momentum, double count_global, const float* __restrict__ packed, float* __restrict__ mean,
                                       float* __restrict__ rstd, float* __restrict__ running_mean, float* __restrict__ running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float m = packed[c];
    float var = packed[C + c] - m * m;
    var = var > 0.f ? var : 0.f;
    mean[c] = m;
    rstd[c] = rsqrtf(var + eps);
    if (running_mean != nullptr) {
        const double unbiased = (double)var * (count_global / (count_global > 1.0 ? count_global - 1.0 : 1.0));
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

}  // namespace b200
