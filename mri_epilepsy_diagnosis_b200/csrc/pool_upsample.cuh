// MaxPool3d (with argmax) and trilinear/nearest upsampling on channels-last tensors.  HBM-bound gathers:
// one thread per (voxel, 16-byte channel vector), consecutive threads on consecutive channel vectors.
#pragma once
#include "common.cuh"

namespace b200 {

// ------------------------------------------------------------------ max pool
// Tie/NaN rule of ATen's CPU max_pool3d: `if (val > max || isnan(val)) take` scanning (d,h,w) in raster order.
template <typename T, int V>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(b200_pool_desc d, const T* __restrict__ x, T* __restrict__ y,
                                                          uint8_t* __restrict__ code, int64_t* __restrict__ indices) {
    const int CV = d.C / V;
    const int64_t total = (int64_t)d.N * d.Do * d.Ho * d.Wo * CV;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = i / CV;
        const int cv = (int)(i - v * CV);
        const int xo = (int)(v % d.Wo); v /= d.Wo;
        const int yo = (int)(v % d.Ho); v /= d.Ho;
        const int zo = (int)(v % d.Do);
        const int n = (int)(v / d.Do);
        float best[V];
        int bcode[V];
#pragma unroll
        for (int k = 0; k < V; ++k) { best[k] = -INFINITY; bcode[k] = 0; }
        const T* xb = x + (int64_t)n * d.Di * d.Hi * d.Wi * d.C + cv * V;
        int local = 0;
        for (int a = 0; a < d.kd; ++a)
            for (int b = 0; b < d.kh; ++b)
                for (int c = 0; c < d.kw; ++c, ++local) {
                    const int zi = zo * d.sd + a, yi = yo * d.sh + b, xi = xo * d.sw + c;
                    float vals[V];
                    Pack<T, V>::load(xb + (((int64_t)zi * d.Hi + yi) * d.Wi + xi) * d.C, vals);
#pragma unroll
                    for (int k = 0; k < V; ++k)
                        if (vals[k] > best[k] || vals[k] != vals[k]) { best[k] = vals[k]; bcode[k] = local; }
                }
        const int64_t o = i * V;
        Pack<T, V>::store(y + o, best);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if (code) code[o + k] = (uint8_t)bcode[k];
            if (indices) {
                const int a = bcode[k] / (d.kh * d.kw), r = bcode[k] % (d.kh * d.kw);
                const int zi = zo * d.sd + a, yi = yo * d.sh + r / d.kw, xi = xo * d.sw + r % d.kw;
                indices[o + k] = ((int64_t)zi * d.Hi + yi) * d.Wi + xi;
            }
        }
    }
}

// gather form of the scatter: every input voxel sums dy of the windows that cover it and chose it (deterministic, writes all of dx)
template <typename T, int V>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(b200_pool_desc d, const T* __restrict__ dy, const uint8_t* __restrict__ code,
                                                          T* __restrict__ dx) {
    const int CV = d.C / V;
    const int64_t total = (int64_t)d.N * d.Di * d.Hi * d.Wi * CV;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = i / CV;
        const int cv = (int)(i - v * CV);
        const int xi = (int)(v % d.Wi); v /= d.Wi;
        const int yi = (int)(v % d.Hi); v /= d.Hi;
        const int zi = (int)(v % d.Di);
        const int n = (int)(v / d.Di);
        float acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = 0.f;
        const int z_lo = max(0, (zi - d.kd + d.sd) / d.sd), z_hi = min(d.Do - 1, zi / d.sd);
        const int y_lo = max(0, (yi - d.kh + d.sh) / d.sh), y_hi = min(d.Ho - 1, yi / d.sh);
        const int x_lo = max(0, (xi - d.kw + d.sw) / d.sw), x_hi = min(d.Wo - 1, xi / d.sw);
        for (int zo = z_lo; zo <= z_hi; ++zo)
            for (int yo = y_lo; yo <= y_hi; ++yo)
                for (int xo = x_lo; xo <= x_hi; ++xo) {
                    const int local = ((zi - zo * d.sd) * d.kh + (yi - yo * d.sh)) * d.kw + (xi - xo * d.sw);
                    const int64_t o = (((((int64_t)n * d.Do + zo) * d.Ho + yo) * d.Wo + xo) * CV + cv) * V;
                    float g[V];
                    Pack<T, V>::load(dy + o, g);
#pragma unroll
                    for (int k = 0; k < V; ++k) acc[k] += (code[o + k] == local) ? g[k] : 0.f;
                }
        Pack<T, V>::store(dx + i * V, acc);
    }
}

// kernel = stride = 2 (every pool the benchmarked models use): each input voxel belongs to exactly one window
template <typename T, int V>
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(b200_pool_desc d, const T* __restrict__ dy, const uint8_t* __restrict__ code,
                                                           T* __restrict__ dx, const T* __restrict__ add = nullptr, int add_ctot = 0) {
    const int CV = d.C / V;
    const int rows = d.N * d.Di * d.Hi, per_row = d.Wi * CV;
    const int tpr = per_row < 256 ? per_row : 256, rpb = 256 / tpr;          // narrow rows share a block (see the forward kernel)
    const int sub = threadIdx.x / tpr, lane_e = threadIdx.x - sub * tpr;
    if (sub >= rpb) return;
    for (int row = blockIdx.x * rpb + sub; row < rows; row += gridDim.x * rpb) {
        const int yi = row % d.Hi, zi = (row / d.Hi) % d.Di, n = row / (d.Hi * d.Di);
        const int zo = zi >> 1, yo = yi >> 1;
        const bool row_in = zo < d.Do && yo < d.Ho;                   // floor mode: a trailing odd plane/row is in no window
        const int64_t orow = (((int64_t)n * d.Do + zo) * d.Ho + yo) * d.Wo;
        const int lzy = ((zi & 1) * 2 + (yi & 1)) * 2;
        T* xr = dx + (int64_t)row * d.Wi * d.C;
        for (int e = lane_e; e < per_row; e += tpr) {
            const int xi = e / CV, cv = e - xi * CV;
            const int xo = xi >> 1;
            float acc[V];
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] = 0.f;
            if (row_in && xo < d.Wo) {
                const int64_t o = ((orow + xo) * CV + cv) * V;
                const int local = lzy + (xi & 1);
                float g[V];
                Pack<T, V>::load(dy + o, g);
#pragma unroll
                for (int k = 0; k < V; ++k) acc[k] = (code[o + k] == local) ? g[k] : 0.f;
            }
            if (add != nullptr) {
                // second consumer of the pooled tensor (U-Net skip connection): its gradient, a channel window of a wider
                // channels-last tensor (row pitch add_ctot), is summed here instead of by a separate add kernel
                float a[V];
                Pack<T, V>::load(add + ((int64_t)row * d.Wi + xi) * add_ctot + cv * V, a);
#pragma unroll
                for (int k = 0; k < V; ++k) acc[k] += a[k];
            }
            Pack<T, V>::store(xr + e * V, acc);
        }
    }
}

// ------------------------------------------------------------------ upsample
// Source-index rules of ATen (UpSample.h): nearest: min(floor(dst*scale), in-1); linear align_corners=False:
// max(scale*(dst+0.5)-0.5, 0); align_corners=True: dst*(in-1)/(out-1).
struct Lin1 { int i0, i1; float w0, w1; };
__device__ __forceinline__ Lin1 lin_src(int dst, int in, int out, int mode) {
    Lin1 r;
    if (mode == B200_UP_NEAREST) {
        const float scale = (float)in / (float)out;
        r.i0 = r.i1 = min((int)floorf(dst * scale), in - 1);
        r.w0 = 1.f; r.w1 = 0.f;
        return r;
    }
    float src;
    if (mode == B200_UP_TRILINEAR_ALIGNED) {
        const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
        src = scale * dst;
    } else {
        const float scale = (float)in / (float)out;
        src = scale * (dst + 0.5f) - 0.5f;
        if (src < 0.f) src = 0.f;
    }
    r.i0 = min((int)src, in - 1);
    r.i1 = min(r.i0 + 1, in - 1);
    r.w1 = src - (float)r.i0;
    r.w0 = 1.f - r.w1;
    return r;
}

template <typename T, int V>
__global__ void __launch_bounds__(256) upsample_fwd_kernel(b200_up_desc d, const T* __restrict__ x, T* __restrict__ y) {
    const int CV = d.C / V;
    const int64_t total = (int64_t)d.N * d.Do * d.Ho * d.Wo * CV;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = i / CV;
        const int cv = (int)(i - v * CV);
        const int64_t vox = v;
        const int xo = (int)(v % d.Wo); v /= d.Wo;
        const int yo = (int)(v % d.Ho); v /= d.Ho;
        const int zo = (int)(v % d.Do);
        const int n = (int)(v / d.Do);
        const Lin1 lz = lin_src(zo, d.Di, d.Do, d.mode), ly = lin_src(yo, d.Hi, d.Ho, d.mode), lx = lin_src(xo, d.Wi, d.Wo, d.mode);
        const T* xb = x + (int64_t)n * d.Di * d.Hi * d.Wi * d.C + cv * V;
        float acc[V];
        if (d.mode == B200_UP_NEAREST) {
            Pack<T, V>::load(xb + (((int64_t)lz.i0 * d.Hi + ly.i0) * d.Wi + lx.i0) * d.C, acc);
        } else {
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] = 0.f;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float w = (a ? lz.w1 : lz.w0) * (b ? ly.w1 : ly.w0) * (c ? lx.w1 : lx.w0);
                        float t[V];
                        Pack<T, V>::load(xb + (((int64_t)(a ? lz.i1 : lz.i0) * d.Hi + (b ? ly.i1 : ly.i0)) * d.Wi + (c ? lx.i1 : lx.i0)) * d.C, t);
#pragma unroll
                        for (int k = 0; k < V; ++k) acc[k] = fmaf(w, t[k], acc[k]);
                    }
        }
        Pack<T, V>::store(y + vox * d.Ctot + d.c_off + cv * V, acc);
    }
}

// For input index j along one dimension: the output indices whose interpolation touches j and their weights.
// Upsampling by an integer factor f touches at most 2f+2 outputs; we support f <= 4 (reference uses 2 and 4).
constexpr int kMaxTouch = 12;
struct Touch { int n; int o[kMaxTouch]; float w[kMaxTouch]; };
__device__ __forceinline__ void touching(int j, int in, int out, int mode, Touch& t) {
    t.n = 0;
    const float inv = (float)out / (float)in;
    int lo = (int)floorf((j - 1) * inv) - 2, hi = (int)ceilf((j + 2) * inv) + 2;
    lo = max(lo, 0); hi = min(hi, out - 1);
    for (int o = lo; o <= hi; ++o) {
        const Lin1 l = lin_src(o, in, out, mode);
        float w = 0.f;
        if (l.i0 == j) w += l.w0;
        if (l.i1 == j && mode != B200_UP_NEAREST) w += l.w1;
        if (w != 0.f && t.n < kMaxTouch) { t.o[t.n] = o; t.w[t.n] = w; ++t.n; }
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) upsample_bwd_kernel(b200_up_desc d, const T* __restrict__ dy, T* __restrict__ dx) {
    const int CV = d.C / V;
    const int64_t total = (int64_t)d.N * d.Di * d.Hi * d.Wi * CV;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = i / CV;
        const int cv = (int)(i - v * CV);
        const int xi = (int)(v % d.Wi); v /= d.Wi;
        const int yi = (int)(v % d.Hi); v /= d.Hi;
        const int zi = (int)(v % d.Di);
        const int n = (int)(v / d.Di);
        Touch tz, ty, tx;
        touching(zi, d.Di, d.Do, d.mode, tz);
        touching(yi, d.Hi, d.Ho, d.mode, ty);
        touching(xi, d.Wi, d.Wo, d.mode, tx);
        float acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = 0.f;
        const T* gb = dy + (int64_t)n * d.Do * d.Ho * d.Wo * d.Ctot + d.c_off + cv * V;
        for (int a = 0; a < tz.n; ++a)
            for (int b = 0; b < ty.n; ++b) {
                const float wab = tz.w[a] * ty.w[b];
                const T* row = gb + ((int64_t)tz.o[a] * d.Ho + ty.o[b]) * d.Wo * d.Ctot;
                for (int c = 0; c < tx.n; ++c) {
                    float g[V];
                    Pack<T, V>::load(row + (int64_t)tx.o[c] * d.Ctot, g);
                    const float w = wab * tx.w[c];
#pragma unroll
                    for (int k = 0; k < V; ++k) acc[k] = fmaf(w, g[k], acc[k]);
                }
            }
        Pack<T, V>::store(dx + i * V, acc);
    }
}

// ------------------------------------------------------------------------------------------------ x2 trilinear fast path
// Exact factor 2, align_corners=False (unet3d.py:73,85; unet.UNet decoder): output 2i+a reads inputs (i-1+a, i+a) with weights
// (0.25, 0.75) / (0.75, 0.25), indices clamped at the borders -- no floating-point index arithmetic, one block per output row,
// 32-bit index math only.  Same arithmetic order as the generic kernel (weights are exact in binary).
struct Lin2 { int i0, i1; float w0, w1; };
__device__ __forceinline__ Lin2 lin2(int o, int in) {
    Lin2 r;
    const int i = o >> 1;
    if (o & 1) { r.i0 = i; r.i1 = min(i + 1, in - 1); r.w0 = 0.75f; r.w1 = 0.25f; }
    else if (i == 0) { r.i0 = 0; r.i1 = min(1, in - 1); r.w0 = 1.f; r.w1 = 0.f; }         // src clamped to 0
    else { r.i0 = i - 1; r.i1 = i; r.w0 = 0.25f; r.w1 = 0.75f; }
    return r;
}

// One thread = the output pair (2i, 2i+1) of one row and one channel vector: the three input columns i-1, i, i+1 of the four
// contributing (z, y) rows are loaded once (12 vector loads for 2 outputs instead of 16), the (z, y) weighted sums are shared.
template <typename T, int V>
__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(b200_up_desc d, const T* __restrict__ x, T* __restrict__ y) {
    const int CV = d.C / V;
    const int rows = d.N * d.Do * d.Ho, per_row = d.Wi * CV;
    // narrow rows (the 2-channel fp32 logits of the deep-supervision heads: 64 work items per row) share a block: 256 / per_row rows
    // per block pass, so no thread idles
    const int tpr = per_row < 256 ? per_row : 256, rpb = 256 / tpr;
    const int sub = threadIdx.x / tpr, lane_e = threadIdx.x - sub * tpr;
    if (sub >= rpb) return;
    for (int row = blockIdx.x * rpb + sub; row < rows; row += gridDim.x * rpb) {
        const int yo = row % d.Ho, zo = (row / d.Ho) % d.Do, n = row / (d.Ho * d.Do);
        const Lin2 lz = lin2(zo, d.Di), ly = lin2(yo, d.Hi);
        const T* xn = x + (int64_t)n * d.Di * d.Hi * d.Wi * d.C;
        const T* r00 = xn + ((int64_t)lz.i0 * d.Hi + ly.i0) * d.Wi * d.C;
        const T* r01 = xn + ((int64_t)lz.i0 * d.Hi + ly.i1) * d.Wi * d.C;
        const T* r10 = xn + ((int64_t)lz.i1 * d.Hi + ly.i0) * d.Wi * d.C;
        const T* r11 = xn + ((int64_t)lz.i1 * d.Hi + ly.i1) * d.Wi * d.C;
        const float w00 = lz.w0 * ly.w0, w01 = lz.w0 * ly.w1, w10 = lz.w1 * ly.w0, w11 = lz.w1 * ly.w1;
        T* yr = y + (int64_t)row * d.Wo * d.Ctot + d.c_off;
        for (int e = lane_e; e < per_row; e += tpr) {
            const int xi = e / CV, cv = e - xi * CV;
            const int xm = max(xi - 1, 0), xp = min(xi + 1, d.Wi - 1);
            const int om = xm * d.C + cv * V, oc = xi * d.C + cv * V, op = xp * d.C + cv * V;
            float a[4][3][V];
            Pack<T, V>::load(r00 + om, a[0][0]); Pack<T, V>::load(r00 + oc, a[0][1]); Pack<T, V>::load(r00 + op, a[0][2]);
            Pack<T, V>::load(r01 + om, a[1][0]); Pack<T, V>::load(r01 + oc, a[1][1]); Pack<T, V>::load(r01 + op, a[1][2]);
            Pack<T, V>::load(r10 + om, a[2][0]); Pack<T, V>::load(r10 + oc, a[2][1]); Pack<T, V>::load(r10 + op, a[2][2]);
            Pack<T, V>::load(r11 + om, a[3][0]); Pack<T, V>::load(r11 + oc, a[3][1]); Pack<T, V>::load(r11 + op, a[3][2]);
            float c[3][V];                                  // (z, y)-interpolated columns i-1, i, i+1
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int k = 0; k < V; ++k) c[j][k] = fmaf(w11, a[3][j][k], fmaf(w10, a[2][j][k], fmaf(w01, a[1][j][k], w00 * a[0][j][k])));
            // even output 2i: 0.25*c[i-1] + 0.75*c[i] (i = 0: the clamped source is exactly column 0); odd output 2i+1:
            // 0.75*c[i] + 0.25*c[i+1] (the clamped i+1 at the right border is column i itself)
            float ev[V], od[V];
            const float we = xi == 0 ? 0.f : 0.25f;
#pragma unroll
            for (int k = 0; k < V; ++k) {
                ev[k] = fmaf(we, c[0][k], (1.f - we) * c[1][k]);
                od[k] = fmaf(0.25f, c[2][k], 0.75f * c[1][k]);
            }
            T* dst = yr + (int64_t)(2 * xi) * d.Ctot + cv * V;
            Pack<T, V>::store(dst, ev);
            Pack<T, V>::store(dst + d.Ctot, od);
        }
    }
}

// Sliding-window variant for vector widths: one thread owns the input cell row (n, zi, yi), one x-segment and one channel vector and
// produces the 2 x 2 output rows (2zi+a, 2yi+b) of that segment.  Per input column it loads the 9 neighbouring rows once, folds
// them into the four (z, y)-interpolated values c[a][b], keeps c of the previous column in registers
// and emits the outputs between consecutive columns: ~1/3 of the instructions of the pair kernel above (which remains the scalar path).
template <typename T, int V, int SEG>
__global__ void __launch_bounds__(128, 3) upsample2x_fwd_slide_kernel(b200_up_desc d, const T* __restrict__ x, T* __restrict__ y) {
    const int CV = d.C / V, nseg = (d.Wi + SEG - 1) / SEG;
    const int64_t items = (int64_t)d.N * d.Di * d.Hi * nseg * CV;
    for (int64_t it = (int64_t)blockIdx.x * 128 + threadIdx.x; it < items; it += (int64_t)gridDim.x * 128) {
        int64_t r = it;
        const int cv = (int)(r % CV); r /= CV;
        const int seg = (int)(r % nseg); r /= nseg;
        const int yi = (int)(r % d.Hi); r /= d.Hi;
        const int zi = (int)(r % d.Di);
        const int n = (int)(r / d.Di);
        // weights of input planes/rows (i-1, i, i+1) for the even (a = 0) and odd (a = 1) output of index i, clamped borders folded in
        float wz[2][3], wy[2][3];
        wz[0][0] = zi > 0 ? 0.25f : 0.f; wz[0][1] = zi > 0 ? 0.75f : 1.f; wz[0][2] = 0.f;
        wz[1][0] = 0.f; wz[1][1] = zi + 1 < d.Di ? 0.75f : 1.f; wz[1][2] = zi + 1 < d.Di ? 0.25f : 0.f;
        wy[0][0] = yi > 0 ? 0.25f : 0.f; wy[0][1] = yi > 0 ? 0.75f : 1.f; wy[0][2] = 0.f;
        wy[1][0] = 0.f; wy[1][1] = yi + 1 < d.Hi ? 0.75f : 1.f; wy[1][2] = yi + 1 < d.Hi ? 0.25f : 0.f;
        const T* rows[3][3];
#pragma unroll
        for (int dz = 0; dz < 3; ++dz)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int z = min(max(zi + dz - 1, 0), d.Di - 1), yy = min(max(yi + dy - 1, 0), d.Hi - 1);
                rows[dz][dy] = x + ((((int64_t)n * d.Di + z) * d.Hi + yy) * d.Wi) * d.C + cv * V;
            }
        auto column = [&](int xi, float (&c)[2][2][V]) {
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int k = 0; k < V; ++k) c[a][b][k] = 0.f;
#pragma unroll
            for (int dz = 0; dz < 3; ++dz)
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    float v[V];
                    Pack<T, V>::load(rows[dz][dy] + (int64_t)xi * d.C, v);
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        if ((a == 0 && dz == 2) || (a == 1 && dz == 0)) continue;          // structural zeros
#pragma unroll
                        for (int b = 0; b < 2; ++b) {
                            if ((b == 0 && dy == 2) || (b == 1 && dy == 0)) continue;
                            const float w = wz[a][dz] * wy[b][dy];
#pragma unroll
                            for (int k = 0; k < V; ++k) c[a][b][k] = fmaf(w, v[k], c[a][b][k]);
                        }
                    }
                }
        };
        const int x0 = seg * SEG, x1 = min(x0 + SEG, d.Wi);
        T* out0 = y + ((((int64_t)n * d.Do + 2 * zi) * d.Ho + 2 * yi) * d.Wo) * d.Ctot + d.c_off + cv * V;
        const int64_t ystride = (int64_t)d.Wo * d.Ctot, zstride = (int64_t)d.Ho * d.Wo * d.Ctot;
        // column xi yields the odd output 2xi-1 (with column xi-1) and the even output 2xi; only two column sets stay live
        auto emit = [&](const float (&p)[2][2][V], const float (&c)[2][2][V], int xi) {
            const float we = xi == 0 ? 0.f : 0.25f;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    T* dst = out0 + a * zstride + b * ystride + (int64_t)(2 * xi) * d.Ctot;
                    float o[V];
                    if (xi > x0) {
#pragma unroll
                        for (int k = 0; k < V; ++k) o[k] = fmaf(0.25f, c[a][b][k], 0.75f * p[a][b][k]);
                        Pack<T, V>::store(dst - d.Ctot, o);
                    }
                    if (xi < x1) {
#pragma unroll
                        for (int k = 0; k < V; ++k) o[k] = fmaf(we, p[a][b][k], (1.f - we) * c[a][b][k]);
                        Pack<T, V>::store(dst, o);
                    }
                }
        };
        float ca[2][2][V], cb[2][2][V];
        column(max(x0 - 1, 0), cb);
        for (int xi = x0; xi <= x1; xi += 2) {
            column(min(xi, d.Wi - 1), ca);
            emit(cb, ca, xi);
            if (xi + 1 <= x1) {
                column(min(xi + 1, d.Wi - 1), cb);
                emit(ca, cb, xi + 1);
            }
        }
    }
}

// adjoint: input index j receives from outputs 2j-1 (0.25), 2j (0.75), 2j+1 (0.75), 2j+2 (0.25), with the clamped borders folded in
struct Touch2 { int o[4]; float w[4]; };
__device__ __forceinline__ Touch2 touch2(int j, int in) {
    Touch2 t;
    const int out = 2 * in;
    t.o[0] = max(2 * j - 1, 0);        t.w[0] = (2 * j - 1 >= 1) ? 0.25f : 0.f;          // odd output 2(j-1)+1, i1 = j
    t.o[1] = 2 * j;                    t.w[1] = j == 0 ? 1.f : 0.75f;                     // even output 2j (clamped: all weight on 0)
    t.o[2] = 2 * j + 1;                t.w[2] = (j == in - 1) ? 1.f : 0.75f;              // odd output 2j+1 (i1 clamps to j at the end)
    t.o[3] = min(2 * j + 2, out - 1);  t.w[3] = (2 * j + 2 <= out - 2) ? 0.25f : 0.f;     // even output 2(j+1), i0 = j
    return t;
}

// Sliding-window adjoint: one thread owns the input cell row (n, zi, yi), one x-segment and one channel vector.  For every
// output column ox it folds the 4 x 4 (z, y) rows that touch the cell into G[ox] (16 loads), and each input column xi combines
// G[2xi-1 .. 2xi+2]; two of those four carry over from the previous column, so a column costs 32 loads instead of 64.
template <typename T, int V, int SEG>
__global__ void __launch_bounds__(128) upsample2x_bwd_slide_kernel(b200_up_desc d, const T* __restrict__ dy, T* __restrict__ dx) {
    const int CV = d.C / V, nseg = (d.Wi + SEG - 1) / SEG;
    const int64_t items = (int64_t)d.N * d.Di * d.Hi * nseg * CV;
    for (int64_t it = (int64_t)blockIdx.x * 128 + threadIdx.x; it < items; it += (int64_t)gridDim.x * 128) {
        int64_t r = it;
        const int cv = (int)(r % CV); r /= CV;
        const int seg = (int)(r % nseg); r /= nseg;
        const int yi = (int)(r % d.Hi); r /= d.Hi;
        const int zi = (int)(r % d.Di);
        const int n = (int)(r / d.Di);
        const Touch2 tz = touch2(zi, d.Di), ty = touch2(yi, d.Hi);
        const T* gn = dy + (int64_t)n * d.Do * d.Ho * d.Wo * d.Ctot + d.c_off + cv * V;
        const T* rows[4][4];
        float wzy[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                rows[a][b] = gn + ((int64_t)tz.o[a] * d.Ho + ty.o[b]) * d.Wo * d.Ctot;
                wzy[a][b] = tz.w[a] * ty.w[b];
            }
        auto gcol = [&](int ox, float (&g)[V]) {
#pragma unroll
            for (int k = 0; k < V; ++k) g[k] = 0.f;
            const int64_t off = (int64_t)ox * d.Ctot;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    float v[V];
                    Pack<T, V>::load(rows[a][b] + off, v);
#pragma unroll
                    for (int k = 0; k < V; ++k) g[k] = fmaf(wzy[a][b], v[k], g[k]);
                }
        };
        const int x0 = seg * SEG, x1 = min(x0 + SEG, d.Wi);
        float ga[V], gb[V], gc[V], gd[V];
        gcol(max(2 * x0 - 1, 0), ga);
        gcol(2 * x0, gb);
        T* xr = dx + ((((int64_t)n * d.Di + zi) * d.Hi + yi) * d.Wi) * d.C + cv * V;
        for (int xi = x0; xi < x1; ++xi) {
            gcol(2 * xi + 1, gc);
            gcol(min(2 * xi + 2, d.Wo - 1), gd);
            const Touch2 tx = touch2(xi, d.Wi);
            float o[V];
#pragma unroll
            for (int k = 0; k < V; ++k) o[k] = fmaf(tx.w[0], ga[k], fmaf(tx.w[1], gb[k], fmaf(tx.w[2], gc[k], tx.w[3] * gd[k])));
            Pack<T, V>::store(xr + (int64_t)xi * d.C, o);
#pragma unroll
            for (int k = 0; k < V; ++k) { ga[k] = gc[k]; gb[k] = gd[k]; }
        }
    }
}

// Tiled adjoint for bf16 (the decoder's upsamples): a 256-thread block owns TY x 16 input cells of one (n, zi) plane and CG
// channel vectors.  Stage 1 folds the four output planes that touch zi into one z-reduced fp32 tile of (2TY+2) x 34 output
// positions in shared memory (coalesced 16-byte loads, every dy element of the tile converted once); stage 2 gathers the 4 x 4
// (y, x) taps of each cell from that tile.  ~30% fewer instructions than the sliding-window kernel and 4x its occupancy.
template <int CG>
struct UpBwdTile {
    static constexpr int TX = 16, TY = 256 / (16 * CG), ROWS = 2 * TY + 2, COLS = 2 * TX + 2;
    static constexpr int COL4 = CG * 2 + CG / 2;                     // float4 per column: [half][cv] + padding against bank conflicts
    static constexpr int SMEM = ROWS * COLS * COL4 * 16;
};
template <int CG>
__global__ void __launch_bounds__(256) upsample2x_bwd_tile_kernel(b200_up_desc d, const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx) {
    using G = UpBwdTile<CG>;
    extern __shared__ float4 up_sm[];
    const int ntx = (d.Wi + G::TX - 1) / G::TX, nty = (d.Hi + G::TY - 1) / G::TY, ncg = d.C / (8 * CG);
    int b = blockIdx.x;
    const int xt = b % ntx; b /= ntx;
    const int yt = b % nty; b /= nty;
    const int cg = b % ncg; b /= ncg;
    const int zi = b % d.Di, n = b / d.Di;
    const int x0 = xt * G::TX, y0 = yt * G::TY;
    const Touch2 tz = touch2(zi, d.Di);
    const int64_t plane = (int64_t)d.Ho * d.Wo * d.Ctot;
    const __nv_bfloat16* gn = dy + (int64_t)n * d.Do * plane + d.c_off + cg * CG * 8;
    for (int it = threadIdx.x; it < G::ROWS * G::COLS * CG; it += 256) {
        const int cv = it % CG, col = (it / CG) % G::COLS, row = it / (CG * G::COLS);
        const int oy = min(max(2 * y0 - 1 + row, 0), d.Ho - 1), ox = min(max(2 * x0 - 1 + col, 0), d.Wo - 1);
        const __nv_bfloat16* src = gn + ((int64_t)oy * d.Wo + ox) * d.Ctot + cv * 8;
        float v[4][8], acc[8];
#pragma unroll
        for (int a = 0; a < 4; ++a) Pack<__nv_bfloat16, 8>::load(src + tz.o[a] * plane, v[a]);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(tz.w[0], v[0][k], fmaf(tz.w[1], v[1][k], fmaf(tz.w[2], v[2][k], tz.w[3] * v[3][k])));
        float4* dst = up_sm + (row * G::COLS + col) * G::COL4 + cv;
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dst[CG] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncthreads();
    const int cv = threadIdx.x % CG, xl = (threadIdx.x / CG) % G::TX, yl = threadIdx.x / (CG * G::TX);
    const int yi = y0 + yl, xi = x0 + xl;
    if (yi >= d.Hi || xi >= d.Wi) return;
    const Touch2 ty = touch2(yi, d.Hi), tx = touch2(xi, d.Wi);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int bb = 0; bb < 4; ++bb)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float w = ty.w[bb] * tx.w[c];
            const float4* src = up_sm + ((2 * yl + bb) * G::COLS + 2 * xl + c) * G::COL4 + cv;
            const float4 lo = src[0], hi = src[CG];
            acc[0] = fmaf(w, lo.x, acc[0]); acc[1] = fmaf(w, lo.y, acc[1]); acc[2] = fmaf(w, lo.z, acc[2]); acc[3] = fmaf(w, lo.w, acc[3]);
            acc[4] = fmaf(w, hi.x, acc[4]); acc[5] = fmaf(w, hi.y, acc[5]); acc[6] = fmaf(w, hi.z, acc[6]); acc[7] = fmaf(w, hi.w, acc[7]);
        }
    Pack<__nv_bfloat16, 8>::store(dx + ((((int64_t)n * d.Di + zi) * d.Hi + yi) * d.Wi + xi) * d.C + (cg * CG + cv) * 8, acc);
}

// bf16 tiled adjoint (upsample2x_bwd_tile_kernel); false when the grid would not fit
template <int CG>
inline bool up_bwd_tile_launch_cg(const b200_up_desc* d, const void* dy, void* dx, void* stream, int* rc) {
    using G = UpBwdTile<CG>;
    const int64_t blocks = (int64_t)d->N * d->Di * ceil_div(d->Hi, G::TY) * ceil_div(d->Wi, G::TX) * (d->C / (8 * CG));
    if (blocks >= (1ll << 31)) return false;
    static cudaError_t attr = cudaFuncSetAttribute(upsample2x_bwd_tile_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
    if (attr != cudaSuccess) { *rc = fail("upsample_bwd: cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(attr)); return true; }
    *rc = [&]() -> int {
        B200_LAUNCH((upsample2x_bwd_tile_kernel<CG>), (int)blocks, 256, G::SMEM, stream, *d, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx);
        return 0;
    }();
    return true;
}
inline bool up_bwd_tile_launch(const b200_up_desc* d, const void* dy, void* dx, void* stream, int* rc) {
    return d->C % 32 == 0 ? up_bwd_tile_launch_cg<4>(d, dy, dx, stream, rc) : up_bwd_tile_launch_cg<2>(d, dy, dx, stream, rc);
}

template <typename T, int V>
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(b200_up_desc d, const T* __restrict__ dy, T* __restrict__ dx) {
    const int CV = d.C / V;
    const int rows = d.N * d.Di * d.Hi, per_row = d.Wi * CV;
    const int tpr = per_row < 256 ? per_row : 256, rpb = 256 / tpr;          // narrow rows share a block (see the forward kernel)
    const int sub = threadIdx.x / tpr, lane_e = threadIdx.x - sub * tpr;
    if (sub >= rpb) return;
    for (int row = blockIdx.x * rpb + sub; row < rows; row += gridDim.x * rpb) {
        const int yi = row % d.Hi, zi = (row / d.Hi) % d.Di, n = row / (d.Hi * d.Di);
        const Touch2 tz = touch2(zi, d.Di), ty = touch2(yi, d.Hi);
        const T* gn = dy + (int64_t)n * d.Do * d.Ho * d.Wo * d.Ctot + d.c_off;
        T* xr = dx + (int64_t)row * d.Wi * d.C;
        for (int e = lane_e; e < per_row; e += tpr) {
            const int xi = e / CV, cv = e - xi * CV;
            const Touch2 tx = touch2(xi, d.Wi);
            float acc[V];
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] = 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if (tz.w[a] == 0.f) continue;                                   // uniform per row
                float g[4][4][V];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const T* grow = gn + ((int64_t)tz.o[a] * d.Ho + ty.o[b]) * d.Wo * d.Ctot + cv * V;
#pragma unroll
                    for (int c = 0; c < 4; ++c) Pack<T, V>::load(grow + (int64_t)tx.o[c] * d.Ctot, g[b][c]);
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float wab = tz.w[a] * ty.w[b];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float w = wab * tx.w[c];
#pragma unroll
                        for (int k = 0; k < V; ++k) acc[k] = fmaf(w, g[b][c][k], acc[k]);
                    }
                }
            }
            Pack<T, V>::store(xr + e * V, acc);
        }
    }
}

inline bool upsample_is_2x_trilinear(const b200_up_desc* d) {
    return d->mode == B200_UP_TRILINEAR && d->Do == 2 * d->Di && d->Ho == 2 * d->Hi && d->Wo == 2 * d->Wi &&
           (int64_t)d->N * d->Do * d->Ho < (1ll << 31) && (int64_t)d->Wo * d->C < (1ll << 30);
}

}  // namespace b200
