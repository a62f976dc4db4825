// Row f-2, the surface half: neighbour codes + EXACT Euclidean distance transform for `compute_surface_distances`
// (segmentation/metrics.py:25-178, called from segmentation/routine.py:206-214 for every validation volume; 7.5 s per volume on
// the reference's CPU path, almost all of it scipy's distance transform).
//
// Geometry.  The reference correlates the (bounding-box-cropped, zero-padded) mask with a 2x2x2 kernel of bit weights: the
// result lives on the CORNER grid (D+1, H+1, W+1) -- corner (i,j,k) sees the eight voxels (i-1..i, j-1..j, k-1..k), voxels
// outside the volume counting as 0.  The crop only removes all-zero corners, so the kernels below work on the corner grid of
// the whole volume: same codes at the same corners, same distances between them.
//   code      u8  = sum of 128,64,32,16,8,4,2,1 over the eight voxels in the kernel's (a,b,c) order          (metrics.py:123-130)
//   border        = code != 0 && code != 255                                                                   (:133-135)
//   edt           = exact squared distance to the nearest border corner, separable min-plus passes:
//                   x: two sweeps per row;  y, z: per corner, outward search with pruning (stop once r^2 >= best)
//                   int32 for unit spacing (exact; sqrt is taken once, in fp64, by the caller), fp64 with spacing^2 otherwise
//   collect       = (dist2 to the OTHER surface, code) of every border corner, appended through one atomic counter (the caller
//                   sorts by (distance, area), so the order of the append does not matter)
#pragma once
#include "common.cuh"

namespace b200 {

__global__ void __launch_bounds__(256) sd_code_kernel(const uint8_t* __restrict__ mask, int D, int H, int W, uint8_t* __restrict__ code,
                                                      unsigned int* __restrict__ nborder) {
    const int W1 = W + 1, H1 = H + 1;
    const int64_t total = (int64_t)(D + 1) * H1 * W1;
    unsigned int local = 0;
    for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < total; id += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = id;
        const int k = (int)(r % W1); r /= W1;
        const int j = (int)(r % H1);
        const int i = (int)(r / H1);
        unsigned int c = 0;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int z = i + a - 1, y = j + b - 1, x = k + cc - 1;
                    const bool in = (unsigned)z < (unsigned)D && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
                    const unsigned int v = in ? (mask[((int64_t)z * H + y) * W + x] != 0 ? 1u : 0u) : 0u;
                    c |= v << (7 - (a * 4 + b * 2 + cc));
                }
        code[id] = (uint8_t)c;
        local += (c != 0 && c != 255) ? 1u : 0u;
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(nborder, local);
}

template <typename T> struct SdInf;
template <> struct SdInf<int32_t> { static __device__ __forceinline__ int32_t v() { return 1 << 29; } };
template <> struct SdInf<double> { static __device__ __forceinline__ double v() { return 1e300; } };

// x pass: one thread per (i, j) row; forward sweep then backward sweep.  s2 = spacing_x^2 (1 for the integer path).
template <typename T>
__global__ void __launch_bounds__(128) sd_edt_x_kernel(const uint8_t* __restrict__ code, int D1, int H1, int W1, T s2, T* __restrict__ g) {
    const int64_t rows = (int64_t)D1 * H1;
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const uint8_t* c = code + row * W1;
    T* o = g + row * W1;
    int last = -(1 << 20);
    for (int k = 0; k < W1; ++k) {
        const unsigned int v = c[k];
        if (v != 0 && v != 255) last = k;
        const int d = k - last;
        o[k] = last < 0 ? SdInf<T>::v() : (T)d * (T)d * s2;
    }
    last = 1 << 20;
    for (int k = W1 - 1; k >= 0; --k) {
        const unsigned int v = c[k];
        if (v != 0 && v != 255) last = k;
        if (last < (1 << 20)) {
            const int d = last - k;
            const T cand = (T)d * (T)d * s2;
            if (cand < o[k]) o[k] = cand;
        }
    }
}

// min-plus pass along one axis with element stride `stride` and extent `n`: out[p] = min_q in[q] + ((p - q) * spacing)^2
template <typename T>
__global__ void __launch_bounds__(256) sd_edt_axis_kernel(const T* __restrict__ in, int64_t total, int64_t stride, int n, T s2, T* __restrict__ out) {
    for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < total; id += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)((id / stride) % n);
        T best = in[id];
        for (int r = 1; r < n; ++r) {
            const T rr = (T)r * (T)r * s2;
            if (rr >= best) break;                                   // nothing further out can win
            if (p - r >= 0) { const T c = in[id - (int64_t)r * stride] + rr; if (c < best) best = c; }
            if (p + r < n) { const T c = in[id + (int64_t)r * stride] + rr; if (c < best) best = c; }
            if (p - r < 0 && p + r >= n) break;
        }
        out[id] = best;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) sd_collect_kernel(const uint8_t* __restrict__ code, const T* __restrict__ dist2_other, int64_t total,
                                                         T* __restrict__ out_d2, uint8_t* __restrict__ out_code, unsigned int* __restrict__ counter) {
    for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < total; id += (int64_t)gridDim.x * blockDim.x) {
        const unsigned int c = code[id];
        if (c != 0 && c != 255) {
            const unsigned int slot = atomicAdd(counter, 1u);
            out_d2[slot] = dist2_other[id];
            out_code[slot] = (uint8_t)c;
        }
    }
}

template <typename T>
inline int surface_edt_run(const uint8_t* code, int D1, int H1, int W1, T sz2, T sy2, T sx2, T* dist2, T* scratch, void* stream) {
    const int64_t total = (int64_t)D1 * H1 * W1, rows = (int64_t)D1 * H1;
    B200_LAUNCH(sd_edt_x_kernel<T>, (int)ceil_div(rows, 128), 128, 0, stream, code, D1, H1, W1, sx2, dist2);
    B200_LAUNCH(sd_edt_axis_kernel<T>, stream_grid(total, 256), 256, 0, stream, (const T*)dist2, total, (int64_t)W1, H1, sy2, scratch);
    B200_LAUNCH(sd_edt_axis_kernel<T>, stream_grid(total, 256), 256, 0, stream, (const T*)scratch, total, (int64_t)W1 * H1, D1, sz2, dist2);
    return 0;
}

}  // namespace b200
