// Row f-4: FCD mask post-processing on the device -- FCDMaskGenerator._get_predictions_per_batches' scatter, ._postprocess and
// ._masking (detection/model_utils.py:136-216) as three small integer kernels over the sliding-window plan (patches.cuh).
//   scatter  patch_map[slot][row0 / h][slice] = label                                  (:160-178; slot: patch_1->0, patch_3->1, patch_4->2, patch_2->3)
//   vote     res = 0.25 * (4-neighbour sum inside a slab, zero padded); pos = res == 1, neg = res == 0            (:183-190)
//            fixed = 1: map[pos] = 1, map[neg] = 0 (the evidently intended boolean-mask vote)
//            fixed = 0: the reference's behaviour -- it indexes the map with the INT64 0/1 arrays (:191-192), i.e. fancy indexing
//                       along axis 0: slab 0 is overwritten when the array contains a 0, slab 1 when it contains a 1; first with
//                       1 (pos), then with 0 (neg).  Needs only four "any" flags, gathered by the same pass.
//   paint    mask[c0 + kx][Y - j - ky][slice] = patch_map[slot][j / h][slice], kx < w, ky < h, j > 0 (`-0:-h:-1` is empty), boxes of a
//            strip in the reference's order (slots 0, 3, 1, 2): later boxes overwrite earlier ones                (:195-216)
#pragma once
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ int fcd_slot(int c0, int X, int w) {
    const int mid = X / 2 - w;
    return c0 == mid ? 1 : (c0 == X - mid - w ? 2 : (c0 < mid ? 0 : 3));
}

__global__ void __launch_bounds__(256) fcd_scatter_kernel(const int32_t* __restrict__ plan, int64_t rows, const int64_t* __restrict__ labels, int X, int h, int w,
                                                          int ny, int Z, int64_t* __restrict__ pm) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= rows) return;
    const int i = plan[p * 5 + 0], j = plan[p * 5 + 1], c0 = plan[p * 5 + 2];
    pm[((int64_t)fcd_slot(c0, X, w) * ny + j / h) * Z + i] = labels[p];
}

// flags[0] = any(pos), flags[1] = any(!pos), flags[2] = any(neg), flags[3] = any(!neg)
__global__ void __launch_bounds__(256) fcd_vote_kernel(const int64_t* __restrict__ pm, int ny, int Z, int fixed, int64_t* __restrict__ out, int* __restrict__ flags) {
    const int64_t total = (int64_t)4 * ny * Z;
    int f0 = 0, f1 = 0, f2 = 0, f3 = 0;
    for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < total; id += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(id % Z), y = (int)((id / Z) % ny);
        double s = 0.0;                                              // scipy.signal.convolve evaluates 0.25 * neighbours in float64
        if (y > 0) s += 0.25 * (double)pm[id - Z];
        if (y + 1 < ny) s += 0.25 * (double)pm[id + Z];
        if (z > 0) s += 0.25 * (double)pm[id - 1];
        if (z + 1 < Z) s += 0.25 * (double)pm[id + 1];
        const bool pos = s == 1.0, neg = s == 0.0;
        f0 |= pos; f1 |= !pos; f2 |= neg; f3 |= !neg;
        out[id] = fixed ? (pos ? 1 : (neg ? 0 : pm[id])) : pm[id];
    }
    if (__any_sync(0xffffffffu, f0) && (threadIdx.x & 31) == 0) atomicOr(flags + 0, 1);
    if (__any_sync(0xffffffffu, f1) && (threadIdx.x & 31) == 0) atomicOr(flags + 1, 1);
    if (__any_sync(0xffffffffu, f2) && (threadIdx.x & 31) == 0) atomicOr(flags + 2, 1);
    if (__any_sync(0xffffffffu, f3) && (threadIdx.x & 31) == 0) atomicOr(flags + 3, 1);
}

// the reference's int-array-as-index behaviour: map[change_to_pos] = 1 ; map[change_to_neg] = 0 with 0/1 INDEX arrays
__global__ void __launch_bounds__(256) fcd_vote_quirk_kernel(int ny, int Z, const int* __restrict__ flags, int64_t* __restrict__ out) {
    const int64_t slab = (int64_t)ny * Z;
    const int any_pos = flags[0], any_npos = flags[1], any_neg = flags[2], any_nneg = flags[3];
    for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < 2 * slab; id += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(id / slab);
        // index value 0 occurs where the condition is FALSE, index value 1 where it is TRUE
        if (s == 0) { if (any_npos) out[id] = 1; if (any_nneg) out[id] = 0; }
        else { if (any_pos) out[id] = 1; if (any_neg) out[id] = 0; }
    }
}

__global__ void __launch_bounds__(256) fcd_paint_kernel(const int32_t* __restrict__ plan, int64_t rows, const int64_t* __restrict__ pm, int X, int Y, int Z,
                                                        int h, int w, int ny, int want_slot, int64_t* __restrict__ mask) {
    const int64_t per = (int64_t)w * h;
    const int64_t total = rows * per;
    for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < total; id += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = id / per;
        const int e = (int)(id % per), kx = e / h, ky = e % h;
        const int i = plan[p * 5 + 0], j = plan[p * 5 + 1], c0 = plan[p * 5 + 2];
        const int slot = fcd_slot(c0, X, w);
        if (slot != want_slot || j == 0) continue;
        const int x = c0 + kx, y = Y - j - ky;
        if (x >= X || y < 0) continue;
        mask[((int64_t)x * Y + y) * Z + i] = pm[((int64_t)slot * ny + j / h) * Z + i];
    }
}

}  // namespace b200
