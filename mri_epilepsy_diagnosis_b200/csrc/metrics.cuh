// Validation overlap metrics on the device (the counting part): one pass over a predicted label volume
// and the ground truth yields every integer the reference's Dice and IoU need --
//   compute_dice_coefficient (segmentation/metrics.py:312-329):  mask_gt.sum(), mask_pred.sum(), (mask_gt & mask_pred).sum()
//   get_iou_score (segmentation/routine.py:198-204):              #(pred > 0 and gt > 0), #(pred > 0 or gt > 0)
// on uint8 volumes exactly as validate_dsc_asd passes them (:216-237).  Integer work: bit-exact; 2 B per voxel, HBM-bound.
// The surface-distance part of f-2 (neighbour codes + exact Euclidean distance transform) is not built.
#pragma once
#include "common.cuh"

namespace b200 {

// out[0] = sum(gt), out[1] = sum(pred), out[2] = sum(gt & pred), out[3] = #(pred>0 && gt>0), out[4] = #(pred>0 || gt>0)
__global__ void __launch_bounds__(256) overlap_counts_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt, int64_t n,
                                                             unsigned long long* __restrict__ out) {
    unsigned long long c[5] = {0, 0, 0, 0, 0};
    const int64_t n16 = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(gt)) & 15) == 0 ? n / 16 : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 p = __ldg(reinterpret_cast<const uint4*>(pred) + i), g = __ldg(reinterpret_cast<const uint4*>(gt) + i);
        const uint32_t pw[4] = {p.x, p.y, p.z, p.w}, gw[4] = {g.x, g.y, g.z, g.w};
        uint32_t s[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t pv = (pw[w] >> (8 * b)) & 255u, gv = (gw[w] >> (8 * b)) & 255u;
                s[0] += gv; s[1] += pv; s[2] += gv & pv;
                s[3] += (pv > 0 && gv > 0) ? 1u : 0u;
                s[4] += (pv > 0 || gv > 0) ? 1u : 0u;
            }
#pragma unroll
        for (int k = 0; k < 5; ++k) c[k] += s[k];
    }
    for (int64_t i = n16 * 16 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t pv = pred[i], gv = gt[i];
        c[0] += gv; c[1] += pv; c[2] += gv & pv;
        c[3] += (pv > 0 && gv > 0) ? 1u : 0u;
        c[4] += (pv > 0 || gv > 0) ? 1u : 0u;
    }
    __shared__ unsigned long long red[5][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        unsigned long long v = c[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        unsigned long long v = 0;
        for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
        if (v) atomicAdd(out + threadIdx.x, v);                // integer sums: order-independent, exact
    }
}

inline int overlap_counts_run(const uint8_t* pred, const uint8_t* gt, int64_t n, uint64_t* out, void* stream) {
    B200_REQUIRE(pred != nullptr && gt != nullptr && out != nullptr && n > 0, "overlap_counts: bad arguments");
    cudaError_t e = cudaMemsetAsync(out, 0, 5 * sizeof(uint64_t), (cudaStream_t)stream);
    B200_REQUIRE(e == cudaSuccess, "overlap_counts: memset failed: %s", cudaGetErrorString(e));
    B200_LAUNCH(overlap_counts_kernel, stream_grid(n / 16 + 1, 256), 256, 0, stream, pred, gt, n, (unsigned long long*)out);
    return 0;
}

}  // namespace b200
