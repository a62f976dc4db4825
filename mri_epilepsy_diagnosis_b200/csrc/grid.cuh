// Row f-3: sliding-grid patch inference and random-patch sampling on the device -- the data movement of torchio's
// GridSampler / GridAggregator / ImageSampler as the reference calls them (segmentation/pretraining_3d_unet.ipynb
// [cell 26, 35]: 64^3 patches, overlap 4; segmentation/routine.py:150-178: Queue of random 64^3 patches).
//
//   grid_gather     out[l][c][dz][dy][dx] = vol[c][z0+dz][y0+dy][x0+dx]        one launch for all L windows of a volume
//   grid_aggregate  vol[z][y][x] = labels[l*][0][z-z0][y-y0][x-x0] for the LAST window l* (in location order) whose
//                   border-cropped box [ini+b, fin-b) contains the voxel -- exactly the result of the reference's
//                   sequential `output[i_ini:i_fin, ...] = window` overwrites; voxels no cropped window covers keep 0.
// Both are pure copies (bit-exact by construction); consecutive threads = consecutive x (coalesced on the volume side).
#pragma once
#include "common.cuh"

namespace b200 {

template <typename T>
__global__ void __launch_bounds__(256) grid_gather_kernel(const T* __restrict__ vol, const int32_t* __restrict__ loc /* [L][6] */, int64_t L, int C,
                                                          int D, int H, int W, int pd, int ph, int pw, T* __restrict__ out) {
    const int64_t per = (int64_t)C * pd * ph * pw;
    const int64_t total = L * per;
    for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < total; id += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = id;
        const int dx = (int)(r % pw); r /= pw;
        const int dy = (int)(r % ph); r /= ph;
        const int dz = (int)(r % pd); r /= pd;
        const int c = (int)(r % C);
        const int64_t l = r / C;
        const int z = loc[l * 6 + 0] + dz, y = loc[l * 6 + 1] + dy, x = loc[l * 6 + 2] + dx;
        out[id] = vol[(((int64_t)c * D + z) * H + y) * W + x];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) grid_aggregate_kernel(const T* __restrict__ labels /* [L][pd][ph][pw] */, const int32_t* __restrict__ loc, int L,
                                                             int D, int H, int W, int pd, int ph, int pw, int bd, int bh, int bw,
                                                             T* __restrict__ vol) {
    const int64_t total = (int64_t)D * H * W;
    for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < total; id += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = id;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H);
        const int z = (int)(r / H);
        for (int l = L - 1; l >= 0; --l) {                 // the last writer wins
            const int z0 = loc[l * 6 + 0], y0 = loc[l * 6 + 1], x0 = loc[l * 6 + 2];
            if (z >= z0 + bd && z < z0 + pd - bd && y >= y0 + bh && y < y0 + ph - bh && x >= x0 + bw && x < x0 + pw - bw) {
                vol[id] = labels[(((int64_t)l * pd + (z - z0)) * ph + (y - y0)) * pw + (x - x0)];
                break;
            }
        }
    }
}

}  // namespace b200
