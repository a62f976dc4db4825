// Sliding-window mirrored patch extraction (detection/patch_utils.py:17-191) on the device.
//
// Volumes are C-order (X,Y,Z) float64: element (c,y,i) at (c*Y + y)*Z + i, so the axial index i is the
// contiguous one and every kernel below maps consecutive threads to consecutive i (coalesced).
// The reference's working slice is rot90(vol[:,:,i]): S[r][c] = vol[c][Y-1-r][i].
//
//   1. first_pos[y][i]  = first column c with gm[c][y][i] > 0 (X when the row is empty)
//   2. slot decisions   : one slot per (pass k, slice i, strip j) in the reference's emission order
//                         -> up to 4 candidate windows + labels, count per slot
//   3. exclusive scan   : order-preserving compaction offsets (single block, deterministic)
//   4. emit plan rows   : {slice, row0, c0, c1, label}
//   5. gather           : out[p][ch][r][k] = target[c0+k | c1-k][Y-1-(row0+r)][slice]
#pragma once
#include "common.cuh"

namespace b200 {

struct PatchGeom {
    int X, Y, Z, h, w, NJ0, NJk, passes;   // strips per slice in the base pass / in the k>=1 passes
    int64_t slots;
};
inline PatchGeom patch_geom(const b200_patch_desc* d) {
    PatchGeom g;
    g.X = d->X; g.Y = d->Y; g.Z = d->Z; g.h = d->h; g.w = d->w;
    g.NJ0 = (d->Y + d->h - 1) / d->h;                         // range(0, Y, h)
    g.NJk = d->Y - d->h > 0 ? (d->Y - d->h + d->h - 1) / d->h : 0;   // range(0, Y-h, h)
    g.passes = (d->with_mask && d->upsample_passes) ? d->h : 1;
    g.slots = (int64_t)d->Z * g.NJ0 + (int64_t)(g.passes - 1) * d->Z * g.NJk;
    return g;
}

__global__ void __launch_bounds__(256) patch_first_pos_kernel(PatchGeom g, const double* __restrict__ gm, int32_t* __restrict__ first_pos) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)g.Y * g.Z) return;
    const int i = (int)(t % g.Z), y = (int)(t / g.Z);
    int c = 0;
    for (; c < g.X; ++c)
        if (gm[((int64_t)c * g.Y + y) * g.Z + i] > 0.0) break;
    first_pos[t] = c;
}

// slot s -> (pass k, slice i, strip j); threads are laid out (k, j, i) with i fastest for coalescing
struct SlotOut { int32_t cnt; int32_t c0[4], c1[4], label[4]; };

__global__ void __launch_bounds__(128) patch_slot_kernel(PatchGeom g, const int32_t* __restrict__ first_pos, const uint8_t* __restrict__ mask,
                                                         int32_t* __restrict__ slot_cnt, int32_t* __restrict__ slot_rows /* [slots][4][3] */,
                                                         int32_t* __restrict__ status) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t base_threads = (int64_t)g.NJ0 * g.Z;
    int k, j, i;
    int64_t slot;
    if (t < base_threads) {
        k = 0; i = (int)(t % g.Z); j = (int)(t / g.Z);
        slot = (int64_t)i * g.NJ0 + j;
    } else {
        const int64_t u = t - base_threads;
        const int64_t per = (int64_t)g.NJk * g.Z;
        if (g.passes <= 1 || per == 0 || u >= (int64_t)(g.passes - 1) * per) return;
        k = 1 + (int)(u / per);
        const int64_t r = u % per;
        i = (int)(r % g.Z); j = (int)(r / g.Z);
        slot = base_threads + ((int64_t)(k - 1) * g.Z + i) * g.NJk + j;
    }
    const int row0 = k + j * g.h;
    const int nrows = min(g.h, g.Y - row0);
    int start = g.X;
    for (int r = 0; r < nrows; ++r) start = min(start, first_pos[(int64_t)(g.Y - 1 - (row0 + r)) * g.Z + i]);
    int cnt = 0;
    if (start < g.X) {                                             // patch_utils.py:153 non-empty strip
        if (start == 0) atomicOr(status, 1);                      // :160 assert start_idx != 0
        const int mid = g.X / 2 - g.w;                            // :158
        int c0s[4], c1s[4], nc = 0;
        if (start < mid) {                                        // :173
            c0s[nc] = start; c1s[nc] = g.X - 1 - start; ++nc;                 // patch_1
            c0s[nc] = g.X - start - g.w; c1s[nc] = start + g.w - 1; ++nc;     // patch_2
        }
        c0s[nc] = mid; c1s[nc] = g.X - 1 - mid; ++nc;                        // patch_3
        c0s[nc] = g.X - mid - g.w; c1s[nc] = mid + g.w - 1; ++nc;            // patch_4
        for (int q = 0; q < nc; ++q) {
            int label = 0;
            if (mask != nullptr) {
                for (int r = 0; r < nrows && !label; ++r) {
                    const int y = g.Y - 1 - (row0 + r);
                    for (int c = c0s[q]; c < c0s[q] + g.w; ++c)
                        if (c >= 0 && c < g.X && mask[((int64_t)c * g.Y + y) * g.Z + i]) { label = 1; break; }
                }
            }
            if (k > 0 && !label) continue;                        // :113-137 positive-only passes
            int32_t* row = slot_rows + (slot * 4 + cnt) * 3;
            row[0] = c0s[q]; row[1] = c1s[q]; row[2] = label;
            ++cnt;
        }
        if (cnt > 0 && nrows != g.h) atomicOr(status, 2);         // ragged strip: reference fails in np.concatenate
    }
    slot_cnt[slot] = cnt;
}

// single-block exclusive scan over slot counts (<= ~40k slots): offsets[s], total -> *count
__global__ void __launch_bounds__(1024) patch_scan_kernel(int64_t slots, const int32_t* __restrict__ cnt, int32_t* __restrict__ offsets,
                                                          int32_t* __restrict__ count) {
    __shared__ int32_t warp_sums[32];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < slots; base += 1024) {
        const int64_t s = base + threadIdx.x;
        const int32_t v = s < slots ? cnt[s] : 0;
        int32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += n;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int32_t w = warp_sums[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t n = __shfl_up_sync(0xffffffffu, wi, o);
                if (threadIdx.x >= o) wi += n;
            }
            warp_sums[threadIdx.x] = wi - w;                      // exclusive warp offsets
        }
        __syncthreads();
        const int32_t excl = carry + warp_sums[threadIdx.x >> 5] + incl - v;
        if (s < slots) offsets[s] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry;
}

__global__ void __launch_bounds__(256) patch_emit_kernel(PatchGeom g, const int32_t* __restrict__ cnt, const int32_t* __restrict__ offsets,
                                                         const int32_t* __restrict__ slot_rows, int32_t* __restrict__ plan) {
    const int64_t slot = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (slot >= g.slots) return;
    const int64_t base_slots = (int64_t)g.Z * g.NJ0;
    int k, i, j;
    if (slot < base_slots) { k = 0; i = (int)(slot / g.NJ0); j = (int)(slot % g.NJ0); }
    else {
        const int64_t u = slot - base_slots;
        const int64_t per = (int64_t)g.Z * g.NJk;
        k = 1 + (int)(u / per);
        const int64_t r = u % per;
        i = (int)(r / g.NJk); j = (int)(r % g.NJk);
    }
    const int n = cnt[slot];
    for (int q = 0; q < n; ++q) {
        int32_t* row = plan + ((int64_t)offsets[slot] + q) * 5;
        const int32_t* src = slot_rows + (slot * 4 + q) * 3;
        row[0] = i; row[1] = k + j * g.h; row[2] = src[0]; row[3] = src[1]; row[4] = src[2];
    }
}

template <typename TO>
__global__ void __launch_bounds__(256) patch_gather_kernel(PatchGeom g, const double* __restrict__ target, const int32_t* __restrict__ plan,
                                                           int64_t rows, TO* __restrict__ out) {
    const int hw = g.h * g.w;
    const int64_t total = rows * 2 * hw;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = e / (2 * hw);
        const int rem = (int)(e - p * 2 * hw);
        const int ch = rem / hw, r = (rem % hw) / g.w, k = rem % g.w;
        const int32_t* row = plan + p * 5;
        const int c = ch == 0 ? row[2] + k : row[3] - k;
        const int y = g.Y - 1 - (row[1] + r);
        double v = 0.0;
        if (c >= 0 && c < g.X && y >= 0 && y < g.Y) v = target[((int64_t)c * g.Y + y) * g.Z + row[0]];
        out[e] = (TO)v;
    }
}

}  // namespace b200

// ------------------------------------------------------------------------------------------------ min-max normalisation
// get_image_patches (detection/patch_utils.py:193-205): target = (target - target.min()) / (target.max() - target.min()) in float64,
// evaluated per element with the same two correctly rounded operations as numpy -> bit-exact.
namespace b200 {

__global__ void __launch_bounds__(256) minmax_partial_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ partial /* [grid][2] */) {
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        lo = fmin(lo, v); hi = fmax(hi, v);
    }
    __shared__ double slo[8], shi[8];
    for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = fmin(lo, slo[w]); hi = fmax(hi, shi[w]); }
        partial[2 * blockIdx.x] = lo; partial[2 * blockIdx.x + 1] = hi;
    }
}

__global__ void __launch_bounds__(256) minmax_apply_kernel(const double* __restrict__ x, int64_t n, const double* __restrict__ partial, int nparts,
                                                           double* __restrict__ out) {
    __shared__ double s[2];
    if (threadIdx.x < 32) {
        double lo = INFINITY, hi = -INFINITY;
        for (int i = threadIdx.x; i < nparts; i += 32) { lo = fmin(lo, partial[2 * i]); hi = fmax(hi, partial[2 * i + 1]); }
        for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
        if (threadIdx.x == 0) { s[0] = lo; s[1] = hi - lo; }
    }
    __syncthreads();
    const double lo = s[0], range = s[1];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (x[i] - lo) / range;
}

}  // namespace b200
