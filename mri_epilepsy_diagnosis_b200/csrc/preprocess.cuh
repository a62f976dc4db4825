// Intensity preprocessing on the device: histogram standardisation of one MRI volume,
// classification/train_ENC_CLF.ipynb [cell 9] `normalize`:
//   percentile_values = np.percentile(data[mask], percentiles)        13 landmarks percentiles (1, 10, 20, 25, ..., 90, 99)
//   piecewise-linear map of the 11 used landmarks onto the trained `landmarks`, np.digitize + slope * x + intercept in float64
// The percentiles need EXACT order statistics of up to 192^3 floats.  They come from a three-level radix select on the
// order-preserving 32-bit key of a float (12 + 10 + 10 bits): each level is one pass over the volume that histograms only the
// elements whose key prefix matches one of the (deduplicated) target ranks' prefixes, followed by a one-block scan that narrows
// every rank to its digit.  Nothing returns to the host: ranks, prefixes and the final affine map live in a small device state.
// Three reads of the volume for the select + one read and one write for the map: 5 x 4 B per voxel, HBM-bound.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kSelMaxRanks = 32;          // 2 order statistics per percentile, <= 16 percentiles
constexpr int kSelMaxQ = 16;
constexpr int kSelBins0 = 4096, kSelBinsL = 1024;
constexpr int kSelIlp = 8;                // independent loads per thread and iteration in the histogram passes

using HsParams = b200_histstd_desc;       // host values, passed to the kernels by value (no H2D copies, capture-safe)
static_assert(sizeof(((b200_histstd_desc*)nullptr)->q) / sizeof(double) == kSelMaxQ, "descriptor arrays are kSelMaxQ long");

struct SelState {
    uint32_t n;                           // number of selected (masked) elements
    int32_t R, nuniq;
    uint32_t rank[kSelMaxRanks], prefix[kSelMaxRanks], resid[kSelMaxRanks], uniq[kSelMaxRanks];
    int32_t umap[kSelMaxRanks];
    float val[kSelMaxRanks];              // the order statistics
    double gamma[kSelMaxQ];               // interpolation weight of each percentile
    double pct[kSelMaxQ];                 // np.percentile(...) result
    double thr[kSelMaxQ], slope[kSelMaxQ], icpt[kSelMaxQ];
    int32_t nbins;                        // nrange - 1
};

struct SelWorkspace {
    uint32_t hist0[kSelBins0];
    uint32_t histL[2][kSelMaxRanks * kSelBinsL];
    SelState st;
};

__device__ __forceinline__ uint32_t sel_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sel_unkey(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// one shared-memory atomic per distinct bin of the warp (MRI volumes are mostly background: whole warps hit one bin)
__device__ __forceinline__ void sel_bump(uint32_t* sh, uint32_t bin, bool take) {
    const unsigned act = __ballot_sync(0xffffffffu, take);
    if (!take) return;
    const unsigned m = __match_any_sync(act, bin);
    if ((threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(sh + bin, (uint32_t)__popc(m));
}

__global__ void __launch_bounds__(512) sel_hist0_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, int64_t n, uint32_t* __restrict__ hist0) {
    __shared__ uint32_t sh[kSelBins0];
    for (int i = threadIdx.x; i < kSelBins0; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    // every thread runs the same number of iterations (the warp votes need whole warps); kSelIlp independent loads per
    // iteration: one load per vote made the pass latency-bound (85 us for 28 MB)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, step = stride * kSelIlp;
    const int64_t nround = (n + step - 1) / step * step;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += step) {
        float v[kSelIlp];
        bool take[kSelIlp];
#pragma unroll
        for (int k = 0; k < kSelIlp; ++k) {
            const int64_t idx = i + k * stride;
            take[k] = idx < n && (mask == nullptr || mask[idx] != 0);
            v[k] = take[k] ? x[idx] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kSelIlp; ++k) sel_bump(sh, sel_key(v[k]) >> 20, take[k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSelBins0; i += blockDim.x)
        if (sh[i]) atomicAdd(hist0 + i, sh[i]);
}

// level 1: prefix = top 12 bits, digit = next 10; level 2: prefix = top 22 bits, digit = last 10
__global__ void __launch_bounds__(1024) sel_histL_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, int64_t n,
                                                        const SelState* __restrict__ st, int prefix_shift, int digit_shift, uint32_t* __restrict__ histL) {
    extern __shared__ uint32_t shl[];                                   // [nuniq][1024]
    __shared__ uint32_t uq[kSelMaxRanks];
    const int nuniq = st->nuniq;
    if (threadIdx.x < kSelMaxRanks) uq[threadIdx.x] = threadIdx.x < nuniq ? st->uniq[threadIdx.x] : 0xFFFFFFFFu;
    for (int i = threadIdx.x; i < nuniq * kSelBinsL; i += blockDim.x) shl[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, step = stride * kSelIlp;
    const int64_t nround = (n + step - 1) / step * step;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += step) {
        float v[kSelIlp];
        bool tk[kSelIlp];
#pragma unroll
        for (int k = 0; k < kSelIlp; ++k) {
            const int64_t idx = i + k * stride;
            tk[k] = idx < n && (mask == nullptr || mask[idx] != 0);
            v[k] = tk[k] ? x[idx] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kSelIlp; ++k) {
            bool take = tk[k];
            uint32_t bin = 0;
            if (take) {
                const uint32_t key = sel_key(v[k]);
                const uint32_t pfx = key >> prefix_shift;
                int u = -1;
                for (int j = 0; j < nuniq; ++j)
                    if (uq[j] == pfx) u = j;
                take = u >= 0;
                bin = (uint32_t)u * kSelBinsL + ((key >> digit_shift) & (kSelBinsL - 1));
            }
            sel_bump(shl, bin, take);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nuniq * kSelBinsL; i += blockDim.x)
        if (shl[i]) atomicAdd(histL + i, shl[i]);
}

// one thread; works on register/shared copies (a chain of ~350 dependent global loads cost 20 us per scan kernel)
__device__ __forceinline__ void sel_dedupe(SelState* st, const uint32_t* prefix /* shared */, int R) {
    uint32_t uq[kSelMaxRanks];
    int nu = 0;
    for (int r = 0; r < R; ++r) {
        const uint32_t p = prefix[r];
        int u = -1;
        for (int k = 0; k < nu; ++k)
            if (uq[k] == p) u = k;
        if (u < 0) { u = nu; uq[nu] = p; st->uniq[nu] = p; ++nu; }
        st->umap[r] = u;
    }
    st->nuniq = nu;
}

// ranks from the element count (numpy's 'linear' method: virtual index q * (n - 1), neighbours floor and floor + 1), then each
// rank's level-0 bin.  One block of 1024 threads.
__global__ void __launch_bounds__(1024) sel_scan0_kernel(const uint32_t* __restrict__ hist0, SelState* __restrict__ st, HsParams hp) {
    __shared__ uint32_t cum[kSelBins0 + 1];
    __shared__ uint32_t part[1024];
    const int t = threadIdx.x;
    uint32_t h[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { h[k] = hist0[t * 4 + k]; s += h[k]; }
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                                  // inclusive scan of the 1024 partial sums
        const uint32_t v = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t base = part[t] - s;
#pragma unroll
    for (int k = 0; k < 4; ++k) { cum[t * 4 + k] = base; base += h[k]; }
    if (t == 1023) cum[kSelBins0] = base;
    __syncthreads();
    const uint32_t n = cum[kSelBins0];
    if (t == 0) { st->n = n; st->R = 2 * hp.nq; st->nbins = hp.nrange - 1; }
    if (t < hp.nq && n > 0) {
        const double vi = hp.q[t] * (double)(n - 1);
        double lo = floor(vi);
        if (lo < 0.0) lo = 0.0;
        if (lo > (double)(n - 1)) lo = (double)(n - 1);
        const uint32_t r0 = (uint32_t)lo, r1 = r0 + 1 < n ? r0 + 1 : n - 1;
        st->rank[2 * t] = r0; st->rank[2 * t + 1] = r1;
        st->gamma[t] = vi - lo;
    }
    __syncthreads();
    if (t < 2 * hp.nq && n > 0) {
        const uint32_t r = st->rank[t];
        int lo = 0, hi = kSelBins0 - 1;                                    // last bin with cum[bin] <= r
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (cum[mid] <= r) lo = mid; else hi = mid - 1;
        }
        st->prefix[t] = (uint32_t)lo;
        st->resid[t] = r - cum[lo];
        part[t] = (uint32_t)lo;                                            // shared copy for the dedupe
    }
    __syncthreads();
    if (t == 0 && n > 0) sel_dedupe(st, part, 2 * hp.nq);
    if (t == 0 && n == 0) st->nuniq = 0;
}

// one warp per rank: find the digit inside the rank's 1024-bin histogram
__global__ void __launch_bounds__(1024) sel_scanL_kernel(const uint32_t* __restrict__ histL, SelState* __restrict__ st, int last) {
    __shared__ uint32_t newp[kSelMaxRanks];
    const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = st->R;
    const bool live = st->n > 0 && r < R;
    if (live) {
        const uint32_t* h = histL + (size_t)st->umap[r] * kSelBinsL + lane * 32;
        uint32_t s = 0, hv[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) hv[k] = h[k];                         // 32 loads in flight
#pragma unroll
        for (int k = 0; k < 32; ++k) s += hv[k];
        uint32_t inc = s;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        const uint32_t resid = st->resid[r], before = inc - s;
        const bool mine = resid >= before && resid < inc;
        if (mine) {
            uint32_t c = before;
            int d = 0;
#pragma unroll
            for (int k = 0; k < 31; ++k) {
                const bool before_k = d == k && resid >= c + hv[k];       // still searching and the rank lies beyond digit k
                c += before_k ? hv[k] : 0u;
                d += before_k ? 1 : 0;
            }
            const uint32_t np = (st->prefix[r] << 10) | (uint32_t)(lane * 32 + d);
            st->prefix[r] = np;
            newp[r] = np;
            st->resid[r] = resid - c;
            if (last) st->val[r] = sel_unkey(np);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && st->n > 0 && !last) sel_dedupe(st, newp, R);
}

// np.percentile's lerp (numpy/lib/_function_base_impl.py `_lerp`: the difference is taken in the array's float32, the rest in
// float64) and the affine maps of train_ENC_CLF.ipynb [cell 9]; products and sums rounded separately like numpy (no FMA)
__global__ void hs_build_map_kernel(SelState* __restrict__ st, HsParams hp) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int j = 0; j < hp.nq; ++j) {
        const float a = st->val[2 * j], b = st->val[2 * j + 1];
        const double diff = (double)(b - a), t = st->gamma[j];
        st->pct[j] = st->n == 0 ? 0.0 : (t >= 0.5 ? __dsub_rn((double)b, __dmul_rn(diff, 1.0 - t)) : __dadd_rn((double)a, __dmul_rn(diff, t)));
    }
    const int nb = hp.nrange - 1;
    for (int k = 0; k < nb; ++k) {
        const double p0 = st->pct[hp.range_idx[k]], p1 = st->pct[hp.range_idx[k + 1]];
        const double m0 = hp.landmarks[hp.range_idx[k]], m1 = hp.landmarks[hp.range_idx[k + 1]];
        double dp = __dsub_rn(p1, p0);
        const double dm = __dsub_rn(m1, m0);
        const double slope = dp < hp.eps ? 0.0 : __ddiv_rn(dm, dp);          // numpy: diff_perc[diff_perc < eps] = inf -> dm / inf = 0 (sign of dm kept below)
        st->slope[k] = dp < hp.eps ? (dm < 0.0 ? -0.0 : 0.0) : slope;
        st->icpt[k] = __dsub_rn(m0, __dmul_rn(st->slope[k], p0));
        if (k >= 1) st->thr[k - 1] = p0;                                     // np.digitize bins = range_perc[1:-1]
    }
}

__global__ void __launch_bounds__(256) hs_apply_kernel(const float* __restrict__ x, int64_t n, const SelState* __restrict__ st, float* __restrict__ out) {
    __shared__ double thr[kSelMaxQ], slope[kSelMaxQ], icpt[kSelMaxQ];
    __shared__ int nb;
    if (threadIdx.x < kSelMaxQ) { thr[threadIdx.x] = st->thr[threadIdx.x]; slope[threadIdx.x] = st->slope[threadIdx.x]; icpt[threadIdx.x] = st->icpt[threadIdx.x]; }
    if (threadIdx.x == 0) nb = st->nbins;
    __syncthreads();
    const int nthr = nb - 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)x[i];
        int bin = 0;
        for (int k = 0; k < nthr; ++k) bin += (v >= thr[k]) ? 1 : 0;           // np.digitize(right=False) on increasing bins
        out[i] = (float)__dadd_rn(__dmul_rn(slope[bin], v), icpt[bin]);
    }
}

inline size_t histstd_workspace_bytes() { return sizeof(SelWorkspace) + 256; }

inline int histstd_run(const float* x, const uint8_t* mask, int64_t n, const HsParams& hp, float* out, double* pct_out, void* workspace, size_t ws_bytes,
                       void* stream) {
    B200_REQUIRE(x != nullptr && n > 0 && n < (1ll << 32), "hist_standardize: need 1 <= n < 2^32 elements");
    B200_REQUIRE(hp.nq >= 2 && hp.nq <= kSelMaxQ && hp.nrange >= 2 && hp.nrange <= hp.nq, "hist_standardize: bad landmark counts");
    for (int k = 0; k < hp.nq; ++k) B200_REQUIRE(hp.q[k] >= 0.0 && hp.q[k] <= 1.0 && (k == 0 || hp.q[k] >= hp.q[k - 1]), "hist_standardize: percentiles must ascend in [0, 1]");
    for (int k = 0; k < hp.nrange; ++k)
        B200_REQUIRE(hp.range_idx[k] >= 0 && hp.range_idx[k] < hp.nq && (k == 0 || hp.range_idx[k] > hp.range_idx[k - 1]), "hist_standardize: bad range index");
    B200_REQUIRE(workspace != nullptr && ws_bytes >= histstd_workspace_bytes(), "hist_standardize: workspace too small");
    SelWorkspace* ws = (SelWorkspace*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(SelWorkspace), s);
    B200_REQUIRE(e == cudaSuccess, "hist_standardize: memset failed: %s", cudaGetErrorString(e));
    static cudaError_t attr = cudaFuncSetAttribute(sel_histL_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelMaxRanks * kSelBinsL * 4);
    B200_REQUIRE(attr == cudaSuccess, "hist_standardize: cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(attr));
    const int grid = stream_grid(n, 512, 2);
    // the filtered passes keep 128 KB of histograms per block: one 1024-thread block per SM, and half as many flushes
    const int gridL = stream_grid(n, 1024, 1);
    const size_t smemL = (size_t)kSelMaxRanks * kSelBinsL * 4;
    B200_LAUNCH(sel_hist0_kernel, grid, 512, 0, s, x, mask, n, ws->hist0);
    B200_LAUNCH(sel_scan0_kernel, 1, 1024, 0, s, ws->hist0, &ws->st, hp);
    B200_LAUNCH(sel_histL_kernel, gridL, 1024, smemL, s, x, mask, n, &ws->st, 20, 10, ws->histL[0]);
    B200_LAUNCH(sel_scanL_kernel, 1, 1024, 0, s, ws->histL[0], &ws->st, 0);
    B200_LAUNCH(sel_histL_kernel, gridL, 1024, smemL, s, x, mask, n, &ws->st, 10, 0, ws->histL[1]);
    B200_LAUNCH(sel_scanL_kernel, 1, 1024, 0, s, ws->histL[1], &ws->st, 1);
    B200_LAUNCH(hs_build_map_kernel, 1, 32, 0, s, &ws->st, hp);
    if (pct_out != nullptr) {
        e = cudaMemcpyAsync(pct_out, ws->st.pct, sizeof(double) * hp.nq, cudaMemcpyDeviceToDevice, s);
        B200_REQUIRE(e == cudaSuccess, "hist_standardize: copy failed: %s", cudaGetErrorString(e));
    }
    if (out != nullptr) B200_LAUNCH(hs_apply_kernel, stream_grid(n, 256), 256, 0, s, x, n, &ws->st, out);
    return 0;
}

}  // namespace b200
