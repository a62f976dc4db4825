// "Row-slab" tcgen05 convolution kernels for sm_100a (second generation; conv_umma.cuh keeps the general 16x8-tile kernels).
//
// Shared-memory image: whole x-rows of the NDHWC activation tensor, [row][voxel][C channels] with C*2 = 32/64/128 bytes
// per voxel, written by ONE TMA tensor load per (plane, row block) in the matching SWIZZLE_32B/64B/128B mode (probe:
// tools/desc_probe.cu -- ~25 B/cycle/SM, HBM-bound, against ~1 B/cycle/SM for the 16-byte rows of the first design).
// The same image is a canonical K-major operand (rows = voxels, forward/dgrad) and a canonical MN-major operand
// (rows = channels, wgrad) of tcgen05.mma.
//
// ---- wgrad: "shifted operand" formulation ------------------------------------------------------------------------
//   dW[co][ci][kz][ky][kx] = sum_u dy[u][co] * x[u + (kz-pd, ky-ph, kx-pw)][ci]
// A tcgen05.mma contracts K = 16 consecutive voxels of one x-row.  In the MN-major swizzled layouts the byte distance
// between consecutive "atoms" of the M (or N) dimension is the descriptor's LBO field -- a free parameter.  Setting it
// to ONE VOXEL for the x operand and to ONE ROW for the dy operand makes the hardware enumerate shifted copies:
//   A[(s, ci), v] = x [z+kz-pd][y      ][16k + v + s - pw][ci]      s = kx in 0..kw-1  (M = 128 = (128/cP) atoms of cP channels)
//   B[(a, co), v] = dy[z      ][y-ph+a ][16k + v         ][co]      a in 0..kh-1       (N = kh*cQ)
//   D[(s, ci), (a, co)] += A * B^T   ==  dW[co][ci][kz][ky = kh-1-a][kx = s]   (one accumulator per kz, resident in TMEM)
// so ONE instruction accumulates 9 filter taps (27 with the three kz accumulators) instead of one tap per instruction with
// 7/8 of the M rows idle: 8-9x fewer tensor instructions for the 16/32-channel layers that dominate the 3-D U-Net.
// A CTA owns a (ci chunk, co chunk) unit, sweeps z over its share of (sample, row block, z segment) items with a ring of x
// planes, and writes its fp32 partial once; a second kernel reduces the partials in a fixed order (deterministic).
#pragma once
#include "conv_umma.cuh"

namespace b200 {

namespace ptx {
__device__ __forceinline__ void tma_load_4d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    tma_load_4d(dst, map, bar, c0, c1, c2, c3);
}
}  // namespace ptx

// UMMA layout-type field (bits 61..63 of the descriptor = bits 29..31 of the high word) for a row of `cbytes` per voxel
__host__ __device__ inline uint32_t row_layout_bits(int cbytes) { return cbytes == 32 ? 6u : (cbytes == 64 ? 4u : 2u); }
inline CUtensorMapSwizzle row_swizzle(int cbytes) {
    return cbytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (cbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
}

// 4-D activation map (C, W, H, planes) with a (cbox, wbox, hbox, 1) box
inline int make_row_map(CUtensorMap* map, const void* ptr, int C, int W, int H, int64_t planes, int cbox, int wbox, int hbox) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    B200_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is unavailable in this driver");
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {(cuuint32_t)cbox, (cuuint32_t)wbox, (cuuint32_t)hbox, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           row_swizzle(cbox * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (C=%d W=%d H=%d box %d,%d,%d)", (int)r, C, W, H, cbox, wbox, hbox);
    return 0;
}

// ================================================================================================ wgrad
constexpr int kRwThreads = 192;          // warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = final epilogue
constexpr int kRwMaxX = 4, kRwMaxQ = 3;

struct RowWgradParams {
    int N, D, H, W, Ci, Co;
    int kd, kh, kw;              // 3 or 1 each (kh == kw); padding = k/2, stride 1
    int cP, cQ;                  // ci / co chunk of one unit
    int n_ci, n_co;              // chunks; units = n_ci * n_co
    int splits;                  // CTAs per unit
    int YB;                      // x rows per block
    int zsegs, zs;               // z segments per sample, planes per segment
    int yblocks;
    int items;                   // N * yblocks * zsegs
    int nsx, nsq;                // ring depths
    int x_bytes, q_bytes;        // bytes of one x plane block / one dy stage (1024-byte multiples)
    int x_tx, q_tx;              // bytes the TMA box delivers
    int NN;                      // UMMA N = kh * cQ
    int tmem_cols;
    uint32_t idesc;
    float* partial;              // [unit][split][kd*NN][kw*cP]
};

struct alignas(128) RowWgradBarriers {
    uint64_t xfull[kRwMaxX], xempty[kRwMaxX];
    uint64_t qfull[kRwMaxQ], qempty[kRwMaxQ];
    uint64_t done;
    uint32_t tmem_base, started;
};

struct RwItem { int n, y0, z0, z1; };
__device__ __forceinline__ RwItem rw_decode(const RowWgradParams& p, int item) {
    RwItem c;
    const int zg = item % p.zsegs; item /= p.zsegs;
    const int yb = item % p.yblocks;
    c.n = item / p.yblocks;
    c.y0 = yb * p.YB;
    c.z0 = zg * p.zs;
    c.z1 = min(p.D, c.z0 + p.zs);
    return c;
}

__global__ void __launch_bounds__(kRwThreads, 1)
row_wgrad_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap dy_map, const RowWgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte aligned base (swizzle patterns are functions of the absolute shared-memory address)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ RowWgradBarriers bars;
    uint8_t* xbuf = smem;
    uint8_t* qbuf = smem + (size_t)p.nsx * p.x_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x % p.splits;
    const int unit = blockIdx.x / p.splits;
    const int cic = unit % p.n_ci, coc = unit / p.n_ci;
    const int pd = p.kd >> 1, ph = p.kh >> 1, pw = p.kw >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.nsx; ++i) { ptx::mbar_init(ptx::smem_u32(&bars.xfull[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars.xempty[i]), 1); }
        for (int i = 0; i < p.nsq; ++i) { ptx::mbar_init(ptx::smem_u32(&bars.qfull[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars.qempty[i]), 1); }
        ptx::mbar_init(ptx::smem_u32(&bars.done), 1);
        bars.started = 0;
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&x_map);
        ptx::prefetch_tmap(&dy_map);
    }
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(&bars.tmem_base), (uint32_t)p.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars.tmem_base;

    if (warp == 0) {
        // ===================================================== TMA producer: one new x plane block per z step + the dy stage
        if (lane == 0) {
            uint32_t xc = 0, qc = 0;                    // ring counters (monotonic)
            for (int item = split; item < p.items; item += p.splits) {
                const RwItem c = rw_decode(p, item);
                int next_plane = max(c.z0 - pd, 0);
                for (int z = c.z0; z < c.z1; ++z) {
                    const int last_needed = min(z + pd, p.D - 1);
                    for (; next_plane <= last_needed; ++next_plane, ++xc) {
                        const uint32_t s = xc % p.nsx, phs = (xc / p.nsx) & 1;
                        ptx::mbar_wait(ptx::smem_u32(&bars.xempty[s]), phs ^ 1);
                        const uint32_t full = ptx::smem_u32(&bars.xfull[s]);
                        ptx::mbar_expect_tx(full, (uint32_t)p.x_tx);
                        ptx::tma_load_4d(ptx::smem_u32(xbuf + (size_t)s * p.x_bytes), &x_map, full, cic * p.cP, -pw, c.y0, c.n * p.D + next_plane);
                    }
                    const uint32_t s = qc % p.nsq, phs = (qc / p.nsq) & 1;
                    ptx::mbar_wait(ptx::smem_u32(&bars.qempty[s]), phs ^ 1);
                    const uint32_t full = ptx::smem_u32(&bars.qfull[s]);
                    ptx::mbar_expect_tx(full, (uint32_t)p.q_tx);
                    ptx::tma_load_4d(ptx::smem_u32(qbuf + (size_t)s * p.q_bytes), &dy_map, full, coc * p.cQ, 0, c.y0 - ph, c.n * p.D + z);
                    ++qc;
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (whole warp converged, one elected lane issues)
        const uint32_t cpb = (uint32_t)p.cP * 2, cqb = (uint32_t)p.cQ * 2;
        const uint32_t a_hi = ((8 * cpb) >> 4) | (1u << 14) | (row_layout_bits((int)cpb) << 29);      // SBO = next 8 voxels
        const uint32_t b_hi = ((8 * cqb) >> 4) | (1u << 14) | (row_layout_bits((int)cqb) << 29);
        const uint32_t a_lbo = (cpb >> 4) << 16;                                                     // LBO = one voxel  -> atom s = kx
        const uint32_t b_lbo = (((uint32_t)p.W * cqb) >> 4) << 16;                                   // LBO = one row    -> atom a
        const uint32_t x16 = ptx::smem_u32(xbuf) >> 4, xs16 = (uint32_t)p.x_bytes >> 4;
        const uint32_t q16 = ptx::smem_u32(qbuf) >> 4, qs16 = (uint32_t)p.q_bytes >> 4;
        const uint32_t a_row16 = ((uint32_t)(p.W + 2 * pw) * cpb) >> 4, b_row16 = ((uint32_t)p.W * cqb) >> 4;
        const uint32_t a_k16 = (16 * cpb) >> 4, b_k16 = (16 * cqb) >> 4;
        const int nk = p.W / 16;
        const uint32_t idesc = p.idesc;
        uint32_t xc = 0, qc = 0, started = 0;
        for (int item = split; item < p.items; item += p.splits) {
            const RwItem c = rw_decode(p, item);
            const int pfirst = max(c.z0 - pd, 0);
            const uint32_t xc0 = xc;                               // ring index of plane pfirst
            int next_plane = pfirst;
            const int rows = min(p.YB, p.H - c.y0);
            for (int z = c.z0; z < c.z1; ++z) {
                const int last_needed = min(z + pd, p.D - 1);
                for (; next_plane <= last_needed; ++next_plane, ++xc) {
                    const uint32_t s = xc % p.nsx, phs = (xc / p.nsx) & 1;
                    ptx::mbar_wait(ptx::smem_u32(&bars.xfull[s]), phs);
                }
                const uint32_t qs = qc % p.nsq, qph = (qc / p.nsq) & 1;
                ptx::mbar_wait(ptx::smem_u32(&bars.qfull[qs]), qph);
                ptx::tc_fence_after();
                // operand bases of this z step
                uint32_t a_base[3], d_tmem[3];
                uint32_t valid = 0;
#pragma unroll
                for (int kz = 0; kz < 3; ++kz) {
                    a_base[kz] = 0; d_tmem[kz] = tmem_base + (uint32_t)(kz * p.NN);
                    const int pl = z + kz - pd;
                    if (kz < p.kd && pl >= 0 && pl < p.D) {
                        const uint32_t ridx = xc0 + (uint32_t)(pl - pfirst);
                        a_base[kz] = ((x16 + (ridx % p.nsx) * xs16) & 0x3FFF) | a_lbo;
                        valid |= 1u << kz;
                    }
                }
                const uint32_t b_base = ((q16 + qs * qs16) & 0x3FFF) | b_lbo;
                if (ptx::elect_one()) {
                    uint32_t acc0 = started & 1, acc1 = (started >> 1) & 1, acc2 = (started >> 2) & 1;
                    for (int r = 0; r < rows; ++r) {
                        uint32_t ao = (uint32_t)r * a_row16, bo = (uint32_t)r * b_row16;
                        for (int k = 0; k < nk; ++k) {
                            const uint32_t b_lo = b_base + bo;
                            if (valid & 1) { ptx::umma_bf16_lohi(d_tmem[0], a_base[0] + ao, a_hi, b_lo, b_hi, idesc, acc0); acc0 = 1; }
                            if (valid & 2) { ptx::umma_bf16_lohi(d_tmem[1], a_base[1] + ao, a_hi, b_lo, b_hi, idesc, acc1); acc1 = 1; }
                            if (valid & 4) { ptx::umma_bf16_lohi(d_tmem[2], a_base[2] + ao, a_hi, b_lo, b_hi, idesc, acc2); acc2 = 1; }
                            ao += a_k16; bo += b_k16;
                        }
                    }
                    ptx::umma_commit(ptx::smem_u32(&bars.qempty[qs]));
                    // the oldest plane of the window is not needed by z+1; at the end of the segment release everything
                    const int lo = z - pd, hi = (z + 1 == c.z1) ? min(z + pd, p.D - 1) : lo;
                    for (int pl = max(lo, pfirst); pl <= hi; ++pl) {
                        const uint32_t ridx = xc0 + (uint32_t)(pl - pfirst);
                        ptx::umma_commit(ptx::smem_u32(&bars.xempty[ridx % p.nsx]));
                    }
                }
                __syncwarp();
                started |= valid;
                ++qc;
            }
        }
        if (ptx::elect_one()) {
            bars.started = started;
            __threadfence_block();
            ptx::umma_commit(ptx::smem_u32(&bars.done));
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue (once): TMEM -> fp32 partial[unit][split][col][lane]
        const int lane_grp = warp & 3;
        const int row = lane_grp * 32 + lane;                               // (s, ci_local) = row / cP, row % cP
        const int mrows = p.kw * p.cP;                                      // useful accumulator rows
        ptx::mbar_wait(ptx::smem_u32(&bars.done), 0);
        ptx::tc_fence_after();
        const uint32_t started = *reinterpret_cast<volatile uint32_t*>(&bars.started);
        if (lane_grp * 32 < mrows) {
            const int ncols = p.kd * p.NN;
            float* dst = p.partial + ((size_t)blockIdx.x * ncols) * mrows + row;
            for (int c0 = 0; c0 < ncols; c0 += 16) {
                float v[16];
                const int kz = c0 / p.NN;                                   // NN is a multiple of 16
                if ((started >> kz) & 1) ptx::tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c0, v);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
                if (row < mrows) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) dst[(size_t)(c0 + i) * mrows] = v[i];
                }
            }
        }
        ptx::tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
}

// dw[(co*Ci + ci)*taps + tap] = sum_split partial[unit][split][kz*NN + a*cQ + co_l][kx*cP + ci_l],  ky = kh-1-a
// Neighbouring threads take neighbouring weight elements (they share cache lines; a warp-per-element variant with the splits on
// the lanes was 5x slower -- 148 scattered sectors per element).  The chain of `splits` dependent loads is what costs (20 us
// for ANY layer size with one thread per element), so for layers with few elements SG thread groups of a block each take every
// SG-th split and the groups are combined through shared memory in a fixed order (deterministic).
template <int SG>
__global__ void __launch_bounds__(256) row_wgrad_reduce_kernel(RowWgradParams p, float* __restrict__ dw) {
    constexpr int EL = 256 / SG;
    __shared__ float red[SG][EL];
    const int el = threadIdx.x % EL, sg = threadIdx.x / EL;
    const int taps = p.kd * p.kh * p.kw;
    const int64_t total = (int64_t)p.Co * p.Ci * taps;
    const int mrows = p.kw * p.cP, ncols = p.kd * p.NN;
    for (int64_t base = (int64_t)blockIdx.x * EL; base < total; base += (int64_t)gridDim.x * EL) {
        const int64_t e = base + el;
        float acc = 0.f;
        if (e < total) {
            const int tap = (int)(e % taps);
            const int64_t r = e / taps;
            const int ci = (int)(r % p.Ci), co = (int)(r / p.Ci);
            const int kz = tap / (p.kh * p.kw), ky = (tap / p.kw) % p.kh, kx = tap % p.kw;
            const int a = p.kh - 1 - ky;
            const int cic = ci / p.cP, cil = ci % p.cP, coc = co / p.cQ, col_ = co % p.cQ;
            const int unit = coc * p.n_ci + cic;
            const float* src = p.partial + ((size_t)unit * p.splits * ncols + (size_t)(kz * p.NN + a * p.cQ + col_)) * mrows + kx * p.cP + cil;
#pragma unroll 8
            for (int s = sg; s < p.splits; s += SG) acc += __ldg(src + (size_t)s * ncols * mrows);      // same order every run; 8 loads in flight
        }
        if (SG == 1) {
            if (e < total) dw[e] = acc;
        } else {
            red[sg][el] = acc;
            __syncthreads();
            if (sg == 0 && e < total) {
                float t = red[0][el];
#pragma unroll
                for (int g = 1; g < SG; ++g) t += red[g][el];
                dw[e] = t;
            }
            __syncthreads();
        }
    }
}

inline bool row_wgrad_supported(const b200_conv_desc* d) {
    if (!d->allow_umma || d->transposed) return false;
    if (d->x_dtype != B200_BF16 || d->y_dtype != B200_BF16) return false;
    if (d->sd != 1 || d->sh != 1 || d->sw != 1 || d->dd != 1 || d->dh != 1 || d->dw != 1) return false;
    if (!((d->kd == 1 || d->kd == 3) && (d->kh == 1 || d->kh == 3) && d->kw == d->kh)) return false;
    if (d->pd != d->kd / 2 || d->ph != d->kh / 2 || d->pw != d->kw / 2) return false;
    if (d->Ci % 16 || d->Co % 16 || d->Ci > 512 || d->Co > 512) return false;
    if (d->Wi % 16 || d->Wi > 240 || d->Wi < 16) return false;
    if ((int64_t)d->N * d->Do * d->Ho * d->Wo < 1024) return false;       // tiny problems: launch-bound either way
    return true;
}

inline int row_wgrad_plan(const b200_conv_desc* d, RowWgradParams* p, size_t* smem_bytes, size_t* partial_bytes) {
    memset(p, 0, sizeof *p);
    p->N = d->N; p->D = d->Di; p->H = d->Hi; p->W = d->Wi; p->Ci = d->Ci; p->Co = d->Co;
    p->kd = d->kd; p->kh = d->kh; p->kw = d->kw;
    const int cmax = d->kh == 3 ? 32 : 64;            // 3x3: atoms of <= 32 channels (>= 3 shifted atoms in M = 128, N = 3*cQ <= 96)
    p->cP = (d->Ci % 64 == 0 && cmax == 64) ? 64 : (d->Ci % 32 == 0 ? 32 : 16);
    p->cQ = (d->Co % 64 == 0 && cmax == 64) ? 64 : (d->Co % 32 == 0 ? 32 : 16);
    p->n_ci = d->Ci / p->cP; p->n_co = d->Co / p->cQ;
    p->NN = p->kh * p->cQ;
    int cols = p->kd * p->NN, pow2 = 32;
    while (pow2 < cols) pow2 <<= 1;
    B200_REQUIRE(pow2 <= 512, "row wgrad: accumulators do not fit TMEM");
    p->tmem_cols = pow2;
    p->nsx = p->kd == 3 ? 4 : 2; p->nsq = 2;
    const int cpb = p->cP * 2, cqb = p->cQ * 2, pw = p->kw / 2, ph = p->kh / 2;
    const size_t budget = 200 * 1024;
    int YB = d->Hi < 8 ? d->Hi : 8;
    for (; YB >= 1; --YB) {
        const size_t xb = (((size_t)YB * (p->W + 2 * pw) * cpb) + 1023) & ~(size_t)1023;
        const size_t qb = (((size_t)(YB + 2 * ph) * p->W * cqb) + 1023) & ~(size_t)1023;
        if ((size_t)p->nsx * xb + (size_t)p->nsq * qb <= budget) break;
    }
    B200_REQUIRE(YB >= 1, "row wgrad: a row block does not fit shared memory");
    p->YB = YB;
    p->x_tx = YB * (p->W + 2 * pw) * cpb;
    p->q_tx = (YB + 2 * ph) * p->W * cqb;
    p->x_bytes = (p->x_tx + 1023) & ~1023;
    p->q_bytes = (p->q_tx + 1023) & ~1023;
    // + 1 KB: shifted atoms of the last row read a few voxels past the block; + 1 KB: manual 1024-byte alignment
    *smem_bytes = (size_t)p->nsx * p->x_bytes + (size_t)p->nsq * p->q_bytes + 2048;
    p->yblocks = (d->Hi + YB - 1) / YB;
    const int units = p->n_ci * p->n_co;
    int splits = kNumSMs / units;
    if (splits < 1) splits = 1;
    const int64_t columns = (int64_t)d->N * p->yblocks;
    int zsegs = 1;
    while (columns * zsegs < (int64_t)splits * 3 && zsegs * 8 <= d->Di) zsegs *= 2;     // >= 3 items per CTA, segments >= 8 planes
    p->zs = (d->Di + zsegs - 1) / zsegs;
    p->zsegs = (d->Di + p->zs - 1) / p->zs;
    const int64_t items = columns * p->zsegs;
    if (splits > items) splits = (int)items;
    p->splits = splits;
    p->items = (int)items;
    p->idesc = make_idesc_bf16(p->NN) | (1u << 15) | (1u << 16);            // A and B are MN-major
    *partial_bytes = (size_t)units * splits * p->kd * p->NN * p->kw * p->cP * sizeof(float);
    return 0;
}

inline size_t row_wgrad_workspace_bytes(const b200_conv_desc* d) {
    RowWgradParams p;
    size_t smem = 0, part = 0;
    if (row_wgrad_plan(d, &p, &smem, &part)) return 0;
    int64_t chunks = ((int64_t)d->N * d->Do * d->Ho * d->Wo + 4095) / 4096;
    if (chunks > 2 * kNumSMs) chunks = 2 * kNumSMs;
    if (chunks < 1) chunks = 1;
    return ((part + 255) & ~(size_t)255) + (size_t)chunks * d->Co * 4 + 512;
}

inline int row_wgrad_run(const b200_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, void* workspace, size_t ws_bytes,
                         void* stream) {
    B200_REQUIRE(row_wgrad_supported(d), "row wgrad: unsupported descriptor");
    B200_REQUIRE(aligned16(x) && aligned16(dy), "row wgrad: pointers must be 16-byte aligned");
    RowWgradParams p;
    size_t smem_bytes = 0, partial_bytes = 0;
    if (row_wgrad_plan(d, &p, &smem_bytes, &partial_bytes)) return 1;
    B200_REQUIRE(ws_bytes >= row_wgrad_workspace_bytes(d), "row wgrad: workspace too small");
    p.partial = (float*)workspace;
    const int pw = p.kw / 2, ph = p.kh / 2;
    CUtensorMap x_map, dy_map;
    if (make_row_map(&x_map, x, d->Ci, d->Wi, d->Hi, (int64_t)d->N * d->Di, p.cP, p.W + 2 * pw, p.YB)) return 1;
    if (make_row_map(&dy_map, dy, d->Co, d->Wo, d->Ho, (int64_t)d->N * d->Do, p.cQ, p.W, p.YB + 2 * ph)) return 1;
    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [] { attr_err = cudaFuncSetAttribute(row_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); });
    B200_REQUIRE(attr_err == cudaSuccess, "row wgrad: cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(attr_err));
    const int units = p.n_ci * p.n_co;
    B200_LAUNCH(row_wgrad_kernel, units * p.splits, kRwThreads, smem_bytes, stream, x_map, dy_map, p);
    const int64_t total = (int64_t)d->kd * d->kh * d->kw * d->Co * d->Ci;
    if (total <= 128 * 1024) B200_LAUNCH(row_wgrad_reduce_kernel<8>, stream_grid(total, 32), 256, 0, stream, p, dw);
    else B200_LAUNCH(row_wgrad_reduce_kernel<1>, stream_grid(total, 256), 256, 0, stream, p, dw);
    if (dbias != nullptr) {
        float* bpart = (float*)((char*)workspace + ((partial_bytes + 255) & ~(size_t)255));
        const int64_t Vy = (int64_t)d->N * d->Do * d->Ho * d->Wo;
        int64_t chunks = (Vy + 4095) / 4096;
        if (chunks > 2 * kNumSMs) chunks = 2 * kNumSMs;
        if (chunks < 1) chunks = 1;
        const int64_t rows_per_chunk = (Vy + chunks - 1) / chunks;
        if (colsum_launch<__nv_bfloat16>((const __nv_bfloat16*)dy, d->Co, Vy, (int)chunks, rows_per_chunk, bpart, stream)) return 1;
        B200_LAUNCH(colsum_final_kernel, (int)((d->Co + 127) / 128), 128, 0, stream, (int)chunks, d->Co, bpart, dbias);
    }
    return 0;
}

}  // namespace b200
