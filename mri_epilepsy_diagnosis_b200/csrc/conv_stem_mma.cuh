// One-channel 3x3x3 stems (Cin = 1 -> Cout = 16) on the warp-level tensor path (mma.sync m16n8k16 bf16, fp32 accumulate).
//
// The contraction is tiny (K = 27 taps) but the FFMA kernels of conv_small.cuh spend 432 multiply-adds per voxel on it: forward
// 0.25 ms and weight gradient 0.34 ms on 4 x 128^3, five to seven times the HBM time of the 36 bytes per voxel they move.  tcgen05
// does not fit (a 128 x 16 x 32 tile per instruction, operands from shared memory in a canonical layout that would have to be
// materialised per voxel); the legacy warp MMA takes its operands from REGISTERS, so the im2col fragment is built on the fly from a
// shared-memory halo tile of the fp32 volume:
//   * the fp32 input is split x = hi + lo (two bf16, |x - hi - lo| <= 2^-17 |x|), the fp32 weights likewise in the forward pass; the
//     products hi*w_hi + lo*w_hi + hi*w_lo are accumulated in fp32 -- the same split the tcgen05 stem wgrad of round 2 used;
//   * halo tile: 3 planes x (rows + 2) x 130 columns of fp32, row pitch 137 and plane pitch 1371 words: with these pitches every
//     fragment load of either kernel (8 voxels x 4 taps, or 8 taps x 4 voxel pairs, per instruction) hits 32 distinct banks;
//   * forward:  D[16 voxels][16 co] += A[16 voxels][32 taps] * B[32 taps][16 co], A from the halo tile, B (weights) in registers;
//               a quad transpose (3 shuffles) turns the accumulator layout into one 16-byte store per lane, 512 B per warp;
//   * wgrad:    D[32 taps][16 co] += A[32 taps][16 voxels] * B[16 voxels][16 co], B = dy staged by cp.async and read with
//               ldmatrix.trans; tap 27 is the constant 1, so row 27 of D is the bias gradient.  Accumulators stay in registers
//               over the whole sweep of a block; per-block partials are reduced by stem3_wgrad_reduce_kernel in a fixed order.
#pragma once
#include <mutex>

#include "conv_small.cuh"

namespace b200 {

constexpr int kSmTX = 128;                 // x chunk of a tile
constexpr int kSmPW = 137;                 // halo row pitch (words)
constexpr int kSmFwdRows = 8, kSmWgRows = 4;
constexpr int kSmFwdPP = 1371;             // plane pitch, forward tile (10 rows)
constexpr int kSmWgPP = 827;               // plane pitch, wgrad tile (6 rows = 822 words; 827 = 1371 mod 32 keeps the bank pattern)
constexpr int kSmWgStage = kSmWgRows * kSmTX * 32 + ((3 * kSmWgPP * 4 + 15) & ~15);      // dy tile + halo tile
constexpr int kSmWgSmem = 2 * kSmWgStage;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (v0, v1) -> packed bf16 hi parts (v0 in the low half) and packed bf16 lo parts of the remainders
__device__ __forceinline__ void split_bf16x2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    hi = *reinterpret_cast<uint32_t*>(&h);
    const float r0 = v0 - __uint_as_float(hi << 16), r1 = v1 - __uint_as_float(hi & 0xffff0000u);
    __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
    lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ int stem_tap_off(int tap, int PP) {           // word offset of tap (kz, ky, kx) in the halo tile; pad taps -> tap 0
    if (tap >= 27) tap = 0;
    return (tap / 9) * PP + ((tap / 3) % 3) * kSmPW + tap % 3;
}

// halo tile of plane z-1..z+1, rows y0-1 .. y0+ROWS, columns x0-1 .. x0+128 (zero outside the volume).  fp32 volumes: 4-byte
// cp.async with zero fill (the caller commits / waits; the tile of item i+1 is in flight while item i is computed -- a synchronous
// fill left the 15 dependent load -> store round trips of a tile exposed and made this path SLOWER than the FFMA kernel);
// bf16 volumes are converted on the way and filled synchronously.
// Index arithmetic is the cost of this loop (the first version, one flat index with two divisions and 64-bit addressing per element,
// was HALF of the kernel's instructions): thread -> (row = 2*pass + tid / 128, column = tid % 128), 32-bit offsets from the tile's
// corner voxel, and one extra pass for the two right-hand halo columns.  256 threads.
template <typename TX, int ROWS, int PP>
__device__ __forceinline__ void stem_fill_halo(float* tile, const TX* __restrict__ x, int64_t n, int z, int y0, int x0, int D, int H, int W) {
    constexpr int HR = ROWS + 2, NR = 3 * HR;
    static_assert(NR % 2 == 0, "row pairs");
    const int tid = threadIdx.x;
    const TX* corner = x + ((n * D + z) * H + y0) * (int64_t)W + x0;          // voxel (z, y0, x0)
    auto put = [&](int row, int c) {
        const int pl = row / HR, r = row - pl * HR;
        const int zz = z + pl - 1, yy = y0 + r - 1, xx = x0 + c - 1;
        const bool ok = (unsigned)zz < (unsigned)D && (unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W;
        const int off = ((pl - 1) * H + (r - 1)) * W + (c - 1);                // |off| <= 2*H*W: 32 bits
        const TX* src = ok ? corner + off : x;
        float* dstp = tile + pl * PP + r * kSmPW + c;
        if constexpr (sizeof(TX) == 4) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dstp)), "l"(src), "r"(ok ? 4 : 0) : "memory");
        } else {
            *dstp = ok ? ldg_f<TX>(src) : 0.f;
        }
    };
#pragma unroll 3
    for (int p = 0; p < NR / 2; ++p) put(2 * p + (tid >> 7), tid & 127);
    if (tid < 2 * NR) put(tid >> 1, kSmTX + (tid & 1));
}
// branch-free select (the ternaries of the quad transpose compiled to divergent branches, 4-way serialised)
__device__ __forceinline__ uint32_t selp_u32(uint32_t a, uint32_t b, uint32_t c) {           // c != 0 ? a : b
    uint32_t d;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tselp.b32 %0, %1, %2, p;\n\t}" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// 4 x 4 transpose inside a quad (lanes t = 0..3, words w[0..3]): afterwards lane t holds w_k[t] in w[k].  Two butterfly stages.
__device__ __forceinline__ void quad_transpose(uint32_t (&w)[4], int t) {
    const uint32_t b0 = t & 1, b1 = t & 2;
#pragma unroll
    for (int k = 0; k < 4; k += 2) {                             // pairs (0,1), (2,3) across lane ^ 1
        const uint32_t r = __shfl_xor_sync(0xffffffffu, selp_u32(w[k], w[k + 1], b0), 1);
        w[k] = selp_u32(r, w[k], b0);
        w[k + 1] = selp_u32(w[k + 1], r, b0);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {                                // pairs (0,2), (1,3) across lane ^ 2
        const uint32_t r = __shfl_xor_sync(0xffffffffu, selp_u32(w[k], w[k + 2], b1), 2);
        w[k] = selp_u32(r, w[k], b1);
        w[k + 2] = selp_u32(w[k + 2], r, b1);
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct StemTile { int64_t n; int z, y0, x0; };
__device__ __forceinline__ StemTile stem_tile(int64_t item, int xchunks, int yblocks, int D, int rows) {
    StemTile t;
    t.x0 = (int)(item % xchunks) * kSmTX; item /= xchunks;
    t.y0 = (int)(item % yblocks) * rows; item /= yblocks;
    t.z = (int)(item % D);
    t.n = item / D;
    return t;
}

// ------------------------------------------------------------------------------------------------ forward
// x [N][D][H][W] fp32 or bf16, w fp32 [27][16], y bf16 [N][D][H][W][16].  256 threads: warp r = output row y0 + r of the tile.
template <typename TX>
__global__ void __launch_bounds__(256, 3) stem3_mma_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                            __nv_bfloat16* __restrict__ y, int N, int D, int H, int W, int64_t items, int xchunks,
                                                            int yblocks) {
    constexpr bool SPLIT = sizeof(TX) == 4;                    // a bf16 input has no lo part
    __shared__ float tiles[2][3 * kSmFwdPP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // B fragments: [k-step][n-tile] x (b0, b1), hi and lo parts of the weights
    uint32_t bh[2][2][2], bl[2][2][2];
    int off[2][4];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
#pragma unroll
        for (int j = 0; j < 4; ++j) off[s][j] = stem_tap_off(16 * s + 2 * t + (j & 1) + (j >> 1) * 8, kSmFwdPP);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k0 = 16 * s + 2 * t + 8 * h, co = 8 * nt + g;
                const float w0 = k0 < 27 ? w[k0 * 16 + co] : 0.f, w1 = k0 + 1 < 27 ? w[(k0 + 1) * 16 + co] : 0.f;
                split_bf16x2(w0, w1, bh[s][nt][h], bl[s][nt][h]);
            }
    }
    float bv[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) { bv[nt][0] = bias ? bias[8 * nt + 2 * t] : 0.f; bv[nt][1] = bias ? bias[8 * nt + 2 * t + 1] : 0.f; }

    {
        const StemTile t0 = stem_tile(blockIdx.x, xchunks, yblocks, D, kSmFwdRows);
        stem_fill_halo<TX, kSmFwdRows, kSmFwdPP>(tiles[0], x, t0.n, t0.z, t0.y0, t0.x0, D, H, W);
        cp_async_commit();
    }
    int buf = 0;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x, buf ^= 1) {
        const StemTile tl = stem_tile(item, xchunks, yblocks, D, kSmFwdRows);
        __syncthreads();                                          // tiles[buf ^ 1] has been consumed (previous iteration)
        if (item + gridDim.x < items) {
            const StemTile tn = stem_tile(item + gridDim.x, xchunks, yblocks, D, kSmFwdRows);
            stem_fill_halo<TX, kSmFwdRows, kSmFwdPP>(tiles[buf ^ 1], x, tn.n, tn.z, tn.y0, tn.x0, D, H, W);
        }
        cp_async_commit();
        cp_async_wait<1>();                                       // this thread's copies of the current tile have landed ...
        __syncthreads();                                          // ... and everybody else's
        const float* tile = tiles[buf];
        const int yy = tl.y0 + warp;
        if (yy >= H) continue;                                    // (warp-uniform; the barriers above are reached by every warp)
        const float* pa[2][4];                                    // fragment source of (k-step, tap slot) for voxel g of m-tile 0
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j) pa[s][j] = tile + warp * kSmPW + g + off[s][j];
        __nv_bfloat16* yrow = y + (((tl.n * D + tl.z) * H + yy) * (int64_t)W + tl.x0) * 16;
        const int xlim = W - tl.x0;                               // valid columns of this chunk
#pragma unroll 4
        for (int mt = 0; mt < kSmTX / 16; ++mt) {
            if (mt * 16 >= xlim) break;
            float acc[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) { acc[nt][0] = acc[nt][2] = bv[nt][0]; acc[nt][1] = acc[nt][3] = bv[nt][1]; }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                uint32_t ah[4], al[4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {                     // taps 2t, 2t+1 (h = 0) / 2t+8, 2t+9 (h = 1)
                    const float* p0 = pa[s][2 * h] + mt * 16, *p1 = pa[s][2 * h + 1] + mt * 16;
                    split_bf16x2(p0[0], p1[0], ah[2 * h], al[2 * h]);             // voxel g
                    split_bf16x2(p0[8], p1[8], ah[2 * h + 1], al[2 * h + 1]);     // voxel g + 8
                }
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    mma_bf16_16816(acc[nt], ah, bh[s][nt][0], bh[s][nt][1]);
                    mma_bf16_16816(acc[nt], ah, bl[s][nt][0], bl[s][nt][1]);
                    if (SPLIT) mma_bf16_16816(acc[nt], al, bh[s][nt][0], bh[s][nt][1]);
                }
            }
            // accumulators: lane (g, t) holds channels 8nt + 2t, +1 of voxels g (c0, c1) and g + 8 (c2, c3).  Quad transpose: lane
            // j = 2*vi + nt ends up with the eight channels 8nt .. 8nt+7 of voxel g + 8*vi -> one 16-byte store per lane
            uint32_t wv[4];
#pragma unroll
            for (int vi = 0; vi < 2; ++vi)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[nt][2 * vi], acc[nt][2 * vi + 1]);
                    wv[2 * vi + nt] = *reinterpret_cast<uint32_t*>(&h2);
                }
            quad_transpose(wv, t);
            const int vox = mt * 16 + g + 8 * (t >> 1);
            if (vox < xlim) *reinterpret_cast<uint4*>(yrow + (int64_t)vox * 16 + 8 * (t & 1)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// partial[block][16][28]: 27 taps + the bias gradient (tap 27 = the constant 1); dy bf16 [N][D][H][W][16].
// 256 threads: warp = (row r = warp % 4 of the tile, x half = warp / 4).
template <typename TX>
__global__ void __launch_bounds__(256, 3) stem3_mma_wgrad_kernel(const TX* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ partial,
                                                              int N, int D, int H, int W, int64_t items, int xchunks, int yblocks) {
    constexpr bool SPLIT = sizeof(TX) == 4;
    // dynamic shared memory (kSmWgSmem): two stages of { dy of the tile [row][x][16 co] bf16 (16 KB), halo tile fp32 }; stage 0's dy
    // buffer is reused for the block reduction
    extern __shared__ __align__(16) uint8_t wg_smem[];
    auto gt = [&](int b) { return reinterpret_cast<__nv_bfloat16*>(wg_smem + (size_t)b * kSmWgStage); };
    auto ht = [&](int b) { return reinterpret_cast<float*>(wg_smem + (size_t)b * kSmWgStage + kSmWgRows * kSmTX * 32); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int r = warp & 3, xh = warp >> 2;
    float acc[2][2][4];                                           // [m-tile: taps 0..15 / 16..31][n-tile][c0..c3]
#pragma unroll
    for (int i = 0; i < 16; ++i) (&acc[0][0][0])[i] = 0.f;
    int off[2][2];                                                // [m-tile][tap g / tap g + 8]
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) off[m][h] = stem_tap_off(16 * m + 8 * h + g, kSmWgPP) + 2 * t;
    const bool one = g == 3;                                      // tap 27 = m-tile 1, row g + 8 with g == 3: the constant 1 (bias gradient)

    auto fill = [&](int b, int64_t item) {
        const StemTile tl = stem_tile(item, xchunks, yblocks, D, kSmWgRows);
        stem_fill_halo<TX, kSmWgRows, kSmWgPP>(ht(b), x, tl.n, tl.z, tl.y0, tl.x0, D, H, W);
        // dy tile: 16-byte chunks by cp.async, zero-filled outside the volume (src-size 0)
        const uint32_t gdst = (uint32_t)__cvta_generic_to_shared(gt(b));
        const __nv_bfloat16* gcorner = dy + (((tl.n * D + tl.z) * H + tl.y0) * (int64_t)W + tl.x0) * 16;
#pragma unroll
        for (int q = 0; q < kSmWgRows; ++q) {                      // 256 threads = one row of 128 voxels x two 16-byte halves
            const int half = threadIdx.x & 1, c = threadIdx.x >> 1;
            const bool ok = tl.y0 + q < H && tl.x0 + c < W;
            const __nv_bfloat16* src = ok ? gcorner + ((q * W + c) * 16 + 8 * half) : dy;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(gdst + (uint32_t)(((q * kSmTX + c) * 16 + 8 * half) * 2)), "l"(src),
                         "r"(ok ? 16 : 0) : "memory");
        }
    };
    fill(0, blockIdx.x);
    cp_async_commit();
    int buf = 0;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x, buf ^= 1) {
        const StemTile tl = stem_tile(item, xchunks, yblocks, D, kSmWgRows);
        __syncthreads();                                          // stage buf ^ 1 has been consumed (previous iteration)
        if (item + gridDim.x < items) fill(buf ^ 1, item + gridDim.x);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (tl.y0 + r >= H) continue;
        const float* rowp = ht(buf) + r * kSmPW + xh * 64;
        const __nv_bfloat16* grow = gt(buf) + (r * kSmTX + xh * 64) * 16;
        const int xlim = W - tl.x0 - xh * 64;
#pragma unroll 2
        for (int kt = 0; kt < 4; ++kt) {                           // 16 voxels per k-step
            if (kt * 16 >= xlim) break;
            // B: ldmatrix.x4.trans over the [16 voxels][16 co] block: matrices (vox 0-7, co 0-7), (vox 8-15, co 0-7), (vox 0-7, co 8-15), (8-15, 8-15)
            uint32_t b[4];
            {
                const int mrow = (lane & 7) + 8 * ((lane >> 3) & 1), mcol = 8 * (lane >> 4);
                const uint32_t addr = (uint32_t)__cvta_generic_to_shared(grow + (kt * 16 + mrow) * 16 + mcol);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(addr));
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                uint32_t ah[4], al[4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {                     // tap row g (h = 0) / g + 8 (h = 1); voxel pairs 2t, 2t+1 and 2t+8, 2t+9
                    const float* p = rowp + kt * 16 + off[m][h];
                    float v0 = p[0], v1 = p[1], v2 = p[8], v3 = p[9];
                    if (m == 1 && h == 1 && one) { v0 = v1 = v2 = v3 = 1.f; }
                    split_bf16x2(v0, v1, ah[h], al[h]);
                    split_bf16x2(v2, v3, ah[2 + h], al[2 + h]);
                }
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    mma_bf16_16816(acc[m][nt], ah, b[2 * nt], b[2 * nt + 1]);
                    if (SPLIT) mma_bf16_16816(acc[m][nt], al, b[2 * nt], b[2 * nt + 1]);
                }
            }
        }
    }
    // block reduction in a fixed order: red[warp][tap 0..31][co 0..15] (fp32, aliases the dy tile: 8 * 512 * 4 = 16 KB)
    cp_async_wait<0>();
    __syncthreads();
    float* red = reinterpret_cast<float*>(gt(0));
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int tap = 16 * m + g + 8 * (i >> 1), co = 8 * nt + 2 * t + (i & 1);
                red[(warp * 32 + tap) * 16 + co] = acc[m][nt][i];
            }
    __syncthreads();
    for (int e = threadIdx.x; e < 16 * 28; e += 256) {
        const int co = e / 28, tap = e % 28;
        float s = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) s += red[(wq * 32 + tap) * 16 + co];
        partial[(int64_t)blockIdx.x * 16 * 28 + e] = s;
    }
}

// ------------------------------------------------------------------------------------------------ host side
inline bool stem3_mma_supported(const b200_conv_desc* d) { return stem3_supported(d) && d->Co == 16; }

inline int stem3_mma_fwd_run(const b200_conv_desc* d, const void* x, const float* w, const float* bias, void* y, void* stream) {
    const int xchunks = (int)ceil_div(d->Wi, kSmTX), yblocks = (int)ceil_div(d->Hi, kSmFwdRows);
    const int64_t items = (int64_t)d->N * d->Di * yblocks * xchunks;
    const int grid = (int)(items < (int64_t)kNumSMs * 8 ? items : (int64_t)kNumSMs * 8);
    if (d->x_dtype == B200_F32)
        B200_LAUNCH(stem3_mma_fwd_kernel<float>, grid, 256, 0, stream, (const float*)x, w, bias, (__nv_bfloat16*)y, d->N, d->Di, d->Hi, d->Wi, items, xchunks, yblocks);
    else
        B200_LAUNCH(stem3_mma_fwd_kernel<__nv_bfloat16>, grid, 256, 0, stream, (const __nv_bfloat16*)x, w, bias, (__nv_bfloat16*)y, d->N, d->Di, d->Hi, d->Wi,
                    items, xchunks, yblocks);
    return 0;
}

// workspace: stem3_wgrad_ws_bytes(d) (kSmallBlocks per-block partials)
inline int stem3_mma_wgrad_run(const b200_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, void* workspace, void* stream) {
    float* partial = (float*)workspace;
    const int xchunks = (int)ceil_div(d->Wi, kSmTX), yblocks = (int)ceil_div(d->Hi, kSmWgRows);
    const int64_t items = (int64_t)d->N * d->Di * yblocks * xchunks;
    const int grid = (int)(items < kSmallBlocks ? items : kSmallBlocks);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(stem3_mma_wgrad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmWgSmem);
        if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(stem3_mma_wgrad_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmWgSmem);
    });
    B200_REQUIRE(attr_err == cudaSuccess, "stem wgrad: cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(attr_err));
    if (d->x_dtype == B200_F32)
        B200_LAUNCH(stem3_mma_wgrad_kernel<float>, grid, 256, kSmWgSmem, stream, (const float*)x, (const __nv_bfloat16*)dy, partial, d->N, d->Di, d->Hi, d->Wi, items,
                    xchunks, yblocks);
    else
        B200_LAUNCH(stem3_mma_wgrad_kernel<__nv_bfloat16>, grid, 256, kSmWgSmem, stream, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, partial, d->N, d->Di, d->Hi,
                    d->Wi, items, xchunks, yblocks);
    B200_LAUNCH(stem3_wgrad_reduce_kernel, (int)ceil_div(16 * 28, 8), 256, 0, stream, partial, grid, 16, dw, dbias);
    return 0;
}

}  // namespace b200
