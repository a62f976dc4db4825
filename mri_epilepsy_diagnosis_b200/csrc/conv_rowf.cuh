// "Row-slab" tcgen05 forward / dgrad convolution for sm_100a (second generation; conv_umma.cuh keeps the 16x8-tile kernel for
// the shapes this one does not take).
//
// Shared-memory image of one input plane block: (YB + 2*ph) whole x-rows of the NDHWC tensor INCLUDING the x halo,
//   [row][x = -pw .. W-1+pw][C channels],  C*2 = 32/64/128 bytes per voxel,
// written by ONE TMA tensor load (box C x (W+2pw) x (YB+2ph) x 1, out-of-volume rows/columns zero-filled = the padding) in
// the matching SWIZZLE_32B/64B/128B mode.  The image is a linear array of "slots" (slot = row*(W+2pw) + x + pw) and, read as a
// K-major swizzled UMMA operand, ANY run of 128 consecutive slots is a valid A operand (probe: tools/desc_probe.cu T2).
// An output tile is 128 consecutive OUTPUT slots; filter tap (kz,ky,kx) is the same run shifted by ky*(W+2pw) + kx slots in
// the image of plane z+kz-pd: 27 descriptor start addresses, no im2col, every input byte staged once per plane block.
// Slots that fall into the x halo produce garbage rows of D that are simply not stored.
//   * planes stream through a ring along z (one new plane per output plane; 3 live for a 3x3x3 filter);
//   * the packed weights of ALL taps stay resident in shared memory, or stream tap by tap through a ring where they do not fit;
//   * kx-folded mode (RowFwdParams::fold): planes WITHOUT the x halo, a tile = 128 / W whole rows, one MMA per (kz, ky) with the
//     three kx weight blocks side by side (N = 3*Cout), the x shift applied to the accumulators in the epilogue;
//   * accumulators: T tiles x Cout fp32 columns of TMEM, two sets, so the epilogue of plane z overlaps the MMAs of z+1;
//   * warp roles (352 threads): 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 and 7..10 = epilogue (TMEM -> bf16 -> global,
//     consecutive lanes = consecutive voxels = fully coalesced stores), optionally accumulating the per-channel sum / sum of
//     squares of the outputs for the BatchNorm that follows (fp32, per-thread partials, one deterministic per-CTA partial; no atomics).
#pragma once
#include "conv_row.cuh"

namespace b200 {

constexpr int kRfThreads = 352;           // warps: 0 TMA, 1 MMA, 2..5 epilogue set 0, 6 weight producer (streaming mode), 7..10 epilogue set 1
constexpr int kRfMaxW = 8;                // weight ring depth (streaming mode)
constexpr int kRfRing = 8;                // deepest plane ring (barrier array size); the planner picks p.ring in [3, 8]
constexpr int kRfMaxTiles = 16;
constexpr int kRfMaxDynSmem = 227 * 1024 - 4096;      // 4 KB reserved for the kernel's static shared memory (barriers, statistics scratch)

struct RowFwdParams {
    int N, D, H, W;              // output == input spatial size ("same" convolution, stride 1)
    int IC, OC;
    int kd, khw;                 // 1 or 3 each (kh == kw == khw); padding = k/2
    int YB;                      // output rows per block
    int pitchW;                  // W + 2*pw slots per image row
    int tpr, rowstride;          // tile t starts at slot (t / tpr) * rowstride + (t % tpr) * 128
    int T;                       // tiles per full block
    int yblocks, zsegs, zs;
    int items;
    int plane_bytes, plane_tx;   // ring slot size (1024-byte multiple) / bytes one TMA box delivers
    int w_bytes;                 // all packed weights
    int stream_w;                // 0: every tap resident in shared memory; 1: taps stream through a ring of nw stages of wtap_bytes
    int nw, wtap_bytes;
    int ring;                    // plane ring depth (3 or 4)
    int tmem_cols;
    uint32_t idesc;
    const __nv_bfloat16* w;      // [tap][IC/8][OC][8]
    const float* bias;           // [OC] or null
    __nv_bfloat16* out;          // [N][D][H][W][OC]
    float* stats;                // [grid][2][OC] fp32 per-CTA (sum, sum of squares) of the outputs (OC <= 64); or null
    int store_c0;                // only channels [store_c0, OC) are stored (out has OC - store_c0 channels); statistics cover all
    // kx-folded mode (small Cout, W = 32 / 64 / 128): ONE tcgen05.mma per (kz, ky) with the UNSHIFTED voxel rows as A and the three kx
    // weight blocks side by side as B (N = 3*OC): TMEM holds P_kx[x] = x[x] . W[kx] in three column blocks per tile and the
    // epilogue forms out[x] = P_0[x-1] + P_1[x] + P_2[x+1] (lane shuffles; warp-boundary lanes through shared memory).
    // 9 instructions of max(32 + 3*OC/4, 3*OC/2) cycles per tile and plane instead of 27 of max(32 + OC/4, OC/2).
    // Folded planes carry NO x halo (pitchW == W, W in {32, 64, 128}): a tile is 128 consecutive slots = 128 / W whole rows.
    int fold;
    int wshift;                  // log2(W) (folded mode)
    int ncol;                    // TMEM columns per tile: OC, or 3*OC when folded
    int dbg;                     // B200_ROWF_DBG (timing experiments only, results are WRONG): 1 = epilogue skips all work, 2 = no shifted sum
};

struct alignas(128) RowFwdBarriers {
    uint64_t pfull[kRfRing], pempty[kRfRing];
    uint64_t wfull[kRfMaxW], wempty[kRfMaxW];
    uint64_t afull[2], aempty[2];
    uint32_t tmem_base;
};

struct RfItem { int n, y0, rows, z0, z1; };
__device__ __forceinline__ RfItem rf_decode(const RowFwdParams& p, int item) {
    RfItem c;
    const int yb = item % p.yblocks; item /= p.yblocks;        // adjacent CTAs take adjacent row blocks: halo rows hit L2
    const int zg = item % p.zsegs;
    c.n = item / p.zsegs;
    c.y0 = yb * p.YB;
    c.rows = min(p.YB, p.H - c.y0);
    c.z0 = zg * p.zs;
    c.z1 = min(p.D, c.z0 + p.zs);
    return c;
}
__device__ __forceinline__ int rf_tiles(const RowFwdParams& p, int rows) {
    // tiles needed to cover output slots [0, (rows-1)*pitchW + W)
    if (p.tpr < (1 << 20)) return rows * p.tpr;
    return ((rows - 1) * p.pitchW + p.W + 127) >> 7;
}

__device__ __forceinline__ void rf_store16(__nv_bfloat16* dst, const float (&v)[16]) {
    uint4 lo, hi;
    __nv_bfloat162* l2 = reinterpret_cast<__nv_bfloat162*>(&lo);
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&hi);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        l2[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        h2[i] = __floats2bfloat162_rn(v[8 + 2 * i], v[8 + 2 * i + 1]);
    }
    *reinterpret_cast<uint4*>(dst) = lo;
    *reinterpret_cast<uint4*>(dst + 8) = hi;
}

template <int KD, int KHW, int NKS, bool FOLD = false>
__global__ void __launch_bounds__(kRfThreads, 1) row_fwd_kernel(const __grid_constant__ CUtensorMap in_map, const RowFwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ RowFwdBarriers bars;
    // folded epilogue: [epilogue set][phase][warp][P_0 row of lane 31 | P_2 row of lane 0][channel] + a row of zeros (the x halo)
    __shared__ __align__(16) float xch[FOLD ? 2 : 1][2][4][2][16];
    __shared__ __align__(16) float xzero[16];
    if (threadIdx.x < 16) xzero[threadIdx.x] = 0.f;
    uint8_t* planes = smem;
    uint8_t* wsm = smem + (size_t)p.ring * p.plane_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int pd = KD / 2, ph = KHW / 2;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kRfRing; ++i) { ptx::mbar_init(ptx::smem_u32(&bars.pfull[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars.pempty[i]), 1); }
        for (int i = 0; i < kRfMaxW; ++i) { ptx::mbar_init(ptx::smem_u32(&bars.wfull[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars.wempty[i]), 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(ptx::smem_u32(&bars.afull[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars.aempty[i]), 8); }
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&in_map);
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
        // ===================================================== TMA producer: resident weights once, then the plane ring
        if (lane == 0) {
            if (!p.stream_w) {
                const uint32_t wfull = ptx::smem_u32(&bars.wfull[0]);
                ptx::mbar_expect_tx(wfull, (uint32_t)p.w_bytes);
                for (int off = 0; off < p.w_bytes; off += 32768) {
                    const int n = min(32768, p.w_bytes - off);
                    ptx::bulk_load(ptx::smem_u32(wsm + off), reinterpret_cast<const uint8_t*>(p.w) + off, (uint32_t)n, wfull);
                }
            }
            const uint32_t ring = (uint32_t)p.ring;
            uint32_t cnt = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                const RfItem c = rf_decode(p, item);
                const int first = max(c.z0 - pd, 0), last = min(c.z1 - 1 + pd, p.D - 1);
                for (int pl = first; pl <= last; ++pl, ++cnt) {
                    const uint32_t s = cnt % ring, phs = (cnt / ring) & 1;
                    ptx::mbar_wait(ptx::smem_u32(&bars.pempty[s]), phs ^ 1);
                    const uint32_t full = ptx::smem_u32(&bars.pfull[s]);
                    ptx::mbar_expect_tx(full, (uint32_t)p.plane_tx);
                    ptx::tma_load_4d(ptx::smem_u32(planes + (size_t)s * p.plane_bytes), &in_map, full, 0, FOLD ? 0 : -ph, c.y0 - ph, c.n * p.D + pl);
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (whole warp converged, one elected lane issues)
        const uint32_t vox16 = ((uint32_t)p.IC * 2) >> 4;                       // 16-byte units per voxel
        const uint32_t a_hi = ((8 * (uint32_t)p.IC * 2) >> 4) | (1u << 14) | (row_layout_bits(p.IC * 2) << 29);   // SBO = 8 voxels
        const uint32_t b_hi = (128u >> 4) | (1u << 14);                          // SBO = next 8 output channels
        const uint32_t b_lbo = ((uint32_t)p.OC) << 16;                           // LBO = next 8 input channels = OC*16 B
        const uint32_t pl16 = ptx::smem_u32(planes) >> 4, plane16 = (uint32_t)p.plane_bytes >> 4;
        const uint32_t w16 = ptx::smem_u32(wsm) >> 4;
        const uint32_t row16 = (uint32_t)p.pitchW * vox16;
        constexpr int nks = NKS;
        const uint32_t wtap16 = ((uint32_t)p.IC * p.OC * 2) >> 4, wks16 = (2 * (uint32_t)p.OC * 16) >> 4;
        const uint32_t idesc = p.idesc;
        const bool leader = ptx::elect_one();                                   // one fixed lane issues every tcgen05 instruction of the CTA
        uint32_t cnt = 0, group = 0;
        if (!p.stream_w) {
        const uint32_t rring = (uint32_t)p.ring;
        ptx::mbar_wait(ptx::smem_u32(&bars.wfull[0]), 0);
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const RfItem c = rf_decode(p, item);
            const int first = max(c.z0 - pd, 0), last = min(c.z1 - 1 + pd, p.D - 1);
            const int ntiles = rf_tiles(p, c.rows);
            int ready = first - 1;                                               // highest plane known to be resident
            for (int z = c.z0; z < c.z1; ++z, ++group) {
                const int need = min(z + pd, p.D - 1);
                for (; ready < need; ++ready) {
                    const uint32_t i = cnt + (uint32_t)(ready + 1 - first);
                    ptx::mbar_wait(ptx::smem_u32(&bars.pfull[i % rring]), (i / rring) & 1);
                }
                const uint32_t set = group & 1;
                ptx::mbar_wait(ptx::smem_u32(&bars.aempty[set]), ((group >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                // ring slot of plane z+kz-pd is (cnt + z-pd-first + kz) % ring.  The modulo is taken ONCE per group (first tap plane that
                // exists) and advanced by compare-and-wrap; tile offsets and accumulator columns advance by addition: the issuing warp
                // is a single dependent instruction stream, and integer divisions / modulos per tile showed up as idle tensor cycles
                // once the kx-folded MMAs were no longer epilogue-bound.
                const int j0 = z - pd - first;                                      // >= -pd
                const int kzf = j0 < 0 ? -j0 : 0;                                   // first tap plane inside the volume
                const uint32_t slot0 = (cnt + (uint32_t)(j0 + kzf)) % rring;
                uint32_t hasmask = 0;
#pragma unroll
                for (int kz = 0; kz < KD; ++kz) { const int pl = z + kz - pd; if (pl >= 0 && pl < p.D) hasmask |= 1u << kz; }
                {
                    // The issue rate of this warp bounds the tensor pipe for N = 16/32 (39-40 cycles per MMA): all 32 lanes run
                    // the loop with warp-uniform values -- descriptors live in uniform registers, no R2UR per MMA -- and only the
                    // tcgen05 instructions are predicated on the elected lane.  The body stays small enough for the instruction
                    // cache (a fully unrolled 27-tap body stalled on instruction fetch).
                    uint32_t off[KHW * KHW];
#pragma unroll
                    for (int j = 0; j < KHW * KHW; ++j) off[j] = (uint32_t)(j / KHW) * row16 + (uint32_t)(j % KHW) * vox16;
                    const uint32_t a_flag = 1u << 16;
                    const uint32_t colstep16 = 128u * vox16, rowstep16 = (uint32_t)p.rowstride * vox16;
                    uint32_t d_tmem = tmem_base + (uint32_t)(set * p.T * p.ncol);
                    uint32_t tile16 = 0, trow16 = 0;
                    int tcol = 0;
                    for (int t = 0; t < ntiles; ++t) {
                        uint32_t acc = 0;
                        uint32_t bb = w16 | (FOLD ? 3u * b_lbo : b_lbo);
                        uint32_t slot = slot0;
#pragma unroll 1
                        for (int kz = 0; kz < KD; ++kz) {
                            if (!((hasmask >> kz) & 1)) { bb += (uint32_t)(KHW * KHW) * wtap16; continue; }
                            const uint32_t a0 = (pl16 + slot * plane16 + tile16) | a_flag;
                            slot = slot + 1 == rring ? 0 : slot + 1;
                            if (FOLD) {
                                // B block of (kz, ky): [ci group][kx][oc][8 ci] -> LBO = 3*OC*16 B, next 16 ci = 2 groups further
#pragma unroll
                                for (int ky = 0; ky < KHW; ++ky) {
                                    uint32_t a = a0 + (uint32_t)ky * row16, b2 = bb;             // unshifted rows (the image has no x halo)
#pragma unroll
                                    for (int ks = 0; ks < nks; ++ks) {
                                        if (leader) ptx::umma_bf16_lohi(d_tmem, a, a_hi, b2, b_hi, idesc, acc);
                                        a += 2; b2 += 3u * wks16; acc = 1;
                                    }
                                    bb += (uint32_t)KHW * wtap16;
                                }
                                continue;
                            }
#pragma unroll
                            for (int j = 0; j < KHW * KHW; ++j) {
                                uint32_t a = a0 + off[j], b2 = bb;
#pragma unroll
                                for (int ks = 0; ks < nks; ++ks) {
                                    if (leader) ptx::umma_bf16_lohi(d_tmem, a, a_hi, b2, b_hi, idesc, acc);
                                    a += 2; b2 += wks16; acc = 1;
                                }
                                bb += wtap16;
                            }
                        }
                        d_tmem += (uint32_t)p.ncol;
                        if (++tcol == p.tpr) { tcol = 0; trow16 += rowstep16; tile16 = trow16; } else tile16 += colstep16;
                    }
                    if (leader) {
                        ptx::umma_commit(ptx::smem_u32(&bars.afull[set]));
                        // plane z-pd is not needed by z+1; at the end of the segment release everything that is left
                        const int lo = z - pd, hi = (z + 1 == c.z1) ? last : lo;
                        for (int pl = max(lo, first); pl <= hi; ++pl) {
                            const uint32_t i = cnt + (uint32_t)(pl - first);
                            ptx::umma_commit(ptx::smem_u32(&bars.pempty[i % rring]));
                        }
                    }
                }
                __syncwarp();
            }
            cnt += (uint32_t)(last - first + 1);
        }
        } else {
        // ---------------------------------------------- streaming mode: tap-outer, tile-inner; one weight stage per tap
        const uint32_t ring = (uint32_t)p.ring;
        const uint32_t wstage16 = (uint32_t)p.wtap_bytes >> 4;
        uint32_t ws = 0, wph = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const RfItem c = rf_decode(p, item);
            const int first = max(c.z0 - pd, 0), last = min(c.z1 - 1 + pd, p.D - 1);
            const int ntiles = rf_tiles(p, c.rows);
            int ready = first - 1;
            for (int z = c.z0; z < c.z1; ++z, ++group) {
                const int need = min(z + pd, p.D - 1);
                for (; ready < need; ++ready) {
                    const uint32_t i = cnt + (uint32_t)(ready + 1 - first);
                    ptx::mbar_wait(ptx::smem_u32(&bars.pfull[i % ring]), (i / ring) & 1);
                }
                const uint32_t set = group & 1;
                ptx::mbar_wait(ptx::smem_u32(&bars.aempty[set]), ((group >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_set = tmem_base + (uint32_t)(set * p.T * p.ncol);
                uint32_t acc = 0;
                bool released = false;
#pragma unroll 1
                for (int kz = 0; kz < KD; ++kz) {
                    const int pl = z + kz - pd;
                    if (pl < 0 || pl >= p.D) continue;                                  // the weight producer skips the same taps
                    const uint32_t a_pl = (pl16 + ((cnt + (uint32_t)(pl - first)) % ring) * plane16) | (1u << 16);
#pragma unroll 1
                    for (int j = 0; j < (FOLD ? KHW : KHW * KHW); ++j) {              // folded: one stage = the three kx blocks of (kz, ky)
                        ptx::mbar_wait(ptx::smem_u32(&bars.wfull[ws]), wph);
                        ptx::tc_fence_after();
                        {
                            const uint32_t a_tap = FOLD ? a_pl + (uint32_t)j * row16 : a_pl + (uint32_t)(j / KHW) * row16 + (uint32_t)(j % KHW) * vox16;
                            const uint32_t b0 = (w16 + ws * wstage16) | (FOLD ? 3u * b_lbo : b_lbo);
                            uint32_t d = d_set;
                            for (int t = 0; t < ntiles; ++t) {
                                uint32_t a = a_tap + (uint32_t)((t / p.tpr) * p.rowstride + (t % p.tpr) * 128) * vox16, b2 = b0;
#pragma unroll
                                for (int ks = 0; ks < nks; ++ks) {
                                    if (leader) ptx::umma_bf16_lohi(d, a, a_hi, b2, b_hi, idesc, ks == 0 ? acc : 1u);
                                    a += 2; b2 += FOLD ? 3u * wks16 : wks16;
                                }
                                d += (uint32_t)p.ncol;
                            }
                            if (leader) ptx::umma_commit(ptx::smem_u32(&bars.wempty[ws]));           // stage free once these MMAs retire
                        }
                        acc = 1;
                        if (++ws == (uint32_t)p.nw) { ws = 0; wph ^= 1; }
                    }
                    // the oldest plane of the window is only read by the kz = 0 pass: free it NOW so that the producer can refill
                    // the slot during the remaining two thirds of this group (this is what makes a ring of 3 enough)
                    if (KD == 3 && kz == 0 && z + 1 < c.z1 && pl >= first) {
                        if (leader) ptx::umma_commit(ptx::smem_u32(&bars.pempty[(cnt + (uint32_t)(pl - first)) % ring]));
                        released = true;
                    }
                }
                if (leader) {
                    ptx::umma_commit(ptx::smem_u32(&bars.afull[set]));
                    const int lo = released ? z - pd + 1 : z - pd, hi = (z + 1 == c.z1) ? last : z - pd;
                    for (int pl = max(lo, first); pl <= hi; ++pl) ptx::umma_commit(ptx::smem_u32(&bars.pempty[(cnt + (uint32_t)(pl - first)) % ring]));
                }
                __syncwarp();
            }
            cnt += (uint32_t)(last - first + 1);
        }
        }
    } else if (warp == 6) {
        // ===================================================== weight producer (streaming mode): one tap per stage, in MMA order
        if (lane == 0 && p.stream_w) {
            uint32_t s = 0, ph = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                const RfItem c = rf_decode(p, item);
                for (int z = c.z0; z < c.z1; ++z)
                    for (int kz = 0; kz < KD; ++kz) {
                        const int pl = z + kz - pd;
                        if (pl < 0 || pl >= p.D) continue;
                        constexpr int stages = FOLD ? KHW : KHW * KHW;             // wtap_bytes = one stage (three taps when folded)
                        for (int j = 0; j < stages; ++j) {
                            ptx::mbar_wait(ptx::smem_u32(&bars.wempty[s]), ph ^ 1);
                            const uint32_t full = ptx::smem_u32(&bars.wfull[s]);
                            ptx::mbar_expect_tx(full, (uint32_t)p.wtap_bytes);
                            ptx::bulk_load(ptx::smem_u32(wsm + (size_t)s * p.wtap_bytes),
                                           reinterpret_cast<const uint8_t*>(p.w) + (size_t)(kz * stages + j) * p.wtap_bytes, (uint32_t)p.wtap_bytes, full);
                            if (++s == (uint32_t)p.nw) { s = 0; ph ^= 1; }
                        }
                    }
            }
        }
    } else {
        // ===================================================== epilogue: TMEM -> (+bias) -> bf16 -> global (+ BN statistics)
        // TWO sets of four warps (set 0 = warps 2..5, set 1 = warps 7..10; a warp reads the TMEM lane quarter warp % 4, so each set
        // covers the 128 rows of a tile).  The sets take alternate tiles of every accumulator group: the epilogue is a long
        // dependent chain (TMEM load -> exchange -> convert -> store), so one warp per scheduler could not hide its latencies --
        // with the kx-folded MMAs the epilogue, not the tensor pipe, was the bound (ncu: tensor pipe 45 % busy, 71 % no eligible warp).
        const int eset = warp >= 7 ? 1 : 0;
        const int lane_grp = warp & 3;
        const int m = lane_grp * 32 + lane;
        const bool want_stats = p.stats != nullptr;
        constexpr int kStatCh = FOLD ? 2 : 4;                    // statistics / fused stores: up to 4 x 16 output channels (2 x 16 when folded)
        float ssum[16 * kStatCh], ssq[16 * kStatCh];             // per-thread partial statistics
#pragma unroll
        for (int i = 0; i < 16 * kStatCh; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
        const int ocs = p.OC - p.store_c0;                       // stored channels per voxel
        // folded with 64 output channels (N = 192, one tile per accumulator set): the two epilogue sets split the CHANNELS of every
        // tile (set s takes [32 s, 32 s + 32)) instead of alternating tiles, so each thread still carries 32 channels of statistics
        const bool csplit = FOLD && p.OC > 32;
        const int cbase = csplit ? eset * 32 : 0;
        const bool lzero = ((lane_grp * 32) & (p.W - 1)) == 0;          // folded: this warp's lane 0 is x == 0 / its lane 31 is x == W-1
        const bool rzero = (((lane_grp + 1) * 32) & (p.W - 1)) == 0;
        uint32_t group = 0, xphase = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const RfItem c = rf_decode(p, item);
            const int ntiles = rf_tiles(p, c.rows);
            for (int z = c.z0; z < c.z1; ++z, ++group) {
                const uint32_t set = group & 1;
                ptx::mbar_wait(ptx::smem_u32(&bars.afull[set]), (group >> 1) & 1);
                ptx::tc_fence_after();
                const int64_t plane_vox = ((int64_t)c.n * p.D + z) * p.H;
                for (int t = csplit ? 0 : (eset ^ (int)(group & 1)); t < ntiles; t += csplit ? 1 : 2) {   // (alternating start: one-tile groups still use both sets)
                    if (p.dbg == 1) continue;
                    int row, x;
                    if (FOLD) { const int slot = t * 128 + m; row = slot >> p.wshift; x = slot & (p.W - 1); }   // 128 / W whole rows per tile
                    else {
                        const int slot = (t / p.tpr) * p.rowstride + (t % p.tpr) * 128 + m;
                        row = slot / p.pitchW; x = slot - row * p.pitchW;
                    }
                    const bool valid = x < p.W && row < c.rows;
                    __nv_bfloat16* dst = p.out + ((plane_vox + c.y0 + row) * p.W + x) * ocs - p.store_c0;
                    const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)((set * p.T + t) * p.ncol);
                    if (FOLD) {
                        // out[x] = P_0[x-1] + P_1[x] + P_2[x+1]; the tile is 128 / W whole x-rows whose ends fall on warp boundaries
                        // (W = 32, 64, 128), where the neighbour is the zero halo (lzero / rzero).  Neighbour lanes by a rotating shuffle; the value that crosses a warp boundary travels
                        // through shared memory: lane 31 publishes its P_0 row and then REPLACES it by the previous warp's lane 31
                        // (lane 0 likewise for P_2), so the rotation delivers every lane its neighbour without a per-value select.
#pragma unroll
                        for (int ch = 0; ch < kStatCh; ++ch) {
                            const int c0 = cbase + ch * 16;                         // first channel of this pass
                            if (c0 < p.OC) {
                                float v[16], l[16], r[16];
                                ptx::tmem_ld16x3(taddr + (uint32_t)c0, taddr + (uint32_t)(p.OC + c0), taddr + (uint32_t)(2 * p.OC + c0), l, v, r);
                                float (&buf)[4][2][16] = xch[eset][xphase & 1];
                                if (p.dbg != 2) {
                                if (lane == 31) {
                                    float4* d4 = reinterpret_cast<float4*>(&buf[lane_grp][0][0]);
#pragma unroll
                                    for (int q = 0; q < 4; ++q) d4[q] = make_float4(l[4 * q], l[4 * q + 1], l[4 * q + 2], l[4 * q + 3]);
                                }
                                if (lane == 0) {
                                    float4* d4 = reinterpret_cast<float4*>(&buf[lane_grp][1][0]);
#pragma unroll
                                    for (int q = 0; q < 4; ++q) d4[q] = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
                                }
                                if (eset == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
                                else asm volatile("bar.sync 2, 128;" ::: "memory");
                                if (lane == 31) {
                                    const float4* s4 = reinterpret_cast<const float4*>(!lzero ? &buf[lane_grp - 1][0][0] : &xzero[0]);
#pragma unroll
                                    for (int q = 0; q < 4; ++q) { const float4 a4 = s4[q]; l[4 * q] = a4.x; l[4 * q + 1] = a4.y; l[4 * q + 2] = a4.z; l[4 * q + 3] = a4.w; }
                                }
                                if (lane == 0) {
                                    const float4* s4 = reinterpret_cast<const float4*>(!rzero ? &buf[lane_grp + 1][1][0] : &xzero[0]);
#pragma unroll
                                    for (int q = 0; q < 4; ++q) { const float4 a4 = s4[q]; r[4 * q] = a4.x; r[4 * q + 1] = a4.y; r[4 * q + 2] = a4.z; r[4 * q + 3] = a4.w; }
                                }
                                const int from_l = (lane + 31) & 31, from_r = (lane + 1) & 31;
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    v[i] += __shfl_sync(0xffffffffu, l[i], from_l) + __shfl_sync(0xffffffffu, r[i], from_r);
                                } else {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) v[i] += l[i] + r[i];
                                }
                                ++xphase;
                                if (p.bias != nullptr) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + c0 + i);
                                }
                                if (valid) {
                                    if (c0 >= p.store_c0) rf_store16(dst + c0, v);
                                    if (want_stats) {
#pragma unroll
                                        for (int i = 0; i < 16; ++i) { ssum[ch * 16 + i] += v[i]; ssq[ch * 16 + i] = fmaf(v[i], v[i], ssq[ch * 16 + i]); }
                                    }
                                }
                            }
                        }
                    } else if (p.OC <= 16 * kStatCh) {
#pragma unroll
                        for (int ch = 0; ch < kStatCh; ++ch) {
                            if (ch * 16 < p.OC) {
                                float v[16];
                                ptx::tmem_ld16(taddr + (uint32_t)(ch * 16), v);
                                if (p.bias != nullptr) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + ch * 16 + i);
                                }
                                if (valid) {
                                    if (ch * 16 >= p.store_c0) rf_store16(dst + ch * 16, v);
                                    if (want_stats) {
#pragma unroll
                                        for (int i = 0; i < 16; ++i) { ssum[ch * 16 + i] += v[i]; ssq[ch * 16 + i] = fmaf(v[i], v[i], ssq[ch * 16 + i]); }
                                    }
                                }
                            }
                        }
                    } else {
                        for (int c0 = 0; c0 < p.OC; c0 += 16) {
                            float v[16];
                            ptx::tmem_ld16(taddr + (uint32_t)c0, v);
                            if (p.bias != nullptr) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + c0 + i);
                            }
                            if (valid && c0 >= p.store_c0) rf_store16(dst + c0, v);
                        }
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars.aempty[set]));
            }
        }
        if (want_stats) {
            // per-CTA partial: lanes -> warp (shuffle), 8 warps -> CTA in a fixed order (deterministic); summed over CTAs by
            // norm_stats_finalize_kernel in double.  The scratch aliases the plane ring: every MMA that reads it has retired (the
            // last accumulator group was waited for above) and no TMA load is in flight.
            float* red = reinterpret_cast<float*>(smem);                 // [8 warps][2][16 * kStatCh]
            const int w8 = eset * 4 + lane_grp;
#pragma unroll
            for (int i = 0; i < 16 * kStatCh; ++i) {
                if (cbase + i < p.OC) {
                    float a = ssum[i], b = ssq[i];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
                    if (lane == 0) { red[(w8 * 2 + 0) * 16 * kStatCh + i] = a; red[(w8 * 2 + 1) * 16 * kStatCh + i] = b; }
                }
            }
            asm volatile("bar.sync 3, 256;" ::: "memory");
            if (eset == 0 && m < 2 * p.OC) {
                const int which = m / p.OC, ch = m - which * p.OC;
                float acc = 0.f;
                if (csplit) {                                       // channel ch was accumulated by the four warps of set ch / 32
#pragma unroll
                    for (int w = 0; w < 4; ++w) acc += red[(((ch >> 5) * 4 + w) * 2 + which) * 16 * kStatCh + (ch & 31)];
                } else {
#pragma unroll
                    for (int w = 0; w < 8; ++w) acc += red[(w * 2 + which) * 16 * kStatCh + ch];
                }
                p.stats[((size_t)blockIdx.x * 2 + which) * p.OC + ch] = acc;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
}

// ------------------------------------------------------------------------------------------------ host side
struct RowFwdGeom { int IC, OC, D, H, W, kd, khw; };

inline bool row_fwd_geom(const b200_conv_desc* d, int pass, RowFwdGeom* g) {
    if (!d->allow_umma || d->transposed) return false;
    if (pass != B200_PASS_FWD && pass != B200_PASS_DGRAD) return false;
    if (d->x_dtype != B200_BF16 || d->y_dtype != B200_BF16) return false;
    if (d->sd != 1 || d->sh != 1 || d->sw != 1 || d->dd != 1 || d->dh != 1 || d->dw != 1) return false;
    if (!((d->kd == 1 || d->kd == 3) && (d->kh == 1 || d->kh == 3) && d->kw == d->kh)) return false;
    if (d->pd != d->kd / 2 || d->ph != d->kh / 2 || d->pw != d->kw / 2) return false;       // "same": dgrad has the same geometry
    g->IC = pass == B200_PASS_FWD ? d->Ci : d->Co;
    g->OC = pass == B200_PASS_FWD ? d->Co : d->Ci;
    g->D = d->Di; g->H = d->Hi; g->W = d->Wi; g->kd = d->kd; g->khw = d->kh;
    if (!(g->IC == 16 || g->IC == 32 || g->IC == 64)) return false;                          // one swizzle atom per voxel
    if (g->OC % 16 || g->OC > 256) return false;
    if (g->W + 2 * (g->khw / 2) > 256 || g->W < 8) return false;                             // TMA box limit
    if ((int64_t)d->N * g->D * g->H * g->W < 4096) return false;                             // tiny problems: launch-bound either way
    return true;
}

// kx-folded mode (see RowFwdParams::fold): 128 / W whole x-rows per tile, small Cout, resident or streamed weights.  B200_ROWF_FOLD=0 disables.
inline bool row_fwd_fold_geom(const RowFwdGeom& g) {
    // Measured on B200, 4 x 128^3 forward.  Round 1 (one epilogue warp set): 32->32 0.684 -> 0.512 ms, 16->16 0.378 -> 0.338 ms, but
    // 16->32 0.395 -> 0.500 ms (two 16-channel epilogue passes per tile against only nine K = 16 instructions), so 16->32 stayed
    // unfolded.  Round 2 (two epilogue sets, cheaper shifted sum): 32->32 0.41 ms, 16->16 0.23 ms and 16->32 0.366 against 0.390 ms
    // unfolded -- folding now pays whenever the geometry allows it.
    // B200_ROWF_FOLD: 0 = never, 1 (default) = whenever the geometry allows.
    static const int mode = [] { const char* e = getenv("B200_ROWF_FOLD"); return e == nullptr ? 1 : atoi(e); }();
    // Cout = 64 (N = 192 = the full MMA rate, ONE tile per accumulator set) only with resident weights (Cin <= 32): streamed, its
    // weights would pass once per 128 voxels.  B200_ROWF_FOLD=2 keeps Cout = 64 unfolded.
    const bool oc_ok = g.OC <= 32 || (g.OC == 64 && g.IC <= 32 && mode != 2);
    return mode != 0 && g.khw == 3 && (g.W == 128 || g.W == 64 || g.W == 32) && oc_ok;
}
inline int row_fwd_plan_mode(const RowFwdGeom& g, int N, RowFwdParams* p, size_t* smem_bytes, bool fold);
// the folded plan when the geometry allows it and it fits, the plain one otherwise
inline int row_fwd_plan(const RowFwdGeom& g, int N, RowFwdParams* p, size_t* smem_bytes) {
    if (row_fwd_fold_geom(g)) {
        const std::string saved = err_slot();
        if (row_fwd_plan_mode(g, N, p, smem_bytes, true) == 0) return 0;
        err_slot() = saved;
    }
    return row_fwd_plan_mode(g, N, p, smem_bytes, false);
}
inline bool row_fwd_folds(const b200_conv_desc* d, int pass) {          // decides the packed-weight layout: must follow the plan
    RowFwdGeom g;
    if (!row_fwd_geom(d, pass, &g)) return false;
    RowFwdParams p; size_t smem;
    const std::string saved = err_slot();
    const bool ok = row_fwd_plan(g, d->N, &p, &smem) == 0;
    err_slot() = saved;
    return ok && p.fold != 0;
}

inline int row_fwd_plan_mode(const RowFwdGeom& g, int N, RowFwdParams* p, size_t* smem_bytes, bool fold) {
    memset(p, 0, sizeof *p);
    p->N = N; p->D = g.D; p->H = g.H; p->W = g.W; p->IC = g.IC; p->OC = g.OC; p->kd = g.kd; p->khw = g.khw;
    const int ph = g.khw / 2, pd = g.kd / 2;
    p->pitchW = fold ? g.W : g.W + 2 * ph;
    p->wshift = g.W == 128 ? 7 : g.W == 64 ? 6 : 5;
    const bool per_row = !fold && (g.W % 128) == 0;
    p->tpr = per_row ? g.W / 128 : (1 << 30);
    p->rowstride = per_row ? p->pitchW : 0;
    p->w_bytes = g.kd * g.khw * g.khw * g.IC * g.OC * 2;
    p->wtap_bytes = (fold ? 3 : 1) * g.IC * g.OC * 2;                     // one streamed stage
    // weights: RESIDENT (every tap in shared memory next to a 4-deep plane ring) or STREAMED tap by tap through a ring (the
    // plane ring then shrinks to 3: the kz = 0 plane is released after its pass, see the MMA loop).  Both modes are costed.
    p->fold = fold ? 1 : 0;
    p->ncol = p->fold ? 3 * g.OC : g.OC;
    int maxT = 256 / p->ncol;
    if (maxT > kRfMaxTiles) maxT = kRfMaxTiles;
    if (maxT < 1) maxT = 1;
    // Search (mode, weight stages, row-block height, z-segment length) with a small cost model of one persistent CTA:
    //   MMA cycles per group  = tiles * taps * ksteps * cycles(N)         (probe: max(48, 32 + N/4, N/2) in situ)
    //   weight cycles / group = bytes of all taps / min(10, bytes in flight / 5000 cycles) B/clk/SM      (streaming mode)
    //   item = zs groups + the z-halo planes it has to load first; total = rounds of the persistent grid * item
    const int taps = g.kd * g.khw * g.khw, nks = g.IC / 16;
    double mma_cyc = 32.0 + p->ncol / 4.0;
    if (mma_cyc < 48.0) mma_cyc = 48.0;
    if (mma_cyc < p->ncol / 2.0) mma_cyc = p->ncol / 2.0;
    int best = 0, best_zs = 1, best_nw = 1, best_stream = 0, best_ring = 4; double best_cost = 1e30;
    static const int ring_cap = [] { const char* e = getenv("B200_ROWF_RING"); const int v = e == nullptr ? kRfRing : atoi(e); return v < 4 ? 4 : (v > kRfRing ? kRfRing : v); }();
    const int stages = fold ? taps / 3 : taps;
    for (int stream_w = 0; stream_w <= 1; ++stream_w) {
        const int nw_lo = stream_w ? 2 : 1, nw_hi = stream_w ? kRfMaxW : 1;
        for (int nw = nw_lo; nw <= nw_hi; ++nw) {
            const size_t wsm_bytes = stream_w ? (size_t)nw * p->wtap_bytes : (size_t)p->w_bytes;
            if (wsm_bytes + 1024 >= (size_t)kRfMaxDynSmem) break;
            const size_t budget = (size_t)kRfMaxDynSmem - 1024 - wsm_bytes;
            for (int YB = 1; YB <= 32 && YB <= g.H; ++YB) {
                const int T = per_row ? YB * p->tpr : ((YB - 1) * p->pitchW + g.W + 127) / 128;
                if (T > maxT) break;
                const size_t pb = (((size_t)(YB + 2 * ph) * p->pitchW * g.IC * 2) + 1023) & ~(size_t)1023;
                const size_t reach = ((size_t)T * 128 + (size_t)(g.khw - 1) * (p->pitchW + 1)) * g.IC * 2;
                const size_t tail = reach > pb ? reach - pb : 0;
                // plane ring: 2*pd + 1 planes are live while a group runs, every further slot is one plane of TMA lookahead.  With a
                // single spare slot the load of plane z+2 has exactly one group time to arrive, and a TMA round trip under load
                // (~2 k cycles + transfer) is LONGER than the MMAs of a kx-folded group: measured 32->32 @128^3 with the epilogue
                // switched off, 0.367 ms against 0.23 ms of MMA time.  Resident mode takes the deepest ring that fits (<= 8);
                // streaming mode keeps 3 (it releases the kz = 0 plane early, see the MMA loop).
                int ring = stream_w ? 3 : ring_cap;
                while (ring > (stream_w ? 3 : 4) && (size_t)ring * pb + (tail > wsm_bytes ? tail - wsm_bytes : 0) > budget) --ring;
                if ((size_t)ring * pb + (tail > wsm_bytes ? tail - wsm_bytes : 0) > budget) continue;
                const int yblocks = (g.H + YB - 1) / YB;
                const double g_mma = (double)T * (p->fold ? taps / 3 : taps) * nks * mma_cyc;
                // measured (64->64 @ 64^3): a bulk copy of a tap takes ~5k cycles under load, so the ring depth bounds the rate
                double w_bw = (double)nw * p->wtap_bytes / 5000.0;
                if (w_bw > 10.0) w_bw = 10.0;
                const double g_w = stream_w ? (double)p->w_bytes / w_bw : 0.0;
                // streaming pays a full-barrier wait + commit per tap and cannot run ahead of the weight ring: measured ~1.5x the
                // MMA time of the resident mode on the layers where both fit
                double group = stream_w ? 1.5 * (g_mma > g_w ? g_mma : g_w) + 150.0 * stages + 400.0 : g_mma + 400.0;
                const double plane_cyc = (double)pb / 20.0;
                const int lookahead = stream_w ? 1 : ring - (2 * pd + 1);
                const double feed = (2000.0 + (double)pb / 16.0) / (lookahead > 0 ? lookahead : 1);      // one plane per group must arrive
                if (!stream_w && group < feed) group = feed;
                for (int zs = g.D; zs >= 1; zs = (zs > 4 ? (zs + 1) / 2 : zs - 1)) {
                    if (g.kd == 1 && zs != 1) continue;
                    const int zsegs = (g.D + zs - 1) / zs;
                    const int64_t items = (int64_t)N * yblocks * zsegs;
                    const double rounds = (double)((items + kNumSMs - 1) / kNumSMs);
                    // resident weights are loaded once per CTA (w_bytes / 10 B/clk)
                    const double cost = rounds * (zs * group + 2 * pd * plane_cyc + 1500.0) + (stream_w ? 0.0 : (double)p->w_bytes / 10.0);
                    if (cost < best_cost) { best_cost = cost; best = YB; best_zs = zs; best_nw = nw; best_stream = stream_w; best_ring = ring; }
                }
            }
        }
    }
    p->stream_w = best_stream;
    p->ring = best_stream ? 3 : best_ring;
    p->nw = best_nw;
    B200_REQUIRE(best >= 1, "row fwd: a row block does not fit shared memory / TMEM");
    p->YB = best;
    p->T = per_row ? best * p->tpr : ((best - 1) * p->pitchW + g.W + 127) / 128;
    p->plane_tx = (best + 2 * ph) * p->pitchW * g.IC * 2;
    p->plane_bytes = (p->plane_tx + 1023) & ~1023;
    p->yblocks = (g.H + best - 1) / best;
    p->zs = best_zs;
    p->zsegs = (g.D + p->zs - 1) / p->zs;
    const int64_t items = (int64_t)N * p->yblocks * p->zsegs;
    B200_REQUIRE(items < (1ll << 31), "row fwd: too many items");
    p->items = (int)items;
    int cols = 2 * p->T * p->ncol, pow2 = 32;
    while (pow2 < cols) pow2 <<= 1;
    B200_REQUIRE(pow2 <= 512, "row fwd: accumulators do not fit TMEM");
    p->tmem_cols = pow2;
    p->idesc = make_idesc_bf16(p->ncol);
    // The MMAs of the last (partial) tile read slots past the end of a plane image: up to T*128 + (k-1)*(pitchW + 1) slots from
    // the plane start.  Those rows of D are never stored, but the reads must stay inside the allocation: whatever follows the
    // last ring slot (resident weights / weight ring) is padded up to that tail.  + 1 KB manual 1024-byte alignment.
    const size_t wsm_bytes = p->stream_w ? (size_t)p->nw * p->wtap_bytes : (size_t)p->w_bytes;
    const size_t reach = ((size_t)p->T * 128 + (size_t)(g.khw - 1) * (p->pitchW + 1)) * g.IC * 2;
    const size_t tail = reach > (size_t)p->plane_bytes ? reach - p->plane_bytes : 0;
    *smem_bytes = (size_t)p->ring * p->plane_bytes + (wsm_bytes > tail ? wsm_bytes : tail) + 1024;
    B200_REQUIRE(*smem_bytes <= (size_t)kRfMaxDynSmem, "row fwd: plan exceeds shared memory");
    return 0;
}

inline bool row_fwd_supported(const b200_conv_desc* d, int pass) {
    RowFwdGeom g;
    if (!row_fwd_geom(d, pass, &g)) return false;
    RowFwdParams p; size_t smem;
    const std::string saved = err_slot();
    const bool ok = row_fwd_plan(g, d->N, &p, &smem) == 0;
    err_slot() = saved;
    return ok;
}

template <int KD, int KHW, int NKS, bool FOLD = false>
inline int row_fwd_launch(const CUtensorMap& map, const RowFwdParams& p, size_t smem_bytes, void* stream) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(row_fwd_kernel<KD, KHW, NKS, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRfMaxDynSmem); });
    B200_REQUIRE(attr_err == cudaSuccess, "row fwd: cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(attr_err));
    const int grid = p.items < kNumSMs ? p.items : kNumSMs;
    B200_LAUNCH((row_fwd_kernel<KD, KHW, NKS, FOLD>), grid, kRfThreads, smem_bytes, stream, map, p);
    return 0;
}

// per-CTA statistics partials: chunks = CTAs of the launch; 0 when (desc, pass) cannot produce them
inline int row_fwd_stats_chunks(const b200_conv_desc* d) {
    RowFwdGeom g;
    if (!row_fwd_geom(d, B200_PASS_FWD, &g) || g.OC > 64) return 0;
    RowFwdParams p; size_t smem;
    const std::string saved = err_slot();
    const bool ok = row_fwd_plan(g, d->N, &p, &smem) == 0;
    err_slot() = saved;
    if (!ok) return 0;
    return p.items < kNumSMs ? p.items : kNumSMs;
}

// `stats` (optional): fp32 [chunks][2][OC] per-CTA sum / sum of squares of the fp32 outputs (row_fwd_stats_chunks(d) > 0)
inline int row_fwd_run(const b200_conv_desc* d, int pass, const void* in, const void* w_packed, const float* bias, void* out, float* stats,
                       void* stream, int store_c0 = 0) {
    RowFwdGeom g;
    B200_REQUIRE(row_fwd_geom(d, pass, &g), "row fwd: unsupported descriptor");
    B200_REQUIRE(aligned16(in) && aligned16(out) && aligned16(w_packed), "row fwd: pointers must be 16-byte aligned");
    B200_REQUIRE(stats == nullptr || g.OC <= 64, "row fwd: fused statistics need Cout <= 64");
    B200_REQUIRE(store_c0 >= 0 && store_c0 < g.OC && store_c0 % 16 == 0 && (store_c0 == 0 || g.OC <= 64), "row fwd: bad first stored channel %d", store_c0);
    RowFwdParams p;
    size_t smem_bytes = 0;
    if (row_fwd_plan(g, d->N, &p, &smem_bytes)) return 1;
    p.w = (const __nv_bfloat16*)w_packed; p.bias = bias; p.out = (__nv_bfloat16*)out; p.stats = stats; p.store_c0 = store_c0;
    static const int dbg = [] { const char* e = getenv("B200_ROWF_DBG"); return e == nullptr ? 0 : atoi(e); }();
    p.dbg = dbg;
    static const bool debug = [] { const char* e = getenv("B200_ROWF_DEBUG"); return e != nullptr && e[0] == '1'; }();
    if (debug)
        fprintf(stderr, "[row_fwd] N=%d %dx%dx%d IC=%d OC=%d k=%d,%d: YB=%d T=%d zs=%d items=%d ring=%d stream=%d nw=%d plane=%dB smem=%zuB tmem=%d fold=%d\n", d->N,
                g.D, g.H, g.W, g.IC, g.OC, g.kd, g.khw, p.YB, p.T, p.zs, p.items, p.ring, p.stream_w, p.nw, p.plane_bytes, smem_bytes, p.tmem_cols, p.fold);
    const int ph = g.khw / 2;
    CUtensorMap map;
    if (make_row_map(&map, in, g.IC, g.W, g.H, (int64_t)d->N * g.D, g.IC, p.pitchW, p.YB + 2 * ph)) return 1;
    if (p.fold) {
        if (g.kd == 3) return g.IC == 16 ? row_fwd_launch<3, 3, 1, true>(map, p, smem_bytes, stream)
                            : g.IC == 32 ? row_fwd_launch<3, 3, 2, true>(map, p, smem_bytes, stream) : row_fwd_launch<3, 3, 4, true>(map, p, smem_bytes, stream);
        return g.IC == 16 ? row_fwd_launch<1, 3, 1, true>(map, p, smem_bytes, stream)
             : g.IC == 32 ? row_fwd_launch<1, 3, 2, true>(map, p, smem_bytes, stream) : row_fwd_launch<1, 3, 4, true>(map, p, smem_bytes, stream);
    }
#define B200_RF_CASE(KD_, KHW_)                                                                              \
    if (g.kd == KD_ && g.khw == KHW_) {                                                                      \
        if (g.IC == 16) return row_fwd_launch<KD_, KHW_, 1>(map, p, smem_bytes, stream);                     \
        if (g.IC == 32) return row_fwd_launch<KD_, KHW_, 2>(map, p, smem_bytes, stream);                     \
        return row_fwd_launch<KD_, KHW_, 4>(map, p, smem_bytes, stream);                                     \
    }
    B200_RF_CASE(3, 3)
    B200_RF_CASE(1, 3)
    B200_RF_CASE(3, 1)
    B200_RF_CASE(1, 1)
#undef B200_RF_CASE
    return fail("row fwd: unreachable");
}

}  // namespace b200
