// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 in, fp32 accumulate in TMEM, bf16 out).
//
// Covers the dense contractions of the path: 3x3x3 / 1x1x1 (and 2-D 3x3) stride-1 convolutions with
// channel counts that are multiples of 16, forward and dgrad (dgrad = the same kernel on flipped,
// transposed weights).  Everything else (Cin=1 stems, Cout=2 heads, strided separable convs) is HBM-bound and
// runs on the SIMT kernels (conv_simt.cuh).
//
// GEMM view per output tile:  D[128 voxels x Cout] += A[128 voxels x 16 ch] * B[16 ch x Cout]   (tcgen05.mma, M=128)
//   * output tile  = 1 z-plane x 16 rows (y) x 8 columns (x) = 128 voxels = the 128 TMEM lanes;
//   * A operand    = a HALO slab of the input plane, (16+kh-1) x (8+kw-1) voxels, staged ONCE by TMA in the
//                    channel-group-major layout [cg = ch/8][voxel][8 ch] (16 B per entry).  That is the canonical
//                    K-major no-swizzle UMMA layout: 8 consecutive x-voxels form a 128-byte core matrix, the next
//                    row group (y+1) is SBO = WW*16 bytes away and the next 8 channels LBO = one cg plane away.
//                    A filter tap (ky,kx) is therefore just a different descriptor START address into the same slab:
//                    the 27 taps re-read shared memory, never HBM/L2 (27x less traffic than im2col-by-TMA);
//   * B operand    = packed weights [tap][kchunk][cg][cout][8 ch] streamed by 1-D bulk copies through a ring;
//   * accumulators = P output planes x Cout fp32 columns of TMEM, double-buffered when they fit, so the epilogue of
//                    one tile group overlaps the MMAs of the next; each streamed weight tap is reused for P planes.
// Warp roles (224 threads): 0 = TMA slab producer, 1 = weight producer, 2 = MMA issuer + TMEM owner, 3..6 = epilogue.
#pragma once
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------ PTX wrappers
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must trap (and fail the launch) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}

// TMA: 4-D tiled tensor load, completes `bytes` on the mbarrier
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// 1-D bulk copy global -> shared
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace ptx

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4   [16,30) LBO>>4 (byte step between the two 8-element K groups)   [32,46) SBO>>4 (byte step between
//   8-row groups)   [46,48) version = 1 (Blackwell)   [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
// kind::f16 instruction descriptor: fp32 accumulate, bf16 A and B, both K-major, M = 128, N = n
__host__ __device__ inline uint32_t make_idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ kernel parameters
constexpr int kTileH = 16, kTileW = 8;          // output tile (y, x); 16*8 = 128 = UMMA M
constexpr int kMaxP = 4;                        // output planes per tile group
constexpr int kUmmaThreads = 224;
constexpr int kMaxSlabs = 16, kMaxWStages = 8;

struct UmmaConvParams {
    // "planes" are the merged (n, z) index when kd == 1; per-sample z otherwise
    int NB;                 // independent plane ranges (N when kd == 3, 1 when kd == 1)
    int Dpi, Dpo;           // input / output planes per range
    int Hi, Wi, Ho, Wo;
    int IC, OC;             // gathered (K) and produced (N) channels
    int kd, kh, kw, pd, ph, pw;
    int KC, NKC;            // K chunk (channels) and number of chunks
    int P;                  // output planes per tile group
    int HH, WW;             // halo slab extent
    int cg_pitch;           // bytes between channel groups inside a slab (128-byte multiple)
    int slab_bytes, nslabs;
    int wstage_bytes, nwstages;
    int nsets;              // TMEM accumulator sets (1 or 2)
    int tmem_cols;
    int tiles_y, tiles_x, zchunks;   // per range
    int items;
    uint32_t idesc;
    const __nv_bfloat16* w;          // [tap][kchunk][cg][oc][8]
    const float* bias;               // [OC] or null
    __nv_bfloat16* out;              // [NB*Dpo][Ho][Wo][OC]
};

struct alignas(128) UmmaBarriers {
    uint64_t slab_full[kMaxSlabs], slab_empty[kMaxSlabs];
    uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};
static_assert(sizeof(UmmaBarriers) % 128 == 0, "barrier block must keep the slabs 128-byte aligned");

struct ItemCoord { int nb, z0, pvalid, y0, x0; };
__device__ __forceinline__ ItemCoord decode_item(const UmmaConvParams& p, int item) {
    ItemCoord c;
    const int tx = item % p.tiles_x; item /= p.tiles_x;
    const int ty = item % p.tiles_y; item /= p.tiles_y;
    const int zc = item % p.zchunks;
    c.nb = item / p.zchunks;
    c.z0 = zc * p.P;
    c.pvalid = min(p.P, p.Dpo - c.z0);
    c.y0 = ty * kTileH;
    c.x0 = tx * kTileW;
    return c;
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap in_map, const UmmaConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    UmmaBarriers* bars = reinterpret_cast<UmmaBarriers*>(smem);
    uint8_t* slabs = smem + sizeof(UmmaBarriers);
    uint8_t* wstages = slabs + (size_t)p.nslabs * p.slab_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.nslabs; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->slab_full[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars->slab_empty[i]), 1); }
        for (int i = 0; i < p.nwstages; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->w_full[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars->w_empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->acc_full[i]), 1); ptx::mbar_init(ptx::smem_u32(&bars->acc_empty[i]), 4); }
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&in_map);
    }
    if (warp == 2) {
        ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), (uint32_t)p.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const int nplanes_in = p.P + p.kd - 1;      // input planes touched by one tile group
    const int ntaps = p.kd * p.kh * p.kw;

    if (warp == 0) {
        // ===================================================== TMA producer: halo slabs
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                const ItemCoord c = decode_item(p, item);
                for (int kc = 0; kc < p.NKC; ++kc)
                    for (int q = 0; q < c.pvalid + p.kd - 1; ++q) {
                        const int zi = c.z0 - p.pd + q;
                        if (zi < 0 || zi >= p.Dpi) continue;
                        const uint32_t s = it % p.nslabs, ph = (it / p.nslabs) & 1;
                        ptx::mbar_wait(ptx::smem_u32(&bars->slab_empty[s]), ph ^ 1);
                        const uint32_t full = ptx::smem_u32(&bars->slab_full[s]);
                        const int ncg = p.KC / 8;
                        ptx::mbar_expect_tx(full, (uint32_t)(ncg * p.HH * p.WW * 16));
                        const uint32_t dst = ptx::smem_u32(slabs + (size_t)s * p.slab_bytes);
                        for (int cg = 0; cg < ncg; ++cg)
                            ptx::tma_load_4d(dst + cg * p.cg_pitch, &in_map, full, kc * p.KC + cg * 8, c.x0 - p.pw, c.y0 - p.ph, c.nb * p.Dpi + zi);
                        ++it;
                    }
            }
        }
    } else if (warp == 1) {
        // ===================================================== weight producer: one (tap, kchunk) block per stage
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x)
                for (int kc = 0; kc < p.NKC; ++kc)
                    for (int tap = 0; tap < ntaps; ++tap) {
                        const uint32_t s = it % p.nwstages, ph = (it / p.nwstages) & 1;
                        ptx::mbar_wait(ptx::smem_u32(&bars->w_empty[s]), ph ^ 1);
                        const uint32_t full = ptx::smem_u32(&bars->w_full[s]);
                        ptx::mbar_expect_tx(full, (uint32_t)p.wstage_bytes);
                        ptx::bulk_load(ptx::smem_u32(wstages + (size_t)s * p.wstage_bytes),
                                       reinterpret_cast<const uint8_t*>(p.w) + ((size_t)tap * p.NKC + kc) * p.wstage_bytes, (uint32_t)p.wstage_bytes, full);
                        ++it;
                    }
        }
    } else if (warp == 2) {
        // ===================================================== MMA issuer (one thread)
        if (lane == 0) {
            uint32_t slab_it = 0, w_it = 0, group = 0;
            const uint32_t sbo_a = (uint32_t)p.WW * 16, lbo_a = (uint32_t)p.cg_pitch;
            const uint32_t sbo_b = 128, lbo_b = (uint32_t)p.OC * 16;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++group) {
                const ItemCoord c = decode_item(p, item);
                const uint32_t set = group % p.nsets, set_phase = (group / p.nsets) & 1;
                ptx::mbar_wait(ptx::smem_u32(&bars->acc_empty[set]), set_phase ^ 1);
                ptx::tc_fence_after();
                uint32_t started = 0;                       // bit p: accumulator p already holds a partial sum
                for (int kc = 0; kc < p.NKC; ++kc) {
                    // slabs of this (item, kchunk), in producer order
                    uint32_t slab_addr[kMaxP + 2];
                    uint32_t slab_stage[kMaxP + 2];
                    uint32_t have = 0;
                    for (int q = 0; q < c.pvalid + p.kd - 1; ++q) {
                        const int zi = c.z0 - p.pd + q;
                        if (zi < 0 || zi >= p.Dpi) continue;
                        const uint32_t s = slab_it % p.nslabs, ph = (slab_it / p.nslabs) & 1;
                        ptx::mbar_wait(ptx::smem_u32(&bars->slab_full[s]), ph);
                        slab_addr[q] = ptx::smem_u32(slabs + (size_t)s * p.slab_bytes);
                        slab_stage[q] = s;
                        have |= 1u << q;
                        ++slab_it;
                    }
                    ptx::tc_fence_after();
                    for (int tap = 0; tap < ntaps; ++tap) {
                        const int kz = tap / (p.kh * p.kw), kr = tap - kz * p.kh * p.kw, ky = kr / p.kw, kx = kr - ky * p.kw;
                        const uint32_t ws = w_it % p.nwstages, wph = (w_it / p.nwstages) & 1;
                        ptx::mbar_wait(ptx::smem_u32(&bars->w_full[ws]), wph);
                        ptx::tc_fence_after();
                        const uint32_t wbase = ptx::smem_u32(wstages + (size_t)ws * p.wstage_bytes);
                        const uint32_t a_off = (uint32_t)(ky * p.WW + kx) * 16;
                        for (int pl = 0; pl < c.pvalid; ++pl) {
                            const int q = pl + kz;
                            if (!((have >> q) & 1)) continue;          // plane outside the volume: zero contribution
                            const uint32_t d_tmem = tmem_base + (uint32_t)((set * p.P + pl) * p.OC);
                            for (int ks = 0; ks < p.KC / 16; ++ks) {
                                const uint64_t a_desc = make_smem_desc(slab_addr[q] + a_off + (uint32_t)(2 * ks) * lbo_a, lbo_a, sbo_a);
                                const uint64_t b_desc = make_smem_desc(wbase + (uint32_t)(2 * ks) * lbo_b, lbo_b, sbo_b);
                                ptx::umma_bf16(d_tmem, a_desc, b_desc, p.idesc, (started >> pl) & 1);
                                started |= 1u << pl;
                            }
                        }
                        ptx::umma_commit(ptx::smem_u32(&bars->w_empty[ws]));      // weight stage free once these MMAs retire
                        ++w_it;
                    }
                    for (int q = 0; q < c.pvalid + p.kd - 1; ++q)
                        if ((have >> q) & 1) ptx::umma_commit(ptx::smem_u32(&bars->slab_empty[slab_stage[q]]));
                }
                ptx::umma_commit(ptx::smem_u32(&bars->acc_full[set]));          // accumulators of this tile group are final
            }
        }
    } else {
        // ===================================================== epilogue: TMEM -> registers -> (+bias) -> bf16 -> global
        const int lane_grp = warp & 3;                        // TMEM lanes this warp may read: [32*lane_grp, +32)
        const int m = lane_grp * 32 + lane;                   // row of the tile = voxel (yy, xx)
        const int yy = m >> 3, xx = m & 7;
        uint32_t group = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++group) {
            const ItemCoord c = decode_item(p, item);
            const uint32_t set = group % p.nsets, set_phase = (group / p.nsets) & 1;
            ptx::mbar_wait(ptx::smem_u32(&bars->acc_full[set]), set_phase);
            ptx::tc_fence_after();
            const int y = c.y0 + yy, x = c.x0 + xx;
            const bool inside = y < p.Ho && x < p.Wo;
            for (int pl = 0; pl < c.pvalid; ++pl) {
                const int64_t vox = (((int64_t)c.nb * p.Dpo + c.z0 + pl) * p.Ho + y) * p.Wo + x;
                __nv_bfloat16* dst = p.out + vox * p.OC;
                const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)((set * p.P + pl) * p.OC);
                for (int c0 = 0; c0 < p.OC; c0 += 16) {
                    float v[16];
                    ptx::tmem_ld16(taddr + (uint32_t)c0, v);
                    if (p.bias != nullptr) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + c0 + i);
                    }
                    if (inside) {
                        uint4 lo, hi;
                        __nv_bfloat162* l2 = reinterpret_cast<__nv_bfloat162*>(&lo);
                        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&hi);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            l2[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                            h2[i] = __floats2bfloat162_rn(v[8 + 2 * i], v[8 + 2 * i + 1]);
                        }
                        *reinterpret_cast<uint4*>(dst + c0) = lo;
                        *reinterpret_cast<uint4*>(dst + c0 + 8) = hi;
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[set]));
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ weight packing
// dst[tap'][kchunk][cg][oc][8] (bf16) from the fp32 PyTorch parameter (Co, Ci, taps).
//   fwd  : ic = ci, oc = co, tap' = tap
//   dgrad: ic = co, oc = ci, tap' = taps-1-tap (all three axes flipped)
__global__ void pack_weights_umma_kernel(int Ci, int Co, int taps, int dgrad, int KC, const float* __restrict__ w, __nv_bfloat16* __restrict__ dst) {
    const int IC = dgrad ? Co : Ci, OC = dgrad ? Ci : Co;
    const int NKC = IC / KC;
    const int64_t total = (int64_t)taps * IC * OC;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e;
        const int j = (int)(r % 8); r /= 8;
        const int oc = (int)(r % OC); r /= OC;
        const int cg = (int)(r % (KC / 8)); r /= (KC / 8);
        const int kc = (int)(r % NKC);
        const int tp = (int)(r / NKC);
        const int ic = kc * KC + cg * 8 + j;
        const int tap = dgrad ? taps - 1 - tp : tp;
        const int ci = dgrad ? oc : ic, co = dgrad ? ic : oc;
        dst[e] = __float2bfloat16_rn(w[((int64_t)co * Ci + ci) * taps + tap]);
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(sym);
    });
    return fn;
}

inline int umma_kchunk(int IC) {
    for (int kc = 64; kc >= 16; kc -= 16)
        if (IC % kc == 0) return kc;
    return 0;
}

// geometry of (desc, pass) as a stride-1 "same-style" convolution gathered from `in` producing `out`
struct UmmaGeom { int IC, OC, Di, Hi, Wi, Do, Ho, Wo, kd, kh, kw, pd, ph, pw; };

inline bool umma_geom(const b200_conv_desc* d, int pass, UmmaGeom* g) {
    if (!d->allow_umma || d->transposed) return false;
    if (pass != B200_PASS_FWD && pass != B200_PASS_DGRAD) return false;
    if (d->x_dtype != B200_BF16 || d->y_dtype != B200_BF16) return false;
    if (d->sd != 1 || d->sh != 1 || d->sw != 1 || d->dd != 1 || d->dh != 1 || d->dw != 1) return false;
    if (!((d->kd == 1 || d->kd == 3) && (d->kh == 1 || d->kh == 3) && (d->kw == 1 || d->kw == 3))) return false;
    if (d->pd > d->kd - 1 || d->ph > d->kh - 1 || d->pw > d->kw - 1) return false;
    if (d->Ci % 16 || d->Co % 16 || d->Ci > 512 || d->Co > 512) return false;
    g->kd = d->kd; g->kh = d->kh; g->kw = d->kw;
    if (pass == B200_PASS_FWD) {
        g->IC = d->Ci; g->OC = d->Co; g->Di = d->Di; g->Hi = d->Hi; g->Wi = d->Wi; g->Do = d->Do; g->Ho = d->Ho; g->Wo = d->Wo;
        g->pd = d->pd; g->ph = d->ph; g->pw = d->pw;
    } else {
        g->IC = d->Co; g->OC = d->Ci; g->Di = d->Do; g->Hi = d->Ho; g->Wi = d->Wo; g->Do = d->Di; g->Ho = d->Hi; g->Wo = d->Wi;
        g->pd = d->kd - 1 - d->pd; g->ph = d->kh - 1 - d->ph; g->pw = d->kw - 1 - d->pw;
    }
    if (g->OC > 256) return false;                         // one UMMA N and <= 512 TMEM columns for 2 planes
    // tiny problems are launch-bound either way; keep them on the SIMT path (also avoids degenerate TMA boxes)
    if ((int64_t)d->N * g->Do * g->Ho * g->Wo < 512) return false;
    return true;
}

inline bool umma_conv_supported(const b200_conv_desc* d, int pass) {
    UmmaGeom g;
    return umma_geom(d, pass, &g);
}

inline size_t umma_packed_bytes(const b200_conv_desc* d, int pass) {
    (void)pass;
    return (size_t)d->kd * d->kh * d->kw * d->Ci * d->Co * sizeof(__nv_bfloat16);
}
inline size_t umma_workspace_bytes(const b200_conv_desc*, int) { return 0; }

inline int umma_pack_weights(const b200_conv_desc* d, int pass, const float* w, void* packed, void* stream) {
    UmmaGeom g;
    B200_REQUIRE(umma_geom(d, pass, &g), "umma pack: unsupported descriptor");
    const int taps = d->kd * d->kh * d->kw;
    const int64_t total = (int64_t)taps * d->Ci * d->Co;
    B200_LAUNCH(pack_weights_umma_kernel, stream_grid(total, 256), 256, 0, stream, d->Ci, d->Co, taps, pass == B200_PASS_DGRAD ? 1 : 0,
                umma_kchunk(g.IC), w, (__nv_bfloat16*)packed);
    return 0;
}

inline int umma_plan(const UmmaGeom& g, int N, UmmaConvParams* p, size_t* smem_bytes) {
    memset(p, 0, sizeof *p);
    p->kd = g.kd; p->kh = g.kh; p->kw = g.kw; p->pd = g.pd; p->ph = g.ph; p->pw = g.pw;
    p->IC = g.IC; p->OC = g.OC; p->Hi = g.Hi; p->Wi = g.Wi; p->Ho = g.Ho; p->Wo = g.Wo;
    if (g.kd == 1) { p->NB = 1; p->Dpi = N * g.Di; p->Dpo = N * g.Do; }      // planes are independent: merge (n, z)
    else { p->NB = N; p->Dpi = g.Di; p->Dpo = g.Do; }
    p->KC = umma_kchunk(g.IC);
    B200_REQUIRE(p->KC > 0, "umma: IC=%d is not a multiple of 16", g.IC);
    p->NKC = g.IC / p->KC;
    p->HH = kTileH + g.kh - 1; p->WW = kTileW + g.kw - 1;
    p->cg_pitch = ((p->HH * p->WW * 16) + 127) & ~127;
    p->slab_bytes = (p->KC / 8) * p->cg_pitch;
    p->wstage_bytes = p->KC * g.OC * 2;
    // accumulators: prefer two sets (epilogue/MMA overlap) of up to 4 planes
    int P = 256 / g.OC;
    p->nsets = 2;
    if (P < 2) { P = 512 / g.OC; p->nsets = 1; }
    if (P > kMaxP) P = kMaxP;
    if (P > p->Dpo) P = p->Dpo;
    if (P < 1) P = 1;
    p->P = P;
    int cols = p->nsets * P * g.OC, pow2 = 32;
    while (pow2 < cols) pow2 <<= 1;
    B200_REQUIRE(pow2 <= 512, "umma: accumulators do not fit TMEM");
    p->tmem_cols = pow2;
    // shared memory: slabs (at least one tile group's worth, two if they fit) + weight ring
    const size_t budget = 227 * 1024 - sizeof(UmmaBarriers) - 1024;
    const int need = P + g.kd - 1;
    int nslabs = need, nw = 2;
    B200_REQUIRE((size_t)nslabs * p->slab_bytes + (size_t)nw * p->wstage_bytes <= budget, "umma: tile does not fit shared memory");
    while (nw < kMaxWStages && nw < 4 && (size_t)nslabs * p->slab_bytes + (size_t)(nw + 1) * p->wstage_bytes <= budget) ++nw;
    if ((size_t)2 * need * p->slab_bytes + (size_t)nw * p->wstage_bytes <= budget && 2 * need <= kMaxSlabs) nslabs = 2 * need;
    while (nw < kMaxWStages && (size_t)nslabs * p->slab_bytes + (size_t)(nw + 1) * p->wstage_bytes <= budget) ++nw;
    p->nslabs = nslabs; p->nwstages = nw;
    *smem_bytes = sizeof(UmmaBarriers) + (size_t)nslabs * p->slab_bytes + (size_t)nw * p->wstage_bytes;
    p->tiles_y = (g.Ho + kTileH - 1) / kTileH;
    p->tiles_x = (g.Wo + kTileW - 1) / kTileW;
    p->zchunks = (p->Dpo + P - 1) / P;
    const int64_t items = (int64_t)p->NB * p->zchunks * p->tiles_y * p->tiles_x;
    B200_REQUIRE(items < (1ll << 31), "umma: too many tiles");
    p->items = (int)items;
    p->idesc = make_idesc_bf16(g.OC);
    return 0;
}

inline int umma_conv_run(const b200_conv_desc* d, int pass, const void* in, const void* w_packed, const float* bias, void* out, void*, size_t,
                         void* stream) {
    UmmaGeom g;
    B200_REQUIRE(umma_geom(d, pass, &g), "umma conv: unsupported descriptor");
    B200_REQUIRE(aligned16(in) && aligned16(out) && aligned16(w_packed), "umma conv: pointers must be 16-byte aligned");
    UmmaConvParams p;
    size_t smem_bytes = 0;
    if (umma_plan(g, d->N, &p, &smem_bytes)) return 1;
    p.w = (const __nv_bfloat16*)w_packed;
    p.bias = bias;
    p.out = (__nv_bfloat16*)out;
    PFN_tmapEncodeTiled enc = tmap_encoder();
    B200_REQUIRE(enc != nullptr, "umma conv: cuTensorMapEncodeTiled is unavailable in this driver");
    CUtensorMap map;
    const cuuint64_t gdim[4] = {(cuuint64_t)g.IC, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)d->N * g.Di};
    const cuuint64_t gstr[3] = {(cuuint64_t)g.IC * 2, (cuuint64_t)g.Wi * g.IC * 2, (cuuint64_t)g.Hi * g.Wi * g.IC * 2};
    const cuuint32_t box[4] = {8, (cuuint32_t)p.WW, (cuuint32_t)p.HH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "umma conv: cuTensorMapEncodeTiled failed with %d", (int)r);
    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [] { attr_err = cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    B200_REQUIRE(attr_err == cudaSuccess, "umma conv: cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(attr_err));
    const int grid = p.items < kNumSMs ? p.items : kNumSMs;       // persistent: one CTA per SM
    B200_LAUNCH(conv_umma_kernel, grid, kUmmaThreads, smem_bytes, stream, map, p);
    return 0;
}

inline int umma_wgrad_run(const b200_conv_desc*, const void*, const void*, float*, float*, void*, size_t, void*) {
    return fail("umma wgrad is not selected by b200_conv_algo");
}

}  // namespace b200
