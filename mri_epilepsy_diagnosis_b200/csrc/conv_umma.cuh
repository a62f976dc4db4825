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

#include <cstdlib>
#include <cstring>
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

// one lane of a converged warp (the CUTLASS elect_one_sync idiom): keeps the surrounding control flow and all operands
// warp-uniform, so descriptors live in uniform registers and UTCHMMA is issued without per-instruction R2UR/vote loops
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred != 0;
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
// 16-byte asynchronous global->shared copy (LDGSTS); src_bytes = 0 zero-fills (padding / out-of-volume halo)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
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
// same, with the two 64-bit shared-memory descriptors given as (lo, hi) halves: the hi halves (SBO, version) are loop
// invariants and the lo halves advance by plain 32-bit adds, which keeps the single issuing thread's loop short
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
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

// three 32x16 blocks with ONE wait (loads and wait in a single asm block: the compiler must not touch the destination registers
// before the wait)
__device__ __forceinline__ void tmem_ld16x3(uint32_t t0, uint32_t t1, uint32_t t2, float (&a)[16], float (&b)[16], float (&c)[16]) {
    uint32_t r[48];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%48];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%49];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47}, [%50];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47])
        : "r"(t0), "r"(t1), "r"(t2) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[16 + i]); c[i] = __uint_as_float(r[32 + i]); }
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
constexpr int kStageThreads = 128;             // LDGSTS producer threads (4 warps)
constexpr int kMaxP = 4;                        // output planes per tile group
constexpr int kUmmaThreads = 224 + kStageThreads;     // + 4 LDGSTS producer warps (7..10)
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
    int w_resident;         // 1: every (tap, kchunk) block is loaded once per CTA and stays in shared memory
    int nsets;              // TMEM accumulator sets (1 or 2)
    int tmem_cols;
    int tiles_y, tiles_x, zchunks;   // per range
    int items;
    uint32_t idesc;
    int use_tma;                     // 1: halo slabs by TMA tensor loads, 0: by LDGSTS producer warps
    // split-K for layers with fewer tile groups than SMs (8^3 / 16^3 levels): an item is (tile group, split) and a split owns
    // NKC/skc channel chunks and, when skz > 1, one kz plane of the filter; fp32 partials go to `partial` and are reduced afterwards
    int skc, skz, nsplit;
    int64_t total_vox;
    float* partial;                  // [nsplit][total_vox][OC] or null
    const __nv_bfloat16* in;         // [NB*Dpi][Hi][Wi][IC]
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

struct ItemCoord { int nb, z0, pvalid, y0, x0, sp, kc0, kc1, kz0, kz1; };
__device__ __forceinline__ ItemCoord decode_item(const UmmaConvParams& p, int item) {
    ItemCoord c;
    c.sp = 0; c.kc0 = 0; c.kc1 = p.NKC; c.kz0 = 0; c.kz1 = p.kd;
    if (p.nsplit > 1) {
        c.sp = item % p.nsplit; item /= p.nsplit;
        const int per = p.NKC / p.skc, ikc = c.sp / p.skz;
        c.kc0 = ikc * per; c.kc1 = c.kc0 + per;
        if (p.skz > 1) { c.kz0 = c.sp % p.skz; c.kz1 = c.kz0 + 1; }
    }
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

// ------------------------------------------------------------------------------------------------ halo staging by LDGSTS
// Measured on B200: TMA tensor loads with the 16-byte inner rows this layout needs cost ~10 cycles per row and capped
// the kernel at ~9% of the tensor peak (profiles/r1_conv_table_tma.json).  128 producer threads issuing 16-byte
// cp.async (one (voxel, channel-group) entry each, coalesced along the channels of a voxel, zero-filled outside the
// volume) move the same slab ~30x faster; completion is published to the MMA thread through the same mbarriers after
// a generic->async proxy fence.  The TMA variant stays available (B200_CONV_STAGING=tma) for comparison.
__device__ __forceinline__ void ldgsts_tile(uint32_t dst, int cg_pitch, const __nv_bfloat16* plane, int C, int H, int W, int c0, int ncg,
                                            int y0, int x0, int HH, int WW, int tid) {
    const int n = ncg * HH * WW;
    for (int e = tid; e < n; e += kStageThreads) {
        const int v = e / ncg, cg = e - v * ncg;                 // channel group fastest: a voxel's channels are contiguous in HBM
        const int hy = v / WW, hx = v - hy * WW;
        const int y = y0 + hy, x = x0 + hx;
        const bool ok = (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
        const __nv_bfloat16* src = ok ? plane + ((int64_t)y * W + x) * C + c0 + cg * 8 : plane;
        ptx::cp_async16(dst + (uint32_t)(cg * cg_pitch + v * 16), src, ok ? 16u : 0u);
    }
}

// Software pipeline of one producer thread: the arrive for a stage is deferred until the next stage's copies have been
// issued (so two stages are in flight), but never across a wait that could block on the consumer.
struct StagePipe {
    uint32_t pending = 0;
    __device__ __forceinline__ void flush() {
        if (pending) {
            ptx::cp_async_wait<0>();
            ptx::fence_proxy_async();
            ptx::mbar_arrive(pending);
            pending = 0;
        }
    }
    __device__ __forceinline__ void acquire(uint32_t empty_bar, uint32_t parity) {
        if (!ptx::mbar_try_wait(empty_bar, parity)) { flush(); ptx::mbar_wait(empty_bar, parity); }
    }
    __device__ __forceinline__ void publish(uint32_t full_bar) {
        ptx::cp_async_commit();
        if (pending) {
            ptx::cp_async_wait<1>();
            ptx::fence_proxy_async();
            ptx::mbar_arrive(pending);
        }
        pending = full_bar;
    }
};

// ------------------------------------------------------------------------------------------------ MMA issue loop
// One thread issues every tcgen05.mma of the CTA, so its instruction count per MMA bounds the tensor pipe for small N.
// KD is a template parameter so that the plane index q = pl + kz is static and the slab descriptors stay in registers.
template <int KD>
__device__ __forceinline__ void mma_issue_loop(const UmmaConvParams& p, UmmaBarriers* bars, uint8_t* slabs, uint8_t* wstages, uint32_t tmem_base) {
    // Executed by ALL 32 lanes of the MMA warp with warp-uniform values; only the tcgen05 instructions are elected.
    constexpr int NQ = kMaxP + KD - 1;
    const uint32_t lbo_a16 = (uint32_t)p.cg_pitch >> 4, lbo_b16 = (uint32_t)p.OC;        // 16-byte units (OC*16 B >> 4)
    const uint32_t a_hi = (((uint32_t)p.WW * 16) >> 4) | (1u << 14);                     // SBO | version 1 (bit 46)
    const uint32_t b_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lbo_field = lbo_a16 << 16, b_lbo_field = lbo_b16 << 16;
    const uint32_t slabs16 = ptx::smem_u32(slabs) >> 4, slab16 = (uint32_t)p.slab_bytes >> 4;
    const uint32_t w16 = ptx::smem_u32(wstages) >> 4, wstage16 = (uint32_t)p.wstage_bytes >> 4;
    const int nks = p.KC / 16, khw = p.kh * p.kw;
    const uint32_t a_kstep = 2 * lbo_a16, b_kstep = 2 * lbo_b16;
    const uint32_t idesc = p.idesc;
    uint32_t ss = 0, sph = 0, ws = 0, wph = 0, group = 0;
    if (p.w_resident) { ptx::mbar_wait(ptx::smem_u32(&bars->w_full[0]), 0); ptx::tc_fence_after(); }
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++group) {
        const ItemCoord c = decode_item(p, item);
        const uint32_t set = group % p.nsets, set_phase = (group / p.nsets) & 1;
        ptx::mbar_wait(ptx::smem_u32(&bars->acc_empty[set]), set_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(set * p.P * p.OC);
        uint32_t started = 0;                                   // bit pl: accumulator pl already holds a partial sum
        for (int kc = c.kc0; kc < c.kc1; ++kc) {
            uint32_t slab_lo[NQ], slab_bar[NQ], have = 0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                slab_lo[q] = 0; slab_bar[q] = 0;
                const int zi = c.z0 - p.pd + q;
                if (q >= c.kz0 && q < c.pvalid + c.kz1 - 1 && zi >= 0 && zi < p.Dpi) {
                    ptx::mbar_wait(ptx::smem_u32(&bars->slab_full[ss]), sph);
                    slab_lo[q] = ((slabs16 + ss * slab16) & 0x3FFF) | a_lbo_field;
                    slab_bar[q] = ptx::smem_u32(&bars->slab_empty[ss]);
                    have |= 1u << q;
                    if (++ss == (uint32_t)p.nslabs) { ss = 0; sph ^= 1; }
                }
            }
            ptx::tc_fence_after();
#pragma unroll
            for (int kz = 0; kz < KD; ++kz) {
                if (kz < c.kz0 || kz >= c.kz1) continue;
                uint32_t a_tap = 0;                             // (ky*WW + kx) in 16-byte units
                int kx = 0;
                for (int kyx = 0; kyx < khw; ++kyx) {
                    uint32_t b_lo;
                    if (p.w_resident) {
                        b_lo = ((w16 + (uint32_t)((kz * khw + kyx) * p.NKC + kc) * wstage16) & 0x3FFF) | b_lbo_field;
                    } else {
                        ptx::mbar_wait(ptx::smem_u32(&bars->w_full[ws]), wph);
                        ptx::tc_fence_after();
                        b_lo = ((w16 + ws * wstage16) & 0x3FFF) | b_lbo_field;
                    }
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int pl = 0; pl < kMaxP; ++pl) {
                            if (pl < c.pvalid && ((have >> (pl + kz)) & 1)) {
                                uint32_t a_lo = slab_lo[pl + kz] + a_tap, bb = b_lo;
                                const uint32_t d_tmem = d0 + (uint32_t)(pl * p.OC);
                                uint32_t acc = (started >> pl) & 1;
                                for (int ks = 0; ks < nks; ++ks) {
                                    ptx::umma_bf16_lohi(d_tmem, a_lo, a_hi, bb, b_hi, idesc, acc);
                                    a_lo += a_kstep; bb += b_kstep; acc = 1;
                                }
                            }
                        }
                        if (!p.w_resident) ptx::umma_commit(ptx::smem_u32(&bars->w_empty[ws]));   // stage free once these MMAs retire
                    }
                    __syncwarp();
#pragma unroll
                    for (int pl = 0; pl < kMaxP; ++pl)
                        if (pl < c.pvalid && ((have >> (pl + kz)) & 1)) started |= 1u << pl;
                    if (!p.w_resident) { if (++ws == (uint32_t)p.nwstages) { ws = 0; wph ^= 1; } }
                    ++a_tap;
                    if (++kx == p.kw) { kx = 0; a_tap += (uint32_t)(p.WW - p.kw); }
                }
            }
            if (ptx::elect_one()) {
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    if ((have >> q) & 1) ptx::umma_commit(slab_bar[q]);
            }
            __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit(ptx::smem_u32(&bars->acc_full[set]));          // accumulators of this tile group are final
        __syncwarp();
    }
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
        for (int i = 0; i < p.nslabs; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->slab_full[i]), p.use_tma ? 1 : kStageThreads);
            ptx::mbar_init(ptx::smem_u32(&bars->slab_empty[i]), 1);
        }
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
        // ===================================================== TMA producer: halo slabs (B200_CONV_STAGING=tma only)
        if (lane == 0 && p.use_tma) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                const ItemCoord c = decode_item(p, item);
                for (int kc = c.kc0; kc < c.kc1; ++kc)
                    for (int q = c.kz0; q < c.pvalid + c.kz1 - 1; ++q) {
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
        // ===================================================== weight producer
        if (lane == 0) {
            if (p.w_resident) {
                // small layers: all taps stay resident -- one bulk copy per (tap, kchunk) block, one barrier, once per CTA
                const uint32_t full = ptx::smem_u32(&bars->w_full[0]);
                const int nblocks = ntaps * p.NKC;
                ptx::mbar_expect_tx(full, (uint32_t)(nblocks * p.wstage_bytes));
                for (int b = 0; b < nblocks; ++b)
                    ptx::bulk_load(ptx::smem_u32(wstages + (size_t)b * p.wstage_bytes), reinterpret_cast<const uint8_t*>(p.w) + (size_t)b * p.wstage_bytes,
                                   (uint32_t)p.wstage_bytes, full);
            } else {
                uint32_t s = 0, ph = 0;
                const int khw = p.kh * p.kw;
                for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                    const ItemCoord c = decode_item(p, item);
                    for (int kc = c.kc0; kc < c.kc1; ++kc)
                        for (int tap = c.kz0 * khw; tap < c.kz1 * khw; ++tap) {
                            ptx::mbar_wait(ptx::smem_u32(&bars->w_empty[s]), ph ^ 1);
                            const uint32_t full = ptx::smem_u32(&bars->w_full[s]);
                            ptx::mbar_expect_tx(full, (uint32_t)p.wstage_bytes);
                            ptx::bulk_load(ptx::smem_u32(wstages + (size_t)s * p.wstage_bytes),
                                           reinterpret_cast<const uint8_t*>(p.w) + ((size_t)tap * p.NKC + kc) * p.wstage_bytes, (uint32_t)p.wstage_bytes, full);
                            if (++s == (uint32_t)p.nwstages) { s = 0; ph ^= 1; }
                        }
                }
            }
        }
    } else if (warp == 2) {
        // ===================================================== MMA issuer (whole warp converged, one elected lane issues)
        if (p.kd == 3) mma_issue_loop<3>(p, bars, slabs, wstages, tmem_base);
        else mma_issue_loop<1>(p, bars, slabs, wstages, tmem_base);
    } else if (warp >= 7) {
        // ===================================================== LDGSTS producers: halo slabs, 128 threads
        if (!p.use_tma) {
            const int tid = threadIdx.x - 7 * 32;
            StagePipe pipe;
            uint32_t ss = 0, sph = 0;
            const int ncg = p.KC / 8;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                const ItemCoord c = decode_item(p, item);
                for (int kc = c.kc0; kc < c.kc1; ++kc)
                    for (int q = c.kz0; q < c.pvalid + c.kz1 - 1; ++q) {
                        const int zi = c.z0 - p.pd + q;
                        if (zi < 0 || zi >= p.Dpi) continue;
                        pipe.acquire(ptx::smem_u32(&bars->slab_empty[ss]), sph ^ 1);
                        const __nv_bfloat16* plane = p.in + (int64_t)(c.nb * p.Dpi + zi) * p.Hi * p.Wi * p.IC;
                        ldgsts_tile(ptx::smem_u32(slabs + (size_t)ss * p.slab_bytes), p.cg_pitch, plane, p.IC, p.Hi, p.Wi, kc * p.KC, ncg,
                                    c.y0 - p.ph, c.x0 - p.pw, p.HH, p.WW, tid);
                        pipe.publish(ptx::smem_u32(&bars->slab_full[ss]));
                        if (++ss == (uint32_t)p.nslabs) { ss = 0; sph ^= 1; }
                    }
            }
            pipe.flush();
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
                if (p.partial != nullptr) {
                    // split-K: fp32 partial of this split; a split whose filter planes all fall outside the volume issued no MMA
                    bool any = false;
                    for (int kz = c.kz0; kz < c.kz1; ++kz) {
                        const int zi = c.z0 + pl - p.pd + kz;
                        any = any || (zi >= 0 && zi < p.Dpi);
                    }
                    float* dstp = p.partial + ((int64_t)c.sp * p.total_vox + vox) * p.OC;
                    for (int c0 = 0; c0 < p.OC; c0 += 16) {
                        float v[16];
                        ptx::tmem_ld16(taddr + (uint32_t)c0, v);
                        if (inside) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                *reinterpret_cast<float4*>(dstp + c0 + 4 * i) =
                                    any ? make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    continue;
                }
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

// out[v][oc] = bf16(bias[oc] + sum over splits of partial[s][v][oc]), fixed order (deterministic); 8 channels per thread
__global__ void umma_split_reduce_kernel(int nsplit, int64_t total, int OC, const float* __restrict__ partial, const float* __restrict__ bias,
                                         __nv_bfloat16* __restrict__ out) {
    const int64_t n8 = total / 8;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n8; e += (int64_t)gridDim.x * blockDim.x) {
        float a[8];
        const int oc = (int)((e * 8) % OC);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = bias != nullptr ? __ldg(bias + oc + i) : 0.f;
        for (int s = 0; s < nsplit; ++s) {
            const float4* src = reinterpret_cast<const float4*>(partial + (int64_t)s * total + e * 8);
            const float4 u = __ldg(src), w = __ldg(src + 1);
            a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w; a[4] += w.x; a[5] += w.y; a[6] += w.z; a[7] += w.w;
        }
        Pack<__nv_bfloat16, 8>::store(out + e * 8, a);
    }
}

// ------------------------------------------------------------------------------------------------ weight packing
// dst[tap'][kchunk][cg][oc][8] (bf16) from the fp32 PyTorch parameter (Co, Ci, taps).
//   fwd  : ic = ci, oc = co, tap' = tap
//   dgrad: ic = co, oc = ci, tap' = taps-1-tap (all three axes flipped)
// fold = 1 (row_fwd_kernel's kx-folded mode, NKC == 1): dst[(kz,ky)][cg][kx][oc][8] -- the three kx blocks of a filter row
// side by side, so one B operand of N = 3*OC covers them
__global__ void pack_weights_umma_kernel(int Ci, int Co, int taps, int dgrad, int KC, const float* __restrict__ w, __nv_bfloat16* __restrict__ dst,
                                         int fold) {
    const int IC = dgrad ? Co : Ci, OC = dgrad ? Ci : Co;
    const int NKC = IC / KC;
    const int64_t total = (int64_t)taps * IC * OC;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e;
        const int j = (int)(r % 8); r /= 8;
        const int oc = (int)(r % OC); r /= OC;
        int cg, kc, tp;
        if (fold) {
            const int kx = (int)(r % 3); r /= 3;
            cg = (int)(r % (KC / 8)); r /= (KC / 8);
            kc = 0;
            tp = (int)r * 3 + kx;
        } else {
            cg = (int)(r % (KC / 8)); r /= (KC / 8);
            kc = (int)(r % NKC);
            tp = (int)(r / NKC);
        }
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

inline bool staging_uses_tma() {
    static const bool v = [] { const char* e = getenv("B200_CONV_STAGING"); return e != nullptr && strcmp(e, "tma") == 0; }();
    return v;
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

inline bool umma_wgrad_geom(const b200_conv_desc* d);
inline bool row_fwd_folds(const b200_conv_desc* d, int pass);          // conv_rowf.cuh: the row-slab kernel wants the kx-folded weight layout
inline bool umma_conv_supported(const b200_conv_desc* d, int pass) {
    if (pass == B200_PASS_WGRAD) return umma_wgrad_geom(d);
    UmmaGeom g;
    return umma_geom(d, pass, &g);
}

inline size_t umma_packed_bytes(const b200_conv_desc* d, int pass) {
    (void)pass;
    return (size_t)d->kd * d->kh * d->kw * d->Ci * d->Co * sizeof(__nv_bfloat16);
}
inline size_t umma_wgrad_workspace_bytes(const b200_conv_desc* d);
inline size_t umma_split_workspace_bytes(const b200_conv_desc* d, int pass);
inline size_t umma_workspace_bytes(const b200_conv_desc* d, int pass) {
    return pass == B200_PASS_WGRAD ? umma_wgrad_workspace_bytes(d) + 256 : umma_split_workspace_bytes(d, pass);
}

inline int umma_pack_weights(const b200_conv_desc* d, int pass, const float* w, void* packed, void* stream) {
    UmmaGeom g;
    B200_REQUIRE(umma_geom(d, pass, &g), "umma pack: unsupported descriptor");
    const int taps = d->kd * d->kh * d->kw;
    const int64_t total = (int64_t)taps * d->Ci * d->Co;
    B200_LAUNCH(pack_weights_umma_kernel, stream_grid(total, 256), 256, 0, stream, d->Ci, d->Co, taps, pass == B200_PASS_DGRAD ? 1 : 0,
                umma_kchunk(g.IC), w, (__nv_bfloat16*)packed, row_fwd_folds(d, pass) ? 1 : 0);
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
    p->use_tma = staging_uses_tma() ? 1 : 0;
    // TMA destinations must be 128-byte aligned; for LDGSTS an odd multiple of 16 B spreads the channel groups over the banks
    p->cg_pitch = p->use_tma ? ((p->HH * p->WW * 16) + 127) & ~127 : p->HH * p->WW * 16 + 16;
    p->slab_bytes = (((p->KC / 8) * p->cg_pitch) + 127) & ~127;
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
    const size_t all_w = (size_t)g.kd * g.kh * g.kw * g.IC * g.OC * 2;
    int nslabs = need, nw = 2;
    size_t wbytes;
    if (all_w <= 96 * 1024 && (size_t)2 * need * p->slab_bytes + all_w <= budget && 2 * need <= kMaxSlabs) {
        p->w_resident = 1;                       // small layers (<= 32x32x27, 64x16x27, ...): weights stay in shared memory
        nslabs = 2 * need; nw = 1;
        wbytes = all_w;
    } else {
        B200_REQUIRE((size_t)nslabs * p->slab_bytes + (size_t)nw * p->wstage_bytes <= budget, "umma: tile does not fit shared memory");
        while (nw < kMaxWStages && nw < 4 && (size_t)nslabs * p->slab_bytes + (size_t)(nw + 1) * p->wstage_bytes <= budget) ++nw;
        if ((size_t)2 * need * p->slab_bytes + (size_t)nw * p->wstage_bytes <= budget && 2 * need <= kMaxSlabs) nslabs = 2 * need;
        while (nw < kMaxWStages && (size_t)nslabs * p->slab_bytes + (size_t)(nw + 1) * p->wstage_bytes <= budget) ++nw;
        wbytes = (size_t)nw * p->wstage_bytes;
    }
    p->nslabs = nslabs; p->nwstages = nw;
    *smem_bytes = sizeof(UmmaBarriers) + (size_t)nslabs * p->slab_bytes + wbytes;
    p->tiles_y = (g.Ho + kTileH - 1) / kTileH;
    p->tiles_x = (g.Wo + kTileW - 1) / kTileW;
    p->zchunks = (p->Dpo + P - 1) / P;
    const int64_t items = (int64_t)p->NB * p->zchunks * p->tiles_y * p->tiles_x;
    B200_REQUIRE(items < (1ll << 31), "umma: too many tiles");
    p->items = (int)items;
    p->idesc = make_idesc_bf16(g.OC);
    // split-K when the tile groups cannot fill the machine: pick the (channel-chunk, kz) split with the fewest rounds x work per item
    p->skc = 1; p->skz = 1; p->nsplit = 1;
    p->total_vox = (int64_t)N * g.Do * g.Ho * g.Wo;
    static const bool allow_split = [] { const char* e = getenv("B200_UMMA_SPLIT"); return e == nullptr || strcmp(e, "0") != 0; }();
    if (allow_split && items * 2 <= kNumSMs) {
        double best = 1.0;
        for (int skz = 1; skz <= g.kd; skz += g.kd > 1 ? g.kd - 1 : 1)
            for (int skc = 1; skc <= p->NKC; ++skc) {
                if (p->NKC % skc) continue;
                const int ns = skc * skz;
                const double cost = (double)ceil_div(items * ns, kNumSMs) / ns + 0.01 * ns;     // the partials are not free
                if (cost < best - 1e-9) { best = cost; p->skc = skc; p->skz = skz; p->nsplit = ns; }
            }
        p->items = (int)(items * p->nsplit);
    }
    return 0;
}

inline size_t umma_split_workspace_bytes(const b200_conv_desc* d, int pass) {
    UmmaGeom g;
    UmmaConvParams p;
    size_t smem = 0;
    if (!umma_geom(d, pass, &g) || umma_plan(g, d->N, &p, &smem) || p.nsplit <= 1) return 0;
    return (size_t)p.nsplit * p.total_vox * g.OC * sizeof(float) + 256;
}

inline int umma_conv_run(const b200_conv_desc* d, int pass, const void* in, const void* w_packed, const float* bias, void* out, void* workspace,
                         size_t ws_bytes, void* stream) {
    UmmaGeom g;
    B200_REQUIRE(umma_geom(d, pass, &g), "umma conv: unsupported descriptor");
    B200_REQUIRE(aligned16(in) && aligned16(out) && aligned16(w_packed), "umma conv: pointers must be 16-byte aligned");
    UmmaConvParams p;
    size_t smem_bytes = 0;
    if (umma_plan(g, d->N, &p, &smem_bytes)) return 1;
    p.in = (const __nv_bfloat16*)in;
    p.w = (const __nv_bfloat16*)w_packed;
    p.bias = bias;
    p.out = (__nv_bfloat16*)out;
    if (p.nsplit > 1) {
        B200_REQUIRE(workspace != nullptr && ws_bytes >= (size_t)p.nsplit * p.total_vox * g.OC * sizeof(float), "umma conv: split-K workspace too small");
        p.partial = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
        B200_REQUIRE((size_t)((char*)p.partial - (char*)workspace) + (size_t)p.nsplit * p.total_vox * g.OC * sizeof(float) <= ws_bytes,
                     "umma conv: split-K workspace too small after alignment");
    }
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
    if (p.nsplit > 1) {
        const int64_t total = p.total_vox * g.OC;
        B200_LAUNCH(umma_split_reduce_kernel, stream_grid(total / 8, 256), 256, 0, stream, p.nsplit, total, g.OC, p.partial, bias, (__nv_bfloat16*)out);
    }
    return 0;
}

// ================================================================================================ wgrad on tcgen05
// dW[tap][co][ci] = sum over voxels v of dy[v][co] * x[v + tap - pad][ci]          (stride 1)
//
// GEMM per filter tap:  D[M = co (128 TMEM lanes)][N = ci] += A[co][K = 16 voxels] * B[16 voxels][ci]
// Both operands are "MN-major": the SAME channel-group-major shared-memory layout [cg][voxel][8 ch] that the forward
// kernel uses is, read the other way round, the canonical MN-major no-swizzle UMMA layout (8 channels contiguous, the
// 8 voxels of one x-row 16 B apart = one core matrix, next x-row at LBO, next channel group at SBO).  One UMMA consumes
// two x-rows (16 voxels); a tap is again only a different start address into the x halo slab.
// A CTA owns a "unit" = (tap range, ci chunk <= 64, co chunk <= 128) whose accumulators (taps x ci columns <= 512) stay
// in TMEM for the CTA's whole life while it marches over its share of the volume plane by plane; at the end the fp32
// partial is written once and a second kernel reduces the partials in a fixed order (deterministic).
constexpr int kWgThreads = 160;          // warps 0..3 = producers, then the final epilogue; warp 4 = MMA issuer + TMEM owner
constexpr int kWgMaxSlabs = 8, kWgMaxDy = 4;

struct WgradParams {
    int NB, Dpi, Dpo, Hi, Wi, Ho, Wo;
    int Ci, Co;
    int kd, kh, kw, pd, ph, pw;
    int NC, MC;                 // ci / co chunk handled by one unit
    int n_ci, n_co, n_tg;       // number of chunks / tap groups; units = n_tg * n_ci * n_co
    int tpu;                    // taps per unit (last group may hold fewer)
    int splits;                 // CTAs per unit
    int HH, WW, slab_cg_pitch, slab_bytes, nslabs;
    int dy_cg_pitch, dy_bytes, ndy, dy_row_pitch;
    int tmem_cols;
    int tiles_y, tiles_x, zsegs, zs;   // z segments per range, planes per segment
    int items;
    uint32_t idesc;
    int use_tma;
    const __nv_bfloat16* x;     // [NB*Dpi][Hi][Wi][Ci]
    const __nv_bfloat16* dy;    // [NB*Dpo][Ho][Wo][Co]
    float* partial;             // [splits][taps][Co][Ci]
};

struct alignas(128) WgradBarriers {
    uint64_t slab_full[kWgMaxSlabs], slab_empty[kWgMaxSlabs];
    uint64_t dy_full[kWgMaxDy], dy_empty[kWgMaxDy];
    uint64_t done;
    uint32_t tmem_base, started;
};

struct WgItem { int nb, z0, z1, y0, x0; };
__device__ __forceinline__ WgItem wg_decode(const WgradParams& p, int item) {
    WgItem c;
    const int tx = item % p.tiles_x; item /= p.tiles_x;
    const int ty = item % p.tiles_y; item /= p.tiles_y;
    const int zg = item % p.zsegs;
    c.nb = item / p.zsegs;
    c.z0 = zg * p.zs;
    c.z1 = min(p.Dpo, c.z0 + p.zs);
    c.y0 = ty * kTileH; c.x0 = tx * kTileW;
    return c;
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_umma_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap dy_map, const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    WgradBarriers* bars = reinterpret_cast<WgradBarriers*>(smem);
    uint8_t* dyst = smem + sizeof(WgradBarriers);                       // dy stages first: M=128 descriptors of a narrow co chunk
    uint8_t* slabs = dyst + (size_t)p.ndy * p.dy_bytes;                 // over-read into the slabs (rows >= Co are discarded)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // unit / split of this CTA
    const int split = blockIdx.x % p.splits;
    int unit = blockIdx.x / p.splits;
    const int tg = unit % p.n_tg; unit /= p.n_tg;
    const int cic = unit % p.n_ci;
    const int coc = unit / p.n_ci;
    const int ntaps = p.kd * p.kh * p.kw;
    const int tap0 = tg * p.tpu, tap1 = min(ntaps, tap0 + p.tpu);

    if (threadIdx.x == 0) {
        const uint32_t nprod = p.use_tma ? 1 : kStageThreads;
        for (int i = 0; i < p.nslabs; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->slab_full[i]), nprod); ptx::mbar_init(ptx::smem_u32(&bars->slab_empty[i]), 1); }
        for (int i = 0; i < p.ndy; ++i) { ptx::mbar_init(ptx::smem_u32(&bars->dy_full[i]), nprod); ptx::mbar_init(ptx::smem_u32(&bars->dy_empty[i]), 1); }
        ptx::mbar_init(ptx::smem_u32(&bars->done), 1);
        bars->started = 0;
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&x_map);
        ptx::prefetch_tmap(&dy_map);
    }
    if (warp == 4) {
        ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), (uint32_t)p.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < 4 && !p.use_tma) {
        // ===================================================== LDGSTS producers (128 threads): x halo slabs (one new plane per z) + dy tiles
        StagePipe pipe;
        uint32_t ss = 0, sph = 0, ds = 0, dph = 0;
        const int ncg_x = p.NC / 8, ncg_y = p.MC / 8;
        for (int item = split; item < p.items; item += p.splits) {
            const WgItem c = wg_decode(p, item);
            int next_plane = c.z0 - p.pd;
            for (int z = c.z0; z < c.z1; ++z) {
                const int last_needed = z - p.pd + p.kd - 1;
                for (; next_plane <= last_needed; ++next_plane) {
                    if (next_plane < 0 || next_plane >= p.Dpi) continue;
                    pipe.acquire(ptx::smem_u32(&bars->slab_empty[ss]), sph ^ 1);
                    const __nv_bfloat16* plane = p.x + (int64_t)(c.nb * p.Dpi + next_plane) * p.Hi * p.Wi * p.Ci;
                    ldgsts_tile(ptx::smem_u32(slabs + (size_t)ss * p.slab_bytes), p.slab_cg_pitch, plane, p.Ci, p.Hi, p.Wi, cic * p.NC, ncg_x,
                                c.y0 - p.ph, c.x0 - p.pw, p.HH, p.WW, (int)threadIdx.x);
                    pipe.publish(ptx::smem_u32(&bars->slab_full[ss]));
                    if (++ss == (uint32_t)p.nslabs) { ss = 0; sph ^= 1; }
                }
                pipe.acquire(ptx::smem_u32(&bars->dy_empty[ds]), dph ^ 1);
                const __nv_bfloat16* plane = p.dy + (int64_t)(c.nb * p.Dpo + z) * p.Ho * p.Wo * p.Co;
                ldgsts_tile(ptx::smem_u32(dyst + (size_t)ds * p.dy_bytes), p.dy_cg_pitch, plane, p.Co, p.Ho, p.Wo, coc * p.MC, ncg_y, c.y0, c.x0,
                            kTileH, kTileW, (int)threadIdx.x);
                pipe.publish(ptx::smem_u32(&bars->dy_full[ds]));
                if (++ds == (uint32_t)p.ndy) { ds = 0; dph ^= 1; }
            }
        }
        pipe.flush();
    }
    if (warp == 0 && p.use_tma) {
        // ===================================================== TMA producer (B200_CONV_STAGING=tma): same schedule, one thread
        if (lane == 0) {
            uint32_t ss = 0, sph = 0, ds = 0, dph = 0;
            const int ncg_x = p.NC / 8, ncg_y = p.MC / 8;
            for (int item = split; item < p.items; item += p.splits) {
                const WgItem c = wg_decode(p, item);
                int next_plane = c.z0 - p.pd;                               // next input plane (relative to the range) to load
                for (int z = c.z0; z < c.z1; ++z) {
                    const int last_needed = z - p.pd + p.kd - 1;
                    for (; next_plane <= last_needed; ++next_plane) {
                        if (next_plane < 0 || next_plane >= p.Dpi) continue;
                        ptx::mbar_wait(ptx::smem_u32(&bars->slab_empty[ss]), sph ^ 1);
                        const uint32_t full = ptx::smem_u32(&bars->slab_full[ss]);
                        ptx::mbar_expect_tx(full, (uint32_t)(ncg_x * p.HH * p.WW * 16));
                        const uint32_t dst = ptx::smem_u32(slabs + (size_t)ss * p.slab_bytes);
                        for (int cg = 0; cg < ncg_x; ++cg)
                            ptx::tma_load_4d(dst + cg * p.slab_cg_pitch, &x_map, full, cic * p.NC + cg * 8, c.x0 - p.pw, c.y0 - p.ph, c.nb * p.Dpi + next_plane);
                        if (++ss == (uint32_t)p.nslabs) { ss = 0; sph ^= 1; }
                    }
                    ptx::mbar_wait(ptx::smem_u32(&bars->dy_empty[ds]), dph ^ 1);
                    const uint32_t full = ptx::smem_u32(&bars->dy_full[ds]);
                    ptx::mbar_expect_tx(full, (uint32_t)(ncg_y * 128 * 16));
                    const uint32_t dst = ptx::smem_u32(dyst + (size_t)ds * p.dy_bytes);
                    for (int cg = 0; cg < ncg_y; ++cg)
                        ptx::tma_load_4d(dst + cg * p.dy_cg_pitch, &dy_map, full, coc * p.MC + cg * 8, c.x0, c.y0, c.nb * p.Dpo + z);
                    if (++ds == (uint32_t)p.ndy) { ds = 0; dph ^= 1; }
                }
            }
        }
    }
    if (warp == 4) {
        // ===================================================== MMA issuer (whole warp converged, one elected lane issues)
        {
            const uint32_t a_hi = ((uint32_t)p.dy_cg_pitch >> 4) | (1u << 14);          // SBO = next 8 output channels
            const uint32_t b_hi = ((uint32_t)p.slab_cg_pitch >> 4) | (1u << 14);        // SBO = next 8 input channels
            const uint32_t a_lbo = (128u >> 4) << 16;                                   // LBO = next x-row of the 16x8 dy tile
            const uint32_t b_lbo = (((uint32_t)p.WW * 16) >> 4) << 16;                  // LBO = next x-row of the halo slab
            const uint32_t dy16 = ptx::smem_u32(dyst) >> 4, dystage16 = (uint32_t)p.dy_bytes >> 4;
            const uint32_t sl16 = ptx::smem_u32(slabs) >> 4, slab16 = (uint32_t)p.slab_bytes >> 4;
            const int khw = p.kh * p.kw;
            const uint32_t idesc = p.idesc, a_rstep = (2 * (uint32_t)p.dy_row_pitch) >> 4, b_rstep = (uint32_t)(2 * p.WW);
            uint32_t ss = 0, sph = 0, ds = 0, dph = 0, started = 0;
            for (int item = split; item < p.items; item += p.splits) {
                const WgItem c = wg_decode(p, item);
                // ring bookkeeping: slot[k] = shared-memory slab of input plane (z - pd + k), or -1 when outside the volume
                int slot[3] = {-1, -1, -1};
                uint32_t slot_bar[3] = {0, 0, 0};
                int next_plane = c.z0 - p.pd;
                for (int z = c.z0; z < c.z1; ++z) {
                    if (z > c.z0) {                                        // slide the window: oldest plane was released last z
#pragma unroll
                        for (int k = 0; k < 2; ++k) { slot[k] = slot[k + 1]; slot_bar[k] = slot_bar[k + 1]; }
                    }
                    const int last_needed = z - p.pd + p.kd - 1;
                    for (; next_plane <= last_needed; ++next_plane) {
                        const int k = next_plane - (z - p.pd);             // position inside the window
                        int sidx = -1;
                        uint32_t sbar = 0;
                        if (next_plane >= 0 && next_plane < p.Dpi) {
                            ptx::mbar_wait(ptx::smem_u32(&bars->slab_full[ss]), sph);
                            sidx = (int)ss; sbar = ptx::smem_u32(&bars->slab_empty[ss]);
                            if (++ss == (uint32_t)p.nslabs) { ss = 0; sph ^= 1; }
                        }
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk) if (kk == k) { slot[kk] = sidx; slot_bar[kk] = sbar; }
                    }
                    ptx::mbar_wait(ptx::smem_u32(&bars->dy_full[ds]), dph);
                    ptx::tc_fence_after();
                    const uint32_t a_base = ((dy16 + ds * dystage16) & 0x3FFF) | a_lbo;
                    for (int tap = tap0; tap < tap1; ++tap) {
                        const int kz = tap / khw, kr = tap - kz * khw, ky = kr / p.kw, kx = kr - ky * p.kw;
                        int sidx = -1;
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk) if (kk == kz) sidx = slot[kk];
                        if (sidx < 0) continue;                            // plane outside the volume: zero contribution
                        const uint32_t d_tmem = tmem_base + (uint32_t)((tap - tap0) * p.NC);
                        if (ptx::elect_one()) {
                            uint32_t a_lo = a_base;
                            uint32_t b_lo = (((sl16 + (uint32_t)sidx * slab16) + (uint32_t)(ky * p.WW + kx)) & 0x3FFF) | b_lbo;
                            uint32_t acc = (started >> (tap - tap0)) & 1;
#pragma unroll
                            for (int r = 0; r < kTileH / 2; ++r) {          // two x-rows (16 voxels) per MMA
                                ptx::umma_bf16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, acc);
                                a_lo += a_rstep;
                                b_lo += b_rstep;
                                acc = 1;
                            }
                        }
                        __syncwarp();
                        started |= 1u << (tap - tap0);
                    }
                    if (ptx::elect_one()) {
                        ptx::umma_commit(ptx::smem_u32(&bars->dy_empty[ds]));
                        // the oldest plane of the window is not needed by z+1 (for kd == 1 that is the only plane)
                        if (slot[0] >= 0) ptx::umma_commit(slot_bar[0]);
                        if (z + 1 == c.z1) {                               // end of the column segment: release the rest
#pragma unroll
                            for (int k = 1; k < 3; ++k) if (k < p.kd && slot[k] >= 0) ptx::umma_commit(slot_bar[k]);
                        }
                    }
                    __syncwarp();
                    if (++ds == (uint32_t)p.ndy) { ds = 0; dph ^= 1; }
                }
            }
            if (ptx::elect_one()) {
                bars->started = started;
                __threadfence_block();
                ptx::umma_commit(ptx::smem_u32(&bars->done));
            }
            __syncwarp();
        }
    } else {
        // ===================================================== epilogue (once): TMEM -> fp32 partial[split][tap][co][ci]
        __syncwarp();
        const int lane_grp = warp & 3;
        const int row = lane_grp * 32 + lane;                               // co inside the chunk
        ptx::mbar_wait(ptx::smem_u32(&bars->done), 0);
        ptx::tc_fence_after();
        const uint32_t started = *reinterpret_cast<volatile uint32_t*>(&bars->started);
        const int co = coc * p.MC + row;
        const bool row_ok = row < p.MC && co < p.Co;
        for (int tap = tap0; tap < tap1; ++tap) {
            const bool has = (started >> (tap - tap0)) & 1;
            float* dst = p.partial + (((size_t)split * ntaps + tap) * p.Co + co) * p.Ci + cic * p.NC;
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)((tap - tap0) * p.NC);
            for (int c0 = 0; c0 < p.NC; c0 += 16) {
                float v[16];
                if (has) ptx::tmem_ld16(taddr + (uint32_t)c0, v);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<float4*>(dst + c0 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        }
        ptx::tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
}

// dw[(co*Ci + ci)*taps + tap] = sum_s partial[s][tap][co][ci]
__global__ void wgrad_umma_reduce_kernel(int splits, int taps, int Co, int Ci, const float* __restrict__ partial, float* __restrict__ dw) {
    const int64_t total = (int64_t)taps * Co * Ci;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        float acc = 0.f;
#pragma unroll 8
        for (int s = 0; s < splits; ++s) acc += __ldg(partial + (int64_t)s * total + e);
        const int ci = (int)(e % Ci);
        const int64_t r = e / Ci;
        const int co = (int)(r % Co), tap = (int)(r / Co);
        dw[((int64_t)co * Ci + ci) * taps + tap] = acc;
    }
}

inline bool umma_wgrad_geom(const b200_conv_desc* d) {
    if (!d->allow_umma || d->transposed) return false;
    if (d->x_dtype != B200_BF16 || d->y_dtype != B200_BF16) return false;
    if (d->sd != 1 || d->sh != 1 || d->sw != 1 || d->dd != 1 || d->dh != 1 || d->dw != 1) return false;
    if (!((d->kd == 1 || d->kd == 3) && (d->kh == 1 || d->kh == 3) && (d->kw == 1 || d->kw == 3))) return false;
    if (d->pd > d->kd - 1 || d->ph > d->kh - 1 || d->pw > d->kw - 1) return false;
    if (d->Ci % 16 || d->Co % 16 || d->Ci > 512 || d->Co > 512) return false;
    if ((int64_t)d->N * d->Do * d->Ho * d->Wo < 512) return false;
    return true;
}

inline int wgrad_plan(const b200_conv_desc* d, WgradParams* p, size_t* smem_bytes, size_t* partial_bytes) {
    memset(p, 0, sizeof *p);
    p->Ci = d->Ci; p->Co = d->Co;
    p->kd = d->kd; p->kh = d->kh; p->kw = d->kw; p->pd = d->pd; p->ph = d->ph; p->pw = d->pw;
    p->Hi = d->Hi; p->Wi = d->Wi; p->Ho = d->Ho; p->Wo = d->Wo;
    if (d->kd == 1) { p->NB = 1; p->Dpi = d->N * d->Di; p->Dpo = d->N * d->Do; }
    else { p->NB = d->N; p->Dpi = d->Di; p->Dpo = d->Do; }
    const int taps = d->kd * d->kh * d->kw;
    p->NC = d->Ci <= 64 ? d->Ci : (d->Ci % 64 == 0 ? 64 : (d->Ci % 48 == 0 ? 48 : (d->Ci % 32 == 0 ? 32 : 16)));
    p->MC = d->Co <= 128 ? d->Co : (d->Co % 128 == 0 ? 128 : (d->Co % 64 == 0 ? 64 : 16));
    p->n_ci = d->Ci / p->NC; p->n_co = d->Co / p->MC;
    p->tpu = 512 / p->NC;
    if (p->tpu > taps) p->tpu = taps;
    p->n_tg = (taps + p->tpu - 1) / p->tpu;
    p->tpu = (taps + p->n_tg - 1) / p->n_tg;                     // balance the groups
    int cols = p->tpu * p->NC, pow2 = 32;
    while (pow2 < cols) pow2 <<= 1;
    B200_REQUIRE(pow2 <= 512, "umma wgrad: accumulators do not fit TMEM");
    p->tmem_cols = pow2;
    p->HH = kTileH + d->kh - 1; p->WW = kTileW + d->kw - 1;
    p->use_tma = staging_uses_tma() ? 1 : 0;
    p->slab_cg_pitch = p->use_tma ? ((p->HH * p->WW * 16) + 127) & ~127 : p->HH * p->WW * 16 + 16;
    p->slab_bytes = (((p->NC / 8) * p->slab_cg_pitch) + 127) & ~127;
    p->dy_row_pitch = kTileW * 16;
    p->dy_cg_pitch = p->use_tma ? 128 * 16 : 128 * 16 + 16;
    p->dy_bytes = (((p->MC / 8) * p->dy_cg_pitch) + 127) & ~127;
    const size_t budget = 227 * 1024 - sizeof(WgradBarriers) - 1024;
    p->ndy = 3; p->nslabs = d->kd + 3;
    while (p->nslabs > d->kd + 1 && (size_t)p->ndy * p->dy_bytes + (size_t)p->nslabs * p->slab_bytes > budget) --p->nslabs;
    while (p->ndy > 2 && (size_t)p->ndy * p->dy_bytes + (size_t)p->nslabs * p->slab_bytes > budget) --p->ndy;
    size_t total = (size_t)p->ndy * p->dy_bytes + (size_t)p->nslabs * p->slab_bytes;
    B200_REQUIRE(total <= budget, "umma wgrad: tile does not fit shared memory");
    // an M=128 descriptor on a narrow co chunk reads 16 channel groups: keep that over-read inside the allocation
    const size_t overread_end = (size_t)(p->ndy - 1) * p->dy_bytes + (size_t)16 * p->dy_cg_pitch + 4096;
    if (total < overread_end) total = overread_end;
    *smem_bytes = sizeof(WgradBarriers) + total;
    p->tiles_y = (d->Ho + kTileH - 1) / kTileH;
    p->tiles_x = (d->Wo + kTileW - 1) / kTileW;
    const int units = p->n_tg * p->n_ci * p->n_co;
    int splits = kNumSMs / units;
    if (splits < 1) splits = 1;
    // column segments: enough items for every split, but long segments so that halo planes are reused along z
    const int64_t columns = (int64_t)p->NB * p->tiles_y * p->tiles_x;
    int zsegs = 1;
    while (columns * zsegs < (int64_t)splits * 2 && zsegs < p->Dpo) zsegs *= 2;
    p->zs = (p->Dpo + zsegs - 1) / zsegs;
    p->zsegs = (p->Dpo + p->zs - 1) / p->zs;
    const int64_t items = columns * p->zsegs;
    if (splits > items) splits = (int)items;
    p->splits = splits;
    p->items = (int)items;
    p->idesc = make_idesc_bf16(p->NC) | (1u << 15) | (1u << 16);            // A and B are MN-major
    *partial_bytes = (size_t)splits * taps * d->Co * d->Ci * sizeof(float);
    return 0;
}

inline size_t umma_wgrad_workspace_bytes(const b200_conv_desc* d) {
    WgradParams p;
    size_t smem = 0, part = 0;
    if (wgrad_plan(d, &p, &smem, &part)) return 0;
    int64_t chunks = ((int64_t)d->N * d->Do * d->Ho * d->Wo + 4095) / 4096;
    if (chunks > 2 * kNumSMs) chunks = 2 * kNumSMs;
    if (chunks < 1) chunks = 1;
    return part + (size_t)chunks * d->Co * 4 + 512;
}

inline int make_act_map(CUtensorMap* map, const void* ptr, int C, int W, int H, int64_t planes, int box_w, int box_h) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    B200_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is unavailable in this driver");
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {8, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d", (int)r);
    return 0;
}

inline int umma_wgrad_run(const b200_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, void* workspace, size_t ws_bytes,
                          void* stream) {
    B200_REQUIRE(umma_wgrad_geom(d), "umma wgrad: unsupported descriptor");
    B200_REQUIRE(aligned16(x) && aligned16(dy), "umma wgrad: pointers must be 16-byte aligned");
    WgradParams p;
    size_t smem_bytes = 0, partial_bytes = 0;
    if (wgrad_plan(d, &p, &smem_bytes, &partial_bytes)) return 1;
    B200_REQUIRE(ws_bytes >= umma_wgrad_workspace_bytes(d), "umma wgrad: workspace too small");
    p.partial = (float*)workspace;
    p.x = (const __nv_bfloat16*)x;
    p.dy = (const __nv_bfloat16*)dy;
    CUtensorMap x_map, dy_map;
    if (make_act_map(&x_map, x, d->Ci, d->Wi, d->Hi, (int64_t)d->N * d->Di, p.WW, p.HH)) return 1;
    if (make_act_map(&dy_map, dy, d->Co, d->Wo, d->Ho, (int64_t)d->N * d->Do, kTileW, kTileH)) return 1;
    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [] { attr_err = cudaFuncSetAttribute(conv_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    B200_REQUIRE(attr_err == cudaSuccess, "umma wgrad: cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(attr_err));
    const int units = p.n_tg * p.n_ci * p.n_co;
    B200_LAUNCH(conv_wgrad_umma_kernel, units * p.splits, kWgThreads, smem_bytes, stream, x_map, dy_map, p);
    const int taps = d->kd * d->kh * d->kw;
    const int64_t total = (int64_t)taps * d->Co * d->Ci;
    B200_LAUNCH(wgrad_umma_reduce_kernel, stream_grid(total, 256), 256, 0, stream, p.splits, taps, d->Co, d->Ci, p.partial, dw);
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
