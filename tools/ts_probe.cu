// Probe for the "A operand from TMEM" formulation of the small-channel convolutions (run on a B200):
//   C1  correctness: a 128 x 16 bf16 A tile in the SWIZZLE_32B K-major row image (what conv_rowf.cuh stages by TMA) is
//       copied to TMEM with tcgen05.cp.128x256b (optionally from a voxel-shifted start address) and used as the A operand of
//       tcgen05.mma ([d], [a_tmem], b_desc); the result must equal the shared-memory-operand MMA on the same data.
//   T1  cycles per TS MMA (A in TMEM) vs SS MMA (A in shared memory) for N = 16/32/64/128
//   T2  cycles per tcgen05.cp.128x256b, alone and interleaved 1 cp : k MMAs
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ts_probe tools/ts_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mri_epilepsy_diagnosis_b200/csrc/conv_simt.cuh"
#include "../mri_epilepsy_diagnosis_b200/csrc/conv_umma.cuh"

using namespace b200;

__device__ __forceinline__ void utccp_128x256b(uint32_t taddr, uint32_t lo, uint32_t hi) {
    asm volatile("{\n\t.reg .b64 d;\n\tmov.b64 d, {%1, %2};\n\ttcgen05.cp.cta_group::1.128x256b [%0], d;\n\t}" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}

struct Cfg { int n, mode, iters, ratio, shift; };   // mode 0: correctness, 1: SS timing, 2: TS timing, 3: cp timing, 4: 1 cp : ratio TS MMAs

// smem: A image (SW32 rows of 32 B, 256 voxels) at 0; B (no-swizzle K-major [cg2][n][8]) at 16 KB
__global__ void __launch_bounds__(128, 1) ts_kernel(Cfg c, const __nv_bfloat16* a_img, const __nv_bfloat16* b_pk, float* out_ss, float* out_ts, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    // A image: voxel v (32 B = 16 bf16) at v*32 with the SWIZZLE_32B pattern: 16-byte chunk index ^= (address bit 7)
    for (int i = threadIdx.x; i < 256 * 2; i += blockDim.x) {
        const int v = i >> 1, chunk = i & 1;
        const uint32_t off = (uint32_t)v * 32;
        const uint32_t sw = off + (uint32_t)((chunk ^ ((off >> 7) & 1)) * 16);
        *reinterpret_cast<uint4*>(smem + sw) = *reinterpret_cast<const uint4*>(a_img + v * 16 + chunk * 8);
    }
    for (int i = threadIdx.x; i < 2 * c.n; i += blockDim.x)
        *reinterpret_cast<uint4*>(smem + 16384 + i * 16) = *reinterpret_cast<const uint4*>(b_pk + i * 8);
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(ptx::smem_u32(&tmem_base_s), 512); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t a_addr = ptx::smem_u32(smem) + (uint32_t)c.shift * 32, b_addr = ptx::smem_u32(smem + 16384);
    const uint32_t a_lo = ((a_addr >> 4) & 0x3FFF) | (1u << 16);
    const uint32_t a_hi = (256u >> 4) | (1u << 14) | (6u << 29);                 // SBO = 8 voxels, SWIZZLE_32B
    const uint32_t b_lo = ((b_addr >> 4) & 0x3FFF) | (((uint32_t)c.n * 16 >> 4) << 16);
    const uint32_t b_hi = (128u >> 4) | (1u << 14);
    const uint32_t idesc = make_idesc_bf16(c.n);
    const uint32_t d_ss = tmem, d_ts = tmem + 128, a_tm = tmem + 256;           // columns
    if (threadIdx.x < 32) {
        __syncwarp();
        const long long t0 = clock64();
        if (c.mode == 0) {
            if (ptx::elect_one()) {
                ptx::umma_bf16_lohi(d_ss, a_lo, a_hi, b_lo, b_hi, idesc, 0);
                utccp_128x256b(a_tm, a_lo, a_hi);
                umma_ts_bf16(d_ts, a_tm, b_lo, b_hi, idesc, 0);
            }
        } else {
            for (int it = 0; it < c.iters; ++it) {
                if (ptx::elect_one()) {
                    if (c.mode == 1) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) ptx::umma_bf16_lohi(d_ss, a_lo + (uint32_t)(j & 3) * 2, a_hi, b_lo, b_hi, idesc, 1);
                    } else if (c.mode == 2) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) umma_ts_bf16(d_ts, a_tm + (uint32_t)(j & 3) * 8, b_lo, b_hi, idesc, 1);
                    } else if (c.mode == 3) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) utccp_128x256b(a_tm + (uint32_t)(j & 7) * 8, a_lo + (uint32_t)(j & 3) * 2, a_hi);
                    } else {
                        for (int g = 0; g < 4; ++g) {
                            utccp_128x256b(a_tm + (uint32_t)(g & 3) * 8, a_lo + (uint32_t)g * 2, a_hi);
                            for (int j = 0; j < c.ratio; ++j) umma_ts_bf16(d_ts + (uint32_t)(j % 3) * (uint32_t)c.n, a_tm + (uint32_t)(g & 3) * 8, b_lo, b_hi, idesc, 1);
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (ptx::elect_one()) ptx::umma_commit(ptx::smem_u32(&bar));
        __syncwarp();
        ptx::mbar_wait(ptx::smem_u32(&bar), 0);
        const long long t1 = clock64();
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    ptx::tc_fence_after();
    if (c.mode == 0 && blockIdx.x == 0) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int c0 = 0; c0 < c.n; c0 += 16) {
            float v[16];
            ptx::tmem_ld16(d_ss + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            for (int i = 0; i < 16; ++i) out_ss[(warp * 32 + lane) * 256 + c0 + i] = v[i];
            ptx::tmem_ld16(d_ts + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            for (int i = 0; i < 16; ++i) out_ts[(warp * 32 + lane) * 256 + c0 + i] = v[i];
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
    srand(3);
    std::vector<__nv_bfloat16> ha(256 * 16), hb(2 * 256 * 8);
    std::vector<float> fa(256 * 16);
    for (size_t i = 0; i < ha.size(); ++i) { fa[i] = (float)(rand() % 9 - 4); ha[i] = __float2bfloat16(fa[i]); }
    __nv_bfloat16 *da, *db; float *dss, *dts; long long* dcyc;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dss, 128 * 256 * 4); cudaMalloc(&dts, 128 * 256 * 4); cudaMalloc(&dcyc, 148 * 8);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int n : {16, 32, 64}) {
        std::vector<float> fw(n * 16);
        for (auto& x : fw) x = (float)(rand() % 5 - 2);
        for (int cg = 0; cg < 2; ++cg) for (int r = 0; r < n; ++r) for (int j = 0; j < 8; ++j) hb[(cg * n + r) * 8 + j] = __float2bfloat16(fw[r * 16 + cg * 8 + j]);
        cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
        for (int shift : {0, 1, 5}) {
            Cfg c{n, 0, 1, 0, shift};
            ts_kernel<<<1, 128, 64 * 1024>>>(c, da, db, dss, dts, dcyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("C1 N=%d: %s\n", n, cudaGetErrorString(e)); return 1; }
            std::vector<float> hss(128 * 256), hts(128 * 256);
            cudaMemcpy(hss.data(), dss, hss.size() * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hts.data(), dts, hts.size() * 4, cudaMemcpyDeviceToHost);
            int bad_ss = 0, bad_ts = 0;
            for (int m = 0; m < 128; ++m) for (int r = 0; r < n; ++r) {
                double ref = 0;
                for (int k = 0; k < 16; ++k) ref += (double)fa[(m + shift) * 16 + k] * fw[r * 16 + k];
                if (fabs(hss[m * 256 + r] - ref) > 1e-3) ++bad_ss;
                if (fabs(hts[m * 256 + r] - ref) > 1e-3) ++bad_ts;
            }
            printf("C1 N=%d shift=%d: SS mismatches %d, TS (tcgen05.cp -> A in TMEM) mismatches %d\n", n, shift, bad_ss, bad_ts);
        }
    }
    const int iters = 400;
    for (int n : {16, 32, 64, 128}) {
        for (int mode : {1, 2, 3}) {
            Cfg c{n, mode, iters, 0, 0};
            ts_kernel<<<148, 128, 64 * 1024>>>(c, da, db, dss, dts, dcyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("T N=%d mode=%d: %s\n", n, mode, cudaGetErrorString(e)); return 1; }
            std::vector<long long> h(148); cudaMemcpy(h.data(), dcyc, 148 * 8, cudaMemcpyDeviceToHost);
            double avg = 0; for (auto v : h) avg += (double)v; avg /= 148.0 * iters * 16;
            printf("T%d N=%-4d %s: %.1f cycles per op\n", mode, n, mode == 1 ? "SS MMA (A in smem)" : mode == 2 ? "TS MMA (A in TMEM)" : "tcgen05.cp 128x256b", avg);
        }
        for (int ratio : {3, 9}) {
            if (3 * n > 128) continue;
            Cfg c{n, 4, iters, ratio, 0};
            ts_kernel<<<148, 128, 64 * 1024>>>(c, da, db, dss, dts, dcyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("T4 N=%d: %s\n", n, cudaGetErrorString(e)); return 1; }
            std::vector<long long> h(148); cudaMemcpy(h.data(), dcyc, 148 * 8, cudaMemcpyDeviceToHost);
            double avg = 0; for (auto v : h) avg += (double)v; avg /= 148.0 * iters * 4 * ratio;
            printf("T4 N=%-4d 1 cp : %d TS MMAs: %.1f cycles per MMA (cp included)\n", n, ratio, avg);
        }
    }
    return 0;
}
