// Correctness + throughput probe for the shared-memory descriptor tricks the v2 conv kernels rely on (run on a B200):
//   T1  MN-major SWIZZLE_32B/64B operands whose M/N "atom" stride (LBO) is ONE VOXEL (A) or ONE ROW (B):
//       D[(s, c), (t, c')] = sum_v P[v + s][c] * Q[t][v][c']      -- the wgrad "shifted operand" formulation
//   T2  K-major SWIZZLE_32B/64B A operand (rows = voxels, C channels contiguous) against K-major no-swizzle packed weights,
//       with the A start address advanced by whole voxels (row shift) and by 32 B (K step inside a 64 B row)
//   T3  TMA tensor-load throughput for [C x 128 voxels] row boxes with 32/64/128-byte inner rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/desc_probe tools/desc_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mri_epilepsy_diagnosis_b200/csrc/conv_simt.cuh"
#include "../mri_epilepsy_diagnosis_b200/csrc/conv_umma.cuh"

using namespace b200;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

static int make_map_2d(CUtensorMap* map, const void* ptr, int C, int64_t rows, int box_rows, CUtensorMapSwizzle sw) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
    const cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    return 0;
}

struct T1Cfg { int C, S, T, layout_bits, mode; int rowpitch; };   // mode 0: T1 (MN-major both), 1: T2 (K-major A, weights B)

// smem: P at 0 (up to 144 voxels), Q at 32 KB (3 rows x 128 voxels), weights at 96 KB
__global__ void __launch_bounds__(128, 1) t1_kernel(const __grid_constant__ CUtensorMap pmap, const __grid_constant__ CUtensorMap qmap,
                                                    T1Cfg c, const __nv_bfloat16* wpk, float* out, int shift_vox) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, done;
    __shared__ uint32_t tmem_base_s;
    uint8_t* P = smem; uint8_t* Q = smem + 32 * 1024; uint8_t* Wt = smem + 96 * 1024;
    const int vox_bytes = c.C * 2;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::mbar_init(ptx::smem_u32(&done), 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(ptx::smem_u32(&tmem_base_s), 256); ptx::tmem_relinquish(); }
    if (c.mode == 1) {   // weights: [kstep][cg2][n][8] no-swizzle K-major, N = c.T * c.C rows... here N = c.T*16 generic
        const int total = (c.C / 16) * 2 * (c.T * 16) * 8;
        for (int i = threadIdx.x; i < total; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(Wt)[i] = wpk[i];
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        ptx::mbar_expect_tx(ptx::smem_u32(&bar), (uint32_t)(144 * vox_bytes + 3 * 128 * vox_bytes));
        tma_load_2d(ptx::smem_u32(P), &pmap, ptx::smem_u32(&bar), 0, 0);
        for (int t = 0; t < 3; ++t) tma_load_2d(ptx::smem_u32(Q + t * c.rowpitch), &qmap, ptx::smem_u32(&bar), 0, t * 128);
        ptx::mbar_wait(ptx::smem_u32(&bar), 0);
        ptx::tc_fence_after();
        const uint32_t lt = (uint32_t)c.layout_bits << 29;
        if (c.mode == 0) {
            const int N = c.T * c.C;
            const uint32_t idesc = make_idesc_bf16(N) | (1u << 15) | (1u << 16);
            const uint32_t kgroup = 8 * vox_bytes;        // 8 voxels along K
            for (int ks = 0; ks < 8; ++ks) {
                const uint32_t a_addr = ptx::smem_u32(P) + (uint32_t)(ks * 16 * vox_bytes);
                const uint32_t b_addr = ptx::smem_u32(Q) + (uint32_t)(ks * 16 * vox_bytes);
                const uint32_t a_lo = ((a_addr >> 4) & 0x3FFF) | (((uint32_t)vox_bytes >> 4) << 16);        // LBO = one voxel
                const uint32_t b_lo = ((b_addr >> 4) & 0x3FFF) | (((uint32_t)c.rowpitch >> 4) << 16);       // LBO = one row
                const uint32_t hi = (kgroup >> 4) | (1u << 14) | lt;                                      // SBO = next 8 voxels
                ptx::umma_bf16_lohi(tmem, a_lo, hi, b_lo, hi, idesc, ks > 0);
            }
        } else {
            // D[v, n] = sum_c P[v + shift][c] * W[n][c];  A K-major swizzled (SBO = 8 voxels), B no-swizzle K-major
            const int N = c.T * 16;
            const uint32_t idesc = make_idesc_bf16(N);
            for (int ks = 0; ks < c.C / 16; ++ks) {
                const uint32_t a_addr = ptx::smem_u32(P) + (uint32_t)(shift_vox * vox_bytes + ks * 32);
                const uint32_t a_lo = ((a_addr >> 4) & 0x3FFF) | (1u << 16);
                const uint32_t a_hi = ((uint32_t)(8 * vox_bytes) >> 4) | (1u << 14) | lt;
                const uint32_t b_addr = ptx::smem_u32(Wt) + (uint32_t)(ks * 2 * N * 16);
                const uint32_t b_lo = ((b_addr >> 4) & 0x3FFF) | (((uint32_t)N * 16 >> 4) << 16);           // LBO = next 8 channels
                const uint32_t b_hi = (128u >> 4) | (1u << 14);
                ptx::umma_bf16_lohi(tmem, a_lo, a_hi, b_lo, b_hi, idesc, ks > 0);
            }
        }
        ptx::umma_commit(ptx::smem_u32(&done));
    }
    __syncwarp();
    ptx::mbar_wait(ptx::smem_u32(&done), 0);
    ptx::tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncols = c.mode == 0 ? c.T * c.C : c.T * 16;
    for (int c0 = 0; c0 < ncols; c0 += 16) {
        float v[16];
        ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 256 + c0 + i] = v[i];
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 256); }
}

// ---------------------------------------------------------------------------------------------- T3: TMA row-box throughput
__global__ void __launch_bounds__(128, 1) t3_kernel(const __grid_constant__ CUtensorMap map, int C, int rows_per_box, int nbox, int64_t total_rows,
                                                    long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[4];
    const int box_bytes = C * 2 * rows_per_box;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) ptx::mbar_init(ptx::smem_u32(&full[i]), 1); ptx::fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        int64_t row = (int64_t)blockIdx.x * rows_per_box;
        for (int i = 0; i < nbox + 4; ++i) {
            if (i >= 4) ptx::mbar_wait(ptx::smem_u32(&full[i & 3]), ((i - 4) >> 2) & 1);
            if (i < nbox) {
                ptx::mbar_expect_tx(ptx::smem_u32(&full[i & 3]), (uint32_t)box_bytes);
                tma_load_2d(ptx::smem_u32(smem + (i & 3) * 32768), &map, ptx::smem_u32(&full[i & 3]), 0, (int)(row % total_rows));
                row += (int64_t)gridDim.x * rows_per_box;
            }
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
    srand(1);
    // ------------------------------------------------------------------ T1 / T2
    for (int C : {16, 32, 64}) {
        const int S = 128 / C, T = 3;
        const int vox_bytes = C * 2;
        std::vector<__nv_bfloat16> hp(144 * C), hq(3 * 128 * C);
        std::vector<float> fp(144 * C), fq(3 * 128 * C);
        for (size_t i = 0; i < hp.size(); ++i) { fp[i] = (float)(rand() % 7 - 3); hp[i] = __float2bfloat16(fp[i]); }
        for (size_t i = 0; i < hq.size(); ++i) { fq[i] = (float)(rand() % 5 - 2); hq[i] = __float2bfloat16(fq[i]); }
        __nv_bfloat16 *dp, *dq, *dw; float* dout;
        cudaMalloc(&dp, hp.size() * 2); cudaMalloc(&dq, hq.size() * 2); cudaMalloc(&dout, 128 * 256 * 4); cudaMalloc(&dw, 64 * 1024);
        cudaMemcpy(dp, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice);
        const CUtensorMapSwizzle sw = C == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : (C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
        const int lbits = C == 16 ? 6 : (C == 32 ? 4 : 2);
        CUtensorMap pmap, qmap;
        if (make_map_2d(&pmap, dp, C, 144, 144, sw) || make_map_2d(&qmap, dq, C, 3 * 128, 128, sw)) return 1;
        cudaFuncSetAttribute(t1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        std::vector<float> hout(128 * 256);
        {
            T1Cfg cfg{C, S, T, lbits, 0, 128 * vox_bytes};
            if (T * C <= 256) {
                t1_kernel<<<1, 128, 160 * 1024>>>(pmap, qmap, cfg, dw, dout, 0);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("T1 C=%d: %s\n", C, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
                int bad = 0; double maxerr = 0;
                for (int s = 0; s < S; ++s) for (int ch = 0; ch < C; ++ch) for (int t = 0; t < T; ++t) for (int c2 = 0; c2 < C; ++c2) {
                    double ref = 0;
                    for (int v = 0; v < 128; ++v) ref += (double)fp[(v + s) * C + ch] * fq[(t * 128 + v) * C + c2];
                    const double got = hout[(s * C + ch) * 256 + t * C + c2];
                    const double err = fabs(got - ref); if (err > maxerr) maxerr = err; if (err > 1e-3) ++bad;
                }
                printf("T1 MN-major shifted operands C=%d (M=%dx%d, N=%dx%d): max err %.3g, %d mismatches\n", C, S, C, T, C, maxerr, bad);
            }
        }
        // T2: K-major A with voxel shift, weights N = 48
        for (int shift : {0, 1, 3, 9}) {
            const int N = T * 16;
            std::vector<float> fw(N * C);
            std::vector<__nv_bfloat16> hw((C / 16) * 2 * N * 8);
            for (auto& x : fw) x = (float)(rand() % 5 - 2);
            for (int ks = 0; ks < C / 16; ++ks) for (int cg = 0; cg < 2; ++cg) for (int n = 0; n < N; ++n) for (int j = 0; j < 8; ++j)
                hw[((ks * 2 + cg) * N + n) * 8 + j] = __float2bfloat16(fw[n * C + ks * 16 + cg * 8 + j]);
            cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
            T1Cfg cfg{C, S, T, lbits, 1, 128 * vox_bytes};
            t1_kernel<<<1, 128, 160 * 1024>>>(pmap, qmap, cfg, dw, dout, shift);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("T2 C=%d: %s\n", C, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0; double maxerr = 0;
            for (int v = 0; v < 128; ++v) for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int ch = 0; ch < C; ++ch) ref += (double)fp[(v + shift) * C + ch] * fw[n * C + ch];
                const double err = fabs(hout[v * 256 + n] - ref); if (err > maxerr) maxerr = err; if (err > 1e-3) ++bad;
            }
            printf("T2 K-major swizzled A C=%d, start shifted by %d voxels: max err %.3g, %d mismatches\n", C, shift, maxerr, bad);
        }
        cudaFree(dp); cudaFree(dq); cudaFree(dout); cudaFree(dw);
    }
    // ------------------------------------------------------------------ T3
    {
        const int64_t rows = 4ll * 128 * 128 * 128;     // voxels of a 4 x 128^3 tensor
        long long* d_out; cudaMalloc(&d_out, 148 * 8);
        cudaFuncSetAttribute(t3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        for (int C : {16, 32, 64}) {
            __nv_bfloat16* d; cudaMalloc(&d, rows * C * 2); cudaMemset(d, 0, rows * C * 2);
            const CUtensorMapSwizzle sw = C == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : (C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
            for (int rpb : {128, 256}) {
                if (rpb * C * 2 > 32768) continue;
                CUtensorMap map;
                if (make_map_2d(&map, d, C, rows, rpb, sw)) return 1;
                const int nbox = 2000;
                t3_kernel<<<148, 128, 160 * 1024>>>(map, C, rpb, nbox, rows, d_out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("T3: %s\n", cudaGetErrorString(e)); return 1; }
                std::vector<long long> h(148); cudaMemcpy(h.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost);
                double avg = 0; for (auto v : h) avg += (double)v; avg /= 148;
                printf("T3 TMA row boxes C=%d (%d B inner) x %d rows: %.1f cycles/box, %.1f B/cycle/SM, %.2f cycles/row\n", C, C * 2, rpb, avg / nbox,
                       (double)rpb * C * 2 * nbox / avg, avg / nbox / rpb);
            }
            cudaFree(d);
        }
    }
    return 0;
}
