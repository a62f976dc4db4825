// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, operand layout and how many
// accumulators the issue stream cycles through.  Operand contents are irrelevant (zeros).  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <vector>

#include "../mri_epilepsy_diagnosis_b200/csrc/conv_simt.cuh"
#include "../mri_epilepsy_diagnosis_b200/csrc/conv_umma.cuh"

using namespace b200;

struct ProbeCfg {
    int n;            // UMMA N
    int layout;       // 0: no-swizzle, A SBO=160 LBO=2896 (conv halo slab); 1: no-swizzle dense SBO=128 LBO=2048; 2: SWIZZLE_128B dense
    int naccs;        // accumulators cycled through
    int shift;        // 1: vary the A start address per MMA like the conv taps do
    int iters;
};

template <int NACCS, int SHIFT>
__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeCfg c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(ptx::smem_u32(&tmem_base_s), 512); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x < 32) {
        const uint32_t a_base = ptx::smem_u32(smem), b_base = ptx::smem_u32(smem + 96 * 1024);
        uint32_t a_hi, b_hi, a_lo0, b_lo0;
        if (c.layout == 0) {
            a_hi = (160u >> 4) | (1u << 14); a_lo0 = ((a_base >> 4) & 0x3FFF) | ((2896u >> 4) << 16);
            b_hi = (128u >> 4) | (1u << 14); b_lo0 = ((b_base >> 4) & 0x3FFF) | (((uint32_t)c.n * 16 >> 4) << 16);
        } else if (c.layout == 1) {
            a_hi = (128u >> 4) | (1u << 14); a_lo0 = ((a_base >> 4) & 0x3FFF) | ((2048u >> 4) << 16);
            b_hi = (128u >> 4) | (1u << 14); b_lo0 = ((b_base >> 4) & 0x3FFF) | (((uint32_t)c.n * 16 >> 4) << 16);
        } else if (c.layout == 2) {
            a_hi = (1024u >> 4) | (1u << 14) | (2u << 29); a_lo0 = ((a_base >> 4) & 0x3FFF) | (1u << 16);
            b_hi = (1024u >> 4) | (1u << 14) | (2u << 29); b_lo0 = ((b_base >> 4) & 0x3FFF) | (1u << 16);
        } else {
            // 3: SWIZZLE_32B rows (32 B per voxel, SBO 256), 4: SWIZZLE_64B rows (SBO 512), 5: SWIZZLE_128B rows (SBO 1024) -- A as the
            // row-slab kernels stage it; B = no-swizzle packed weights.  shift=1 moves the A start by whole 32-byte steps like the taps do
            const uint32_t sbo = c.layout == 3 ? 256u : (c.layout == 4 ? 512u : 1024u), lbits = c.layout == 3 ? 6u : (c.layout == 4 ? 4u : 2u);
            a_hi = (sbo >> 4) | (1u << 14) | (lbits << 29); a_lo0 = ((a_base >> 4) & 0x3FFF) | (1u << 16);
            b_hi = (128u >> 4) | (1u << 14); b_lo0 = ((b_base >> 4) & 0x3FFF) | (((uint32_t)c.n * 16 >> 4) << 16);
        }
        const uint32_t idesc = make_idesc_bf16(c.n);
        __syncwarp();
        const long long t0 = clock64();
        for (int it = 0; it < c.iters; ++it) {
            if (ptx::elect_one()) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const uint32_t acc = (uint32_t)(j % NACCS);
                    const uint32_t a_lo = a_lo0 + (SHIFT ? (c.layout >= 3 ? (uint32_t)((j % 9) / 3 * 260 + (j % 3) * 2) : (uint32_t)((j % 9) / 3 * 10 + (j % 3))) : 0u);
                    ptx::umma_bf16_lohi(tmem + acc * (uint32_t)c.n, a_lo, a_hi, b_lo0, b_hi, idesc, 1);
                }
            }
            __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit(ptx::smem_u32(&bar));
        __syncwarp();
        ptx::mbar_wait(ptx::smem_u32(&bar), 0);
        const long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 148 * sizeof(long long));
    const int iters = 400;
    printf("%-8s %-5s %-6s %-6s %12s\n", "layout", "N", "naccs", "shift", "cyc/UMMA");
    const char* names[6] = {"ns160", "ns128", "sw128", "row32", "row64", "row128"};
    for (int layout = 0; layout < 6; ++layout)
        for (int n : {16, 32, 48, 64, 96, 128, 144, 192, 256})
            for (int naccs : {1, 2, 4})
                for (int shift : {0, 1}) {
                    if (naccs * n > 512) continue;
                    if (layout != 0 && layout < 3 && shift) continue;
                    if (layout >= 3 && ((n > 64 && layout != 3) || naccs == 2)) continue;
                    ProbeCfg c{n, layout, naccs, shift, iters};
                    auto launch = [&](auto kern) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); kern<<<148, 128, 200 * 1024>>>(c, d_out); };
                    if (naccs == 1 && !shift) launch(probe_kernel<1, 0>);
                    else if (naccs == 1) launch(probe_kernel<1, 1>);
                    else if (naccs == 2 && !shift) launch(probe_kernel<2, 0>);
                    else if (naccs == 2) launch(probe_kernel<2, 1>);
                    else if (!shift) launch(probe_kernel<4, 0>);
                    else launch(probe_kernel<4, 1>);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("%s N=%d: %s\n", names[layout], n, cudaGetErrorString(e)); return 1; }
                    std::vector<long long> h(148);
                    cudaMemcpy(h.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
                    double avg = 0;
                    for (auto v : h) avg += (double)v;
                    avg /= 148.0 * iters * 32;
                    printf("%-8s %-5d %-6d %-6d %12.1f\n", names[layout], n, naccs, shift, avg);
                }
    return 0;
}
