// tcgen05 / TMEM / TMA implicit-GEMM convolution (bf16 in, fp32 accumulate).  See DESIGN.md.
#pragma once
#include "common.cuh"

namespace b200 {

inline bool umma_conv_supported(const b200_conv_desc* d, int pass) { (void)d; (void)pass; return false; }
inline size_t umma_packed_bytes(const b200_conv_desc*, int) { return 0; }
inline size_t umma_workspace_bytes(const b200_conv_desc*, int) { return 0; }
inline int umma_pack_weights(const b200_conv_desc*, int, const float*, void*, void*) { return fail("umma path not built"); }
inline int umma_conv_run(const b200_conv_desc*, int, const void*, const void*, const float*, void*, void*, size_t, void*) { return fail("umma path not built"); }
inline int umma_wgrad_run(const b200_conv_desc*, const void*, const void*, float*, float*, void*, size_t, void*) { return fail("umma path not built"); }

}  // namespace b200
