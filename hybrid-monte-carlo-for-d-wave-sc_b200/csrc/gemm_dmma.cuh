// gemm_dmma.cuh -- device helpers shared by the FP64 tensor-core (DMMA) GEMM kernels.
// FP64 has no tcgen05/wgmma kind on sm_100a; the FP64 tensor path is mma.sync m8n8k4
// (SASS DMMA.8x8x4).  Tiles are staged global -> shared with 16-byte cp.async
// (zero-fill on the edges), three stages deep, and fed to the DMMA from padded,
// bank-conflict-free shared-memory layouts.
#pragma once
#include <cuda_runtime.h>

namespace dwg {

// D(8x8) += A(8x4, row) * B(4x8, col), all f64.
// fragment layout: a: row = lane/4, k = lane%4 ; b: k = lane%4, col = lane/4 ;
//                  c0,c1: row = lane/4, col = 2*(lane%4) + {0,1}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace dwg
