// Microbenchmark: latencies that bound the phases of the bulge-chase kernels (one CTA, clock64 around loops).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_chain(double* out, long long* clk, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9, c = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) c = fma(c, a, b);
  }
  long long t1 = clock64();
  out[threadIdx.x] = c;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
template <int NCH>
__global__ void k_dfma_ilp(double* out, long long* clk, int iters) {
  double c[NCH];
  for (int i = 0; i < NCH; ++i) c[i] = threadIdx.x + i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) c[j] = fma(c[j], a, b);
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < NCH; ++i) s += c[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_lds_chain(double* out, long long* clk, int iters) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 1) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) p = idx[p];
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_bar(double* out, long long* clk, int iters) {
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_shfl_dadd(double* out, long long* clk, int iters) {
  double v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
// the pattern of a partial-sum phase: LDS.128 -> 4 DFMA -> ... (ne elements per thread, acc chains NA), STS, barrier
template <int NE, int NA>
__global__ void k_phase(double* out, long long* clk, int iters) {
  extern __shared__ double2 sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = make_double2(i * 1e-3, 1.0);
  __syncthreads();
  double2 acc[NA];
  for (int i = 0; i < NA; ++i) acc[i] = make_double2(0, 0);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const double2 a = sm[(threadIdx.x + e * 256) & 8191];
      const double2 v = sm[(e * 37 + it) & 8191];
      double2& c = acc[e % NA];
      c.x = fma(a.x, v.x, c.x); c.x = fma(-a.y, v.y, c.x);
      c.y = fma(a.x, v.y, c.y); c.y = fma(a.y, v.x, c.y);
    }
    sm[(threadIdx.x + it * 64) & 8191] = acc[0];
    __syncthreads();
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < NA; ++i) s += acc[i].x + acc[i].y;
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
  double* out; long long* clk; long long h;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&clk, 64);
  auto rd = [&]() { cudaDeviceSynchronize(); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); return (double)h; };
  k_dfma_chain<<<1, 32>>>(out, clk, 1000); printf("dependent DFMA latency: %.1f clk\n", rd() / 16000);
  for (int w : {1, 2, 4, 7, 8, 11, 16}) {
    k_dfma_ilp<1><<<1, 32 * w>>>(out, clk, 4000); double a = rd() / 4000;
    k_dfma_ilp<4><<<1, 32 * w>>>(out, clk, 4000); double b = rd() / 4000 / 4;
    k_dfma_ilp<10><<<1, 32 * w>>>(out, clk, 4000); double c = rd() / 4000 / 10;
    printf("warps %2d: clk per DFMA per thread: 1 chain %.2f, 4 chains %.2f, 10 chains %.2f  (per SM DFMA/clk: %.1f)\n", w, a, b, c, 32.0 * w / c);
  }
  k_lds_chain<<<1, 32>>>(out, clk, 10000); printf("dependent LDS latency: %.1f clk\n", rd() / 10000);
  for (int w : {1, 4, 7, 8, 11, 12, 16}) { k_bar<<<1, 32 * w>>>(out, clk, 10000); printf("bar.sync %2d warps: %.1f clk\n", w, rd() / 10000); }
  k_shfl_dadd<<<1, 32>>>(out, clk, 1000); printf("warp_sum (5 x shfl + dadd): %.1f clk\n", rd() / 1000);
  cudaFuncSetAttribute(k_phase<50, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
  cudaFuncSetAttribute(k_phase<30, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
  cudaFuncSetAttribute(k_phase<20, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
  k_phase<50, 5><<<1, 224, 8192 * 16>>>(out, clk, 200); printf("phase 7 warps x 50 elem (2 LDS.128 + 4 DFMA each): %.0f clk\n", rd() / 200);
  k_phase<30, 5><<<1, 352, 8192 * 16>>>(out, clk, 200); printf("phase 11 warps x 30 elem: %.0f clk\n", rd() / 200);
  k_phase<20, 4><<<1, 512, 8192 * 16>>>(out, clk, 200); printf("phase 16 warps x 20 elem: %.0f clk\n", rd() / 200);
  return 0;
}
