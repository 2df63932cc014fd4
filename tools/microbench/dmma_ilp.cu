// Microbenchmark: DMMA (mma.sync.m8n8k4.f64) throughput as a function of warps per SM and independent accumulator
// chains per warp (what the register-resident back-transformation kernel can count on with 8 warps per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_ilp dmma_ilp.cu && ./dmma_ilp
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k(double* out, int iters) {
  double c0[NACC], c1[NACC];
  for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16 / NACC; ++r)
#pragma unroll
      for (int i = 0; i < NACC; ++i) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
void run(int sms, double* out, int warps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  float ms;
  k<NACC><<<sms, warps * 32>>>(out, 10);
  cudaEventRecord(e0); k<NACC><<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1);
  cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)iters * 16;          // DMMAs per warp
  const double fl = (double)sms * warps * n * 512.0;
  printf("warps/SM %2d  chains %d : %6.2f TFLOP/s   %.1f clk per DMMA per warp (at 1.965 GHz)\n", warps, NACC,
         fl / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.965e9 / n);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
  for (int warps : {1, 4, 8, 16, 32}) {
    run<1>(sms, out, warps); run<2>(sms, out, warps); run<4>(sms, out, warps); run<8>(sms, out, warps); run<16>(sms, out, warps);
  }
  return 0;
}
