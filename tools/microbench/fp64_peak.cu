// Microbenchmark: sustained FP64 rate of DMMA (mma.sync.m8n8k4.f64) vs DFMA on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_kernel(double* out, int iters) {
  double c0[NACC], c1[NACC];
  for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_kernel(double* out, int iters) {
  double c[NACC];
  for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x * 1e-3 + i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int blocks_per_sm : {1, 2}) {
      if (threads * blocks_per_sm > 2048) continue;
      float ms;
      dmma_kernel<8><<<sms * blocks_per_sm, threads>>>(out, 100);
      cudaEventRecord(e0); dmma_kernel<8><<<sms * blocks_per_sm, threads>>>(out, iters); cudaEventRecord(e1);
      cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      double fl = (double)sms * blocks_per_sm * (threads / 32) * iters * 8.0 * 512.0;
      printf("DMMA  threads=%4d blocks/SM=%d : %.2f TFLOP/s\n", threads, blocks_per_sm, fl / (ms * 1e-3) / 1e12);
      dfma_kernel<8><<<sms * blocks_per_sm, threads>>>(out, 100);
      cudaEventRecord(e0); dfma_kernel<8><<<sms * blocks_per_sm, threads>>>(out, iters); cudaEventRecord(e1);
      cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      fl = (double)sms * blocks_per_sm * threads * iters * 8.0 * 2.0;
      printf("DFMA  threads=%4d blocks/SM=%d : %.2f TFLOP/s\n", threads, blocks_per_sm, fl / (ms * 1e-3) / 1e12);
    }
  }
  printf("SMs %d, clock %d kHz\n", sms, p.clockRate);
  return 0;
}
