// Microbenchmark: DFMA throughput with three distinct register operands per instruction (the pattern of a
// complex multiply-add on register-resident blocks) vs the two-constant pattern of fp64_peak.cu.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_ops dfma_ops.cu && ./dfma_ops
#include <cstdio>
#include <cuda_runtime.h>
template <int NE>
__global__ void k_cfma(double* out, long long* clk, int iters) {
  double2 a[NE], acc[5];
  for (int i = 0; i < NE; ++i) a[i] = make_double2(threadIdx.x + i, 0.5 * i);
  for (int i = 0; i < 5; ++i) acc[i] = make_double2(0, 0);
  double2 v = make_double2(1.0 + threadIdx.x * 1e-9, 1e-9);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      double2& c = acc[e % 5];
      c.x = fma(a[e].x, v.x, c.x); c.x = fma(-a[e].y, v.y, c.x);
      c.y = fma(a[e].x, v.y, c.y); c.y = fma(a[e].y, v.x, c.y);
    }
    v.x += 1e-9;
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 5; ++i) s += acc[i].x + acc[i].y;
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
// rank-1 update pattern: a[e] -= t[q] * w[c]  (all operands distinct registers, result written back)
template <int NR, int NCOL>
__global__ void k_upd(double* out, long long* clk, int iters) {
  double2 a[NR][NCOL], t[NR], w[NCOL];
  for (int i = 0; i < NR; ++i) { t[i] = make_double2(1e-3 * i, 1e-4); for (int j = 0; j < NCOL; ++j) a[i][j] = make_double2(i + threadIdx.x, j); }
  for (int j = 0; j < NCOL; ++j) w[j] = make_double2(1e-3 * j, 1e-5 * threadIdx.x);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int j = 0; j < NCOL; ++j) {
        double2& c = a[i][j];
        c.x = fma(-t[i].x, w[j].x, c.x); c.x = fma(t[i].y, w[j].y, c.x);
        c.y = fma(-t[i].x, w[j].y, c.y); c.y = fma(-t[i].y, w[j].x, c.y);
      }
    w[0].x += 1e-9;
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < NR; ++i) for (int j = 0; j < NCOL; ++j) s += a[i][j].x + a[i][j].y;
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  double* out; long long* clk; long long h;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&clk, 64);
  auto rd = [&]() { cudaDeviceSynchronize(); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); return (double)h; };
  for (int w : {4, 7, 8, 11, 12, 16}) {
    k_cfma<30><<<1, 32 * w>>>(out, clk, 200); double a = rd() / 200;
    k_upd<5, 6><<<1, 32 * w>>>(out, clk, 200); double b = rd() / 200;
    printf("warps %2d: matvec pattern %.1f DFMA/clk/SM, rank-1 update pattern %.1f DFMA/clk/SM\n", w, 120.0 * 32 * w / a, 120.0 * 32 * w / b);
  }
  return 0;
}
