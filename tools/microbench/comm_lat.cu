// Microbenchmark: round trips a polling helper warp pays -- bulk-copy (TMA) fetch of 16 B / 1.6 KB from L2, relaxed LDG,
// fence + release store -- and the one-way latency of a flag between two SMs (ping-pong).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o comm_lat comm_lat.cu && ./comm_lat
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok = 0;
  do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory"); } while (!ok);
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int ld_relaxed(const int* p) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int ld_acquire(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed(int* p, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void k_local(int* g, long long* clk, int iters) {
  __shared__ __align__(128) unsigned char buf[2048];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned ph = 0; long long t0, t1;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { asm volatile("fence.proxy.async;" ::: "memory"); mbar_expect_tx(&bar, 16); bulk_load(buf, g, 16, &bar); mbar_wait(&bar, ph); ph ^= 1; }
    t1 = clock64(); clk[0] = (t1 - t0) / iters;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { asm volatile("fence.proxy.async;" ::: "memory"); mbar_expect_tx(&bar, 1616); bulk_load(buf, g, 1600, &bar); bulk_load(buf + 1600, g + 512, 16, &bar); mbar_wait(&bar, ph); ph ^= 1; }
    t1 = clock64(); clk[1] = (t1 - t0) / iters;
    int acc = 0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) acc += ld_relaxed(g + (acc & 1));
    t1 = clock64(); clk[2] = (t1 - t0) / iters;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) acc += ld_acquire(g + (acc & 1));
    t1 = clock64(); clk[3] = (t1 - t0) / iters;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { __threadfence(); st_release(g + 64, i); }
    t1 = clock64(); clk[4] = (t1 - t0) / iters;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { st_release(g + 64, i); }
    t1 = clock64(); clk[5] = (t1 - t0) / iters;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { __threadfence(); st_relaxed(g + 64, i); }
    t1 = clock64(); clk[6] = (t1 - t0) / iters;
    g[100] = acc;
  }
}
// ping-pong between block 0 and block 1 (different SMs): round trip / 2 = one-way flag latency
template <int MODE>
__global__ void k_pingpong(int* g, long long* clk, int iters) {
  __shared__ __align__(16) int pollbuf[4];
  __shared__ unsigned long long bar;
  if (threadIdx.x != 0) return;
  mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  unsigned ph = 0;
  int* mine = g + 32 * blockIdx.x; int* other = g + 32 * (1 - blockIdx.x);
  auto wait_ge = [&](int* f, int need) {
    if (MODE == 0) { while (ld_relaxed(f) < need) {} asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
    else { for (;;) { asm volatile("fence.proxy.async;" ::: "memory"); mbar_expect_tx(&bar, 16); bulk_load(pollbuf, f, 16, &bar); mbar_wait(&bar, ph); ph ^= 1; if (((volatile int*)pollbuf)[0] >= need) break; } }
  };
  long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    if (blockIdx.x == 0) { st_release(mine, i); wait_ge(other, i); }
    else { wait_ge(other, i); st_release(mine, i); }
  }
  long long t1 = clock64();
  if (blockIdx.x == 0) clk[0] = (t1 - t0) / iters / 2;
}
int main() {
  int* g; long long* clk; long long h[8];
  cudaMalloc(&g, 1 << 16); cudaMemset(g, 0, 1 << 16); cudaMalloc(&clk, 64);
  k_local<<<1, 32>>>(g, clk, 2000); cudaDeviceSynchronize(); cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
  printf("TMA fetch 16 B: %lld clk; TMA fetch 1600+16 B: %lld clk; ld.relaxed.gpu: %lld clk; ld.acquire.gpu: %lld clk\n", h[0], h[1], h[2], h[3]);
  printf("__threadfence + st.release: %lld clk; st.release: %lld clk; __threadfence + st.relaxed: %lld clk\n", h[4], h[5], h[6]);
  cudaMemset(g, 0, 1 << 16);
  k_pingpong<0><<<2, 32>>>(g, clk, 2000); cudaDeviceSynchronize(); cudaMemcpy(h, clk, 8, cudaMemcpyDeviceToHost);
  printf("flag one-way, LDG polling: %lld clk\n", h[0]);
  cudaMemset(g, 0, 1 << 16);
  k_pingpong<1><<<2, 32>>>(g, clk, 2000); cudaDeviceSynchronize(); cudaMemcpy(h, clk, 8, cudaMemcpyDeviceToHost);
  printf("flag one-way, TMA polling: %lld clk\n", h[0]);
  return 0;
}
