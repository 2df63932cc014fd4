// Experiment: does the HBM-bound tridiagonalisation of one half-batch overlap the DMMA-bound
// D&C + back-transform of the other?  Links the library objects directly.
//   make -C hybrid-monte-carlo-for-d-wave-sc_b200/csrc && nvcc ... (see tools/microbench/build_overlap.sh)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include "dwhmc.h"
#include "internal.h"

static Handle* mk(int B, int L, int seed) {
  const int N = L * L;
  std::vector<int64_t> nn(4 * N), nnn(4 * N);
  auto idx = [&](int x, int y) { return ((y + L) % L) * L + ((x + L) % L) + 1; };
  for (int y = 0; y < L; ++y) for (int x = 0; x < L; ++x) {
    int i = y * L + x;
    nn[0 * N + i] = idx(x + 1, y); nn[1 * N + i] = idx(x, y + 1); nn[2 * N + i] = idx(x - 1, y); nn[3 * N + i] = idx(x, y - 1);
    nnn[0 * N + i] = idx(x + 1, y + 1); nnn[1 * N + i] = idx(x - 1, y + 1); nnn[2 * N + i] = idx(x - 1, y - 1); nnn[3 * N + i] = idx(x + 1, y - 1);
  }
  dwhmc_handle hh;
  if (dwhmc_create(&hh, 0, B, L, L, nn.data(), nnn.data())) { printf("create: %s\n", dwhmc_last_error(nullptr)); exit(1); }
  std::vector<double> t(B, 1.0), tp(B, -0.35), mu(B, -1.08), beta(B, 20.0), J(B, 0.8), m(B, 1.0);
  dwhmc_set_params(hh, t.data(), tp.data(), mu.data(), beta.data(), J.data(), m.data());
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> u(-0.05, 0.05);
  std::vector<double> w((size_t)N * B, 0.0), d((size_t)4 * N * B);
  for (auto& x : w) x = (rng() % 20 == 0) ? 1.0 : 0.0;
  for (auto& x : d) x = u(rng);
  dwhmc_set_disorder(hh, w.data());
  dwhmc_set_field(hh, d.data());
  dwhmc_init_static_H(hh); dwhmc_update_H(hh);
  if (dwhmc_diagonalize(hh)) { printf("diag: %s\n", dwhmc_last_error(hh)); exit(1); }
  return reinterpret_cast<Handle*>(hh);
}

int main(int argc, char** argv) {
  const int L = argc > 1 ? atoi(argv[1]) : 24, B = argc > 2 ? atoi(argv[2]) : 32;
  const int prio = argc > 3 ? atoi(argv[3]) : 0;   // priority of the GEMM-side stream (0 low, -5 high)
  Handle* h1 = mk(B, L, 1);
  Handle* h2 = mk(B, L, 2);
  cudaStream_t s2;
  cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, prio);
  h2->stream = s2;
  cudaEvent_t a0, a1, b0, b1;
  cudaEventCreate(&a0); cudaEventCreate(&a1); cudaEventCreate(&b0); cudaEventCreate(&b1);
  auto tri = [&]() {
    dw_assemble(h1, h1->Hs_w, h1->Hs_par, h1->delta, h1->A, no_mask());
    dw_hetrd(h1, h1->U_prop, no_mask());
  };
  auto gem = [&]() {
    dw_stedc(h2, no_mask());
    dw_stedc_output(h2, h2->E_prop, h2->U_cur, no_mask());
    dw_backtransform(h2, h2->U_cur, no_mask());
  };
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaDeviceSynchronize();
    cudaEventRecord(a0, h1->stream); tri(); cudaEventRecord(a1, h1->stream);
    cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, a0, a1); printf("tridiagonalise alone: %.2f ms\n", ms);
    cudaEventRecord(b0, h2->stream); gem(); cudaEventRecord(b1, h2->stream);
    cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, b0, b1); printf("stedc+backtransform alone: %.2f ms\n", ms);
    cudaEventRecord(a0, h1->stream); cudaEventRecord(b0, h2->stream);
    // interleave the enqueue so neither stream's queue runs dry on the host side
    gem(); tri();
    cudaEventRecord(a1, h1->stream); cudaEventRecord(b1, h2->stream);
    cudaDeviceSynchronize();
    float ta, tb; cudaEventElapsedTime(&ta, a0, a1); cudaEventElapsedTime(&tb, b0, b1);
    printf("concurrent: tridiagonalise %.2f ms, stedc+backtransform %.2f ms\n", ta, tb);
  }
  printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
