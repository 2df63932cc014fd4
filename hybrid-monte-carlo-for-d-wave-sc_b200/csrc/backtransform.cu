// backtransform.cu -- eigenvectors of the BdG matrix from those of the tridiagonal matrix:
// U <- Q U with Q = H_0 ... H_{n-2} (third stage of diagonalize_H_BdG!,
// /root/reference src/Hamiltonian.jl:96-114).  The reflectors are grouped into blocks of DW_NB,
// Q_k = I - V_k T_k V_k^H (T_k from hetrd.cu), applied last block first as three batched
// complex GEMMs on the FP64 tensor cores:  W1 = V_k^H U ; W2 = T_k W1 ; U -= V_k W2.
#include "dwhmc.h"
#include "internal.h"

int dw_backtransform(Handle* h, cplx* U, Mask mask) {
  const int n = h->n, B = h->B;
  for (int k = h->nblk - 1; k >= 0; --k) {
    const int j0 = k * DW_NB;
    const int pn = (n - 1 - j0 < DW_NB) ? n - 1 - j0 : DW_NB;
    const int mk = n - (j0 + 1);
    if (pn <= 0 || mk <= 0) continue;
    const cplx* Vk = h->V + (size_t)j0 * n + (j0 + 1);   // rows j0+1.., columns j0..j0+pn-1
    cplx* Uk = U + (j0 + 1);                             // rows j0+1.. of every column
    ZgemmArgs a;
    a.nseg = 1; a.A[1] = nullptr; a.Bm[1] = nullptr; a.lower = 0; a.batch = B; a.mask = mask;
    // W1 (pn x n) = V_k^H U_k
    a.M = pn; a.N = n; a.K = mk;
    a.A[0] = Vk; a.lda = n; a.sA = (long long)n * n; a.opA = 1;
    a.Bm[0] = Uk; a.ldb = n; a.sB = (long long)n * n; a.opB = 0;
    a.C = h->Wbt; a.ldc = DW_NB; a.sC = (long long)DW_NB * n;
    a.alpha = 1.0; a.beta = 0.0;
    DW_TRY(dw_zgemm(h, a));
    // W2 (pn x n) = T_k W1
    a.M = pn; a.N = n; a.K = pn;
    a.A[0] = h->Tf + (size_t)k * DW_NB * DW_NB; a.lda = DW_NB; a.sA = (long long)h->nblk * DW_NB * DW_NB; a.opA = 0;
    a.Bm[0] = h->Wbt; a.ldb = DW_NB; a.sB = (long long)DW_NB * n; a.opB = 0;
    a.C = h->Wbt2; a.ldc = DW_NB; a.sC = (long long)DW_NB * n;
    a.alpha = 1.0; a.beta = 0.0;
    DW_TRY(dw_zgemm(h, a));
    // U_k -= V_k W2
    a.M = mk; a.N = n; a.K = pn;
    a.A[0] = Vk; a.lda = n; a.sA = (long long)n * n; a.opA = 0;
    a.Bm[0] = h->Wbt2; a.ldb = DW_NB; a.sB = (long long)DW_NB * n; a.opB = 0;
    a.C = Uk; a.ldc = n; a.sC = (long long)n * n;
    a.alpha = -1.0; a.beta = 1.0;
    DW_TRY(dw_zgemm(h, a));
  }
  return DWHMC_OK;
}

int dw_eigensolve(Handle* h, double* E_out, cplx* U_out, Mask mask) {
  // U_out doubles as the W panel scratch of the tridiagonalisation (it is only written with
  // eigenvectors after that stage has finished)
  cudaEvent_t& e0 = h->ev0;
  cudaEvent_t& e1 = h->ev1;
  auto tic = [&]() { if (h->profiling) cudaEventRecord(e0, h->stream); };
  auto toc = [&](int slot) {
    if (!h->profiling) return;
    cudaEventRecord(e1, h->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    h->timers[slot] += ms;
  };
  tic();
  DW_TRY(dw_hetrd(h, U_out, mask));
  toc(1);
  tic();
  DW_TRY(dw_stedc(h, mask));
  DW_TRY(dw_stedc_output(h, E_out, U_out, mask));
  toc(2);
  tic();
  DW_TRY(dw_backtransform(h, U_out, mask));
  toc(3);
  h->eigensolves++;
  return DWHMC_OK;
}
