// backtransform.cu -- eigenvectors of the BdG matrix from those of the tridiagonal matrix:
// U <- Q U with Q = H_0 ... H_{n-2} (third stage of diagonalize_H_BdG!,
// /root/reference src/Hamiltonian.jl:96-114).  The reflectors are grouped into blocks of DW_NB,
// Q_k = I - V_k T_k V_k^H of DW_NBT reflectors (T_k from the Gram matrix V_k^H V_k), applied last
// block first as three batched
// complex GEMMs on the FP64 tensor cores:  W1 = V_k^H U ; W2 = T_k W1 ; U -= V_k W2.
#include "dwhmc.h"
#include "internal.h"

namespace {

__device__ __forceinline__ void cfma_(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}

// T factor of one block reflector from its Gram matrix G = V^H V (forward, columnwise):
// T[i,i] = tau_i, T[0:i, i] = -tau_i T[0:i,0:i] G[0:i, i].  Thread r owns row r of T, which only
// depends on earlier entries of the same row, so the recurrence needs no barrier.
__global__ void __launch_bounds__(DW_NBT) bt_larft_kernel(const cplx* __restrict__ tau, const cplx* __restrict__ Gall,
                                                          cplx* __restrict__ Tall, int n, int nbt, size_t gsplit_stride,
                                                          Mask mask) {
  const int k = blockIdx.x, b = blockIdx.y;
  if (!mask.on(b)) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* T = reinterpret_cast<cplx*>(smem_raw);          // [NBT][NBT + 1] row r at T + r * (NBT + 1)
  cplx* G = T + DW_NBT * (DW_NBT + 1);                   // [NBT][NBT] column-major
  const int r = threadIdx.x;
  const int j0 = k * DW_NBT;
  const int pn = min(DW_NBT, n - 1 - j0);
  const cplx* Gk = Gall + ((size_t)b * nbt + k) * DW_NBT * DW_NBT;
  for (int c = 0; c < DW_NBT; ++c) {
    cplx gs = Gk[c * DW_NBT + r];
    for (int sp = 1; sp < DW_GSPLIT; ++sp) {        // K pieces of the Gram GEMM, fixed order
      const cplx t = Gk[(size_t)sp * gsplit_stride + c * DW_NBT + r];
      gs.x += t.x; gs.y += t.y;
    }
    G[c * DW_NBT + r] = gs;
    T[r * (DW_NBT + 1) + c] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  cplx* Tr = T + r * (DW_NBT + 1);
  for (int i = 0; i < pn; ++i) {
    const cplx t = tau[(size_t)b * n + j0 + i];
    if (r < i) {
      cplx s = make_double2(0.0, 0.0);
      for (int l = r; l < i; ++l) cfma_(s, Tr[l], G[i * DW_NBT + l]);
      Tr[i] = make_double2(-(t.x * s.x - t.y * s.y), -(t.x * s.y + t.y * s.x));
    } else if (r == i) {
      Tr[i] = t;
    }
  }
  __syncthreads();
  cplx* out = Tall + ((size_t)b * nbt + k) * DW_NBT * DW_NBT;
  for (int c = 0; c < DW_NBT; ++c) out[c * DW_NBT + r] = Tr[c];
}

// Partner columns of the particle-hole symmetric spectrum: (u, v) with energy E  ->  (-conj v, conj u)
// with energy -E (tau_y H^* tau_y = -H for H = [[h, D], [D^*, -h]], h real, D symmetric).
__global__ void __launch_bounds__(256) ph_mirror_kernel(double* __restrict__ Eall, cplx* __restrict__ Uall,
                                                        const int* __restrict__ halfflag, int N, Mask mask) {
  const int b = blockIdx.y;
  if (!mask.on(b) || halfflag[b] == 0) return;
  const int n = 2 * N, k = blockIdx.x, src = n - 1 - k;
  cplx* U = Uall + (size_t)b * n * n;
  const cplx* s = U + (size_t)src * n;
  cplx* d = U + (size_t)k * n;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const cplx u = s[i], v = s[i + N];
    d[i] = make_double2(-v.x, v.y);
    d[i + N] = make_double2(u.x, -u.y);
  }
  if (threadIdx.x == 0) Eall[(size_t)b * n + k] = -Eall[(size_t)b * n + src];
}

}  // namespace

int dw_ph_mirror(Handle* h, double* E, cplx* U, Mask mask) {
  dim3 grid(h->N, h->B);
  ph_mirror_kernel<<<grid, 256, 0, h->stream>>>(E, U, h->halfflag, h->N, mask);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

int dw_backtransform(Handle* h, cplx* U, Mask mask, bool ph) {
  const int n = h->n, B = h->B, nbt = h->nbt;
  const long long sT = (long long)nbt * DW_NBT * DW_NBT;
  ZgemmArgs a;
  a.nseg = 1; a.A[1] = nullptr; a.Bm[1] = nullptr; a.lower = 0; a.batch = B; a.mask = mask;
  const bool half = ph && h->ph_mode;
  // Gram matrices of all blocks, then their T factors
  for (int k = 0; k < nbt; ++k) {
    const int j0 = k * DW_NBT;
    const int pn = (n - 1 - j0 < DW_NBT) ? n - 1 - j0 : DW_NBT;
    const int mk = n - (j0 + 1);
    const cplx* Vk = h->V + (size_t)j0 * n + (j0 + 1);
    a.M = pn; a.N = pn; a.K = mk;
    a.A[0] = Vk; a.lda = n; a.sA = (long long)n * n; a.opA = 1;
    a.Bm[0] = Vk; a.ldb = n; a.sB = (long long)n * n; a.opB = 0;
    // long K, tiny output: cut K into DW_GSPLIT pieces so the launch fills the machine; bt_larft_kernel sums them
    a.C = h->Gb + (size_t)k * DW_NBT * DW_NBT; a.ldc = DW_NBT; a.sC = sT;
    a.ksplit = DW_GSPLIT; a.sCk = sT * B;
    a.alpha = 1.0; a.beta = 0.0;
    DW_TRY(dw_zgemm(h, a));
  }
  a.ksplit = 1; a.sCk = 0;
  {
    static bool attr_set[64] = {false};
    const size_t smem = sizeof(cplx) * (DW_NBT * (DW_NBT + 1) + DW_NBT * DW_NBT);
    if (!attr_set[h->device & 63]) {
      DW_CUDA(h, cudaFuncSetAttribute(bt_larft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set[h->device & 63] = true;
    }
    dim3 grid(nbt, B);
    bt_larft_kernel<<<grid, DW_NBT, smem, h->stream>>>(h->tau, h->Gb, h->Tf, n, nbt, (size_t)sT * B, mask);
    DW_LAUNCH_CHECK(h);
  }
  for (int k = nbt - 1; k >= 0; --k) {
    const int j0 = k * DW_NBT;
    const int pn = (n - 1 - j0 < DW_NBT) ? n - 1 - j0 : DW_NBT;
    const int mk = n - (j0 + 1);
    if (pn <= 0 || mk <= 0) continue;
    const cplx* Vk = h->V + (size_t)j0 * n + (j0 + 1);   // rows j0+1.., columns j0..j0+pn-1
    cplx* Uk = U + (j0 + 1);                             // rows j0+1.. of every column
    // W1 (pn x n) = V_k^H U_k
    if (half) { a.skip_flag = h->halfflag; a.skip_cols = h->N; }
    a.M = pn; a.N = n; a.K = mk;
    a.A[0] = Vk; a.lda = n; a.sA = (long long)n * n; a.opA = 1;
    a.Bm[0] = Uk; a.ldb = n; a.sB = (long long)n * n; a.opB = 0;
    a.C = h->Wbt; a.ldc = DW_NBT; a.sC = (long long)DW_NBT * n;
    a.alpha = 1.0; a.beta = 0.0;
    DW_TRY(dw_zgemm(h, a));
    // W2 (pn x n) = T_k W1
    a.M = pn; a.N = n; a.K = pn;
    a.A[0] = h->Tf + (size_t)k * DW_NBT * DW_NBT; a.lda = DW_NBT; a.sA = sT; a.opA = 0;
    a.Bm[0] = h->Wbt; a.ldb = DW_NBT; a.sB = (long long)DW_NBT * n; a.opB = 0;
    a.C = h->Wbt2; a.ldc = DW_NBT; a.sC = (long long)DW_NBT * n;
    a.alpha = 1.0; a.beta = 0.0;
    DW_TRY(dw_zgemm(h, a));
    // U_k -= V_k W2
    a.M = mk; a.N = n; a.K = pn;
    a.A[0] = Vk; a.lda = n; a.sA = (long long)n * n; a.opA = 0;
    a.Bm[0] = h->Wbt2; a.ldb = DW_NBT; a.sB = (long long)DW_NBT * n; a.opB = 0;
    a.C = Uk; a.ldc = n; a.sC = (long long)n * n;
    a.alpha = -1.0; a.beta = 1.0;
    DW_TRY(dw_zgemm(h, a));
  }
  return DWHMC_OK;
}

int dw_assemble_for_solve(Handle* h, const double* w, const double* par3, const cplx* delta, Mask mask) {
  if (h->band_b > 0) return dw_band_assemble(h, w, par3, delta, mask);
  return dw_assemble(h, w, par3, delta, h->A, mask);
}

int dw_eigensolve(Handle* h, double* E_out, cplx* U_out, Mask mask, bool ph) {
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
  const bool band = ph && h->band_b > 0;
  tic();
  if (band) {
    DW_TRY(dw_band_chase(h, mask));
    // the T factors of the back-transformation only need the reflectors: compute them beside the D&C stage
    DW_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
    DW_CUDA(h, cudaStreamWaitEvent(h->gstream[0], h->ev_fork, 0));
    DW_TRY(dw_band_tfactors(h, mask, h->gstream[0]));
    DW_CUDA(h, cudaEventRecord(h->ev_join[0], h->gstream[0]));
  } else {
    DW_TRY(dw_hetrd(h, U_out, mask));
  }
  toc(1);
  tic();
  // first eigenvector column (by rank) a chain flagged for the particle-hole shortcut needs: the back-transformation
  // of the band route works on strips of 8 columns from (N / 16) * 16, the dense route's GEMM tiles are up to 128 wide
  const int c_lo = (ph && h->ph_mode) ? (band ? (h->N / 16) * 16 : (h->N / 128) * 128) : 0;
  DW_TRY(dw_stedc(h, mask, ph, c_lo));
  // band route: the tridiagonal eigenvectors go to h->A (the band is dead by now), in band row order
  DW_TRY(dw_stedc_output(h, E_out, band ? h->A : U_out, mask, ph, c_lo));
  toc(2);
  tic();
  if (band) {
    DW_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join[0], 0));
    DW_TRY(dw_band_backtransform(h, U_out, mask, ph));
  }
  else DW_TRY(dw_backtransform(h, U_out, mask, ph));
  if (ph && h->ph_mode) DW_TRY(dw_ph_mirror(h, E_out, U_out, mask));
  toc(3);
  h->eigensolves++;
  return DWHMC_OK;
}
