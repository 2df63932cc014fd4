// assemble.cu -- BdG matrix assembly (init_static_H! + update_H_BdG!,
// /root/reference src/Hamiltonian.jl:10-47, :55-86) for every chain of the batch.
//
// H = [[h, D], [D^H, -h]],  h_ii = w_i - mu, h_ij = -t (NN), -t' (NNN),
// D_{i,j} = D_{j,i} = Delta_b / 2 for bond b = (i -> j).  The reference stores the upper
// triangle only and lets LAPACK read it through Hermitian(:U); here the eigensolver works on
// full storage, so both triangles are written.  HBM-bound: one 16n^2-byte clear (memset at
// copy bandwidth) plus an O(N) scatter.
#include "dwhmc.h"
#include "internal.h"

namespace {

// one thread per site: rows i and i+N of the matrix (and their mirror entries in the pairing block)
__global__ void scatter_kernel(cplx* __restrict__ Aall, const double* __restrict__ w, const double* __restrict__ par3,
                               const cplx* __restrict__ delta, const int* __restrict__ nn, const int* __restrict__ nnn,
                               int N, int B, int upper_only, Mask mask) {
  const int b = blockIdx.y;
  if (!mask.on(b)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int n = 2 * N;
  cplx* A = Aall + (size_t)b * n * n;
  const double t = par3[b], tp = par3[B + b], mu = par3[2 * B + b];
  const double term = w[(size_t)b * N + i] - mu;
  auto put = [&](int r, int c, double re, double im) {
    if (upper_only && r > c) return;
    A[(size_t)c * n + r] = make_double2(re, im);
  };
  put(i, i, term, 0.0);
  put(i + N, i + N, -term, 0.0);
  // NN first, then NNN: a later assignment wins, as in the reference loop order (:28-43)
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const int j = nn[d * N + i];
    if (j != i) { put(i, j, -t, 0.0); put(i + N, j + N, t, 0.0); }
  }
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const int j = nnn[d * N + i];
    if (j != i) { put(i, j, -tp, 0.0); put(i + N, j + N, tp, 0.0); }
  }
  // pairing block (:68-83): H[i, j+N] = H[j, i+N] = Delta[i,dir]/2 ; lower-left = conjugate transpose
  const cplx* dl = delta + (size_t)b * 2 * N;
#pragma unroll
  for (int dir = 0; dir < 2; ++dir) {
    const int j = nn[dir * N + i];
    const cplx v = dl[dir * N + i];
    const double re = 0.5 * v.x, im = 0.5 * v.y;
    put(i, j + N, re, im);
    put(j, i + N, re, im);
    put(j + N, i, re, -im);
    put(i + N, j, re, -im);
  }
}

}  // namespace

static int assemble_impl(Handle* h, const double* w, const double* par3, const cplx* delta, cplx* A, int upper,
                         Mask mask) {
  const int n = h->n, N = h->N, B = h->B;
  // A is a work matrix: clearing the slices of inactive chains too is harmless
  DW_CUDA(h, cudaMemsetAsync(A, 0, sizeof(cplx) * (size_t)n * n * B, h->stream));
  dim3 grid((N + 127) / 128, B);
  scatter_kernel<<<grid, 128, 0, h->stream>>>(A, w, par3, delta, h->nn, h->nnn, N, B, upper, mask);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

int dw_assemble(Handle* h, const double* w, const double* par3, const cplx* delta, cplx* A, Mask mask) {
  return assemble_impl(h, w, par3, delta, A, 0, mask);
}

int dw_assemble_upper(Handle* h, const double* w, const double* par3, const cplx* delta, cplx* out) {
  return assemble_impl(h, w, par3, delta, out, 1, no_mask());
}
