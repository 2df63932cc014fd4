// hetrd.cu -- batched blocked Householder tridiagonalisation of the Hermitian BdG matrices
// (first stage of diagonalize_H_BdG!, /root/reference src/Hamiltonian.jl:96-114, where the
// reference calls LAPACK zheevr through eigen!).
//
// A = Q T Q^H, Q = H_0 H_1 ... H_{n-2}, H_j = I - tau_j v_j v_j^H, v_j[j+1] = 1, T real.
// Panels of DW_NB columns.  Inside a panel the trailing matrix is not updated; each column costs
//   colstep  (one CTA per chain)  finish w_{j-1}; update column j with the panel's V/W; reflector;
//                                 the small products W^H v, V^H v
//   hemv     (row blocks x column splits x chains)  y = A[j+1:, j+1:] v   -- HBM-bound, 16 m^2 bytes
// and each panel ends with the rank-2k update A -= V W^H + W V^H on the FP64 tensor cores
// (gemm_dmma.cu).  The formulas are the ones prototyped and checked against LAPACK in
// tests/algo_proto.py.
#include "dwhmc.h"
#include "internal.h"

namespace {

constexpr int CT = 512;     // threads of the column-step kernel
constexpr int HT = 128;     // threads (= rows) of a hemv CTA

__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// acc += a * b
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc -= a * b
__device__ __forceinline__ void cfms(cplx& acc, cplx a, cplx b) {
  acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(-a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
// acc += conj(a) * b
__device__ __forceinline__ void cfmac(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}

__device__ __forceinline__ cplx warp_sum(cplx v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}

// sum over the block; result returned to every thread.  red: >= 32 cplx of shared memory.
__device__ __forceinline__ cplx block_sum(cplx v, cplx* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  cplx t = make_double2(0.0, 0.0);
  for (int i = 0; i < nw; ++i) t = cadd(t, red[i]);   // fixed order: identical in every thread
  return t;
}

struct ColArgs {
  cplx* A; cplx* V; cplx* W; cplx* ypart; cplx* P1; cplx* P2; cplx* tau;
  double* d; double* e;
  int n, B, j, j0, finish_prev, make_ref;
  Mask mask;
};

__global__ void __launch_bounds__(CT) colstep_kernel(ColArgs g) {
  const int b = blockIdx.x;
  if (!g.mask.on(b)) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* sv = reinterpret_cast<cplx*>(smem_raw);   // [n] reflector
  cplx* sw = sv + g.n;                            // [n] w, then the updated column a
  cplx* rowW = sw + g.n;                          // [NB]
  cplx* rowV = rowW + DW_NB;                      // [NB]
  cplx* red = rowV + DW_NB;                       // [32]
  __shared__ cplx s_scale;

  const int n = g.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t mat = (size_t)b * n * n;
  cplx* A = g.A + mat;
  cplx* V = g.V + mat;
  cplx* W = g.W + mat;
  const int j0 = g.j0;

  if (g.finish_prev) {
    // ---- finish w for column jp = j - 1
    const int jp = g.j - 1, ip = jp - j0;
    if (tid < ip) {
      rowW[tid] = g.P1[(size_t)b * DW_NB + tid];                   // W_panel^H v
      rowV[tid] = g.P2[((size_t)b * n + jp) * DW_NB + tid];        // V_panel^H v
    }
    const cplx tau = g.tau[(size_t)b * n + jp];
    __syncthreads();
    cplx dot = make_double2(0.0, 0.0);
    for (int r = jp + 1 + tid; r < n; r += CT) {
      cplx acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int s = 0; s < DW_NSPLIT; ++s) acc = cadd(acc, g.ypart[((size_t)s * g.B + b) * n + r]);
      for (int k = 0; k < ip; ++k) {
        cfms(acc, V[(size_t)(j0 + k) * n + r], rowW[k]);
        cfms(acc, W[(size_t)(j0 + k) * n + r], rowV[k]);
      }
      const cplx wv = cmul(tau, acc);
      const cplx vv = V[(size_t)jp * n + r];
      sw[r] = wv;
      sv[r] = vv;
      cfmac(dot, wv, vv);
    }
    dot = block_sum(dot, red);
    cplx alpha = cmul(tau, dot);
    alpha.x *= -0.5; alpha.y *= -0.5;
    for (int r = jp + 1 + tid; r < n; r += CT) {
      cplx wv = sw[r];
      cfma(wv, alpha, sv[r]);
      W[(size_t)jp * n + r] = wv;
    }
    __syncthreads();
  }

  if (g.make_ref) {
    const int j = g.j, i = j - j0;
    if (tid < i) {
      const cplx a = W[(size_t)(j0 + tid) * n + j], c = V[(size_t)(j0 + tid) * n + j];
      rowW[tid] = make_double2(a.x, -a.y);
      rowV[tid] = make_double2(c.x, -c.y);
    }
    __syncthreads();
    cplx nrm = make_double2(0.0, 0.0);
    for (int r = j + tid; r < n; r += CT) {
      cplx a = A[(size_t)j * n + r];
      for (int k = 0; k < i; ++k) {
        cfms(a, V[(size_t)(j0 + k) * n + r], rowW[k]);
        cfms(a, W[(size_t)(j0 + k) * n + r], rowV[k]);
      }
      sw[r] = a;
      if (r >= j + 2) nrm.x += a.x * a.x + a.y * a.y;
    }
    nrm = block_sum(nrm, red);   // also orders the sw writes before the reads below
    if (tid == 0) {
      const cplx a0 = sw[j], alpha = sw[j + 1];
      const double xn2 = nrm.x;
      double beta;
      cplx tau, scale;
      if (xn2 == 0.0 && alpha.y == 0.0) {
        beta = alpha.x;
        tau = make_double2(0.0, 0.0);
        scale = make_double2(0.0, 0.0);
      } else {
        beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn2), alpha.x);
        tau = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
        // scale = 1 / (alpha - beta)
        const double dr = alpha.x - beta, di = alpha.y;
        const double den = dr * dr + di * di;
        scale = make_double2(dr / den, -di / den);
      }
      g.d[(size_t)b * n + j] = a0.x;
      g.e[(size_t)b * n + j] = beta;
      g.tau[(size_t)b * n + j] = tau;
      s_scale = scale;
    }
    __syncthreads();
    const cplx scale = s_scale;
    for (int r = j + 1 + tid; r < n; r += CT) {
      const cplx v = (r == j + 1) ? make_double2(1.0, 0.0) : cmul(sw[r], scale);
      sv[r] = v;
      V[(size_t)j * n + r] = v;
    }
    __syncthreads();
    // small products: P1[k] = W[:, j0+k]^H v, P2[k] = V[:, j0+k]^H v  (k < i), one warp per product
    const int nw = CT / 32;
    for (int q = warp; q < 2 * i; q += nw) {
      const int k = q >> 1, which = q & 1;
      const cplx* src = (which ? V : W) + (size_t)(j0 + k) * n;
      cplx acc = make_double2(0.0, 0.0);
      for (int r = j + 1 + lane; r < n; r += 32) cfmac(acc, src[r], sv[r]);
      acc = warp_sum(acc);
      if (lane == 0) {
        if (which) g.P2[((size_t)b * n + j) * DW_NB + k] = acc;
        else g.P1[(size_t)b * DW_NB + k] = acc;
      }
    }
  }
}

// y = A[q0:, q0:] v for the column split blockIdx.y; q0 = j + 1
__global__ void __launch_bounds__(HT) hemv_kernel(const cplx* __restrict__ Aall, const cplx* __restrict__ Vall,
                                                  cplx* __restrict__ ypart, int n, int B, int j, Mask mask) {
  const int b = blockIdx.z;
  if (!mask.on(b)) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* sv = reinterpret_cast<cplx*>(smem_raw);
  const int q0 = j + 1, m = n - q0;
  const int cs = (m + DW_NSPLIT - 1) / DW_NSPLIT;
  const int s = blockIdx.y;
  const int c_lo = s * cs, c_hi = min(m, c_lo + cs);
  const size_t mat = (size_t)b * n * n;
  const cplx* v = Vall + mat + (size_t)j * n + q0;
  for (int c = c_lo + threadIdx.x; c < c_hi; c += HT) sv[c - c_lo] = v[c];
  __syncthreads();
  const int rl = blockIdx.x * HT + threadIdx.x;
  if (rl >= m) return;
  const cplx* Ar = Aall + mat + (size_t)q0 * n + q0 + rl;   // A[q0 + rl, q0 + c] = Ar[c * n]
  cplx a0 = make_double2(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0;
  int c = c_lo;
  for (; c + 4 <= c_hi; c += 4) {
    const cplx x0 = Ar[(size_t)c * n], x1 = Ar[(size_t)(c + 1) * n];
    const cplx x2 = Ar[(size_t)(c + 2) * n], x3 = Ar[(size_t)(c + 3) * n];
    cfma(a0, x0, sv[c - c_lo]);
    cfma(a1, x1, sv[c + 1 - c_lo]);
    cfma(a2, x2, sv[c + 2 - c_lo]);
    cfma(a3, x3, sv[c + 3 - c_lo]);
  }
  for (; c < c_hi; ++c) cfma(a0, Ar[(size_t)c * n], sv[c - c_lo]);
  a0 = cadd(cadd(a0, a1), cadd(a2, a3));
  ypart[((size_t)s * B + b) * n + q0 + rl] = a0;
}

__global__ void lastd_kernel(const cplx* __restrict__ A, double* __restrict__ d, int n, int B, Mask mask) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || !mask.on(b)) return;
  d[(size_t)b * n + n - 1] = A[(size_t)b * n * n + (size_t)(n - 1) * n + (n - 1)].x;
}

// T factors of the block reflectors (forward, columnwise): T[i,i] = tau_i,
// T[0:i, i] = -tau_i T[0:i,0:i] (V^H v_i); one warp per (block, chain)
__global__ void __launch_bounds__(32) larft_kernel(const cplx* __restrict__ tau, const cplx* __restrict__ P2,
                                                   cplx* __restrict__ Tf, int n, int nblk, Mask mask) {
  const int k = blockIdx.x, b = blockIdx.y;
  if (!mask.on(b)) return;
  __shared__ cplx T[DW_NB][DW_NB + 1];
  const int lane = threadIdx.x;
  const int j0 = k * DW_NB;
  const int pn = min(DW_NB, n - 1 - j0);
  for (int c = 0; c < DW_NB; ++c) T[lane][c] = make_double2(0.0, 0.0);
  __syncwarp();
  for (int i = 0; i < pn; ++i) {
    const cplx t = tau[(size_t)b * n + j0 + i];
    if (lane < i) {
      const cplx* p = P2 + ((size_t)b * n + j0 + i) * DW_NB;
      cplx s = make_double2(0.0, 0.0);
      for (int l = lane; l < i; ++l) cfma(s, T[lane][l], p[l]);
      cplx r = cmul(t, s);
      T[lane][i] = make_double2(-r.x, -r.y);
    } else if (lane == i) {
      T[lane][i] = t;
    }
    __syncwarp();
  }
  cplx* out = Tf + ((size_t)b * nblk + k) * DW_NB * DW_NB;
  for (int c = 0; c < DW_NB; ++c) out[c * DW_NB + lane] = T[lane][c];
}

}  // namespace

int dw_hetrd(Handle* h, cplx* W, Mask mask) {
  const int n = h->n, B = h->B;
  if (n < 2) {
    lastd_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(h->A, h->d, n, B, mask);
    DW_LAUNCH_CHECK(h);
    return DWHMC_OK;
  }
  const size_t col_smem = sizeof(cplx) * (2 * (size_t)n + 2 * DW_NB + 32);
  static bool attr_set[64] = {false};
  if (!attr_set[h->device & 63]) {
    DW_CUDA(h, cudaFuncSetAttribute(colstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set[h->device & 63] = true;
  }
  if (col_smem > 200 * 1024) { h->err = "dw_hetrd: matrix too large for the column-step kernel"; return DWHMC_E_BADARG; }
  ColArgs g;
  g.A = h->A; g.V = h->V; g.W = W; g.ypart = h->ypart; g.P1 = h->P1; g.P2 = h->P2; g.tau = h->tau;
  g.d = h->d; g.e = h->e; g.n = n; g.B = B; g.mask = mask;
  for (int j0 = 0; j0 < n - 1; j0 += DW_NB) {
    const int pn = (n - 1 - j0 < DW_NB) ? n - 1 - j0 : DW_NB;
    g.j0 = j0;
    for (int i = 0; i < pn; ++i) {
      const int j = j0 + i;
      g.j = j; g.finish_prev = (i > 0); g.make_ref = 1;
      colstep_kernel<<<B, CT, col_smem, h->stream>>>(g);
      DW_LAUNCH_CHECK(h);
      const int m = n - j - 1;
      const int cs = (m + DW_NSPLIT - 1) / DW_NSPLIT;
      dim3 grid((m + HT - 1) / HT, DW_NSPLIT, B);
      if (h->profiling) cudaEventRecord(h->ev_begin, h->stream);
      hemv_kernel<<<grid, HT, sizeof(cplx) * cs, h->stream>>>(h->A, h->V, h->ypart, n, B, j, mask);
      DW_LAUNCH_CHECK(h);
      if (h->profiling) {
        cudaEventRecord(h->ev_end, h->stream);
        cudaEventSynchronize(h->ev_end);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev_begin, h->ev_end);
        h->timers[7] += ms;
      }
    }
    const int j1 = j0 + pn;
    g.j = j1; g.finish_prev = 1; g.make_ref = 0;
    colstep_kernel<<<B, CT, col_smem, h->stream>>>(g);
    DW_LAUNCH_CHECK(h);
    // trailing update A[j1:, j1:] -= V W^H + W V^H  (rows >= j1 of the panel columns)
    ZgemmArgs a;
    a.M = n - j1; a.N = n - j1; a.K = pn; a.nseg = 2;
    a.A[0] = h->V + (size_t)j0 * n + j1; a.Bm[0] = W + (size_t)j0 * n + j1;
    a.A[1] = W + (size_t)j0 * n + j1;    a.Bm[1] = h->V + (size_t)j0 * n + j1;
    a.lda = n; a.ldb = n; a.ldc = n;
    a.sA = (long long)n * n; a.sB = (long long)n * n; a.sC = (long long)n * n;
    a.C = h->A + (size_t)j1 * n + j1;
    a.alpha = -1.0; a.beta = 1.0; a.opA = 0; a.opB = 1; a.lower = 0; a.batch = B; a.mask = mask;
    DW_TRY(dw_zgemm(h, a));
  }
  lastd_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(h->A, h->d, n, B, mask);
  DW_LAUNCH_CHECK(h);
  dim3 tg(h->nblk, B);
  larft_kernel<<<tg, 32, 0, h->stream>>>(h->tau, h->P2, h->Tf, n, h->nblk, mask);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}
