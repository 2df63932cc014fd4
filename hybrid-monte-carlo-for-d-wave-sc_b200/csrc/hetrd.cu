// hetrd.cu -- batched blocked Householder tridiagonalisation of the Hermitian BdG matrices
// (first stage of diagonalize_H_BdG!, /root/reference src/Hamiltonian.jl:96-114, where the
// reference calls LAPACK zheevr through eigen!).
//
// A = Q T Q^H, Q = H_0 H_1 ... H_{n-2}, H_j = I - tau_j v_j v_j^H, v_j[j+1] = 1, T real.
// Panels of DW_NB columns.  Inside a panel the trailing matrix is not updated; each column costs
//   colstep  (one CTA per chain)  finish w_{j-1}; update column j with the panel's V/W; reflector;
//                                 the small products W^H v, V^H v
//   hemv     (lower-triangle 64x64 tiles x chains)  y = A[j+1:, j+1:] v   -- HBM-bound, 8 m^2 bytes
// and each panel ends with the rank-2k update A -= V W^H + W V^H on the FP64 tensor cores
// (gemm_dmma.cu).  The formulas are the ones prototyped and checked against LAPACK in
// tests/algo_proto.py.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "dwhmc.h"
#include "internal.h"

namespace cg = cooperative_groups;

namespace {

constexpr int CC = DW_CC;    // CTAs per cluster (= per chain) of the column-step kernel
constexpr int CT = 256;     // threads per CTA of the column-step kernel

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

// L2 residency hints: the panel columns of V and W are re-read by every column step of the panel
// (evict_last), the trailing matrix streamed by hemv is not (evict_first).
__device__ __forceinline__ unsigned long long policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ cplx ld_keep(const cplx* p, unsigned long long pol) {
  cplx v;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_keep(cplx* p, cplx v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}

// One pass over the panel rows V[r, :], W[r, :] (k < cnt) feeding two accumulators:
//   accF -= V[r,k] f1[k] + W[r,k] f2[k]     (finish w_{j-1})
//   accR -= V[r,k] c1[k] + W[r,k] c2[k]     (update column j)
// Loads are issued in batches of 8 + 8 so a thread has 16 independent requests in flight.
__device__ __forceinline__ void panel_row_update2(cplx& accF, cplx& accR, const cplx* Vr, const cplx* Wr, size_t ld,
                                                  const cplx* f1, const cplx* f2, const cplx* c1, const cplx* c2,
                                                  int cnt, bool doR, unsigned long long pol) {
  for (int k = 0; k < cnt; k += 8) {
    cplx v[8], w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool p = k + u < cnt;
      v[u] = p ? ld_keep(Vr + (size_t)(k + u) * ld, pol) : make_double2(0.0, 0.0);
      w[u] = p ? ld_keep(Wr + (size_t)(k + u) * ld, pol) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (k + u < cnt) {
        cfms(accF, v[u], f1[k + u]);
        cfms(accF, w[u], f2[k + u]);
        if (doR) {
          cfms(accR, v[u], c1[k + u]);
          cfms(accR, w[u], c2[k + u]);
        }
      }
    }
  }
}

struct ColArgs {
  cplx* A; cplx* V; cplx* W; cplx* ypart; cplx* P1; cplx* P2; cplx* tau;
  double* d; double* e;
  int n, B, b0, j, j0, finish_prev, make_ref;
  int skip_dots;      // the products W^H v, V^H v are computed by the dot CTAs of the hemv launch instead
  Mask mask;
};

// Column step of the panel factorisation: a cluster of CC CTAs per chain, each owning a contiguous
// slice of the rows j..n-1.  Block sums are exchanged through distributed shared memory (one slot
// array per CTA, read by every CTA of the cluster after a cluster barrier, summed in rank order).
//   finish_prev:  w_{j-1} = tau (y - V (W^H v) - W (V^H v)),  w -= tau/2 (w^H v) v
//   make_ref:     a = A[j:, j] - V conj(W[j,:]) - W conj(V[j,:]);  reflector v_j, tau_j, e_j, d_j;
//                 partial products W^H v_j, V^H v_j of this CTA's rows (summed by the next launch)
__global__ void __cluster_dims__(CC, 1, 1) __launch_bounds__(CT) colstep_kernel(ColArgs g) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = g.b0 + blockIdx.x / CC;
  if (!g.mask.on(b)) return;          // the whole cluster leaves together
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = g.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = g.j, j0 = g.j0;
  const int chunk = (n - j + CC - 1) / CC;          // rows j..n-1 in CC slices
  const int chunk_max = (n + CC - 1) / CC;
  cplx* sv = reinterpret_cast<cplx*>(smem_raw);   // [chunk_max] reflector rows of this CTA
  cplx* sw = sv + chunk_max;                      // [chunk_max] w, then the updated column a
  cplx* sa = sw + chunk_max;                      // [chunk_max] column j minus the panel terms k < i - 1
  cplx* rowW = sa + chunk_max;                    // [NB] W_panel^H v_{j-1}
  cplx* rowV = rowW + DW_NB;                      // [NB] V_panel^H v_{j-1}
  cplx* cW = rowV + DW_NB;                        // [NB] conj(W[j, panel])
  cplx* cV = cW + DW_NB;                          // [NB] conj(V[j, panel])
  cplx* red = cV + DW_NB;                         // [32]
  __shared__ cplx xch[8];                         // DSMEM exchange: 0 dot, 1 norm, 2 a[j], 3 a[j+1], 4 w[j]
  __shared__ cplx s_scale;
  const int lo = j + rank * chunk, hi = min(n, lo + chunk);
  const size_t mat = (size_t)b * n * n;
  cplx* A = g.A + mat;
  cplx* V = g.V + mat;
  cplx* W = g.W + mat;
  const unsigned long long pol = policy_evict_last();
  const int i = j - j0;

  if (g.finish_prev) {
    const int jp = j - 1, ip = jp - j0;
    if (tid < ip) {
      cplx p1 = make_double2(0.0, 0.0), p2 = p1;
      for (int c = 0; c < CC; ++c) {
        p1 = cadd(p1, g.P1[((size_t)b * CC + c) * DW_NB + tid]);     // W_panel^H v
        p2 = cadd(p2, g.P2[((size_t)b * CC + c) * DW_NB + tid]);     // V_panel^H v
      }
      rowW[tid] = p1;
      rowV[tid] = p2;
      if (g.make_ref) {
        const cplx a = W[(size_t)(j0 + tid) * n + j], c = V[(size_t)(j0 + tid) * n + j];
        cW[tid] = make_double2(a.x, -a.y);
        cV[tid] = make_double2(c.x, -c.y);
      }
    }
    const cplx tau = g.tau[(size_t)b * n + jp];
    const int nslots = (n - jp - 1 + 63) / 64;     // tiles per dimension of the hemv of column jp
    __syncthreads();
    cplx dot = make_double2(0.0, 0.0);
    for (int r = lo + tid; r < hi; r += CT) {
      cplx acc = make_double2(0.0, 0.0);
      for (int s0 = 0; s0 < nslots; s0 += 8) {       // up to 8 independent loads in flight, fixed summation order
        cplx y[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          y[u] = (s0 + u < nslots) ? g.ypart[((size_t)(s0 + u) * g.B + b) * n + r] : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = cadd(acc, y[u]);
      }
      cplx accR = g.make_ref ? A[(size_t)j * n + r] : make_double2(0.0, 0.0);
      // one pass over the panel rows serves both w_{j-1} and the update of column j (terms k < i - 1)
      panel_row_update2(acc, accR, V + (size_t)j0 * n + r, W + (size_t)j0 * n + r, n, rowW, rowV, cW, cV, ip,
                        g.make_ref != 0, pol);
      const cplx wv = cmul(tau, acc);
      const cplx vv = ld_keep(V + (size_t)jp * n + r, pol);
      sw[r - lo] = wv;
      sv[r - lo] = vv;
      sa[r - lo] = accR;
      cfmac(dot, wv, vv);
    }
    dot = block_sum(dot, red);
    if (tid == 0) xch[0] = dot;
    cluster.sync();
    dot = make_double2(0.0, 0.0);
    for (int c = 0; c < CC; ++c) dot = cadd(dot, cluster.map_shared_rank(xch, c)[0]);
    cplx alpha = cmul(tau, dot);
    alpha.x *= -0.5; alpha.y *= -0.5;
    for (int r = lo + tid; r < hi; r += CT) {
      cplx wv = sw[r - lo];
      cfma(wv, alpha, sv[r - lo]);
      sw[r - lo] = wv;
      st_keep(W + (size_t)jp * n + r, wv, pol);
      if (r == j) xch[4] = wv;
    }
    if (!g.make_ref) {
      cluster.sync();       // nobody may exit while its exchange slots can still be read
      return;
    }
  }

  {
    cluster.sync();         // xch[4] of rank 0 (= W[j, j-1]) valid
    cplx wj = make_double2(0.0, 0.0);
    if (g.finish_prev) {
      const cplx t = cluster.map_shared_rank(xch, 0)[4];
      wj = make_double2(t.x, -t.y);
    }
    cplx nrm = make_double2(0.0, 0.0);
    for (int r = lo + tid; r < hi; r += CT) {
      cplx a;
      if (g.finish_prev) {
        // last panel term k = i - 1: V[r, j-1] conj(W[j, j-1]) + W[r, j-1] conj(V[j, j-1]), V[j, j-1] = 1
        a = sa[r - lo];
        cfms(a, sv[r - lo], wj);
        const cplx wr = sw[r - lo];
        a.x -= wr.x; a.y -= wr.y;
      } else {
        a = A[(size_t)j * n + r];       // first column of a panel: the trailing matrix is up to date
      }
      sw[r - lo] = a;
      if (r >= j + 2) nrm.x += a.x * a.x + a.y * a.y;
      if (r == j) xch[2] = a;
      if (r == j + 1) xch[3] = a;
    }
    nrm = block_sum(nrm, red);
    if (tid == 0) xch[1] = nrm;
    cluster.sync();
    if (tid == 0) {
      double xn2 = 0.0;
      for (int c = 0; c < CC; ++c) xn2 += cluster.map_shared_rank(xch, c)[1].x;
      const cplx a0 = cluster.map_shared_rank(xch, 0)[2];
      const cplx alpha = cluster.map_shared_rank(xch, 1 / chunk)[3];   // rank owning row j + 1
      double beta;
      cplx tau, scale;
      if (xn2 == 0.0 && alpha.y == 0.0) {
        beta = alpha.x;
        tau = make_double2(0.0, 0.0);
        scale = make_double2(0.0, 0.0);
      } else {
        beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn2), alpha.x);
        tau = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
        const double dr = alpha.x - beta, di = alpha.y;     // scale = 1 / (alpha - beta)
        const double den = dr * dr + di * di;
        scale = make_double2(dr / den, -di / den);
      }
      if (rank == 0) {
        g.d[(size_t)b * n + j] = a0.x;
        g.e[(size_t)b * n + j] = beta;
        g.tau[(size_t)b * n + j] = tau;
      }
      s_scale = scale;
    }
    __syncthreads();
    const cplx scale = s_scale;
    for (int r = max(lo, j + 1) + tid; r < hi; r += CT) {
      const cplx v = (r == j + 1) ? make_double2(1.0, 0.0) : cmul(sw[r - lo], scale);
      sv[r - lo] = v;
      st_keep(V + (size_t)j * n + r, v, pol);
    }
    __syncthreads();
    // partial products over this CTA's rows: P1[k] = W[:, j0+k]^H v, P2[k] = V[:, j0+k]^H v (k < i)
    // Warp w accumulates the products q = w, w + 8, ... (up to 8 of the 2 i <= 64) at once, so each
    // pass over the rows keeps 8-16 independent loads in flight instead of one product at a time.
    const int nw = CT / 32;
    const int r0 = max(lo, j + 1);
    const int ndots = 2 * i;
    if (warp < ndots && !g.skip_dots) {
      const cplx* src[8];
      cplx acc[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int q = min(warp + nw * t, ndots - 1);
        src[t] = ((q & 1) ? V : W) + (size_t)(j0 + (q >> 1)) * n;
        acc[t] = make_double2(0.0, 0.0);
      }
      for (int r = r0 + lane; r < hi; r += 64) {
        const bool two = r + 32 < hi;
        const cplx v0 = sv[r - lo];
        const cplx v1 = two ? sv[r + 32 - lo] : make_double2(0.0, 0.0);
        cplx x0[8], x1[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          x0[t] = ld_keep(src[t] + r, pol);
          x1[t] = two ? ld_keep(src[t] + r + 32, pol) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          cfmac(acc[t], x0[t], v0);
          cfmac(acc[t], x1[t], v1);
        }
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int q = warp + nw * t;
        const cplx sum = warp_sum(acc[t]);
        if (lane == 0 && q < ndots) {
          if (q & 1) g.P2[((size_t)b * CC + rank) * DW_NB + (q >> 1)] = sum;
          else g.P1[((size_t)b * CC + rank) * DW_NB + (q >> 1)] = sum;
        }
      }
    }
    cluster.sync();         // exchange slots stay alive until every CTA has read them
  }
}

// y = A[q0:, q0:] v reading only the lower triangle (q0 = j + 1).  One CTA per 64x64 tile (I >= J) of
// the trailing block: the tile is staged once in shared memory (cp.async, 16 B) and used twice, for
// y[rows of I] += T v[cols of J] and y[rows of J] += T^H v[rows of I].  Partial results go to slot J
// (row pass) / slot I (column pass) of ypart, so every row of block X receives exactly one
// contribution in each slot 0..nt-1 and the reduction in colstep_kernel has a fixed order.
constexpr int TS = 64;
// Register path (the shared-memory staged and the persistent warp-ring variants of round 1 measured slower and were
// removed in round 2; same tiles, same ypart slots, same result layout):
// no shared-memory staging.  Warp w of the CTA owns columns 8w..8w+7 of the 64x64 tile, lane l the
// rows l and l+32; the 16 elements of a lane are loaded straight into registers (16 independent
// 16-byte requests per thread, each warp request one contiguous 512-byte segment) and used for both
// passes.  The row pass accumulates in the lane; the column pass is reduced across the lanes with a
// transposing butterfly (8 -> 4 -> 2 -> 1 values per lane, 9 complex exchanges instead of 40).
// Diagonal tiles neither load nor use the strict upper triangle.
__device__ __forceinline__ cplx shfl_xor_c(cplx v, int o) {
  return make_double2(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o));
}

// The last CC CTAs of each chain's row of the grid are "dot CTAs": they compute this column's products
// W_panel^H v and V_panel^H v (needed only by the next column step) while the tile CTAs stream the
// trailing matrix, which takes one pass over the panel off the critical path of the column step.
__global__ void __launch_bounds__(256, 2) hemv_reg_kernel(const cplx* __restrict__ Aall, const cplx* __restrict__ Vall,
                                                          cplx* __restrict__ ypart, int n, int B, int b0, int j,
                                                          Mask mask, const cplx* __restrict__ Wall, cplx* __restrict__ P1,
                                                          cplx* __restrict__ P2, int j0, int ntiles) {
  const int b = b0 + blockIdx.y;
  if (!mask.on(b)) return;
  if ((int)blockIdx.x >= ntiles) {
    const int rank = blockIdx.x - ntiles;
    const int i = j - j0, ndots = 2 * i;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = 8;
    if (warp >= ndots) return;
    const int chunk = (n - j + CC - 1) / CC;            // same row slices as the column-step cluster
    const int lo = j + rank * chunk, hi = min(n, lo + chunk);
    const int r0 = max(lo, j + 1);
    const size_t mat = (size_t)b * n * n;
    const cplx* V = Vall + mat;
    const cplx* W = Wall + mat;
    const cplx* v = V + (size_t)j * n;
    const unsigned long long pol = policy_evict_last();
    const cplx* src[8];
    cplx acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int q = min(warp + nw * t, ndots - 1);
      src[t] = ((q & 1) ? V : W) + (size_t)(j0 + (q >> 1)) * n;
      acc[t] = make_double2(0.0, 0.0);
    }
    for (int r = r0 + lane; r < hi; r += 64) {
      const bool two = r + 32 < hi;
      const cplx v0 = v[r];
      const cplx v1 = two ? v[r + 32] : make_double2(0.0, 0.0);
      cplx x0[8], x1[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        x0[t] = ld_keep(src[t] + r, pol);
        x1[t] = two ? ld_keep(src[t] + r + 32, pol) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        cfmac(acc[t], x0[t], v0);
        cfmac(acc[t], x1[t], v1);
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int q = warp + nw * t;
      const cplx sum = warp_sum(acc[t]);
      if (lane == 0 && q < ndots) {
        if (q & 1) P2[((size_t)b * CC + rank) * DW_NB + (q >> 1)] = sum;
        else P1[((size_t)b * CC + rank) * DW_NB + (q >> 1)] = sum;
      }
    }
    return;
  }
  __shared__ cplx red[8 * TS];
  __shared__ cplx cs[TS];
  const int q0 = j + 1, m = n - q0;
  const int t = blockIdx.x;
  int I = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((I + 1) * (I + 2) / 2 <= t) ++I;
  while (I * (I + 1) / 2 > t) --I;
  const int J = t - I * (I + 1) / 2;
  const int R0 = I * TS, C0 = J * TS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool diag = (I == J);
  const size_t mat = (size_t)b * n * n;
  const int cl0 = warp * 8;                                  // first local column of this warp
  const cplx* A = Aall + mat + (size_t)(q0 + C0 + cl0) * n + q0 + R0;
  const cplx* v = Vall + mat + (size_t)j * n + q0;
  const unsigned long long pol = policy_evict_first();
  const cplx zero = make_double2(0.0, 0.0);
  cplx a[2][8];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int rl = lane + 32 * hh, cl = cl0 + c;
      const bool p = (R0 + rl < m) && (C0 + cl < m) && (!diag || rl >= cl);
      a[hh][c] = p ? ld_keep(A + (size_t)c * n + rl, pol) : zero;
    }
  cplx vr[2];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) vr[hh] = (R0 + lane + 32 * hh < m) ? v[R0 + lane + 32 * hh] : zero;
  cplx accR[2] = {zero, zero};
  cplx accC[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int cl = cl0 + c;
    const cplx vc = (C0 + cl < m) ? v[C0 + cl] : zero;
    cplx s = zero;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const cplx x = a[hh][c];
      cfma(accR[hh], x, vc);
      const bool ondiag = diag && (lane + 32 * hh == cl);
      if (!ondiag) cfmac(s, x, vr[hh]);
    }
    accC[c] = s;
  }
  // transposing butterfly: afterwards the lanes with lane >> 2 == c hold the sum of column c
  cplx r4[4], r2[2], r1;
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const cplx send = up ? accC[k] : accC[k + 4];
      const cplx keep = up ? accC[k + 4] : accC[k];
      r4[k] = cadd(keep, shfl_xor_c(send, 16));
    }
  }
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const cplx send = up ? r4[k] : r4[k + 2];
      const cplx keep = up ? r4[k + 2] : r4[k];
      r2[k] = cadd(keep, shfl_xor_c(send, 8));
    }
  }
  {
    const bool up = (lane & 4) != 0;
    const cplx send = up ? r2[0] : r2[1];
    const cplx keep = up ? r2[1] : r2[0];
    r1 = cadd(keep, shfl_xor_c(send, 4));
  }
  r1 = cadd(r1, shfl_xor_c(r1, 2));
  r1 = cadd(r1, shfl_xor_c(r1, 1));
  if ((lane & 3) == 0) cs[cl0 + (lane >> 2)] = r1;
  red[warp * TS + lane] = accR[0];
  red[warp * TS + lane + 32] = accR[1];
  __syncthreads();
  if (tid < TS) {
    cplx sum = red[tid];
#pragma unroll
    for (int w = 1; w < 8; ++w) sum = cadd(sum, red[w * TS + tid]);
    if (diag) sum = cadd(sum, cs[tid]);
    if (R0 + tid < m) ypart[((size_t)(diag ? I : J) * B + b) * n + q0 + R0 + tid] = sum;
  } else if (tid < 2 * TS && !diag) {
    const int x = tid - TS;
    if (C0 + x < m) ypart[((size_t)I * B + b) * n + q0 + C0 + x] = cs[x];
  }
}

__global__ void lastd_kernel(const cplx* __restrict__ A, double* __restrict__ d, int n, int B, Mask mask) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || !mask.on(b)) return;
  d[(size_t)b * n + n - 1] = A[(size_t)b * n * n + (size_t)(n - 1) * n + (n - 1)].x;
}

}  // namespace

int dw_hetrd(Handle* h, cplx* W, Mask mask) {
  const int n = h->n, B = h->B;
  if (n < 2) {
    lastd_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(h->A, h->d, n, B, mask);
    DW_LAUNCH_CHECK(h);
    return DWHMC_OK;
  }
  const size_t col_smem = sizeof(cplx) * (3 * (size_t)((n + CC - 1) / CC) + 4 * DW_NB + 32);
  static bool attr_set[64] = {false};
  if (!attr_set[h->device & 63]) {
    DW_CUDA(h, cudaFuncSetAttribute(colstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set[h->device & 63] = true;
  }
  if (col_smem > 200 * 1024) { h->err = "dw_hetrd: matrix too large for the column-step kernel"; return DWHMC_E_BADARG; }
  // The batch is split into groups.  Each group has a high-priority stream for the latency-bound
  // column step (few CTAs) and a low-priority stream for the bulk kernels (hemv, her2k), chained by
  // events, so the column step of one group is scheduled ahead of the queued hemv tiles of the
  // others and overlaps them instead of waiting for their tail.
  const int G = (B >= 2 * h->ngroups && n >= 256 && h->profiling < 2) ? h->ngroups : 1;
  cudaStream_t hp[DW_NGROUP], lp[DW_NGROUP];
  int gb0[DW_NGROUP], gB[DW_NGROUP];
  for (int g = 0; g < G; ++g) {
    gb0[g] = (int)((long long)B * g / G);
    gB[g] = (int)((long long)B * (g + 1) / G) - gb0[g];
    hp[g] = (G == 1) ? h->stream : h->gstream_hi[g];
    lp[g] = (G == 1) ? h->stream : h->gstream[g];
  }
  if (G > 1) {
    DW_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
    for (int g = 0; g < G; ++g) {
      DW_CUDA(h, cudaStreamWaitEvent(hp[g], h->ev_fork, 0));
      DW_CUDA(h, cudaStreamWaitEvent(lp[g], h->ev_fork, 0));
    }
  }
  // bulk -> column step and column step -> bulk hand-over (no-ops when both are the same stream)
  auto bulk_done = [&](int g) {
    if (G > 1) { cudaEventRecord(h->ev_bulk[g], lp[g]); cudaStreamWaitEvent(hp[g], h->ev_bulk[g], 0); }
  };
  auto col_done = [&](int g) {
    if (G > 1) { cudaEventRecord(h->ev_col[g], hp[g]); cudaStreamWaitEvent(lp[g], h->ev_col[g], 0); }
  };
  ColArgs ca;
  ca.A = h->A; ca.V = h->V; ca.W = W; ca.ypart = h->ypart; ca.P1 = h->P1; ca.P2 = h->P2; ca.tau = h->tau;
  ca.d = h->d; ca.e = h->e; ca.n = n; ca.B = B; ca.mask = mask;
  for (int j0 = 0; j0 < n - 1; j0 += DW_NB) {
    const int pn = (n - 1 - j0 < DW_NB) ? n - 1 - j0 : DW_NB;
    ca.j0 = j0;
    for (int i = 0; i < pn; ++i) {
      const int j = j0 + i;
      const int m = n - j - 1;
      const int nt = (m + TS - 1) / TS;
      for (int g = 0; g < G; ++g) {
        ca.j = j; ca.finish_prev = (i > 0); ca.make_ref = 1; ca.b0 = gb0[g];
        ca.skip_dots = 0;
        colstep_kernel<<<gB[g] * CC, CT, col_smem, hp[g]>>>(ca);
        DW_LAUNCH_CHECK(h);
        col_done(g);
        dim3 grid(nt * (nt + 1) / 2, gB[g]);
        if (h->profiling >= 2) cudaEventRecord(h->ev_begin, lp[g]);
        hemv_reg_kernel<<<grid, 256, 0, lp[g]>>>(h->A, h->V, h->ypart, n, B, gb0[g], j, mask, W, h->P1, h->P2, j0, (int)grid.x);
        DW_LAUNCH_CHECK(h);
        if (h->profiling >= 2) {
          cudaEventRecord(h->ev_end, lp[g]);
          cudaEventSynchronize(h->ev_end);
          float ms = 0.f;
          cudaEventElapsedTime(&ms, h->ev_begin, h->ev_end);
          h->timers[7] += ms;
        }
        bulk_done(g);
      }
    }
    const int j1 = j0 + pn;
    for (int g = 0; g < G; ++g) {
      ca.j = j1; ca.finish_prev = 1; ca.make_ref = 0; ca.b0 = gb0[g]; ca.skip_dots = 0;
      colstep_kernel<<<gB[g] * CC, CT, col_smem, hp[g]>>>(ca);
      DW_LAUNCH_CHECK(h);
      col_done(g);
      // trailing update A[j1:, j1:] -= V W^H + W V^H, lower-triangle tiles only
      ZgemmArgs a;
      a.M = n - j1; a.N = n - j1; a.K = pn; a.nseg = 2;
      a.A[0] = h->V + (size_t)j0 * n + j1; a.Bm[0] = W + (size_t)j0 * n + j1;
      a.A[1] = W + (size_t)j0 * n + j1;    a.Bm[1] = h->V + (size_t)j0 * n + j1;
      a.lda = n; a.ldb = n; a.ldc = n;
      a.sA = (long long)n * n; a.sB = (long long)n * n; a.sC = (long long)n * n;
      a.C = h->A + (size_t)j1 * n + j1;
      a.alpha = -1.0; a.beta = 1.0; a.opA = 0; a.opB = 1; a.lower = 1; a.batch = gB[g]; a.mask = mask;
      a.b0 = gb0[g]; a.stream = lp[g];
      DW_TRY(dw_zgemm(h, a));
      bulk_done(g);
    }
  }
  if (G > 1) {
    for (int g = 0; g < G; ++g) {
      DW_CUDA(h, cudaEventRecord(h->ev_join[g], hp[g]));
      DW_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join[g], 0));
    }
  }
  lastd_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(h->A, h->d, n, B, mask);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}
