// stedc.cu -- batched divide-and-conquer eigensolver for the real symmetric tridiagonal
// matrices produced by hetrd.cu (second stage of diagonalize_H_BdG!,
// /root/reference src/Hamiltonian.jl:96-114; the reference reaches LAPACK's tridiagonal
// solver through eigen!).  Cuppen's method with Gu/Eisenstat eigenvectors; the scalar
// numerics live in stedc_core.h and are unit-tested on the CPU (tests/hostcheck).
//
// Every matrix of the batch shares one tree (stedc_tree.h): leaves of <= DW_LEAF rows solved by
// implicit QL (one warp each), then one batch of merges per level:
//   prepare   z vector, merged order, deflation scan, Givens rotations of deflated pairs
//   copy      deflated eigenvectors to the tail of the output block
//   secular   one thread per root of the secular equation
//   zhat      Gu/Eisenstat recomputation of the rank-one vector
//   vectors   normalised eigenvectors S of the rank-one modified diagonal matrix
//   gemm      Q_out = Q_in[:, nondeflated] * S on the FP64 tensor cores (DMMA m8n8k4)
//   finish    new eigenvalue order
#include "dwhmc.h"
#include "gemm_dmma.cuh"
#include "internal.h"
#include "stedc_core.h"

using namespace dwcore;
using namespace dwg;

namespace {

struct LeafRot {
  double* Z;   // shared, column-major, leading dimension ldz
  int ldz, lane, sz;
  __device__ __forceinline__ void operator()(int i, double c, double s) {
    for (int k = lane; k < sz; k += 32) {
      const double f = Z[(i + 1) * ldz + k];
      const double g = Z[i * ldz + k];
      Z[(i + 1) * ldz + k] = s * g + c * f;
      Z[i * ldz + k] = c * g - s * f;
    }
  }
};

constexpr int LDZ = DW_LEAF + 1;

// one warp per (leaf, chain).  Every lane runs the scalar QL recurrence on private copies of
// d and e (identical arithmetic, so identical results) and applies the rotations to its own rows.
__global__ void __launch_bounds__(32) leaf_kernel(double* __restrict__ Dall, const double* __restrict__ Eall,
                                                  double* __restrict__ Zall, int* __restrict__ permall,
                                                  const int* __restrict__ leaf_off, const int* __restrict__ leaf_size,
                                                  int n, int* status, Mask mask) {
  const int b = blockIdx.y;
  if (!mask.on(b)) return;
  const int off = leaf_off[blockIdx.x], sz = leaf_size[blockIdx.x];
  const int lane = threadIdx.x;
  __shared__ double Zs[DW_LEAF * LDZ];
  __shared__ double ds[DW_LEAF];
  double dd[DW_LEAF + 1], ee[DW_LEAF + 1];
  double* D = Dall + (size_t)b * n + off;
  const double* E = Eall + (size_t)b * n + off;
  for (int i = 0; i < sz; ++i) {
    dd[i] = D[i];
    ee[i] = (i + 1 < sz) ? E[i] : 0.0;
  }
  // tear: the coupling to the neighbouring leaves is removed from the diagonal ends
  if (off > 0) dd[0] -= fabs(E[-1]);
  if (off + sz < n) dd[sz - 1] -= fabs(E[sz - 1]);
  for (int c = 0; c < sz; ++c)
    for (int k = lane; k < sz; k += 32) Zs[c * LDZ + k] = (c == k) ? 1.0 : 0.0;
  __syncwarp();
  LeafRot rot{Zs, LDZ, lane, sz};
  const int info = tql_implicit(sz, dd, ee, rot);
  if (info && lane == 0) atomicAdd(&status[0], 1);
  __syncwarp();
  if (lane == 0)
    for (int i = 0; i < sz; ++i) ds[i] = dd[i];
  __syncwarp();
  int* perm = permall + (size_t)b * n + off;
  for (int k = lane; k < sz; k += 32) perm[k] = k;
  __syncwarp();
  for (int k = lane; k < sz; k += 32) {
    const double v = ds[k];
    int rank = 0;
    for (int q = 0; q < sz; ++q) rank += (ds[q] < v) || (ds[q] == v && q < k);
    perm[rank] = k;
    D[k] = v;
  }
  double* Z = Zall + (size_t)b * n * n + (size_t)off * n + off;
  for (int c = 0; c < sz; ++c)
    for (int k = lane; k < sz; k += 32) Z[(size_t)c * n + k] = Zs[c * LDZ + k];
}

struct LevelArgs {
  const int* off; const int* n1; const int* n2;
  int n, B;
  double* D; const double* E; double* Zin; double* Zout; double* S;
  int* perm; int* nd; int* df; DeflRot* rots; int* kcnt; int* nrot; double* rho;
  int* pos; int* ndc; int* kcls;     // class order of the non-deflated columns (top-only | mixed | bottom-only)
  double* dl; double* wv; double* dnew; double* zhat; double* stau; int* sorg;
  int* status;
  Mask mask;
};

__device__ __forceinline__ double block_max(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = red[0];
  for (int i = 1; i < nw; ++i) t = fmax(t, red[i]);
  return t;
}

__global__ void __launch_bounds__(256) dc_prepare_kernel(LevelArgs g) {
  const int b = blockIdx.y;
  if (!g.mask.on(b)) return;
  const int off = g.off[blockIdx.x], n1 = g.n1[blockIdx.x], n2 = g.n2[blockIdx.x], m = n1 + n2;
  const int n = g.n, tid = threadIdx.x;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sd = reinterpret_cast<double*>(smem_raw);
  double* sz = sd + m;
  int* sord = reinterpret_cast<int*>(sz + m);
  int* smix = sord + m;
  __shared__ double red[32];
  __shared__ int s_k, s_nrot;

  const size_t vo = (size_t)b * n + off;
  double* D = g.D + vo;
  double* Q = g.Zin + (size_t)b * n * n + (size_t)off * n + off;
  const double beta = g.E[vo + n1 - 1];
  const double sgn = beta < 0.0 ? -1.0 : 1.0;
  const double rho = 2.0 * fabs(beta);
  const double isq2 = 0.70710678118654752440;
  double dmax = 0.0, zmax = 0.0;
  for (int i = tid; i < m; i += blockDim.x) {
    const double dv = D[i];
    const double zv = (i < n1 ? Q[(size_t)i * n + (n1 - 1)] : sgn * Q[(size_t)i * n + n1]) * isq2;
    sd[i] = dv;
    sz[i] = zv;
    sord[i] = i;
    smix[i] = 0;
    dmax = fmax(dmax, fabs(dv));
    zmax = fmax(zmax, fabs(zv));
  }
  dmax = block_max(dmax, red);
  zmax = block_max(zmax, red);
  // merged ascending order of the two sorted halves (ties: first half first)
  const int* p1 = g.perm + vo;
  const int* p2 = g.perm + vo + n1;
  for (int a = tid; a < n1; a += blockDim.x) {
    const int ia = p1[a];
    const double key = sd[ia];
    int lo = 0, hi = n2;            // count of second-half entries strictly below key
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sd[n1 + p2[mid]] < key) lo = mid + 1; else hi = mid;
    }
    sord[a + lo] = ia;
  }
  for (int c = tid; c < n2; c += blockDim.x) {
    const int ic = n1 + p2[c];
    const double key = sd[ic];
    int lo = 0, hi = n1;            // count of first-half entries <= key
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sd[p1[mid]] <= key) lo = mid + 1; else hi = mid;
    }
    sord[c + lo] = ic;
  }
  __syncthreads();
  int* nd = g.nd + vo;
  int* df = g.df + vo;
  DeflRot* rots = g.rots + vo;
  if (tid == 0) {
    const double tol = 8.0 * DW_EPS * fmax(dmax, zmax);
    int k = 0, nrot = 0;
    if (rho * zmax <= tol) {
      for (int r = 0; r < m; ++r) df[r] = sord[r];
    } else {
      k = deflate_scan(m, rho, tol, sord, sd, sz, nd, df, rots, &nrot);
    }
    s_k = k;
    s_nrot = nrot;
    g.kcnt[vo] = k;
    g.nrot[vo] = nrot;
    g.rho[vo] = rho;
    // The input eigenvector block is block diagonal (n1 x n1, n2 x n2); only columns touched by a
    // rotation across the halves are dense.  Order the non-deflated columns top-only | mixed |
    // bottom-only so the eigenvector GEMM can skip the zero half of every row tile.
    for (int r = 0; r < nrot; ++r) {
      const int a = rots[r].a, c2 = rots[r].b;
      if (((a < n1) != (c2 < n1)) || smix[a] || smix[c2]) { smix[a] = 1; smix[c2] = 1; }
    }
    int cnt[3] = {0, 0, 0};
    for (int q = 0; q < k; ++q) { const int j = nd[q]; cnt[smix[j] ? 1 : (j < n1 ? 0 : 2)]++; }
    int start[3] = {0, cnt[0], cnt[0] + cnt[1]};
    g.kcls[vo] = cnt[0];
    g.kcls[vo + 1] = cnt[0] + cnt[1];
    for (int q = 0; q < k; ++q) {
      const int j = nd[q];
      const int cls = smix[j] ? 1 : (j < n1 ? 0 : 2);
      const int at = start[cls]++;
      g.pos[vo + q] = at;
      g.ndc[vo + at] = j;
    }
  }
  __syncthreads();
  const int k = s_k, nrot = s_nrot;
  for (int i = tid; i < m; i += blockDim.x) D[i] = sd[i];
  for (int p = tid; p < k; p += blockDim.x) {
    const int src = nd[p];
    g.dl[vo + p] = sd[src];
    g.wv[vo + p] = sz[src];
  }
  for (int r = tid; r < m - k; r += blockDim.x) g.dnew[vo + k + r] = sd[df[r]];
  // Givens rotations of the deflated pairs; a thread always owns the same rows, so no barrier
  for (int r = 0; r < nrot; ++r) {
    const DeflRot rt = rots[r];
    double* qa = Q + (size_t)rt.a * n;
    double* qb = Q + (size_t)rt.b * n;
    for (int i = tid; i < m; i += blockDim.x) {
      const double x = qa[i], y = qb[i];
      qa[i] = rt.c * x + rt.s * y;
      qb[i] = rt.c * y - rt.s * x;
    }
  }
}

// deflated columns go to positions k..m-1 of the output block
__global__ void __launch_bounds__(256) dc_copy_kernel(LevelArgs g) {
  const int b = blockIdx.z;
  if (!g.mask.on(b)) return;
  const int off = g.off[blockIdx.y], m = g.n1[blockIdx.y] + g.n2[blockIdx.y];
  const int n = g.n;
  const size_t vo = (size_t)b * n + off;
  const int k = g.kcnt[vo];
  const int ndf = m - k;
  const int r0 = blockIdx.x * 32;
  if (r0 >= ndf) return;
  const int r1 = min(ndf, r0 + 32);
  const double* Qi = g.Zin + (size_t)b * n * n + (size_t)off * n + off;
  double* Qo = g.Zout + (size_t)b * n * n + (size_t)off * n + off;
  const int* df = g.df + vo;
  for (int r = r0; r < r1; ++r) {
    const double* src = Qi + (size_t)df[r] * n;
    double* dst = Qo + (size_t)(k + r) * n;
    for (int i = threadIdx.x; i < m; i += blockDim.x) dst[i] = src[i];
  }
}

// warp-parallel reduction policy for secular_root (same interface as dwcore::SerialPar): lanes
// stride over the poles; the butterfly sum leaves bit-identical totals in every lane, so the
// iteration's control flow stays uniform across the warp
struct WarpPar {
  int lane;
  __device__ __forceinline__ int begin() const { return lane; }
  __device__ __forceinline__ int stride() const { return 32; }
  __device__ __forceinline__ void sum4(double& a, double& b, double& c, double& d) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
      d += __shfl_xor_sync(0xffffffffu, d, o);
    }
  }
};

// one warp per root of the secular equation; 8 roots per CTA share the staged poles
__global__ void __launch_bounds__(256) dc_secular_kernel(LevelArgs g) {
  const int b = blockIdx.z;
  if (!g.mask.on(b)) return;
  const int off = g.off[blockIdx.y];
  const int n = g.n;
  const size_t vo = (size_t)b * n + off;
  const int k = g.kcnt[vo];
  if ((int)(blockIdx.x * 8) >= k) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sdl = reinterpret_cast<double*>(smem_raw);
  double* sw = sdl + k;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    sdl[i] = g.dl[vo + i];
    sw[i] = g.wv[vo + i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 8 + warp;
  if (j >= k) return;
  WarpPar par{lane};
  int org;
  double tau;
  const int it = secular_root(k, j, sdl, sw, g.rho[vo], par, &org, &tau);
  if (lane == 0) {
    if (it < 0) atomicAdd(&g.status[1], 1);
    g.sorg[vo + j] = org;
    g.stau[vo + j] = tau;
    g.dnew[vo + j] = sdl[org] + tau;
  }
}

// Gu/Eisenstat: zhat_i^2 = prod_j (lambda_j - d_i) / prod_{j != i} (d_j - d_i)
__global__ void __launch_bounds__(128) dc_zhat_kernel(LevelArgs g) {
  const int b = blockIdx.z;
  if (!g.mask.on(b)) return;
  const int off = g.off[blockIdx.y];
  const int n = g.n;
  const size_t vo = (size_t)b * n + off;
  const int k = g.kcnt[vo];
  if ((int)(blockIdx.x * blockDim.x) >= k) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sdl = reinterpret_cast<double*>(smem_raw);
  double* sdo = sdl + k;     // d[org_j]
  double* sta = sdo + k;     // tau_j
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    sdl[i] = g.dl[vo + i];
    sta[i] = g.stau[vo + i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x) sdo[i] = sdl[g.sorg[vo + i]];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  const double di = sdl[i];
  double prod = -((di - sdo[i]) - sta[i]);
  for (int j = 0; j < k; ++j) {
    if (j == i) continue;
    const double del = (di - sdo[j]) - sta[j];
    prod *= del / (di - sdl[j]);
  }
  g.zhat[vo + i] = copysign(sqrt(fabs(prod)), g.wv[vo + i]);
}

// S[:, j] = normalised ( zhat_i / (d_i - lambda_j) )_i ; one warp per column
__global__ void __launch_bounds__(256) dc_vectors_kernel(LevelArgs g) {
  const int b = blockIdx.z;
  if (!g.mask.on(b)) return;
  const int off = g.off[blockIdx.y];
  const int n = g.n;
  const size_t vo = (size_t)b * n + off;
  const int k = g.kcnt[vo];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 8 + warp;
  if (j >= k) return;
  const double* dl = g.dl + vo;
  const double* zh = g.zhat + vo;
  const double dorg = dl[g.sorg[vo + j]], tau = g.stau[vo + j];
  double nrm = 0.0;
  for (int i = lane; i < k; i += 32) {
    const double v = zh[i] / ((dl[i] - dorg) - tau);
    nrm += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
  const double sc = 1.0 / sqrt(nrm);
  double* Sc = g.S + (size_t)b * n * n + (size_t)(off + j) * n + off;
  const int* pos = g.pos + vo;
  for (int i = lane; i < k; i += 32) Sc[pos[i]] = zh[i] / ((dl[i] - dorg) - tau) * sc;
}

// ---- Q_out[:, 0:k] = Q_in[:, nd[0:k]] * S[0:k, 0:k]  (real, DMMA m8n8k4) ------------------
constexpr int GBM = 128, GBN = 64, GBK = 16, GST = 3;
constexpr int GLDA = GBM + 4, GLDB = GBK + 4;
constexpr int G_STAGE = GBK * GLDA + GBN * GLDB;
constexpr size_t G_SMEM = sizeof(double) * GST * G_STAGE;

struct GemmArgs {
  LevelArgs lv;
  int ntiles_n;   // column tiles per merge
  // last level only: chains flagged for the particle-hole shortcut need the eigenvectors of rank >= c_lo only.
  // Root j of the secular equation has rank <= j + (deflated values), so column tiles wholly below are skipped.
  const int* halfflag;
  int c_lo;
};

__global__ void __launch_bounds__(256) dc_gemm2_kernel(GemmArgs ga) {
  const LevelArgs& g = ga.lv;
  const int b = blockIdx.z;
  if (!g.mask.on(b)) return;
  const int mi = blockIdx.y / ga.ntiles_n, tn = blockIdx.y - mi * ga.ntiles_n;
  const int off = g.off[mi], m = g.n1[mi] + g.n2[mi];
  const int n = g.n;
  const size_t vo = (size_t)b * n + off;
  const int k = g.kcnt[vo];
  const int m0 = blockIdx.x * GBM, n0 = tn * GBN;
  if (m0 >= m || n0 >= k) return;
  if (ga.c_lo > 0 && ga.halfflag[b] != 0 && n0 + GBN - 1 + (m - k) < ga.c_lo) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp & 3) * 32, wn0 = (warp >> 2) * 32;
  const size_t blk = (size_t)b * n * n + (size_t)off * n + off;
  const double* Qi = g.Zin + blk;
  const double* S = g.S + blk;
  double* Qo = g.Zout + blk;
  const int* nd = g.ndc + vo;
  // rows of the first input block see top-only and mixed columns, rows of the second mixed and bottom-only
  const int n1 = g.n1[mi];
  const int kT = g.kcls[vo], kTM = g.kcls[vo + 1];
  const int kb = (m0 >= n1) ? kT : 0;
  const int ke = (m0 + GBM <= n1) ? kTM : k;
  const int KT = (ke - kb + GBK - 1) / GBK;

  auto load_tile = [&](int kt, int stage) {
    const int k0 = kb + kt * GBK;
    double* As = smem + (size_t)stage * G_STAGE;
    double* Bs = As + GBK * GLDA;
#pragma unroll
    for (int i = 0; i < (GBM * GBK) / 256; ++i) {
      const int idx = tid + i * 256;
      const int mm = idx % GBM, kk = idx / GBM;
      const bool p = (m0 + mm < m) && (k0 + kk < ke);
      const double* src = p ? Qi + (size_t)nd[k0 + kk] * n + (m0 + mm) : Qi;
      cp_async8(As + kk * GLDA + mm, src, p);
    }
#pragma unroll
    for (int i = 0; i < (GBN * GBK) / 256; ++i) {
      const int idx = tid + i * 256;
      const int kk = idx % GBK, nn = idx / GBK;
      const bool p = (n0 + nn < k) && (k0 + kk < ke);
      const double* src = p ? S + (size_t)(n0 + nn) * n + (k0 + kk) : S;
      cp_async8(Bs + nn * GLDB + kk, src, p);
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < GST - 1; ++s) {
    if (s < KT) load_tile(s, s);
    cp_async_commit();
  }
  const int fr = lane >> 2, fk = lane & 3;
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<GST - 2>();
    __syncthreads();
    {
      const int nk = kt + GST - 1;
      if (nk < KT) load_tile(nk, nk % GST);
      cp_async_commit();
    }
    const double* As = smem + (size_t)(kt % GST) * G_STAGE;
    const double* Bs = As + GBK * GLDA;
#pragma unroll
    for (int k4 = 0; k4 < GBK / 4; ++k4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = As[(k4 * 4 + fk) * GLDA + wm0 + i * 8 + fr];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[(wn0 + j * 8 + fr) * GLDB + k4 * 4 + fk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int row = m0 + wm0 + i * 8 + fr;
        const int col = n0 + wn0 + j * 8 + 2 * fk + q;
        if (row < m && col < k) Qo[(size_t)col * n + row] = acc[i][j][q];
      }
}

__global__ void __launch_bounds__(256) dc_finish_kernel(LevelArgs g) {
  const int b = blockIdx.y;
  if (!g.mask.on(b)) return;
  const int off = g.off[blockIdx.x], m = g.n1[blockIdx.x] + g.n2[blockIdx.x];
  const int n = g.n, tid = threadIdx.x;
  const size_t vo = (size_t)b * n + off;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sd = reinterpret_cast<double*>(smem_raw);
  int* perm = g.perm + vo;
  for (int i = tid; i < m; i += blockDim.x) {
    const double v = g.dnew[vo + i];
    sd[i] = v;
    g.D[vo + i] = v;
    perm[i] = i;
  }
  __syncthreads();
  for (int i = tid; i < m; i += blockDim.x) {
    const double v = sd[i];
    int rank = 0;
    for (int q = 0; q < m; ++q) rank += (sd[q] < v) || (sd[q] == v && q < i);
    perm[rank] = i;
  }
}

// Last level, after the secular equation: all eigenvalues of the chain are known (unsorted).  The particle-hole
// decision of dc_evals_kernel, made early so that the eigenvector GEMM of the last level can skip the columns of
// the lower half of the spectrum: flag = (the level of rank n/2 + 1 is resolved from zero).
__global__ void __launch_bounds__(256) dc_halfflag_kernel(const double* __restrict__ dnew, int* __restrict__ halfflag, int n,
                                                          int ph, Mask mask) {
  const int b = blockIdx.x;
  if (!mask.on(b)) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sd = reinterpret_cast<double*>(smem_raw);
  __shared__ double red[32];
  __shared__ double s_val;
  double emax = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = dnew[(size_t)b * n + i];
    sd[i] = v;
    emax = fmax(emax, fabs(v));
  }
  emax = block_max(emax, red);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = sd[i];
    int rank = 0;
    for (int q = 0; q < n; ++q) rank += (sd[q] < v) || (sd[q] == v && q < i);
    if (rank == n / 2 + 1) s_val = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) halfflag[b] = (ph && n >= 8 && s_val > 1e-4 * emax) ? 1 : 0;
}

// eigenvalues ascending; decide whether the particle-hole shortcut is safe for this chain: the
// partner construction needs the second-smallest level of the upper half to be resolved from zero
// (a pair (E, -E) alone is always fine, <psi, C psi> = 0 exactly; two pairs closer to zero than the
// solver's resolution are not)
__global__ void __launch_bounds__(256) dc_evals_kernel(const double* __restrict__ D, const int* __restrict__ perm,
                                                       double* __restrict__ E_out, int* __restrict__ halfflag, int n,
                                                       int ph, Mask mask) {
  const int b = blockIdx.x;
  if (!mask.on(b)) return;
  double* E = E_out + (size_t)b * n;
  for (int r = threadIdx.x; r < n; r += blockDim.x) E[r] = D[(size_t)b * n + perm[(size_t)b * n + r]];
  __syncthreads();
  if (threadIdx.x == 0) {
    int flag = 0;
    if (ph && n >= 8) {
      const double emax = fmax(fabs(E[0]), fabs(E[n - 1]));
      flag = (E[n / 2 + 1] > 1e-4 * emax) ? 1 : 0;
    }
    halfflag[b] = flag;
  }
}

// eigenvectors as complex columns of U (flagged chains: columns >= c_lo only)
__global__ void __launch_bounds__(256) dc_output_kernel(const int* __restrict__ perm, const double* __restrict__ Z,
                                                        cplx* __restrict__ U_out, const int* __restrict__ halfflag,
                                                        int c_lo, int n, Mask mask) {
  const int b = blockIdx.y;
  if (!mask.on(b)) return;
  const int r = blockIdx.x;
  if (r < c_lo && halfflag[b] != 0) return;
  const int c = perm[(size_t)b * n + r];
  const double* src = Z + (size_t)b * n * n + (size_t)c * n;
  cplx* dst = U_out + (size_t)b * n * n + (size_t)r * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = make_double2(src[i], 0.0);
}

}  // namespace

int dw_stedc(Handle* h, Mask mask, bool ph, int c_lo) {
  const int n = h->n, B = h->B;
  // both ping-pong buffers start from zero: merges read the off-diagonal blocks of their input
  DW_CUDA(h, cudaMemsetAsync(h->Z0, 0, sizeof(double) * (size_t)n * n * B, h->stream));
  DW_CUDA(h, cudaMemsetAsync(h->Z1, 0, sizeof(double) * (size_t)n * n * B, h->stream));
  {
    dim3 grid(h->nleaves, B);
    leaf_kernel<<<grid, 32, 0, h->stream>>>(h->d, h->e, h->Z0, h->perm, h->leaf_off, h->leaf_size, n, h->status, mask);
    DW_LAUNCH_CHECK(h);
  }
  static bool attr_set[64] = {false};
  if (!attr_set[h->device & 63]) {
    DW_CUDA(h, cudaFuncSetAttribute(dc_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM));
    DW_CUDA(h, cudaFuncSetAttribute(dc_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    DW_CUDA(h, cudaFuncSetAttribute(dc_zhat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    DW_CUDA(h, cudaFuncSetAttribute(dc_secular_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    DW_CUDA(h, cudaFuncSetAttribute(dc_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    DW_CUDA(h, cudaFuncSetAttribute(dc_halfflag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set[h->device & 63] = true;
  }
  double* Zin = h->Z0;
  double* Zout = h->Z1;
  for (size_t l = 0; l < h->levels.size(); ++l) {
    const DcLevelDev& lv = h->levels[l];
    LevelArgs g;
    g.off = lv.off; g.n1 = lv.n1; g.n2 = lv.n2; g.n = n; g.B = B;
    g.D = h->d; g.E = h->e; g.Zin = Zin; g.Zout = Zout; g.S = h->S;
    g.perm = h->perm; g.nd = h->ndl; g.df = h->dfl; g.rots = reinterpret_cast<DeflRot*>(h->rots);
    g.kcnt = h->kcnt; g.nrot = h->nrot; g.rho = h->rho;
    g.pos = h->ord; g.ndc = h->ndc; g.kcls = h->kcls;
    g.dl = h->dl; g.wv = h->wv; g.dnew = h->dnew; g.zhat = h->zhat; g.stau = h->stau; g.sorg = h->sorg;
    g.status = h->status; g.mask = mask;
    const int mm = lv.max_m;
    if ((size_t)mm * 24 > 96 * 1024) { h->err = "dw_stedc: matrix too large"; return DWHMC_E_BADARG; }
    {
      dim3 grid(lv.nmerge, B);
      dc_prepare_kernel<<<grid, 256, (size_t)mm * 24 + 16, h->stream>>>(g);
      DW_LAUNCH_CHECK(h);
    }
    {
      dim3 grid((mm + 31) / 32, lv.nmerge, B);
      dc_copy_kernel<<<grid, 256, 0, h->stream>>>(g);
      DW_LAUNCH_CHECK(h);
    }
    {
      dim3 sgrid((mm + 7) / 8, lv.nmerge, B);
      dc_secular_kernel<<<sgrid, 256, (size_t)mm * 16, h->stream>>>(g);
      DW_LAUNCH_CHECK(h);
      dim3 grid((mm + 127) / 128, lv.nmerge, B);
      dc_zhat_kernel<<<grid, 128, (size_t)mm * 24, h->stream>>>(g);
      DW_LAUNCH_CHECK(h);
    }
    const bool last = (l + 1 == h->levels.size()) && lv.nmerge == 1;
    const bool skip_low = last && ph && h->ph_mode && c_lo > 0;
    if (skip_low) {
      dc_halfflag_kernel<<<B, 256, (size_t)n * 8, h->stream>>>(h->dnew, h->halfflag, n, 1, mask);
      DW_LAUNCH_CHECK(h);
    }
    {
      dim3 grid((mm + 7) / 8, lv.nmerge, B);
      dc_vectors_kernel<<<grid, 256, 0, h->stream>>>(g);
      DW_LAUNCH_CHECK(h);
    }
    {
      GemmArgs ga;
      ga.lv = g;
      ga.ntiles_n = (mm + GBN - 1) / GBN;
      ga.halfflag = h->halfflag;
      ga.c_lo = skip_low ? c_lo : 0;
      dim3 grid((mm + GBM - 1) / GBM, lv.nmerge * ga.ntiles_n, B);
      dc_gemm2_kernel<<<grid, 256, G_SMEM, h->stream>>>(ga);
      DW_LAUNCH_CHECK(h);
    }
    {
      dim3 grid(lv.nmerge, B);
      dc_finish_kernel<<<grid, 256, (size_t)mm * 8, h->stream>>>(g);
      DW_LAUNCH_CHECK(h);
    }
    double* t = Zin; Zin = Zout; Zout = t;
  }
  h->Zfinal = Zin;
  return DWHMC_OK;
}

int dw_stedc_output(Handle* h, double* E_out, cplx* U_out, Mask mask, bool ph, int c_lo) {
  dc_evals_kernel<<<h->B, 256, 0, h->stream>>>(h->d, h->perm, E_out, h->halfflag, h->n, (ph && h->ph_mode) ? 1 : 0, mask);
  DW_LAUNCH_CHECK(h);
  dim3 grid(h->n, h->B);
  dc_output_kernel<<<grid, 256, 0, h->stream>>>(h->perm, h->Zfinal, U_out, h->halfflag, c_lo, h->n, mask);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}
