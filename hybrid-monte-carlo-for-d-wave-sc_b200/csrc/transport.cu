// transport.cu -- measure_transport_and_spectra + build_current_operator!
// (/root/reference src/Observables.jl:237-283, :314-526) for every chain of the batch, on the
// eigensystem (E_cur, U_cur) and the Fermi factors left by the last compute_forces! /
// measure_observables call (the reference reads cache.fermi_factors the same way, :321).
//
//   A. J = U^H (Jx U)       Jx = blockdiag(Jp, Jp), Jp[i, i+x] += i t, Jp[i, i+x+y] += i t', Jp[i, i+x-y] += i t'
//                           and the conjugates transposed; Jx U is a 6-point gather, U^H (.) one batched
//                           complex GEMM on the FP64 tensor cores (lower tiles; |J|^2 is symmetric)
//   B. stiffness            diamagnetic sum over E_n > 0 minus Lambda_xx = (1/N) sum_nm ratio(n,m) |J_nm|^2
//   C. conductivities       dc and Re sigma(omega_k): Lorentzian-broadened sums over all ordered pairs --
//                           the O(n^2 n_omega) part, one thread per frequency, pairs staged in shared memory
//   D. DOS, antinodal DOS, A(k, 0)   per-state weights, Lorentzian sums, one 2-D DFT per state
// Reductions use fixed orders (partials + a second pass), so results are reproducible.
#include <cmath>

#include "dwhmc.h"
#include "internal.h"

namespace {

constexpr double PI = 3.14159265358979323846;

__device__ __forceinline__ double lorentz(double x, double eta) { return (1.0 / PI) * (eta / (x * x + eta * eta)); }

template <int K>
__device__ __forceinline__ void block_sum_t(double (&v)[K], double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int q = 0; q < K; ++q)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < K; ++q) red[q * 32 + warp] = v[q];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < K; ++q) {
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += red[q * 32 + i];
    v[q] = t;
  }
}

// T[:, c] = Jx U[:, c].  Row i (< N) of Jp collects +i t from column i+x and -i t from column i-x, etc.
// Entries are accumulated like sparse(I, J, V) does when neighbours coincide.
__global__ void __launch_bounds__(256) jx_apply_kernel(const cplx* __restrict__ Uall, cplx* __restrict__ Tall,
                                                       const double* __restrict__ par, const int* __restrict__ nn,
                                                       const int* __restrict__ nnn, int N, int B) {
  const int b = blockIdx.y, c = blockIdx.x;
  const int n = 2 * N;
  const double t = par[b], tp = par[B + b];
  const cplx* u = Uall + (size_t)b * n * n + (size_t)c * n;
  cplx* o = Tall + (size_t)b * n * n + (size_t)c * n;
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const int i = (r < N) ? r : r - N, off = (r < N) ? 0 : N;
    // (i t)(u[i+x] - u[i-x]) + (i t')(u[i+x+y] - u[i-x-y]) + (i t')(u[i+x-y] - u[i-x+y])
    const cplx a = u[off + nn[0 * N + i]], am = u[off + nn[2 * N + i]];
    const cplx p = u[off + nnn[0 * N + i]], pm = u[off + nnn[2 * N + i]];
    const cplx q = u[off + nnn[3 * N + i]], qm = u[off + nnn[1 * N + i]];
    const double sr = t * (a.x - am.x) + tp * ((p.x - pm.x) + (q.x - qm.x));
    const double si = t * (a.y - am.y) + tp * ((p.y - pm.y) + (q.y - qm.y));
    o[r] = make_double2(-si, sr);            // multiply by i
  }
}

// per eigenstate: w_n = sum_i |u_i|^2, antinodal weight, diamagnetic weight (src/Observables.jl:352-368, :450-497)
__global__ void __launch_bounds__(256) state_weights_kernel(const cplx* __restrict__ Uall, const double* __restrict__ par,
                                                            const int* __restrict__ nn, const int* __restrict__ nnn,
                                                            double* __restrict__ wts, int N, int Lx, int B) {
  const int b = blockIdx.y;
  const int n = 2 * N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 8 + warp;
  if (c >= n) return;
  const double t = par[b], tp = par[B + b];
  const cplx* u = Uall + (size_t)b * n * n + (size_t)c * n;
  double w = 0.0, dia = 0.0, s1r = 0.0, s1i = 0.0, s2r = 0.0, s2i = 0.0;
  for (int i = lane; i < N; i += 32) {
    const cplx ui = u[i], vi = u[i + N];
    w += ui.x * ui.x + ui.y * ui.y;
    const int x = i % Lx + 1, y = i / Lx + 1;          // 1-based, mod1 / cld
    const double sx = (x % 2 == 0) ? 1.0 : -1.0, sy = (y % 2 == 0) ? 1.0 : -1.0;
    s1r += sx * ui.x; s1i += sx * ui.y;
    s2r += sy * ui.x; s2i += sy * ui.y;
    const int nb[3] = {nn[i], nnn[i], nnn[3 * N + i]};
    const double hop[3] = {t, tp, tp};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const cplx uj = u[nb[k]], vj = u[nb[k] + N];
      // real( v_i conj(v_j) - conj(u_i) u_j )
      const double re = (vi.x * vj.x + vi.y * vj.y) - (ui.x * uj.x + ui.y * uj.y);
      dia += hop[k] * 2.0 * re;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    w += __shfl_xor_sync(0xffffffffu, w, o);
    dia += __shfl_xor_sync(0xffffffffu, dia, o);
    s1r += __shfl_xor_sync(0xffffffffu, s1r, o);
    s1i += __shfl_xor_sync(0xffffffffu, s1i, o);
    s2r += __shfl_xor_sync(0xffffffffu, s2r, o);
    s2i += __shfl_xor_sync(0xffffffffu, s2i, o);
  }
  if (lane == 0) {
    double* o3 = wts + ((size_t)b * n + c) * 3;
    o3[0] = w;
    o3[1] = 0.5 * ((s1r * s1r + s1i * s1i) + (s2r * s2r + s2i * s2i)) / (double)N;
    o3[2] = dia;
  }
}

// |J|^2 as a full symmetric real matrix from the lower tiles of J, plus the pair sums that do not depend
// on omega: Lambda_xx and the dc conductivity (src/Observables.jl:373-412).  One CTA per column m.
__global__ void __launch_bounds__(256) pair_static_kernel(const cplx* __restrict__ Jall, const double* __restrict__ Eall,
                                                          const double* __restrict__ fall, const double* __restrict__ par,
                                                          double* __restrict__ J2all, double* __restrict__ part, int n,
                                                          int B, double eta) {
  const int b = blockIdx.y, m = blockIdx.x;
  const cplx* J = Jall + (size_t)b * n * n;
  const double* E = Eall + (size_t)b * n;
  const double* f = fall + (size_t)b * n;
  double* J2 = J2all + (size_t)b * n * n;
  const double beta = par[3 * B + b];
  const double Em = E[m], fm = f[m];
  __shared__ double red[2 * 32];
  double v[2] = {0.0, 0.0};
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    // entry (r, m) of the reference's J_mn[n = r, m]; the lower triangle holds r >= m
    const cplx z = (r >= m) ? J[(size_t)m * n + r] : J[(size_t)r * n + m];
    const double j2 = z.x * z.x + z.y * z.y;
    J2[(size_t)m * n + r] = j2;
    const double dE = Em - E[r], fr = f[r];
    const double mdf = beta * fr * (1.0 - fr);
    const double ratio = (fabs(dE) < 1e-8) ? mdf : (fr - fm) / dE;
    v[0] += ratio * j2;
    v[1] += mdf * j2 * lorentz(dE, eta);
  }
  block_sum_t<2>(v, red);
  if (threadIdx.x == 0) {
    part[((size_t)b * n + m) * 2 + 0] = v[0];
    part[((size_t)b * n + m) * 2 + 1] = v[1];
  }
}

// Re sigma(omega_k) partial sums: CTA = (256 frequencies, chunk of PCH columns m, chain).  The pairs of the
// chunk with |f_n - f_m| >= 1e-12 (src/Observables.jl:415) are compacted (deterministic block scan, original
// order kept) into shared memory as (E_m - E_n, (f_n - f_m) |J_nm|^2) and broadcast to the threads.  Four terms
// share one division:  sum_i w_i / d_i = (N12 P34 + N34 P12) / (P12 P34),  d_i = (omega - Delta_i)^2 + eta^2
// (d_i in [eta^2, ~100], so the products stay far inside the FP64 range).
constexpr int PCH = 4;          // columns m per CTA -> 4 n pairs
__global__ void __launch_bounds__(256) sigma_partial_kernel(const double* __restrict__ J2all, const double* __restrict__ Eall,
                                                            const double* __restrict__ fall, const double* __restrict__ omega,
                                                            double* __restrict__ spart, int n, int nw, int nchunk,
                                                            double eta) {
  const int b = blockIdx.z, chunk = blockIdx.y;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* pr = reinterpret_cast<double2*>(smem_raw);       // [PCH * n + 4] (delta, weight), compacted
  __shared__ int wsum[8];
  __shared__ int s_tot;
  const double* E = Eall + (size_t)b * n;
  const double* f = fall + (size_t)b * n;
  const double* J2 = J2all + (size_t)b * n * n;
  const int m0 = chunk * PCH;
  const int tot_in = PCH * n;
  const int per = (tot_in + 255) / 256;                     // contiguous slice per thread keeps the order
  const int lo = threadIdx.x * per, hi = min(tot_in, lo + per);
  int cnt = 0;
  for (int idx = lo; idx < hi; ++idx) {
    const int mm = m0 + idx / n, r = idx % n;
    if (mm < n && fabs(f[r] - f[mm]) >= 1e-12) ++cnt;
  }
  // exclusive scan of cnt over the block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  int base = 0;
  for (int w8 = 0; w8 < warp; ++w8) base += wsum[w8];
  if (threadIdx.x == 255) s_tot = base + inc;
  int at = base + inc - cnt;
  for (int idx = lo; idx < hi; ++idx) {
    const int mm = m0 + idx / n, r = idx % n;
    if (mm < n) {
      const double df = f[r] - f[mm];
      if (fabs(df) >= 1e-12) pr[at++] = make_double2(E[mm] - E[r], df * J2[(size_t)mm * n + r]);
    }
  }
  __syncthreads();
  const int tot = s_tot;
  const int tot4 = (tot + 3) & ~3;
  if (threadIdx.x < tot4 - tot) pr[tot + threadIdx.x] = make_double2(0.0, 0.0);     // neutral padding
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const double w = (k < nw) ? omega[k] : 1.0;
  const double eta2 = eta * eta;
  double acc0 = 0.0, acc1 = 0.0;
  for (int idx = 0; idx < tot4; idx += 8) {
    {
      const double2 v0 = pr[idx], v1 = pr[idx + 1], v2 = pr[idx + 2], v3 = pr[idx + 3];
      const double x0 = w - v0.x, x1 = w - v1.x, x2 = w - v2.x, x3 = w - v3.x;
      const double d0 = fma(x0, x0, eta2), d1 = fma(x1, x1, eta2), d2 = fma(x2, x2, eta2), d3 = fma(x3, x3, eta2);
      const double p01 = d0 * d1, p23 = d2 * d3;
      const double n01 = fma(v0.y, d1, v1.y * d0), n23 = fma(v2.y, d3, v3.y * d2);
      acc0 += fma(n01, p23, n23 * p01) / (p01 * p23);
    }
    if (idx + 4 < tot4) {
      const double2 v0 = pr[idx + 4], v1 = pr[idx + 5], v2 = pr[idx + 6], v3 = pr[idx + 7];
      const double x0 = w - v0.x, x1 = w - v1.x, x2 = w - v2.x, x3 = w - v3.x;
      const double d0 = fma(x0, x0, eta2), d1 = fma(x1, x1, eta2), d2 = fma(x2, x2, eta2), d3 = fma(x3, x3, eta2);
      const double p01 = d0 * d1, p23 = d2 * d3;
      const double n01 = fma(v0.y, d1, v1.y * d0), n23 = fma(v2.y, d3, v3.y * d2);
      acc1 += fma(n01, p23, n23 * p01) / (p01 * p23);
    }
  }
  // (fn_fm / w) J2 (1/pi) eta / (x^2 + eta^2)
  if (k < nw) spart[((size_t)b * nchunk + chunk) * nw + k] = (acc0 + acc1) * (eta / PI) / w;
}

__global__ void sigma_reduce_kernel(const double* __restrict__ spart, double* __restrict__ sigma, int nw, int nchunk,
                                    int N) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nw) return;
  double s = 0.0;
  for (int c = 0; c < nchunk; ++c) s += spart[((size_t)b * nchunk + c) * nw + k];
  sigma[(size_t)b * nw + k] = s * (PI / (double)N);
}

// dos(w) = (1/N) sum_n w_n L(w - E_n), dos_AN(w) = sum_n wAN_n L(w - E_n)
__global__ void __launch_bounds__(256) dos_kernel(const double* __restrict__ Eall, const double* __restrict__ wts,
                                                  const double* __restrict__ grid, double* __restrict__ dos,
                                                  double* __restrict__ dosAN, int n, int nd, int N, double eta) {
  const int b = blockIdx.y;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sE = reinterpret_cast<double*>(smem_raw);
  double* sw = sE + n;
  double* sa = sw + n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    sE[i] = Eall[(size_t)b * n + i];
    sw[i] = wts[((size_t)b * n + i) * 3 + 0];
    sa[i] = wts[((size_t)b * n + i) * 3 + 1];
  }
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nd) return;
  const double w = grid[k];
  double d0 = 0.0, d1 = 0.0;
  for (int i = 0; i < n; ++i) {
    const double l = lorentz(w - sE[i], eta);
    d0 += sw[i] * l;
    d1 += sa[i] * l;
  }
  dos[(size_t)b * nd + k] = d0 / (double)N;
  dosAN[(size_t)b * nd + k] = d1;
}

// scalars: stiffness = (1/N) sum_{E_n>0} dia_n tanh(beta E_n / 2) - Lambda_xx / N ; dc = (pi / N) sum
__global__ void __launch_bounds__(256) transport_scalar_kernel(const double* __restrict__ Eall, const double* __restrict__ wts,
                                                               const double* __restrict__ part, const double* __restrict__ par,
                                                               double* __restrict__ out, int n, int N, int B) {
  const int b = blockIdx.x;
  __shared__ double red[3 * 32];
  const double beta = par[3 * B + b];
  double v[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double E = Eall[(size_t)b * n + i];
    if (E > 0.0) v[0] += wts[((size_t)b * n + i) * 3 + 2] * tanh(0.5 * beta * E) / (double)N;
    v[1] += part[((size_t)b * n + i) * 2 + 0];
    v[2] += part[((size_t)b * n + i) * 2 + 1];
  }
  block_sum_t<3>(v, red);
  if (threadIdx.x == 0) {
    out[2 * b + 0] = v[0] - v[1] / (double)N;
    out[2 * b + 1] = v[2] * (PI / (double)N);
  }
}

// A(k, 0): sum over states with L(-E_n) > 1e-6 of |FFT2(u_n)|^2 L(-E_n) / N (src/Observables.jl:499-519).
// CTA = (chunk of 8 states, chain); u_n[x, y] -> row DFT over x, then column DFT over y, by direct sums
// with a twiddle table (the lattices are small: Lx, Ly <= 64).
constexpr int AKC = 8;
__global__ void __launch_bounds__(256) ak_partial_kernel(const cplx* __restrict__ Uall, const double* __restrict__ Eall,
                                                         double* __restrict__ akpart, int N, int Lx, int Ly, int nchunk,
                                                         double eta) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int n = 2 * N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* su = reinterpret_cast<cplx*>(smem_raw);     // [N] u(x, y), index y * Lx + x
  cplx* st = su + N;                                // [N] after the x transform: t(kx, y)
  cplx* twx = st + N;                               // [Lx] exp(-2 pi i k / Lx)
  cplx* twy = twx + Lx;                             // [Ly]
  double* acc = reinterpret_cast<double*>(twy + Ly);   // [N]
  for (int k = threadIdx.x; k < Lx; k += blockDim.x) { double s, c; sincospi(-2.0 * k / (double)Lx, &s, &c); twx[k] = make_double2(c, s); }
  for (int k = threadIdx.x; k < Ly; k += blockDim.x) { double s, c; sincospi(-2.0 * k / (double)Ly, &s, &c); twy[k] = make_double2(c, s); }
  for (int i = threadIdx.x; i < N; i += blockDim.x) acc[i] = 0.0;
  __syncthreads();
  for (int cc = 0; cc < AKC; ++cc) {
    const int c = chunk * AKC + cc;
    if (c >= n) break;
    const double w0 = lorentz(0.0 - Eall[(size_t)b * n + c], eta);
    if (!(w0 > 1e-6)) continue;                      // uniform across the CTA
    const cplx* u = Uall + (size_t)b * n * n + (size_t)c * n;
    for (int i = threadIdx.x; i < N; i += blockDim.x) su[i] = u[i];
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {      // i = y * Lx + kx
      const int kx = i % Lx, y = i / Lx;
      double re = 0.0, im = 0.0;
      for (int x = 0; x < Lx; ++x) {
        const cplx a = su[y * Lx + x], tw = twx[(kx * x) % Lx];
        re += a.x * tw.x - a.y * tw.y;
        im += a.x * tw.y + a.y * tw.x;
      }
      st[i] = make_double2(re, im);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {      // i = ky * Lx + kx
      const int kx = i % Lx, ky = i / Lx;
      double re = 0.0, im = 0.0;
      for (int y = 0; y < Ly; ++y) {
        const cplx a = st[y * Lx + kx], tw = twy[(ky * y) % Ly];
        re += a.x * tw.x - a.y * tw.y;
        im += a.x * tw.y + a.y * tw.x;
      }
      acc[i] += (re * re + im * im) * w0;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) akpart[((size_t)b * nchunk + chunk) * N + i] = acc[i];
}

__global__ void ak_reduce_kernel(const double* __restrict__ akpart, double* __restrict__ ak, int N, int nchunk) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double s = 0.0;
  for (int c = 0; c < nchunk; ++c) s += akpart[((size_t)b * nchunk + c) * N + i];
  ak[(size_t)b * N + i] = s / (double)N;      // index ky * Lx + kx = column-major A_k[kx, ky] of the reference
}

}  // namespace

// out_dev layout: scal [2B] | sigma [nw B] | dos [nd B] | dosAN [nd B] | ak [N B]; omega_dev [nw], dosgrid_dev [nd]
int dw_transport(Handle* h, double eta, const double* omega_dev, int nw, const double* dosgrid_dev, int nd,
                 double* scal, double* sigma, double* dos, double* dosAN, double* ak, double* work, size_t work_count) {
  const int n = h->n, N = h->N, B = h->B;
  cplx* T = h->A;              // Jx U, then free
  cplx* J = h->U_prop;         // U^H Jx U (lower tiles); U_prop is free between trajectories (V must keep its zeros)
  double* J2 = h->Z0;          // |J|^2, full
  const int nchunk = (n + PCH - 1) / PCH;
  const int akchunk = (n + AKC - 1) / AKC;
  // workspace: wts [3 n B] | part [2 n B] | spart [nchunk nw B] | akpart [akchunk N B]
  const size_t need = (size_t)5 * n * B + (size_t)nchunk * nw * B + (size_t)akchunk * N * B;
  if (work_count < need) { h->err = "dw_transport: workspace too small"; return DWHMC_E_BADARG; }
  double* wts = work;
  double* part = wts + (size_t)3 * n * B;
  double* spart = part + (size_t)2 * n * B;
  double* akpart = spart + (size_t)nchunk * nw * B;
  {
    dim3 grid(n, B);
    jx_apply_kernel<<<grid, 256, 0, h->stream>>>(h->U_cur, T, h->par, h->nn, h->nnn, N, B);
    DW_LAUNCH_CHECK(h);
  }
  {
    ZgemmArgs a;
    a.nseg = 1; a.A[1] = nullptr; a.Bm[1] = nullptr; a.lower = 1; a.batch = B; a.mask = no_mask();
    a.M = n; a.N = n; a.K = n;
    a.A[0] = h->U_cur; a.lda = n; a.sA = (long long)n * n; a.opA = 1;
    a.Bm[0] = T; a.ldb = n; a.sB = (long long)n * n; a.opB = 0;
    a.C = J; a.ldc = n; a.sC = (long long)n * n;
    a.alpha = 1.0; a.beta = 0.0;
    DW_TRY(dw_zgemm(h, a));
  }
  {
    dim3 grid((n + 7) / 8, B);
    state_weights_kernel<<<grid, 256, 0, h->stream>>>(h->U_cur, h->par, h->nn, h->nnn, wts, N, h->Lx, B);
    DW_LAUNCH_CHECK(h);
  }
  {
    dim3 grid(n, B);
    pair_static_kernel<<<grid, 256, 0, h->stream>>>(J, h->E_cur, h->fermi, h->par, J2, part, n, B, eta);
    DW_LAUNCH_CHECK(h);
  }
  transport_scalar_kernel<<<B, 256, 0, h->stream>>>(h->E_cur, wts, part, h->par, scal, n, N, B);
  DW_LAUNCH_CHECK(h);
  if (nw > 0) {
    const size_t smem = sizeof(double2) * ((size_t)PCH * n + 4);
    static bool attr_set[64] = {false};
    if (!attr_set[h->device & 63]) {
      DW_CUDA(h, cudaFuncSetAttribute(sigma_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set[h->device & 63] = true;
    }
    if (smem > 200 * 1024) { h->err = "dw_transport: lattice too large for the conductivity kernel"; return DWHMC_E_BADARG; }
    dim3 grid((nw + 255) / 256, nchunk, B);
    sigma_partial_kernel<<<grid, 256, smem, h->stream>>>(J2, h->E_cur, h->fermi, omega_dev, spart, n, nw, nchunk, eta);
    DW_LAUNCH_CHECK(h);
    dim3 g2((nw + 255) / 256, B);
    sigma_reduce_kernel<<<g2, 256, 0, h->stream>>>(spart, sigma, nw, nchunk, N);
    DW_LAUNCH_CHECK(h);
  }
  if (nd > 0) {
    dim3 grid((nd + 255) / 256, B);
    dos_kernel<<<grid, 256, sizeof(double) * 3 * n, h->stream>>>(h->E_cur, wts, dosgrid_dev, dos, dosAN, n, nd, N, eta);
    DW_LAUNCH_CHECK(h);
  }
  {
    const size_t smem = sizeof(cplx) * (2 * (size_t)N + h->Lx + h->Ly) + sizeof(double) * N;
    static bool attr_set[64] = {false};
    if (!attr_set[h->device & 63]) {
      DW_CUDA(h, cudaFuncSetAttribute(ak_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set[h->device & 63] = true;
    }
    if (smem > 200 * 1024) { h->err = "dw_transport: lattice too large for the A(k, 0) kernel"; return DWHMC_E_BADARG; }
    dim3 grid(akchunk, B);
    ak_partial_kernel<<<grid, 256, smem, h->stream>>>(h->U_cur, h->E_cur, akpart, N, h->Lx, h->Ly, akchunk, eta);
    DW_LAUNCH_CHECK(h);
    dim3 g2((N + 255) / 256, B);
    ak_reduce_kernel<<<g2, 256, 0, h->stream>>>(akpart, ak, N, akchunk);
    DW_LAUNCH_CHECK(h);
  }
  return DWHMC_OK;
}

size_t dw_transport_work_count(const Handle* h, int nw) {
  const int n = h->n, N = h->N, B = h->B;
  const int nchunk = (n + PCH - 1) / PCH;
  const int akchunk = (n + AKC - 1) / AKC;
  return (size_t)5 * n * B + (size_t)nchunk * nw * B + (size_t)akchunk * N * B;
}
