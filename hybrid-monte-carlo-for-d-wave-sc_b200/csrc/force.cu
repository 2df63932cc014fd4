// force.cu -- the O(n^2) and O(N) kernels of the molecular-dynamics step, batched over chains:
//   bond correlators / force       compute_forces!        /root/reference src/Observables.jl:14-62
//   leapfrog kick and drift        hmc_sweep!             src/HMC.jl:92,101,111-113,118
//   H_HMC                          compute_total_energy   src/HMC.jl:12-41
//   momentum refresh               refresh_momentum!      src/HMC.jl:51-61 (Philox instead of the task RNG)
//   Metropolis / restore           hmc_sweep!             src/HMC.jl:124-138
//   light observables              measure_observables    src/Observables.jl:88-222
// All HBM-bound: the correlator kernel streams U once (16 n^2 bytes per chain).
#include "dwhmc.h"
#include "internal.h"

namespace {

// LogExpFunctions 0.3.29, Float64 branches
__device__ __forceinline__ double logistic(double x) {
  if (x < -744.4400719213812) return 0.0;
  if (x > 36.7368005696771) return 1.0;
  const double e = exp(x);
  return e / (1.0 + e);
}
__device__ __forceinline__ double log1pexp(double x) {
  if (x < -745.1332191019412) return 0.0;
  if (x < -36.7368005696771) return exp(x);
  if (x < 18.021826694558577) return log1p(exp(x));
  if (x < 33.23111882352963) return x + exp(-x);
  return x;
}

// deterministic block sum of K doubles (result in every thread); red: K * 32 doubles of shared memory
template <int K>
__device__ __forceinline__ void block_sum_k(double (&v)[K], double* red) {
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

// ---- P_b partial sums over a chunk of eigenvector columns -----------------------------------
// rho1 = sum_n U[i,n] f_n conj(U[j+N,n]), rho2 = sum_n U[j,n] f_n conj(U[i+N,n]), P = -rho1 - rho2
__global__ void __launch_bounds__(256) bond_partial_kernel(const cplx* __restrict__ Uall, const double* __restrict__ Eall,
                                                           const double* __restrict__ par, double* __restrict__ fermi,
                                                           cplx* __restrict__ Ppart, const int* __restrict__ nn, int N,
                                                           int B, int fchunks, Mask mask) {
  const int b = blockIdx.y;
  if (!mask.on(b)) return;
  const int n = 2 * N, chunk = blockIdx.x;
  const int c0 = chunk * DW_FCHUNK, c1 = min(n, c0 + DW_FCHUNK);
  __shared__ double sf[DW_FCHUNK];
  const double beta = par[3 * B + b];
  if (threadIdx.x < c1 - c0) {
    const double f = logistic(-beta * Eall[(size_t)b * n + c0 + threadIdx.x]);
    sf[threadIdx.x] = f;
    fermi[(size_t)b * n + c0 + threadIdx.x] = f;
  }
  __syncthreads();
  const cplx* U = Uall + (size_t)b * n * n;
  cplx* out = Ppart + ((size_t)b * fchunks + chunk) * n;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int jx = nn[i], jy = nn[N + i];
    double xr = 0.0, xi = 0.0, yr = 0.0, yi = 0.0;
    for (int c = c0; c < c1; ++c) {
      const cplx* col = U + (size_t)c * n;
      const double f = sf[c - c0];
      const cplx ui = col[i], uin = col[i + N];
      const cplx ujx = col[jx], ujxn = col[jx + N];
      const cplx ujy = col[jy], ujyn = col[jy + N];
      // a * conj(b) = (ar br + ai bi) + i (ai br - ar bi)
      xr += f * ((ui.x * ujxn.x + ui.y * ujxn.y) + (ujx.x * uin.x + ujx.y * uin.y));
      xi += f * ((ui.y * ujxn.x - ui.x * ujxn.y) + (ujx.y * uin.x - ujx.x * uin.y));
      yr += f * ((ui.x * ujyn.x + ui.y * ujyn.y) + (ujy.x * uin.x + ujy.y * uin.y));
      yi += f * ((ui.y * ujyn.x - ui.x * ujyn.y) + (ujy.y * uin.x - ujy.x * uin.y));
    }
    out[i] = make_double2(-xr, -xi);
    out[N + i] = make_double2(-yr, -yi);
  }
}

// reduce the chunk partials, F = -(beta/2J)(Delta - J P), and (mode 1) the leapfrog updates that
// follow force evaluation number `step` of a trajectory
__global__ void __launch_bounds__(256) force_finish_kernel(const cplx* __restrict__ Ppart, cplx* __restrict__ Pbond,
                                                           cplx* __restrict__ force, cplx* __restrict__ delta,
                                                           cplx* __restrict__ pi, const double* __restrict__ par,
                                                           const int* __restrict__ nt, const double* __restrict__ dtv,
                                                           int N, int B, int fchunks, int mode, int step, Mask mask) {
  const int b = blockIdx.y;
  if (!mask.on(b)) return;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int nb2 = 2 * N;
  if (q >= nb2) return;
  const cplx* pp = Ppart + (size_t)b * fchunks * nb2 + q;
  double pr = 0.0, pim = 0.0;
  for (int c = 0; c < fchunks; ++c) {
    const cplx v = pp[(size_t)c * nb2];
    pr += v.x;
    pim += v.y;
  }
  const double beta = par[3 * B + b], J = par[4 * B + b], mass = par[5 * B + b];
  const size_t o = (size_t)b * nb2 + q;
  Pbond[o] = make_double2(pr, pim);
  if (mode < 0) return;                       // correlators only (observables)
  cplx d = delta[o];
  const double pref = -(beta / (2.0 * J));
  const cplx F = make_double2(pref * (d.x - J * pr), pref * (d.y - J * pim));
  force[o] = F;
  if (mode == 1) {
    const int ntb = nt[b];
    const double dt = dtv[b];
    const bool last = (step == ntb);
    const double kick = (step == 0 || last) ? 0.5 * dt : dt;
    cplx p = pi[o];
    p.x += kick * F.x;
    p.y += kick * F.y;
    pi[o] = p;
    if (!last) {
      const double cf = dt / (2.0 * mass);
      d.x += cf * p.x;
      d.y += cf * p.y;
      delta[o] = d;
    }
  }
}

// H_HMC = E_k + E_b + E_f
__global__ void __launch_bounds__(256) energy_kernel(const double* __restrict__ Eall, const cplx* __restrict__ delta,
                                                     const cplx* __restrict__ pi, const double* __restrict__ par,
                                                     double* __restrict__ out, int N, int B) {
  const int b = blockIdx.x;
  const int n = 2 * N;
  __shared__ double red[3 * 32];
  const double beta = par[3 * B + b], J = par[4 * B + b], mass = par[5 * B + b];
  double v[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double E = Eall[(size_t)b * n + i];
    if (E > 0.0) {
      const double x = beta * E;
      v[0] -= x + 2.0 * log1pexp(-x);
    }
    const cplx d = delta[(size_t)b * n + i];
    const cplx p = pi[(size_t)b * n + i];
    v[1] += d.x * d.x + d.y * d.y;
    v[2] += p.x * p.x + p.y * p.y;
  }
  block_sum_k<3>(v, red);
  if (threadIdx.x == 0) out[b] = 1.0 / (2.0 * mass) * v[2] + beta / (2.0 * J) * v[1] + v[0];
}

// ---- Philox4x32-10 -----------------------------------------------------------------------------
__device__ __forceinline__ void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                       uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {   // (0, 1)
  const unsigned long long x = (((unsigned long long)hi << 32) | lo) >> 11;
  return ((double)x + 0.5) * 1.1102230246251565e-16;
}

// pi: Re, Im ~ N(0, m)  (randn!(ComplexF64) has variance 1/2 per component, then * sqrt(2m))
__global__ void __launch_bounds__(256) momentum_kernel(cplx* __restrict__ pi, const double* __restrict__ par, int N, int B,
                                                       unsigned long long seed, unsigned long long counter) {
  const int b = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= 2 * N) return;
  uint32_t r[4];
  philox((uint32_t)q, (uint32_t)b, (uint32_t)counter, (uint32_t)(counter >> 32) ^ 0x1u, (uint32_t)seed,
         (uint32_t)(seed >> 32), r);
  const double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
  const double rad = sqrt(-2.0 * log(u1)) * sqrt(par[5 * B + b]);
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  pi[(size_t)b * 2 * N + q] = make_double2(rad * c, rad * s);
}

__global__ void uniform_kernel(double* __restrict__ u, int B, unsigned long long seed, unsigned long long counter) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  uint32_t r[4];
  philox((uint32_t)b, 0xFFFFFFFFu, (uint32_t)counter, (uint32_t)(counter >> 32) ^ 0x2u, (uint32_t)seed,
         (uint32_t)(seed >> 32), r);
  const unsigned long long x = (((unsigned long long)r[0] << 32) | r[1]) >> 11;
  u[b] = (double)x * 1.1102230246251565e-16;   // [0, 1)
}

// initialize_state (src/Types.jl:118-134) on the device: round(N n_imp) distinct uniformly random sites get the
// potential W (partial Fisher-Yates, one thread per chain: N is at most a few thousand), Delta0 has
// Re, Im ~ U[-0.05, 0.05) independently on every bond; pi = 0.
__global__ void __launch_bounds__(256) init_state_kernel(double* __restrict__ w, cplx* __restrict__ delta, cplx* __restrict__ pi,
                                                         const double* __restrict__ Wv, const double* __restrict__ nimp,
                                                         int N, int B, unsigned long long seed, unsigned long long counter) {
  const int b = blockIdx.x;
  extern __shared__ int sperm[];
  for (int i = threadIdx.x; i < N; i += blockDim.x) { sperm[i] = i; w[(size_t)b * N + i] = 0.0; }
  for (int q = threadIdx.x; q < 2 * N; q += blockDim.x) {
    uint32_t r[4];
    philox((uint32_t)q, (uint32_t)b, (uint32_t)counter, (uint32_t)(counter >> 32) ^ 0x3u, (uint32_t)seed,
           (uint32_t)(seed >> 32), r);
    const double u1 = (double)((((unsigned long long)r[0] << 32) | r[1]) >> 11) * 1.1102230246251565e-16;
    const double u2 = (double)((((unsigned long long)r[2] << 32) | r[3]) >> 11) * 1.1102230246251565e-16;
    delta[(size_t)b * 2 * N + q] = make_double2((u1 - 0.5) * 0.1, (u2 - 0.5) * 0.1);
    pi[(size_t)b * 2 * N + q] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nimp_sites = (int)rint((double)N * nimp[b]);          // Julia round: ties to even
    for (int k = 0; k < nimp_sites && k < N; ++k) {
      uint32_t r[4];
      philox((uint32_t)k, (uint32_t)b, (uint32_t)counter, (uint32_t)(counter >> 32) ^ 0x4u, (uint32_t)seed,
             (uint32_t)(seed >> 32), r);
      const unsigned long long x = (((unsigned long long)r[0] << 32) | r[1]);
      const int j = k + (int)(x % (unsigned long long)(N - k));
      const int t = sperm[k]; sperm[k] = sperm[j]; sperm[j] = t;
      w[(size_t)b * N + sperm[k]] = Wv[b];
    }
  }
}

__global__ void dH_kernel(const double* __restrict__ Hold, const double* __restrict__ Hnew, double* __restrict__ dH, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) dH[b] = Hnew[b] - Hold[b];
}

// accept iff dH < 0 || u < exp(-dH); NaN rejects (src/HMC.jl:128)
__global__ void metropolis_kernel(const double* __restrict__ dH, const double* __restrict__ u, int* __restrict__ accept,
                                  int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double d = dH[b];
  accept[b] = (d < 0.0 || u[b] < exp(-d)) ? 1 : 0;
}

// accepted: (E, U) <- proposal ; rejected: Delta <- backup.  Either way the pairing block of
// H_base follows Delta (update_H_BdG! inside the loop / after the restore).
__global__ void __launch_bounds__(256) commit_kernel(const int* __restrict__ accept, int* __restrict__ nacc,
                                                     cplx* __restrict__ delta, const cplx* __restrict__ delta_backup,
                                                     cplx* __restrict__ Hs_delta, double* __restrict__ E_cur,
                                                     const double* __restrict__ E_prop, cplx* __restrict__ U_cur,
                                                     const cplx* __restrict__ U_prop, int n) {
  const int b = blockIdx.y;
  const bool acc = accept[b] != 0;
  const size_t nn2 = (size_t)n * n;
  if (acc) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const double4* src = reinterpret_cast<const double4*>(U_prop + (size_t)b * nn2);
    double4* dst = reinterpret_cast<double4*>(U_cur + (size_t)b * nn2);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nn2 / 2; i += stride) dst[i] = src[i];
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const size_t o = (size_t)b * n + i;
      if (acc) E_cur[o] = E_prop[o];
      else delta[o] = delta_backup[o];
      Hs_delta[o] = delta[o];
    }
    if (threadIdx.x == 0 && acc) nacc[b] += 1;
  }
}

// hole_conc partial: sum over columns of the chunk with E > 0 of tanh(beta E / 2) * sum_i (|u_i|^2 - |v_i|^2)
__global__ void __launch_bounds__(256) hole_partial_kernel(const cplx* __restrict__ Uall, const double* __restrict__ Eall,
                                                           const double* __restrict__ par, double* __restrict__ hpart,
                                                           int N, int B, int fchunks) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int n = 2 * N;
  const int c0 = chunk * DW_FCHUNK, c1 = min(n, c0 + DW_FCHUNK);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double beta = par[3 * B + b];
  __shared__ double red[32];
  double tot = 0.0;
  for (int c = c0 + warp; c < c1; c += 8) {
    const double E = Eall[(size_t)b * n + c];
    if (!(E > 0.0)) continue;
    const cplx* col = Uall + (size_t)b * n * n + (size_t)c * n;
    double s = 0.0;
    for (int i = lane; i < N; i += 32) {
      const cplx u = col[i], v = col[i + N];
      s += (u.x * u.x + u.y * u.y) - (v.x * v.x + v.y * v.y);
    }
    tot += s * tanh(0.5 * beta * E);
  }
  double v[1] = {tot};
  block_sum_k<1>(v, red);
  if (threadIdx.x == 0) hpart[(size_t)b * fchunks + chunk] = v[0];
}

// the nine scalars of ObservablesResult (src/Observables.jl:70-80) from Delta, E, P, hole partials
__global__ void __launch_bounds__(256) obs_finish_kernel(const cplx* __restrict__ delta, const cplx* __restrict__ Pbond,
                                                         const double* __restrict__ Eall, const double* __restrict__ hpart,
                                                         const double* __restrict__ par, double* __restrict__ out, int N,
                                                         int B, int fchunks) {
  const int b = blockIdx.x;
  const int n = 2 * N;
  __shared__ double red[11 * 32];
  const double beta = par[3 * B + b], J = par[4 * B + b];
  // 0 amp, 1 local, 2/3 global re/im, 4 |Delta|^2, 5 E_f, 6 hole, 7 diff, 8/9 pair re/im, 10 localpair
  double v[11];
#pragma unroll
  for (int q = 0; q < 11; ++q) v[q] = 0.0;
  const cplx* dl = delta + (size_t)b * n;
  const cplx* P = Pbond + (size_t)b * n;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const cplx dx = dl[i], dy = dl[N + i];
    const cplx px = P[i], py = P[N + i];
    v[0] += 0.5 * (hypot(dx.x, dx.y) + hypot(dy.x, dy.y));
    const double lr = 0.5 * (dx.x - dy.x), li = 0.5 * (dx.y - dy.y);
    v[1] += hypot(lr, li);
    v[2] += lr;
    v[3] += li;
    v[4] += dx.x * dx.x + dx.y * dx.y + dy.x * dy.x + dy.y * dy.y;
    v[7] += 0.5 * (hypot(dx.x - J * px.x, dx.y - J * px.y) + hypot(dy.x - J * py.x, dy.y - J * py.y));
    const double tr = J * 0.5 * (px.x - py.x), ti = J * 0.5 * (px.y - py.y);
    v[8] += tr;
    v[9] += ti;
    v[10] += hypot(tr, ti);
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double E = Eall[(size_t)b * n + i];
    if (E > 0.0) {
      const double x = beta * E;
      v[5] -= x + 2.0 * log1pexp(-x);
    }
  }
  for (int c = threadIdx.x; c < fchunks; c += blockDim.x) v[6] += hpart[(size_t)b * fchunks + c];
  block_sum_k<11>(v, red);
  if (threadIdx.x == 0) {
    const double invN = 1.0 / (double)N;
    double* o = out + (size_t)b * DWHMC_NOBS;
    const double gr = v[2] * invN, gi = v[3] * invN;
    const double g = hypot(gr, gi);
    o[0] = (v[5] + beta / (2.0 * J) * v[4]) * invN;   // total_energy = (E_f + E_b) / N
    o[1] = v[0] * invN;                               // Delta_amp
    o[2] = v[1] * invN;                               // Delta_local
    o[3] = g;                                         // Delta_global
    o[4] = g * g;                                     // S_Delta
    o[5] = v[6] * invN;                               // hole concentration
    o[6] = v[7] * invN;                               // Delta_diff
    o[7] = hypot(v[8] * invN, v[9] * invN);           // Delta_pair
    o[8] = v[10] * invN;                              // Delta_localpair
  }
}

__global__ void copy_delta_kernel(const cplx* __restrict__ src, cplx* __restrict__ dst, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = src[i];
}

}  // namespace

static int bond_correlators(Handle* h, const double* E, const cplx* U, Mask mask) {
  dim3 grid(h->fchunks, h->B);
  bond_partial_kernel<<<grid, 256, 0, h->stream>>>(U, E, h->par, h->fermi, h->Ppart, h->nn, h->N, h->B, h->fchunks, mask);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

int dw_forces(Handle* h, const double* E, const cplx* U, int mode, int step) {
  Mask mask = no_mask();
  if (mode == 1) { mask.nt = h->nt_dev; mask.step = step; }
  DW_TRY(bond_correlators(h, E, U, mask));
  dim3 grid((2 * h->N + 255) / 256, h->B);
  force_finish_kernel<<<grid, 256, 0, h->stream>>>(h->Ppart, h->Pbond, h->force, h->delta, h->pi, h->par, h->nt_dev,
                                                   h->dt_dev, h->N, h->B, h->fchunks, mode, step, mask);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

int dw_total_energy(Handle* h, const double* E, double* out_dev) {
  energy_kernel<<<h->B, 256, 0, h->stream>>>(E, h->delta, h->pi, h->par, out_dev, h->N, h->B);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

int dw_observables(Handle* h, double* out_dev) {
  DW_TRY(bond_correlators(h, h->E_cur, h->U_cur, no_mask()));
  dim3 grid((2 * h->N + 255) / 256, h->B);
  force_finish_kernel<<<grid, 256, 0, h->stream>>>(h->Ppart, h->Pbond, h->force, h->delta, h->pi, h->par, h->nt_dev,
                                                   h->dt_dev, h->N, h->B, h->fchunks, -1, 0, no_mask());
  DW_LAUNCH_CHECK(h);
  dim3 hg(h->fchunks, h->B);
  hole_partial_kernel<<<hg, 256, 0, h->stream>>>(h->U_cur, h->E_cur, h->par, h->hpart, h->N, h->B, h->fchunks);
  DW_LAUNCH_CHECK(h);
  obs_finish_kernel<<<h->B, 256, 0, h->stream>>>(h->delta, h->Pbond, h->E_cur, h->hpart, h->par, out_dev, h->N, h->B,
                                                 h->fchunks);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

int dw_refresh_momentum(Handle* h) {
  dim3 grid((2 * h->N + 255) / 256, h->B);
  momentum_kernel<<<grid, 256, 0, h->stream>>>(h->pi, h->par, h->N, h->B, h->seed, h->rng_counter++);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

int dw_init_state(Handle* h, const double* W_dev, const double* nimp_dev) {
  init_state_kernel<<<h->B, 256, sizeof(int) * h->N, h->stream>>>(h->w, h->delta, h->pi, W_dev, nimp_dev, h->N, h->B, h->seed,
                                                                   h->rng_counter++);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

int dw_uniforms(Handle* h) {
  uniform_kernel<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->unif_dev, h->B, h->seed, h->rng_counter++);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

int dw_begin_trajectory(Handle* h) {
  const size_t count = (size_t)h->n * h->B;
  copy_delta_kernel<<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(h->delta, h->delta_backup, count);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

int dw_dH(Handle* h) {
  dH_kernel<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->Hold_dev, h->Hnew_dev, h->dH_dev, h->B);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

int dw_metropolis(Handle* h, bool) {
  metropolis_kernel<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->dH_dev, h->unif_dev, h->accept_dev, h->B);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}

int dw_commit_dev(Handle* h) {
  dim3 grid(64, h->B);
  commit_kernel<<<grid, 256, 0, h->stream>>>(h->accept_dev, h->nacc_dev, h->delta, h->delta_backup, h->Hs_delta,
                                             h->E_cur, h->E_prop, h->U_cur, h->U_prop, h->n);
  DW_LAUNCH_CHECK(h);
  return DWHMC_OK;
}
