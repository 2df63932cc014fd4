// api.cu -- the C ABI of libdwhmc.so (include/dwhmc.h).  Each entry point names the reference
// operator it stands in for; the batched trajectory is hmc_sweep! (/root/reference src/HMC.jl:71-144)
// for B chains at once.  No CPU fallback: every call needs the CUDA device of its handle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "dwhmc.h"
#include "internal.h"
#include "stedc_core.h"

struct dwhmc_handle_s : Handle {};

static thread_local std::string g_create_err;

#define H_ENTER(hh)                                                              \
  if (!(hh)) return DWHMC_E_BADARG;                                              \
  Handle* h = static_cast<Handle*>(hh);                                          \
  do {                                                                           \
    cudaError_t e__ = cudaSetDevice(h->device);                                  \
    if (e__ != cudaSuccess) { h->err = cudaGetErrorString(e__); return DWHMC_E_CUDA; } \
  } while (0)

#define BADARG(msg) do { h->err = (msg); return DWHMC_E_BADARG; } while (0)

template <class T>
static int dalloc(Handle* h, T*& p, size_t count, bool zero = true) {
  void* q = nullptr;
  if (count == 0) count = 1;
  DW_CUDA(h, cudaMalloc(&q, sizeof(T) * count));
  h->allocs.push_back(q);
  if (zero) DW_CUDA(h, cudaMemsetAsync(q, 0, sizeof(T) * count, h->stream));
  p = static_cast<T*>(q);
  return DWHMC_OK;
}

static int h2d(Handle* h, void* dst, const void* src, size_t bytes) {
  DW_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
  DW_CUDA(h, cudaStreamSynchronize(h->stream));   // host pointers are borrowed for the call only
  return DWHMC_OK;
}
static int d2h(Handle* h, void* dst, const void* src, size_t bytes) {
  DW_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  DW_CUDA(h, cudaStreamSynchronize(h->stream));
  return DWHMC_OK;
}

// eigensolver status words -> error code
static int check_status(Handle* h) {
  int st[4] = {0, 0, 0, 0};
  DW_TRY(d2h(h, st, h->status, sizeof(st)));
  if (st[0] || st[1]) {
    h->err = "eigensolver did not converge (leaf QL failures " + std::to_string(st[0]) + ", secular roots " +
             std::to_string(st[1]) + ")";
    cudaMemsetAsync(h->status, 0, sizeof(int) * 4, h->stream);
    return DWHMC_E_NOCONV;
  }
  if (st[2]) {
    h->err = "band route: a wait on another CTA's progress timed out (results invalid)";
    cudaMemsetAsync(h->status, 0, sizeof(int) * 4, h->stream);
    return DWHMC_E_CUDA;
  }
  return DWHMC_OK;
}

struct StageTimer {
  Handle* h; int slot;
  StageTimer(Handle* hh, int s) : h(hh), slot(s) { if (h->profiling) cudaEventRecord(h->ev0, h->stream); }
  ~StageTimer() {
    if (!h->profiling) return;
    cudaEventRecord(h->ev1, h->stream);
    cudaEventSynchronize(h->ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->timers[slot] += ms;
  }
};

extern "C" {

const char* dwhmc_version(void) { return "dwhmc-b200 0.3 (sm_100a; FP64 DMMA m8n8k4; band route: position-owning bulge chase + D&C + fused block reflectors; dense route: hetrd + D&C + blocked back-transform)"; }

const char* dwhmc_last_error(dwhmc_handle hh) {
  if (!hh) return g_create_err.c_str();
  return static_cast<Handle*>(hh)->err.c_str();
}

int dwhmc_create(dwhmc_handle* out, int device, int B, int Lx, int Ly, const int64_t* nn_table,
                 const int64_t* nnn_table) {
  if (!out) { g_create_err = "dwhmc_create: out is NULL"; return DWHMC_E_BADARG; }
  *out = nullptr;
  if (B < 1 || B > 32767 || Lx < 3 || Ly < 3 || !nn_table || !nnn_table) {
    g_create_err = "dwhmc_create: need 1 <= B <= 32767, Lx, Ly >= 3 (distinct neighbours) and both neighbour tables";
    return DWHMC_E_BADARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_create_err = "dwhmc_create: no CUDA device (libdwhmc has no CPU fallback)";
    return DWHMC_E_NODEVICE;
  }
  if (device < 0 || device >= ndev) { g_create_err = "dwhmc_create: device index out of range"; return DWHMC_E_BADARG; }
  const int N = Lx * Ly, n = 2 * N;
  std::vector<int> nn(4 * (size_t)N), nnn(4 * (size_t)N);
  for (size_t i = 0; i < 4 * (size_t)N; ++i) {
    if (nn_table[i] < 1 || nn_table[i] > N || nnn_table[i] < 1 || nnn_table[i] > N) {
      g_create_err = "dwhmc_create: neighbour table entry out of 1..N";
      return DWHMC_E_BADARG;
    }
    nn[i] = (int)(nn_table[i] - 1);
    nnn[i] = (int)(nnn_table[i] - 1);
  }
  dwhmc_handle_s* hs = new dwhmc_handle_s();
  Handle* h = hs;
  h->device = device; h->B = B; h->Lx = Lx; h->Ly = Ly; h->N = N; h->n = n;
  if (const char* eg = getenv("DWHMC_NGROUP")) {
    const int v = atoi(eg);
    if (v >= 1 && v <= DW_NGROUP) h->ngroups = v;
  }
  if (const char* ep = getenv("DWHMC_PH")) h->ph_mode = atoi(ep) ? 1 : 0;
  cudaDeviceGetAttribute(&h->nsm, cudaDevAttrMultiProcessorCount, device);
  auto fail = [&](int rc) { g_create_err = h->err; dwhmc_destroy(hs); return rc; };
  if (cudaSetDevice(device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return fail(DWHMC_E_CUDA); }
  {
    // the main stream runs at the high priority: work forked to the low-priority side streams (T factors beside the
    // D&C stage) only fills the SMs the main stream leaves free
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { h->err = "stream create failed"; return fail(DWHMC_E_CUDA); }
  }
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);
  cudaEventCreate(&h->ev_begin);
  cudaEventCreate(&h->ev_end);
  cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  {
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);     // numerically lower = higher priority
    for (int g = 0; g < DW_NGROUP; ++g) {
      cudaStreamCreateWithPriority(&h->gstream[g], cudaStreamNonBlocking, prio_lo);
      cudaStreamCreateWithPriority(&h->gstream_hi[g], cudaStreamNonBlocking, prio_hi);
      cudaEventCreateWithFlags(&h->ev_join[g], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&h->ev_bulk[g], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&h->ev_col[g], cudaEventDisableTiming);
    }
  }
  const size_t nB = (size_t)n * B, nnB = (size_t)n * n * B;
  h->nblk = (n - 1 + DW_NB - 1) / DW_NB;
  h->nbt = (n - 1 + DW_NBT - 1) / DW_NBT;
  h->fchunks = (n + DW_FCHUNK - 1) / DW_FCHUNK;
  h->h_par.assign(6 * (size_t)B, 0.0);
  int rc = DWHMC_OK;
#define AL(p, cnt) if ((rc = dalloc(h, p, (cnt))) != DWHMC_OK) return fail(rc)
  AL(h->nn, 4 * (size_t)N); AL(h->nnn, 4 * (size_t)N);
  AL(h->par, 6 * (size_t)B);
  AL(h->w, (size_t)N * B); AL(h->delta, nB); AL(h->pi, nB); AL(h->force, nB); AL(h->delta_backup, nB);
  AL(h->Hs_w, (size_t)N * B); AL(h->Hs_par, 3 * (size_t)B); AL(h->Hs_delta, nB);
  AL(h->E_cur, nB); AL(h->E_prop, nB); AL(h->U_cur, nnB); AL(h->U_prop, nnB); AL(h->fermi, nB);
  AL(h->nt_dev, (size_t)B); AL(h->dt_dev, (size_t)B);
  AL(h->Hold_dev, (size_t)B); AL(h->Hnew_dev, (size_t)B); AL(h->dH_dev, (size_t)B);
  AL(h->accept_dev, (size_t)B); AL(h->nacc_dev, (size_t)B); AL(h->unif_dev, (size_t)B);
  AL(h->obs_dev, (size_t)DWHMC_NOBS * B);
  AL(h->Ppart, (size_t)h->fchunks * nB); AL(h->hpart, (size_t)h->fchunks * B); AL(h->Pbond, nB);
  AL(h->A, nnB); AL(h->V, nnB);
  AL(h->ypart, (size_t)((n + 63) / 64) * nB); AL(h->P1, (size_t)DW_CC * DW_NB * B); AL(h->P2, (size_t)DW_CC * DW_NB * B);
  AL(h->Tf, (size_t)h->nbt * DW_NBT * DW_NBT * B); AL(h->Gb, (size_t)h->nbt * DW_NBT * DW_NBT * B * DW_GSPLIT); AL(h->tau, nB); AL(h->d, nB); AL(h->e, nB);
  AL(h->Wbt, (size_t)DW_NBT * nB); AL(h->Wbt2, (size_t)DW_NBT * nB);
  AL(h->Z0, nnB); AL(h->Z1, nnB); AL(h->S, nnB);
  AL(h->perm, nB); AL(h->ord, nB);
  AL(h->zvec, nB); AL(h->dl, nB); AL(h->wv, nB); AL(h->dnew, nB); AL(h->zhat, nB); AL(h->stau, nB);
  AL(h->ndl, nB); AL(h->dfl, nB); AL(h->sorg, nB); AL(h->ndc, nB); AL(h->kcls, nB + 2);
  {
    dwcore::DeflRot* r = nullptr;
    if ((rc = dalloc(h, r, nB)) != DWHMC_OK) return fail(rc);
    h->rots = r;
  }
  AL(h->kcnt, nB); AL(h->nrot, nB); AL(h->rho, nB); AL(h->status, 4);
  AL(h->halfflag, (size_t)B);
  if ((rc = dw_band_setup(h, nn, nnn)) != DWHMC_OK) return fail(rc);
  if (h->band_b > 0) {
    const size_t nblk = h->band_blk_s0.size();
    AL(h->band_pos, (size_t)n); AL(h->band_prog, nB + (size_t)B); AL(h->band_tau, nB * h->band_KT);
    AL(h->band_rowbox, (size_t)B * h->band_KT * 2 * (h->band_b + 2));
    AL(h->band_bbox, (size_t)B * h->band_KT * 2);
    AL(h->band_VT, (nblk * DW_APPLY_BLOCK_DOUBLES * B + 1) / 2);
    AL(h->band_blk_s0_dev, nblk); AL(h->band_blk_k_dev, nblk);
    {
      std::vector<int> items4;
      dw_band_apply_items(h, items4);
      AL(h->band_items_dev, items4.size());
      if ((rc = h2d(h, h->band_items_dev, items4.data(), sizeof(int) * items4.size())) != DWHMC_OK) return fail(rc);
      AL(h->band_sync, 1 + (size_t)B * (h->band_wave_start.size() - 1));
    }
    if ((rc = h2d(h, h->band_pos, h->band_pos_host.data(), sizeof(int) * n)) != DWHMC_OK) return fail(rc);
    if ((rc = h2d(h, h->band_blk_s0_dev, h->band_blk_s0.data(), sizeof(int) * nblk)) != DWHMC_OK) return fail(rc);
    if ((rc = h2d(h, h->band_blk_k_dev, h->band_blk_k.data(), sizeof(int) * nblk)) != DWHMC_OK) return fail(rc);
  }
  // D&C tree (shared by all chains)
  h->tree = build_dc_tree(n, DW_LEAF);
  h->nleaves = (int)h->tree.leaves.size();
  {
    std::vector<int> lo(h->nleaves), ls(h->nleaves);
    for (int i = 0; i < h->nleaves; ++i) { lo[i] = h->tree.leaves[i].off; ls[i] = h->tree.leaves[i].size; }
    AL(h->leaf_off, (size_t)h->nleaves); AL(h->leaf_size, (size_t)h->nleaves);
    if ((rc = h2d(h, h->leaf_off, lo.data(), sizeof(int) * lo.size())) != DWHMC_OK) return fail(rc);
    if ((rc = h2d(h, h->leaf_size, ls.data(), sizeof(int) * ls.size())) != DWHMC_OK) return fail(rc);
  }
  for (auto& lvl : h->tree.levels) {
    DcLevelDev d;
    d.nmerge = (int)lvl.size(); d.max_m = 0; d.off = d.n1 = d.n2 = nullptr;
    std::vector<int> o(lvl.size()), a(lvl.size()), c(lvl.size());
    for (size_t i = 0; i < lvl.size(); ++i) {
      o[i] = lvl[i].off; a[i] = lvl[i].n1; c[i] = lvl[i].n2;
      if (a[i] + c[i] > d.max_m) d.max_m = a[i] + c[i];
    }
    AL(d.off, lvl.size()); AL(d.n1, lvl.size()); AL(d.n2, lvl.size());
    if ((rc = h2d(h, d.off, o.data(), sizeof(int) * o.size())) != DWHMC_OK) return fail(rc);
    if ((rc = h2d(h, d.n1, a.data(), sizeof(int) * a.size())) != DWHMC_OK) return fail(rc);
    if ((rc = h2d(h, d.n2, c.data(), sizeof(int) * c.size())) != DWHMC_OK) return fail(rc);
    h->levels.push_back(d);
  }
#undef AL
  if ((rc = h2d(h, h->nn, nn.data(), sizeof(int) * nn.size())) != DWHMC_OK) return fail(rc);
  if ((rc = h2d(h, h->nnn, nnn.data(), sizeof(int) * nnn.size())) != DWHMC_OK) return fail(rc);
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) { h->err = "initialisation failed"; return fail(DWHMC_E_CUDA); }
  *out = hs;
  return DWHMC_OK;
}

int dwhmc_destroy(dwhmc_handle hh) {
  if (!hh) return DWHMC_OK;
  Handle* h = static_cast<Handle*>(hh);
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs) cudaFree(p);
  if (h->tr_work) cudaFree(h->tr_work);
  if (h->tr_out) cudaFree(h->tr_out);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_begin) cudaEventDestroy(h->ev_begin);
  if (h->ev_end) cudaEventDestroy(h->ev_end);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (int g = 0; g < DW_NGROUP; ++g) {
    if (h->ev_join[g]) cudaEventDestroy(h->ev_join[g]);
    if (h->ev_bulk[g]) cudaEventDestroy(h->ev_bulk[g]);
    if (h->ev_col[g]) cudaEventDestroy(h->ev_col[g]);
    if (h->gstream[g]) cudaStreamDestroy(h->gstream[g]);
    if (h->gstream_hi[g]) cudaStreamDestroy(h->gstream_hi[g]);
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete static_cast<dwhmc_handle_s*>(hh);
  return DWHMC_OK;
}

int dwhmc_eigensolver_route(dwhmc_handle hh, int* half_bandwidth) {
  H_ENTER(hh);
  if (!half_bandwidth) BADARG("dwhmc_eigensolver_route: NULL");
  *half_bandwidth = h->band_b;
  return DWHMC_OK;
}

int dwhmc_dims(dwhmc_handle hh, int* B, int* N, int* n) {
  H_ENTER(hh);
  if (B) *B = h->B;
  if (N) *N = h->N;
  if (n) *n = h->n;
  return DWHMC_OK;
}

int dwhmc_set_params(dwhmc_handle hh, const double* t, const double* tp, const double* mu, const double* beta,
                     const double* J, const double* mass) {
  H_ENTER(hh);
  if (!t || !tp || !mu || !beta || !J || !mass) BADARG("dwhmc_set_params: NULL array");
  const int B = h->B;
  const double* src[6] = {t, tp, mu, beta, J, mass};
  for (int b = 0; b < B; ++b)                       // validate before touching the host mirror
    if (!(J[b] != 0.0) || !(mass[b] != 0.0)) BADARG("dwhmc_set_params: J and mass must be nonzero");
  for (int k = 0; k < 6; ++k)
    for (int b = 0; b < B; ++b) h->h_par[(size_t)k * B + b] = src[k][b];
  h->params_set = true;
  return h2d(h, h->par, h->h_par.data(), sizeof(double) * 6 * B);
}

int dwhmc_set_disorder(dwhmc_handle hh, const double* w) {
  H_ENTER(hh);
  if (!w) BADARG("dwhmc_set_disorder: NULL");
  return h2d(h, h->w, w, sizeof(double) * (size_t)h->N * h->B);
}
int dwhmc_get_disorder(dwhmc_handle hh, double* w) {
  H_ENTER(hh);
  if (!w) BADARG("dwhmc_get_disorder: NULL");
  return d2h(h, w, h->w, sizeof(double) * (size_t)h->N * h->B);
}
int dwhmc_set_field(dwhmc_handle hh, const double* delta) {
  H_ENTER(hh);
  if (!delta) BADARG("dwhmc_set_field: NULL");
  return h2d(h, h->delta, delta, sizeof(cplx) * (size_t)h->n * h->B);
}
int dwhmc_get_field(dwhmc_handle hh, double* delta) {
  H_ENTER(hh);
  if (!delta) BADARG("dwhmc_get_field: NULL");
  return d2h(h, delta, h->delta, sizeof(cplx) * (size_t)h->n * h->B);
}
int dwhmc_set_momentum(dwhmc_handle hh, const double* pi) {
  H_ENTER(hh);
  if (!pi) BADARG("dwhmc_set_momentum: NULL");
  return h2d(h, h->pi, pi, sizeof(cplx) * (size_t)h->n * h->B);
}
int dwhmc_get_momentum(dwhmc_handle hh, double* pi) {
  H_ENTER(hh);
  if (!pi) BADARG("dwhmc_get_momentum: NULL");
  return d2h(h, pi, h->pi, sizeof(cplx) * (size_t)h->n * h->B);
}
int dwhmc_init_state(dwhmc_handle hh, const double* W, const double* n_imp) {
  H_ENTER(hh);
  if (!W || !n_imp) BADARG("dwhmc_init_state: NULL");
  for (int b = 0; b < h->B; ++b)
    if (!(n_imp[b] >= 0.0 && n_imp[b] <= 1.0)) BADARG("dwhmc_init_state: n_imp must lie in [0, 1]");
  // unif_dev / dH_dev double as the parameter staging area ([B] doubles each)
  DW_TRY(h2d(h, h->unif_dev, W, sizeof(double) * h->B));
  DW_TRY(h2d(h, h->dH_dev, n_imp, sizeof(double) * h->B));
  DW_TRY(dw_init_state(h, h->unif_dev, h->dH_dev));
  DW_CUDA(h, cudaStreamSynchronize(h->stream));
  return DWHMC_OK;
}

int dwhmc_seed(dwhmc_handle hh, uint64_t seed) {
  H_ENTER(hh);
  h->seed = seed;
  h->rng_counter = 0;
  return DWHMC_OK;
}

// ---- per-operator entry points -------------------------------------------------------------------

int dwhmc_init_static_H(dwhmc_handle hh) {
  H_ENTER(hh);
  if (!h->params_set) { h->err = "dwhmc_init_static_H: parameters not set"; return DWHMC_E_STATE; }
  // fill!(H, 0) clears the pairing block too (src/Hamiltonian.jl:15)
  DW_CUDA(h, cudaMemcpyAsync(h->Hs_w, h->w, sizeof(double) * (size_t)h->N * h->B, cudaMemcpyDeviceToDevice, h->stream));
  DW_CUDA(h, cudaMemcpyAsync(h->Hs_par, h->par, sizeof(double) * 3 * h->B, cudaMemcpyDeviceToDevice, h->stream));
  DW_CUDA(h, cudaMemsetAsync(h->Hs_delta, 0, sizeof(cplx) * (size_t)h->n * h->B, h->stream));
  h->static_set = true;
  return DWHMC_OK;
}

int dwhmc_update_H(dwhmc_handle hh) {
  H_ENTER(hh);
  DW_CUDA(h, cudaMemcpyAsync(h->Hs_delta, h->delta, sizeof(cplx) * (size_t)h->n * h->B, cudaMemcpyDeviceToDevice,
                             h->stream));
  return DWHMC_OK;
}

int dwhmc_diagonalize(dwhmc_handle hh) {
  H_ENTER(hh);
  {
    StageTimer t(h, 0);
    DW_TRY(dw_assemble_for_solve(h, h->Hs_w, h->Hs_par, h->Hs_delta, no_mask()));
  }
  DW_TRY(dw_eigensolve(h, h->E_cur, h->U_cur, no_mask(), true));
  DW_CUDA(h, cudaStreamSynchronize(h->stream));
  return check_status(h);
}

int dwhmc_compute_forces(dwhmc_handle hh) {
  H_ENTER(hh);
  if (!h->params_set) { h->err = "dwhmc_compute_forces: parameters not set"; return DWHMC_E_STATE; }
  StageTimer t(h, 4);
  DW_TRY(dw_forces(h, h->E_cur, h->U_cur, 0, 0));
  return DWHMC_OK;
}

int dwhmc_total_energy(dwhmc_handle hh, double* out) {
  H_ENTER(hh);
  if (!out) BADARG("dwhmc_total_energy: NULL");
  if (!h->params_set) { h->err = "dwhmc_total_energy: parameters not set"; return DWHMC_E_STATE; }
  DW_TRY(dw_total_energy(h, h->E_cur, h->Hold_dev));
  return d2h(h, out, h->Hold_dev, sizeof(double) * h->B);
}

int dwhmc_measure_observables(dwhmc_handle hh, double* out) {
  H_ENTER(hh);
  if (!out) BADARG("dwhmc_measure_observables: NULL");
  if (!h->params_set) { h->err = "dwhmc_measure_observables: parameters not set"; return DWHMC_E_STATE; }
  DW_TRY(dw_observables(h, h->obs_dev));
  return d2h(h, out, h->obs_dev, sizeof(double) * DWHMC_NOBS * h->B);
}

int dwhmc_measure_transport(dwhmc_handle hh, double eta, const double* omega_grid, int n_omega, const double* dos_grid,
                            int n_dos, double* scalars, double* sigma, double* dos, double* dos_AN, double* A_k0) {
  H_ENTER(hh);
  if (!h->params_set) { h->err = "dwhmc_measure_transport: parameters not set"; return DWHMC_E_STATE; }
  if (h->pending) { h->err = "dwhmc_measure_transport: a trajectory proposal is pending (commit first)"; return DWHMC_E_STATE; }
  if (!(eta > 0.0) || n_omega < 0 || n_dos < 0 || (n_omega > 0 && !omega_grid) || (n_dos > 0 && !dos_grid))
    BADARG("dwhmc_measure_transport: need eta > 0 and the frequency grids");
  const int B = h->B, N = h->N;
  const size_t wc = dw_transport_work_count(h, n_omega);
  if (wc > h->tr_work_count) {
    if (h->tr_work) cudaFree(h->tr_work);
    h->tr_work = nullptr; h->tr_work_count = 0;
    DW_CUDA(h, cudaMalloc(&h->tr_work, sizeof(double) * wc));
    h->tr_work_count = wc;
  }
  const size_t oc = (size_t)2 * B + (size_t)n_omega * B + (size_t)2 * n_dos * B + (size_t)N * B + n_omega + n_dos + 8;
  if (oc > h->tr_out_count) {
    if (h->tr_out) cudaFree(h->tr_out);
    h->tr_out = nullptr; h->tr_out_count = 0;
    DW_CUDA(h, cudaMalloc(&h->tr_out, sizeof(double) * oc));
    h->tr_out_count = oc;
  }
  double* scal_d = h->tr_out;
  double* sigma_d = scal_d + (size_t)2 * B;
  double* dos_d = sigma_d + (size_t)n_omega * B;
  double* dosAN_d = dos_d + (size_t)n_dos * B;
  double* ak_d = dosAN_d + (size_t)n_dos * B;
  double* omega_d = ak_d + (size_t)N * B;
  double* grid_d = omega_d + n_omega;
  if (n_omega > 0) DW_TRY(h2d(h, omega_d, omega_grid, sizeof(double) * n_omega));
  if (n_dos > 0) DW_TRY(h2d(h, grid_d, dos_grid, sizeof(double) * n_dos));
  DW_TRY(dw_transport(h, eta, omega_d, n_omega, grid_d, n_dos, scal_d, sigma_d, dos_d, dosAN_d, ak_d, h->tr_work,
                      h->tr_work_count));
  if (scalars) DW_TRY(d2h(h, scalars, scal_d, sizeof(double) * 2 * B));
  if (sigma && n_omega > 0) DW_TRY(d2h(h, sigma, sigma_d, sizeof(double) * (size_t)n_omega * B));
  if (dos && n_dos > 0) DW_TRY(d2h(h, dos, dos_d, sizeof(double) * (size_t)n_dos * B));
  if (dos_AN && n_dos > 0) DW_TRY(d2h(h, dos_AN, dosAN_d, sizeof(double) * (size_t)n_dos * B));
  if (A_k0) DW_TRY(d2h(h, A_k0, ak_d, sizeof(double) * (size_t)N * B));
  DW_CUDA(h, cudaStreamSynchronize(h->stream));
  return DWHMC_OK;
}

int dwhmc_get_H(dwhmc_handle hh, double* out) {
  H_ENTER(hh);
  if (!out) BADARG("dwhmc_get_H: NULL");
  if (h->pending) { h->err = "dwhmc_get_H: a trajectory proposal is pending (commit first)"; return DWHMC_E_STATE; }
  DW_TRY(dw_assemble_upper(h, h->Hs_w, h->Hs_par, h->Hs_delta, h->A));
  return d2h(h, out, h->A, sizeof(cplx) * (size_t)h->n * h->n * h->B);
}
int dwhmc_get_eigenvalues(dwhmc_handle hh, double* out) {
  H_ENTER(hh);
  if (!out) BADARG("dwhmc_get_eigenvalues: NULL");
  return d2h(h, out, h->E_cur, sizeof(double) * (size_t)h->n * h->B);
}
int dwhmc_get_eigenvectors(dwhmc_handle hh, double* out) {
  H_ENTER(hh);
  if (!out) BADARG("dwhmc_get_eigenvectors: NULL");
  return d2h(h, out, h->U_cur, sizeof(cplx) * (size_t)h->n * h->n * h->B);
}
int dwhmc_get_forces(dwhmc_handle hh, double* out) {
  H_ENTER(hh);
  if (!out) BADARG("dwhmc_get_forces: NULL");
  return d2h(h, out, h->force, sizeof(cplx) * (size_t)h->n * h->B);
}
int dwhmc_get_fermi(dwhmc_handle hh, double* out) {
  H_ENTER(hh);
  if (!out) BADARG("dwhmc_get_fermi: NULL");
  return d2h(h, out, h->fermi, sizeof(double) * (size_t)h->n * h->B);
}

// ---- batched trajectory --------------------------------------------------------------------------

// enqueue src/HMC.jl:77-124 for all chains; h_nt: host copy of Nt (for the loop bound)
static int trajectory_enqueue(Handle* h, int max_nt, bool device_momentum) {
  if (device_momentum) DW_TRY(dw_refresh_momentum(h));
  DW_TRY(dw_total_energy(h, h->E_cur, h->Hold_dev));
  DW_TRY(dw_begin_trajectory(h));
  {
    StageTimer t(h, 4);
    DW_TRY(dw_forces(h, h->E_cur, h->U_cur, 1, 0));
  }
  for (int s = 1; s <= max_nt; ++s) {
    Mask mask; mask.nt = h->nt_dev; mask.step = s;
    {
      StageTimer t(h, 0);
      DW_TRY(dw_assemble_for_solve(h, h->Hs_w, h->Hs_par, h->delta, mask));
    }
    DW_TRY(dw_eigensolve(h, h->E_prop, h->U_prop, mask, true));
    StageTimer t(h, 4);
    DW_TRY(dw_forces(h, h->E_prop, h->U_prop, 1, s));
  }
  DW_TRY(dw_total_energy(h, h->E_prop, h->Hnew_dev));
  DW_TRY(dw_dH(h));
  return DWHMC_OK;
}

static int load_steps(Handle* h, const int32_t* Nt, const double* dt, int* max_nt) {
  if (!Nt || !dt) BADARG("trajectory: Nt and dt are required (hmc_sweep! has no defaults)");
  if (!h->params_set) { h->err = "trajectory: parameters not set"; return DWHMC_E_STATE; }
  int mx = 0;
  for (int b = 0; b < h->B; ++b) {
    if (Nt[b] < 1) BADARG("trajectory: Nt must be >= 1");
    if (Nt[b] > mx) mx = Nt[b];
  }
  *max_nt = mx;
  DW_CUDA(h, cudaMemcpyAsync(h->nt_dev, Nt, sizeof(int) * h->B, cudaMemcpyHostToDevice, h->stream));
  DW_CUDA(h, cudaMemcpyAsync(h->dt_dev, dt, sizeof(double) * h->B, cudaMemcpyHostToDevice, h->stream));
  DW_CUDA(h, cudaStreamSynchronize(h->stream));
  return DWHMC_OK;
}

int dwhmc_trajectory(dwhmc_handle hh, const int32_t* Nt, const double* dt, const double* pi0, double* H_old,
                     double* H_new, double* dH) {
  H_ENTER(hh);
  if (h->pending) { h->err = "dwhmc_trajectory: previous proposal not committed"; return DWHMC_E_STATE; }
  int max_nt = 0;
  DW_TRY(load_steps(h, Nt, dt, &max_nt));
  if (pi0) DW_TRY(h2d(h, h->pi, pi0, sizeof(cplx) * (size_t)h->n * h->B));
  // On any failure after the trajectory has been enqueued the proposal is dropped: Delta goes back to the backup
  // and the handle stays usable, like the reference's cache after a LAPACKException (src/Hamiltonian.jl:106).
  auto run = [&]() -> int {
    DW_TRY(trajectory_enqueue(h, max_nt, pi0 == nullptr));
    if (H_old) DW_TRY(d2h(h, H_old, h->Hold_dev, sizeof(double) * h->B));
    if (H_new) DW_TRY(d2h(h, H_new, h->Hnew_dev, sizeof(double) * h->B));
    if (dH) DW_TRY(d2h(h, dH, h->dH_dev, sizeof(double) * h->B));
    DW_CUDA(h, cudaStreamSynchronize(h->stream));
    return check_status(h);
  };
  const int rc = run();
  if (rc == DWHMC_OK) { h->pending = true; return DWHMC_OK; }
  const std::string why = h->err;
  if (cudaMemsetAsync(h->accept_dev, 0, sizeof(int) * h->B, h->stream) == cudaSuccess && dw_commit_dev(h) == DWHMC_OK)
    cudaStreamSynchronize(h->stream);
  h->pending = false;
  h->err = why;
  return rc;
}

int dwhmc_commit(dwhmc_handle hh, const int32_t* accept) {
  H_ENTER(hh);
  if (!accept) BADARG("dwhmc_commit: NULL");
  if (!h->pending) { h->err = "dwhmc_commit: no pending trajectory"; return DWHMC_E_STATE; }
  DW_TRY(h2d(h, h->accept_dev, accept, sizeof(int) * h->B));
  DW_TRY(dw_commit_dev(h));
  h->pending = false;
  DW_CUDA(h, cudaStreamSynchronize(h->stream));
  return DWHMC_OK;
}

int dwhmc_hmc_sweep(dwhmc_handle hh, const int32_t* Nt, const double* dt, const double* pi0, const double* uniforms,
                    int32_t* accepted, double* dH) {
  H_ENTER(hh);
  if (h->pending) { h->err = "dwhmc_hmc_sweep: previous proposal not committed"; return DWHMC_E_STATE; }
  int max_nt = 0;
  DW_TRY(load_steps(h, Nt, dt, &max_nt));
  if (pi0) DW_TRY(h2d(h, h->pi, pi0, sizeof(cplx) * (size_t)h->n * h->B));
  if (uniforms) DW_TRY(h2d(h, h->unif_dev, uniforms, sizeof(double) * h->B));
  else DW_TRY(dw_uniforms(h));
  DW_TRY(trajectory_enqueue(h, max_nt, pi0 == nullptr));
  DW_TRY(dw_metropolis(h, true));
  DW_TRY(dw_commit_dev(h));
  if (accepted) DW_TRY(d2h(h, accepted, h->accept_dev, sizeof(int) * h->B));
  if (dH) DW_TRY(d2h(h, dH, h->dH_dev, sizeof(double) * h->B));
  DW_CUDA(h, cudaStreamSynchronize(h->stream));
  return check_status(h);
}

int dwhmc_run_sweeps(dwhmc_handle hh, int n_sweeps, const int32_t* Nt, const double* dt, int32_t* n_accepted,
                     double* last_dH, double* obs) {
  H_ENTER(hh);
  if (n_sweeps < 0) BADARG("dwhmc_run_sweeps: n_sweeps < 0");
  if (h->pending) { h->err = "dwhmc_run_sweeps: previous proposal not committed"; return DWHMC_E_STATE; }
  int max_nt = 0;
  DW_TRY(load_steps(h, Nt, dt, &max_nt));
  DW_CUDA(h, cudaMemsetAsync(h->nacc_dev, 0, sizeof(int) * h->B, h->stream));
  DW_CUDA(h, cudaEventRecord(h->ev_begin, h->stream));
  double* obs_all = nullptr;
  const size_t per = (size_t)DWHMC_NOBS * h->B;
  if (obs && n_sweeps > 0) DW_CUDA(h, cudaMalloc(&obs_all, sizeof(double) * per * n_sweeps));
  int rc = DWHMC_OK;
  for (int s = 0; s < n_sweeps && rc == DWHMC_OK; ++s) {
    rc = dw_uniforms(h);
    if (rc == DWHMC_OK) rc = trajectory_enqueue(h, max_nt, true);
    if (rc == DWHMC_OK) rc = dw_metropolis(h, true);
    if (rc == DWHMC_OK) rc = dw_commit_dev(h);
    if (rc == DWHMC_OK && obs_all) rc = dw_observables(h, obs_all + per * s);
  }
  if (rc == DWHMC_OK) {
    cudaEventRecord(h->ev_end, h->stream);
    cudaEventSynchronize(h->ev_end);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_begin, h->ev_end);
    h->last_ms = ms;
  }
  if (rc == DWHMC_OK && obs_all) rc = d2h(h, obs, obs_all, sizeof(double) * per * n_sweeps);
  if (obs_all) { cudaStreamSynchronize(h->stream); cudaFree(obs_all); }
  DW_TRY(rc);
  if (n_accepted) DW_TRY(d2h(h, n_accepted, h->nacc_dev, sizeof(int) * h->B));
  if (last_dH) DW_TRY(d2h(h, last_dH, h->dH_dev, sizeof(double) * h->B));
  DW_CUDA(h, cudaStreamSynchronize(h->stream));
  return check_status(h);
}

// ---- instrumentation -----------------------------------------------------------------------------

int dwhmc_get_timers(dwhmc_handle hh, double* out) {
  H_ENTER(hh);
  if (!out) BADARG("dwhmc_get_timers: NULL");
  for (int i = 0; i < 8; ++i) out[i] = h->timers[i];
  out[5] = (double)h->eigensolves;
  out[6] = (double)h->launches;
  return DWHMC_OK;
}
int dwhmc_last_elapsed_ms(dwhmc_handle hh, double* out) {
  H_ENTER(hh);
  if (!out) BADARG("dwhmc_last_elapsed_ms: NULL");
  *out = h->last_ms;
  return DWHMC_OK;
}
int dwhmc_reset_timers(dwhmc_handle hh) {
  H_ENTER(hh);
  for (int i = 0; i < 8; ++i) h->timers[i] = 0.0;
  h->eigensolves = 0;
  h->launches = 0;
  return DWHMC_OK;
}
int dwhmc_set_profiling(dwhmc_handle hh, int on) {
  H_ENTER(hh);
  h->profiling = on < 0 ? 0 : on;
  return DWHMC_OK;
}

int dwhmc_debug_tridiagonalize(dwhmc_handle hh, double* d, double* e) {
  H_ENTER(hh);
  if (!d || !e) BADARG("dwhmc_debug_tridiagonalize: NULL");
  DW_TRY(dw_assemble_for_solve(h, h->Hs_w, h->Hs_par, h->Hs_delta, no_mask()));
  if (h->band_b > 0) DW_TRY(dw_band_chase(h, no_mask()));
  else DW_TRY(dw_hetrd(h, h->U_prop, no_mask()));
  std::vector<double> tmp((size_t)h->n * h->B);
  DW_TRY(d2h(h, d, h->d, sizeof(double) * (size_t)h->n * h->B));
  DW_TRY(d2h(h, tmp.data(), h->e, sizeof(double) * (size_t)h->n * h->B));
  for (int b = 0; b < h->B; ++b)
    for (int i = 0; i + 1 < h->n; ++i) e[(size_t)b * (h->n - 1) + i] = tmp[(size_t)b * h->n + i];
  return DWHMC_OK;
}

int dwhmc_debug_stedc(dwhmc_handle hh, const double* d, const double* e, double* w, double* Z) {
  H_ENTER(hh);
  if (!d || !e || !w || !Z) BADARG("dwhmc_debug_stedc: NULL");
  const int n = h->n, B = h->B;
  std::vector<double> tmp((size_t)n * B, 0.0);
  for (int b = 0; b < B; ++b)
    for (int i = 0; i + 1 < n; ++i) tmp[(size_t)b * n + i] = e[(size_t)b * (n - 1) + i];
  DW_TRY(h2d(h, h->d, d, sizeof(double) * (size_t)n * B));
  DW_TRY(h2d(h, h->e, tmp.data(), sizeof(double) * (size_t)n * B));
  DW_TRY(dw_stedc(h, no_mask()));
  DW_TRY(dw_stedc_output(h, h->E_prop, h->U_prop, no_mask(), false, 0));
  DW_TRY(d2h(h, w, h->E_prop, sizeof(double) * (size_t)n * B));
  std::vector<double> zc((size_t)2 * n * n * B);
  DW_TRY(d2h(h, zc.data(), h->U_prop, sizeof(cplx) * (size_t)n * n * B));
  for (size_t i = 0; i < (size_t)n * n * B; ++i) Z[i] = zc[2 * i];
  return check_status(h);
}

int dwhmc_debug_heev(dwhmc_handle hh, const double* A, double* E, double* U) {
  H_ENTER(hh);
  if (!A || !E || !U) BADARG("dwhmc_debug_heev: NULL");
  if (h->pending) { h->err = "dwhmc_debug_heev: a trajectory proposal is pending"; return DWHMC_E_STATE; }
  const int n = h->n, B = h->B;
  DW_TRY(h2d(h, h->A, A, sizeof(cplx) * (size_t)n * n * B));
  DW_TRY(dw_eigensolve(h, h->E_prop, h->U_prop, no_mask(), false));
  DW_TRY(d2h(h, E, h->E_prop, sizeof(double) * (size_t)n * B));
  DW_TRY(d2h(h, U, h->U_prop, sizeof(cplx) * (size_t)n * n * B));
  return check_status(h);
}

}  // extern "C"
