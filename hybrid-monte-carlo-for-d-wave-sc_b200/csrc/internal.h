// internal.h -- shared declarations of libdwhmc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "stedc_tree.h"

#ifndef DW_NB
#define DW_NB 32          // Householder panel width of the tridiagonalisation
#endif
#ifndef DW_CC
#define DW_CC 4
#endif
//      DW_CC:           // CTAs per cluster in the column-step kernel
#ifndef DW_GSPLIT
#define DW_GSPLIT 8
#endif
// DW_GSPLIT: K pieces of the Gram-matrix GEMMs of the back-transformation
#define DW_NBT 64         // reflectors per block of the eigenvector back-transformation
#ifndef DW_LEAF
#define DW_LEAF 18        // largest D&C leaf (18: one more merge level than 36, but the QL leaf stage is 4x shorter; 11.08 -> 10.84 ms at L = 24)
#endif
#define DW_NGROUP 4       // max chain groups (streams) of the tridiagonalisation
#define DW_APPLY_G 32      // reflectors per staircase block of the band route's back-transformation
#define DW_APPLY_ROWS 136  // rows of such a block (half-bandwidth + DW_APPLY_G - 1 at most)
#define DW_APPLY_BLOCK_DOUBLES (2 * 2 * DW_APPLY_G * DW_APPLY_ROWS + DW_APPLY_G * DW_APPLY_ROWS + DW_APPLY_ROWS * (DW_APPLY_G + 8))  // operand planes per block
#define DW_FCHUNK 32      // eigenvector columns per CTA in the bond-correlator kernel

typedef double2 cplx;

// per-chain activity mask: chain b takes part in leapfrog step `step` iff step <= nt[b]
// (steps are 1-based; nt == nullptr means every chain is active)
struct Mask {
  const int* nt;
  int step;
  __host__ __device__ bool on(int b) const { return nt == nullptr || step <= nt[b]; }
};
static inline Mask no_mask() { Mask m; m.nt = nullptr; m.step = 0; return m; }

struct DcLevelDev {
  int nmerge;
  int max_m;
  int* off;   // device [nmerge]
  int* n1;
  int* n2;
};

struct Handle {
  int device = 0;
  int B = 0, Lx = 0, Ly = 0, N = 0, n = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int profiling = 0;             // 1: stage timers; 2: also per-launch hemv timing (serialises the groups)
  double timers[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_begin = nullptr, ev_end = nullptr;
  cudaStream_t gstream[DW_NGROUP] = {};      // bulk kernels of a chain group (low priority)
  cudaStream_t gstream_hi[DW_NGROUP] = {};   // column steps of a chain group (high priority)
  cudaEvent_t ev_fork = nullptr, ev_join[DW_NGROUP] = {}, ev_bulk[DW_NGROUP] = {}, ev_col[DW_NGROUP] = {};
  double last_ms = 0.0;          // device time of the last dwhmc_run_sweeps (events on `stream`)
  std::vector<void*> allocs;    // everything cudaMalloc'ed, for destroy

  // lattice (device, 0-based int32, [dir * N + site])
  int* nn = nullptr;
  int* nnn = nullptr;
  // per-chain parameters, device [6*B]: t, tp, mu, beta, J, mass (each a [B] slice) + host mirror
  double* par = nullptr;
  std::vector<double> h_par;
  bool params_set = false;
  // state
  double* w = nullptr;          // disorder [N*B]
  cplx* delta = nullptr;        // [2N*B]
  cplx* pi = nullptr;
  cplx* force = nullptr;
  cplx* delta_backup = nullptr;
  // what cache.H_base holds in the reference: static part as of init_static_H!, pairing as of update_H_BdG!
  double* Hs_w = nullptr;       // [N*B] disorder snapshot
  double* Hs_par = nullptr;     // [3*B] (t, tp, mu) snapshot, slices of [B]
  cplx* Hs_delta = nullptr;     // [2N*B]
  bool static_set = false;
  // eigensystems
  double *E_cur = nullptr, *E_prop = nullptr;   // [n*B]
  cplx *U_cur = nullptr, *U_prop = nullptr;     // [n*n*B]
  double* fermi = nullptr;                      // [n*B]
  bool pending = false;                         // a trajectory proposal awaits commit
  int* nt_dev = nullptr;                        // [B]
  double* dt_dev = nullptr;                     // [B]
  double *Hold_dev = nullptr, *Hnew_dev = nullptr, *dH_dev = nullptr;  // [B]
  int* accept_dev = nullptr;                    // [B]
  int* nacc_dev = nullptr;                      // [B]
  double* unif_dev = nullptr;                   // [B]
  double* obs_dev = nullptr;                    // [9*B]
  unsigned long long seed = 0x9E3779B97F4A7C15ull;
  unsigned long long rng_counter = 0;
  // force / observable partial sums
  int fchunks = 0;              // column chunks per chain
  cplx* Ppart = nullptr;        // [fchunks * 2N * B]
  double* hpart = nullptr;      // [fchunks * B]
  cplx* Pbond = nullptr;        // [2N*B] reduced pair correlators
  // eigensolver workspace
  cplx* A = nullptr;            // [n*n*B] work matrix (full Hermitian storage)
  cplx* V = nullptr;            // [n*n*B] Householder vectors, explicit unit entries, zeros above
  cplx* ypart = nullptr;        // [ceil(n/64)*n*B] hemv partial results, one slot per tile row/column
  cplx* P1 = nullptr;           // [CC*NB*B] W_panel^H v, one partial per cluster rank
  cplx* P2 = nullptr;           // [CC*NB*B] V_panel^H v
  cplx* Gb = nullptr;           // [nbt*NBT*NBT*B] Gram matrices of the back-transform blocks
  cplx* Tf = nullptr;           // [nbt*NBT*NBT*B] block reflector T factors
  cplx* tau = nullptr;          // [n*B]
  double *d = nullptr, *e = nullptr;            // [n*B]
  cplx* Wbt = nullptr;          // [NBT*n*B] back-transform workspace
  cplx* Wbt2 = nullptr;
  int nblk = 0, nbt = 0;
  // D&C workspace
  DcTree tree;
  std::vector<DcLevelDev> levels;
  int* leaf_off = nullptr; int* leaf_size = nullptr; int nleaves = 0;
  double *Z0 = nullptr, *Z1 = nullptr, *S = nullptr;   // [n*n*B]
  double* Zfinal = nullptr;
  int* perm = nullptr;          // [n*B]
  int* ord = nullptr;           // [n*B]
  double *zvec = nullptr, *dl = nullptr, *wv = nullptr, *dnew = nullptr, *zhat = nullptr, *stau = nullptr;  // [n*B]
  int *ndl = nullptr, *dfl = nullptr, *sorg = nullptr;   // [n*B]
  int *ndc = nullptr, *kcls = nullptr;                   // [n*B] class-ordered non-deflated columns, class counts
  void* rots = nullptr;         // [n*B] DeflRot
  int* kcnt = nullptr;          // [n*B] indexed by b*n + merge offset
  int* nrot = nullptr;          // [n*B]
  double* rho = nullptr;        // [n*B]
  int* status = nullptr;        // [4] device: [0] leaf failures, [1] secular non-convergence
  // band route (band.cu): half-bandwidth of the BdG matrix in the folded site order; 0 = dense route
  int band_b = 0, band_LD = 0, band_KT = 0, band_g = 0;
  std::vector<int> band_pos_host, band_blk_s0, band_blk_k;
  std::vector<int> band_wave_blk, band_wave_start;   // blocks sorted into wavefronts of independent blocks
  int* band_pos = nullptr;      // device [n]: band index of row r of the reference's matrix
  int* band_prog = nullptr;     // device [n*B]: steps completed per sweep (pipelining of the bulge chase); then [B]: next sweep to hand out
  cplx* band_tau = nullptr;     // device [n*KT*B]: tau of reflector (sweep, step)
  cplx* band_rowbox = nullptr;  // device [B][KT][2][b+2]: row messages of the position-owning chase (band_systolic.cu)
  cplx* band_bbox = nullptr;    // device [B][KT][2]: beta of a step and, in the same 32-byte sector, the counter that publishes it
  cplx* band_VT = nullptr;      // device, DW_APPLY_BLOCK_DOUBLES doubles per (chain, block): conj(V) and -V T, operand planes of the back-transformation
  int band_nitems = 0;                    // work items per chain of the back-transformation (blocks x column parts, wavefront order)
  int band_apply_attr = 0;                // row tiles of the apply kernel instance whose shared-memory attribute is set
  int* band_items_dev = nullptr;          // device [4*band_nitems]
  int* band_sync = nullptr;               // device [1 + B*nwave]: ticket, published items per (chain, wavefront)
  int* band_blk_s0_dev = nullptr; int* band_blk_k_dev = nullptr;
  alignas(64) unsigned char band_tmap_a[128] = {};  // CUtensorMaps of the band storage (TMA chase kernel): column pieces of the carried block
  alignas(64) unsigned char band_tmap_b[128] = {};
  bool band_tmap_set = false;
  // transport / spectra workspace (transport.cu), allocated on first use
  double* tr_work = nullptr; size_t tr_work_count = 0;
  double* tr_out = nullptr; size_t tr_out_count = 0;     // scal | sigma | dos | dosAN | ak | omega | dosgrid
  int ngroups = 3;              // chain groups in use (env DWHMC_NGROUP, 1..DW_NGROUP)
  // particle-hole symmetry of the BdG matrix (tau_y H^* tau_y = -H): only the N eigenvectors of the
  // upper half of the spectrum are back-transformed, the rest are their conjugate partners
  int nsm = 148;                // SMs of the device
  int ph_mode = 1;              // env DWHMC_PH=0 switches the shortcut off
  int* halfflag = nullptr;      // device [B]: 1 = this chain's last eigensolve used the shortcut
  long long launches = 0;
  long long eigensolves = 0;
};

// error helpers --------------------------------------------------------------------
#define DW_CUDA(h, call)                                                              \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
      return DWHMC_E_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define DW_LAUNCH_CHECK(h)                                                            \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    (h)->launches++;                                                                  \
    if (e__ != cudaSuccess) {                                                         \
      (h)->err = std::string("kernel launch: ") + cudaGetErrorString(e__) + " at " +  \
                 __FILE__ + ":" + std::to_string(__LINE__);                           \
      return DWHMC_E_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define DW_TRY(expr)                    \
  do {                                  \
    int rc__ = (expr);                  \
    if (rc__ != DWHMC_OK) return rc__;  \
  } while (0)

// stage entry points (each returns a DWHMC_* code) ----------------------------------
// assemble.cu: full Hermitian BdG matrix of every active chain into h->A
int dw_assemble(Handle* h, const double* w, const double* par3, const cplx* delta, cplx* A, Mask mask);
// reference-layout H_base (upper triangle, lower = 0) into out
int dw_assemble_upper(Handle* h, const double* w, const double* par3, const cplx* delta, cplx* out);
// hetrd.cu: h->A -> h->d, h->e, h->V, h->tau, h->Tf ; Wscratch is an n*n*B complex scratch
int dw_hetrd(Handle* h, cplx* Wscratch, Mask mask);
// stedc.cu: h->d, h->e -> eigenvalues E_out (ascending) and real eigenvectors written as complex into U_out.
// ph: decide per chain (h->halfflag) whether the particle-hole shortcut applies; flagged chains only get
// the columns of the upper half of the spectrum.
// c_lo: first eigenvector column (rank) flagged chains need (0: all)
int dw_stedc(Handle* h, Mask mask, bool ph = false, int c_lo = 0);
int dw_stedc_output(Handle* h, double* E_out, cplx* U_out, Mask mask, bool ph, int c_lo);
// backtransform.cu: U <- Q U (flagged chains: upper-half columns only, then the partner columns)
int dw_backtransform(Handle* h, cplx* U, Mask mask, bool ph);
int dw_ph_mirror(Handle* h, double* E, cplx* U, Mask mask);
// A (assembled, destroyed) -> E, U.  ph: the matrix is the BdG matrix assembled by dw_assemble_for_solve
// (band storage when the band route is active); otherwise a dense Hermitian matrix in h->A.
int dw_eigensolve(Handle* h, double* E_out, cplx* U_out, Mask mask, bool ph);
int dw_assemble_for_solve(Handle* h, const double* w, const double* par3, const cplx* delta, Mask mask);
// band.cu
int dw_band_setup(Handle* h, const std::vector<int>& nn, const std::vector<int>& nnn);
int dw_band_assemble(Handle* h, const double* w, const double* par3, const cplx* delta, Mask mask);
int dw_band_chase(Handle* h, Mask mask);
int dw_band_tfactors(Handle* h, Mask mask, cudaStream_t stream);
int dw_band_backtransform(Handle* h, cplx* U, Mask mask, bool ph);
void dw_band_apply_items(Handle* h, std::vector<int>& items4);

// generic batched complex GEMM on FP64 tensor cores (gemm_dmma.cu)
// C = beta*C + alpha * sum_seg opA(A_seg) * opB(B_seg);  opX: 0 = N, 1 = C (conjugate transpose)
struct ZgemmArgs {
  int M, N, K;                 // K per segment
  int nseg;                    // 1 or 2
  const cplx* A[2]; const cplx* Bm[2];
  int lda, ldb, ldc;
  long long sA, sB, sC;        // batch strides (elements)
  cplx* C;
  double alpha, beta;
  int opA, opB;
  int lower;                   // only tiles touching the lower triangle
  int batch;
  Mask mask;
  int b0 = 0;                  // first chain of the launch (chain groups)
  int ksplit = 1;              // > 1: the K range is cut into ksplit pieces, piece s of chain b writes its partial
  long long sCk = 0;           //   product to C + b * sC + s * sCk (beta must be 0; the consumer sums the pieces)
  int stairA = 0;              // > 0: the A operand is a staircase block: entry (row r, column c) of the
                               //   source is used only if 0 <= r - c < stairA (band.cu block reflectors)
  const int* skip_flag = nullptr;  // device [B] or null: chains with flag != 0 skip the column tiles that
  int skip_cols = 0;               //   lie entirely below column skip_cols
  cudaStream_t stream = nullptr;   // nullptr = the handle's stream
};
int dw_zgemm(Handle* h, const ZgemmArgs& a);

// transport.cu
int dw_transport(Handle* h, double eta, const double* omega_dev, int nw, const double* dosgrid_dev, int nd,
                 double* scal, double* sigma, double* dos, double* dosAN, double* ak, double* work, size_t work_count);
size_t dw_transport_work_count(const Handle* h, int nw);

// force.cu ---------------------------------------------------------------------------
// bond correlators of (U, E) -> h->Pbond, h->fermi; then F = -(beta/2J)(Delta - J P) -> h->force.
// mode 0: force only.  mode 1 (trajectory): also the leapfrog kick / drift that follows force
// evaluation number `step` (0 = start of trajectory) for chains with step <= nt[b].
int dw_forces(Handle* h, const double* E, const cplx* U, int mode, int step);
int dw_total_energy(Handle* h, const double* E, double* out_dev);   // out_dev [B]
int dw_observables(Handle* h, double* out_dev);                     // out_dev [9*B], uses E_cur/U_cur
int dw_init_state(Handle* h, const double* W_dev, const double* nimp_dev);   // initialize_state on the device
int dw_refresh_momentum(Handle* h);                                 // Philox N(0, m)
int dw_uniforms(Handle* h);                                         // Philox U[0,1) -> h->unif_dev
int dw_begin_trajectory(Handle* h);                                 // backup Delta
int dw_metropolis(Handle* h, bool use_uniforms);                    // dH, unif -> accept_dev
int dw_commit_dev(Handle* h);                                       // accept_dev -> swap / restore
int dw_dH(Handle* h);                                               // dH = Hnew - Hold
