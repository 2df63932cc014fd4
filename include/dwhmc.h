/*
 * dwhmc.h -- C ABI of libdwhmc.so: the B200 (sm_100a) implementation of the
 * molecular-dynamics force path of DwaveHMC.jl, batched over independent chains.
 *
 * The reference has no FFI layer; its operator API for this path is the set of
 * exported Julia functions in src/DwaveHMC.jl:3-9 acting on (cache, p, state).
 * Each entry point below names the reference function (file:line under
 * /root/reference) it replaces.  A thin Julia `ccall` shim (see INTEGRATION.md
 * and hybrid-monte-carlo-for-d-wave-sc_b200/julia/DwaveHMCB200.jl) keeps those
 * signatures; the Python host package `dwhmc` binds the same symbols by ctypes.
 *
 * Conventions
 *   - every function returns int: 0 = ok, nonzero = error (DWHMC_E_*); the text
 *     is available from dwhmc_last_error().  The Julia shim turns nonzero into
 *     error(...) -- the reference's error convention is exceptions.
 *   - one handle = one CUDA device + one stream + B chains.  Calls on a handle
 *     are serialised by the caller; different handles are independent.
 *   - host pointers are borrowed for the duration of the call only.
 *   - layouts are the reference's (src/Types.jl): column-major, complex numbers
 *     as interleaved (re, im) doubles, neighbour tables Int64 N x 4 column-major
 *     and 1-based, fields/momenta/forces N x 2 (column 1 = +x bond, column 2 =
 *     +y bond), chain index slowest.  n = 2N is the BdG dimension.
 *   - there is no CPU fallback: every entry point needs a CUDA device.
 */
#ifndef DWHMC_H
#define DWHMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dwhmc_handle_s* dwhmc_handle;

enum {
  DWHMC_OK = 0,
  DWHMC_E_BADARG = 1,      /* invalid argument                                   */
  DWHMC_E_CUDA = 2,        /* CUDA runtime error (text in dwhmc_last_error)       */
  DWHMC_E_NOCONV = 3,      /* eigensolver did not converge (LAPACKException twin) */
  DWHMC_E_NODEVICE = 4,    /* no usable CUDA device                               */
  DWHMC_E_STATE = 5        /* call sequence error (e.g. commit without trajectory)*/
};

/* number of scalars measure_observables returns per chain
 * (ObservablesResult, src/Observables.jl:70-80, same field order) */
#define DWHMC_NOBS 9

/* ---- lifecycle ---------------------------------------------------------- */

/* ModelParameters constructor + initialize_cache (src/Types.jl:49-91,182-212):
 * B chains on an Lx x Ly lattice; nn_table / nnn_table exactly as the reference
 * builds them (Int64, N x 4, column-major, 1-based).  All chains share the
 * lattice; physics parameters are per chain (dwhmc_set_params).  E_n, U, fields
 * start at zero like the reference cache. */
int dwhmc_create(dwhmc_handle* out, int device, int B, int Lx, int Ly,
                 const int64_t* nn_table, const int64_t* nnn_table);
int dwhmc_destroy(dwhmc_handle h);
/* which eigensolver route this handle uses: *half_bandwidth = 0 for the dense route (blocked
 * tridiagonalisation), else the half-bandwidth of the BdG matrix in the folded site order (band route:
 * bulge chase + staircase block reflectors).  Same results either way; instrumentation only. */
int dwhmc_eigensolver_route(dwhmc_handle h, int* half_bandwidth);
/* last error text of this handle (or of the failed create when h == NULL) */
const char* dwhmc_last_error(dwhmc_handle h);
/* library / build information string */
const char* dwhmc_version(void);
/* B, N, n = 2N of a handle */
int dwhmc_dims(dwhmc_handle h, int* B, int* N, int* n);

/* ---- parameters and state ---------------------------------------------- */

/* ModelParameters fields t, tp, mu, beta, J, mass (src/Types.jl:21-33), one
 * value per chain (arrays of length B). */
int dwhmc_set_params(dwhmc_handle h, const double* t, const double* tp, const double* mu,
                     const double* beta, const double* J, const double* mass);
/* SimulationState.disorder_pot (src/Types.jl:104): double[N * B]. */
int dwhmc_set_disorder(dwhmc_handle h, const double* w);
int dwhmc_get_disorder(dwhmc_handle h, double* w);
/* SimulationState.Delta (src/Types.jl:111): complex[N * 2 * B]. */
int dwhmc_set_field(dwhmc_handle h, const double* delta);
int dwhmc_get_field(dwhmc_handle h, double* delta);
/* SimulationState.pi (src/Types.jl:115). */
int dwhmc_set_momentum(dwhmc_handle h, const double* pi);
int dwhmc_get_momentum(dwhmc_handle h, double* pi);
/* seed of the on-device Philox generator used when momenta / uniforms are not
 * injected (throughput mode; the reference uses Julia's unseeded task RNG). */
int dwhmc_seed(dwhmc_handle h, uint64_t seed);
/* initialize_state(p) src/Types.jl:118-134 for every chain, on the device (Philox stream of dwhmc_seed):
 * disorder = W[b] on round(N * n_imp[b]) distinct uniformly random sites (0 elsewhere), Delta0 with
 * Re, Im ~ U[-0.05, 0.05), pi = 0.  W, n_imp: double[B].  (Parity runs inject host-drawn state through
 * dwhmc_set_disorder / dwhmc_set_field instead.) */
int dwhmc_init_state(dwhmc_handle h, const double* W, const double* n_imp);

/* ---- per-operator entry points (1:1 with the reference operators) ------- */

/* init_static_H!(cache, p, state)      src/Hamiltonian.jl:10-47  */
int dwhmc_init_static_H(dwhmc_handle h);
/* update_H_BdG!(cache, p, state)       src/Hamiltonian.jl:55-86  */
int dwhmc_update_H(dwhmc_handle h);
/* diagonalize_H_BdG!(cache, p)         src/Hamiltonian.jl:96-114 (eigen! = zheevr)
 * Diagonalises the matrix as last written by init_static_H / update_H. */
int dwhmc_diagonalize(dwhmc_handle h);
/* compute_forces!(cache, p, state)     src/Observables.jl:14-62  */
int dwhmc_compute_forces(dwhmc_handle h);
/* compute_total_energy(cache, p, state) src/HMC.jl:12-41 ; out: double[B] */
int dwhmc_total_energy(dwhmc_handle h, double* out);
/* measure_observables(cache, p, state) src/Observables.jl:88-222 ; out: double[9 * B] */
int dwhmc_measure_observables(dwhmc_handle h, double* out);

/* measure_transport_and_spectra(cache, p) + build_current_operator!(cache, p)
 * src/Observables.jl:314-526, :237-283, for every chain, from the current (E_n, U) and the
 * fermi_factors left by the last compute_forces / measure_observables (as the reference, :321).
 * eta = p.eta; omega_grid = collect(p.omega_min : p.d_omega : p.omega_max) (double[n_omega]) and
 * dos_grid = collect(-p.omega_max : p.d_omega : p.omega_max) (double[n_dos]) are built by the caller.
 * Outputs (SpectrumResult, :293-311): scalars double[2 * B] = (superfluid_stiffness, dc_conductivity)
 * per chain; sigma double[n_omega * B]; dos, dos_AN double[n_dos * B]; A_k0 double[Lx * Ly * B]
 * (column-major Lx x Ly per chain).  Any output may be NULL. */
int dwhmc_measure_transport(dwhmc_handle h, double eta, const double* omega_grid, int n_omega,
                            const double* dos_grid, int n_dos, double* scalars, double* sigma,
                            double* dos, double* dos_AN, double* A_k0);

/* cache getters (debug / parity): H_base as the reference stores it (upper
 * triangle, lower = 0; complex[n * n * B]), E_n (double[n * B]), U
 * (complex[n * n * B]), forces (complex[N * 2 * B]), fermi_factors (double[n * B]). */
int dwhmc_get_H(dwhmc_handle h, double* out);
int dwhmc_get_eigenvalues(dwhmc_handle h, double* out);
int dwhmc_get_eigenvectors(dwhmc_handle h, double* out);
int dwhmc_get_forces(dwhmc_handle h, double* out);
int dwhmc_get_fermi(dwhmc_handle h, double* out);

/* ---- batched trajectory (hmc_sweep!, src/HMC.jl:71-144) ------------------ */

/* Lines :77-124 for every chain: momentum refresh (pi0 = complex[N*2*B] injected,
 * or NULL = on-device Philox N(0, m)), H_old, leapfrog with Nt[b] steps of size
 * dt[b], H_new, dH.  Leaves a pending proposal; the pre-trajectory state is kept.
 * H_old / H_new / dH: double[B] outputs, any may be NULL. */
int dwhmc_trajectory(dwhmc_handle h, const int32_t* Nt, const double* dt, const double* pi0,
                     double* H_old, double* H_new, double* dH);
/* Lines :128-138: accept[b] != 0 keeps the proposal, 0 restores Delta, E_n, U and
 * the pairing block of H.  Splitting trajectory/commit lets the host keep the
 * reference's lazy rand() (drawn only when dH >= 0). */
int dwhmc_commit(dwhmc_handle h, const int32_t* accept);
/* Whole hmc_sweep!: trajectory + Metropolis + commit.  uniforms: double[B]
 * injected u (used only where dH >= 0), or NULL = on-device Philox.  NaN dH
 * rejects, as `dH < 0 || rand() < exp(-dH)` does.  accepted: int32[B], dH: double[B]. */
int dwhmc_hmc_sweep(dwhmc_handle h, const int32_t* Nt, const double* dt, const double* pi0,
                    const double* uniforms, int32_t* accepted, double* dH);
/* n_sweeps whole sweeps back to back with on-device RNG and no host transfer
 * inside (throughput mode); optional outputs: accepted counts int32[B], last dH
 * double[B], observables after every sweep double[9 * B * n_sweeps] (or NULL). */
int dwhmc_run_sweeps(dwhmc_handle h, int n_sweeps, const int32_t* Nt, const double* dt,
                     int32_t* n_accepted, double* last_dH, double* obs);

/* ---- instrumentation ------------------------------------------------------ */

/* device time (ms, CUDA events on the handle's stream) spent in each stage since
 * the last reset: [0] assemble, [1] tridiagonalise, [2] tridiagonal D&C,
 * [3] back-transform, [4] force/energy/update kernels, [5] total eigensolves
 * (count), [6] kernel launches (count), [7] the trailing-matrix hemv kernels alone
 * (part of [1]).  out: double[8].  Stage times accumulate only while profiling is on. */
int dwhmc_get_timers(dwhmc_handle h, double* out);
int dwhmc_reset_timers(dwhmc_handle h);
/* device time (ms) of the sweeps of the last dwhmc_run_sweeps call, measured with CUDA events
 * recorded on the handle's stream around the enqueued work (excludes the final D2H copies) */
int dwhmc_last_elapsed_ms(dwhmc_handle h, double* out);
/* 0: off; 1: per-stage event timing (adds a stream synchronisation per stage); 2: additionally
 * time every hemv launch on its own (slot [7]; serialises the chain groups of the
 * tridiagonalisation, so stage times at level 2 are not representative) */
int dwhmc_set_profiling(dwhmc_handle h, int on);

/* stage-level entry points used by the parity tests of the eigensolver:
 * tridiagonalise the current matrix and return d (double[n*B]), e (double[(n-1)*B]);
 * solve a caller-supplied tridiagonal problem (d, e) -> w (double[n*B]), Z (double[n*n*B]). */
int dwhmc_debug_tridiagonalize(dwhmc_handle h, double* d, double* e);
int dwhmc_debug_stedc(dwhmc_handle h, const double* d, const double* e, double* w, double* Z);
/* diagonalise caller-supplied Hermitian matrices (complex[n*n*B], column-major; only the
 * lower triangle is read) -> E (double[n*B]), U (complex[n*n*B]); does not touch Delta, pi,
 * E_n or U of the chains (it uses the proposal buffers). */
int dwhmc_debug_heev(dwhmc_handle h, const double* A, double* E, double* U);

#ifdef __cplusplus
}
#endif
#endif /* DWHMC_H */
