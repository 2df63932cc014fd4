"""ChainBatch: B independent HMC chains (temperature point x disorder seed x chain) on one GPU,
advanced in lock-step by libdwhmc.  This is the batched form of the loop body of
scripts/batch_scan_T.jl:54-74 / src/Simulation.jl:104-228 of the reference; each method is the
batched twin of the reference operator it names.

Array conventions (NumPy, C order): fields / momenta / forces have shape (B, 2, N) -- per chain
the reference's N x 2 column-major matrix; eigenvectors and H have shape (B, n, n) indexed
[chain, column, row], i.e. ``U_b = arr[b].T``."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import check, dptr, iptr, lib

OBS_NAMES = ("total_energy", "Delta_amp", "Delta_local", "Delta_global", "S_Delta", "hole_concentration",
             "Delta_diff", "Delta_pair", "Delta_localpair")


def neighbour_tables(Lx: int, Ly: int):
    """nn_table / nnn_table exactly as ModelParameters builds them (src/Types.jl:53-80):
    Int64, shape (N, 4), 1-based; returned in Fortran (column-major) order like Julia's Matrix{Int}."""
    N = Lx * Ly
    nn = np.zeros((N, 4), dtype=np.int64, order="F")
    nnn = np.zeros((N, 4), dtype=np.int64, order="F")

    def idx(x, y):   # 1-based x, y with mod1
        return ((y - 1) % Ly) * Lx + ((x - 1) % Lx) + 1

    for y in range(1, Ly + 1):
        for x in range(1, Lx + 1):
            i = idx(x, y) - 1
            nn[i] = (idx(x + 1, y), idx(x, y + 1), idx(x - 1, y), idx(x, y - 1))
            nnn[i] = (idx(x + 1, y + 1), idx(x - 1, y + 1), idx(x - 1, y - 1), idx(x + 1, y - 1))
    return nn, nnn


def calc_optimal_dt(beta: float, J: float, mass: float, Nt: int) -> float:
    """src/Simulation.jl:11-14."""
    T = 2.0 * math.pi * math.sqrt(mass * J / beta)
    return T / (2 * Nt)


def julia_range(start: float, step: float, stop: float) -> np.ndarray:
    """collect(start:step:stop) for Float64 steps (the grids of src/Observables.jl:402, :433)."""
    nsteps = int(math.floor((stop - start) / step + 1e-10)) + 1
    return np.ascontiguousarray(start + step * np.arange(max(nsteps, 0), dtype=np.float64))


def _vec(x, B):
    a = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (B,)))
    return a


class ChainBatch:
    def __init__(self, B: int, Lx: int, Ly: int, *, device: int = 0, nn_table=None, nnn_table=None):
        if nn_table is None or nnn_table is None:
            nn_table, nnn_table = neighbour_tables(Lx, Ly)
        nn = np.asfortranarray(nn_table, dtype=np.int64)
        nnn = np.asfortranarray(nnn_table, dtype=np.int64)
        self._h = C.c_void_p()
        i64p = C.POINTER(C.c_int64)
        rc = lib.dwhmc_create(C.byref(self._h), device, B, Lx, Ly, nn.ctypes.data_as(i64p), nnn.ctypes.data_as(i64p))
        if rc != _lib.OK:
            self._h = None
            check(rc, None)
        self.B, self.Lx, self.Ly, self.N, self.n = B, Lx, Ly, Lx * Ly, 2 * Lx * Ly
        self.device = device

    # ---- lifecycle
    def close(self):
        if getattr(self, "_h", None):
            lib.dwhmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- parameters / state
    def set_params(self, t, tp, mu, beta, J, mass):
        B = self.B
        self.t, self.tp, self.mu, self.beta, self.J, self.mass = (_vec(v, B) for v in (t, tp, mu, beta, J, mass))
        check(lib.dwhmc_set_params(self._h, dptr(self.t), dptr(self.tp), dptr(self.mu), dptr(self.beta), dptr(self.J),
                                   dptr(self.mass)), self._h)

    def set_disorder(self, w):
        w = np.ascontiguousarray(np.broadcast_to(np.asarray(w, dtype=np.float64), (self.B, self.N)))
        check(lib.dwhmc_set_disorder(self._h, dptr(w)), self._h)

    def _field_in(self, a):
        a = np.asarray(a, dtype=np.complex128)
        if a.shape == (self.B, self.N, 2):       # reference orientation per chain -> (B, 2, N)
            a = a.transpose(0, 2, 1)
        assert a.shape == (self.B, 2, self.N), a.shape
        return np.ascontiguousarray(a)

    def set_field(self, delta):
        check(lib.dwhmc_set_field(self._h, dptr(self._field_in(delta))), self._h)

    def get_field(self):
        out = np.empty((self.B, 2, self.N), dtype=np.complex128)
        check(lib.dwhmc_get_field(self._h, dptr(out)), self._h)
        return out

    def set_momentum(self, pi):
        check(lib.dwhmc_set_momentum(self._h, dptr(self._field_in(pi))), self._h)

    def get_momentum(self):
        out = np.empty((self.B, 2, self.N), dtype=np.complex128)
        check(lib.dwhmc_get_momentum(self._h, dptr(out)), self._h)
        return out

    def seed(self, seed: int):
        check(lib.dwhmc_seed(self._h, C.c_uint64(seed)), self._h)

    def init_state(self, W, n_imp):
        """initialize_state (src/Types.jl:118-134) for every chain on the device (Philox stream of seed())."""
        check(lib.dwhmc_init_state(self._h, dptr(_vec(W, self.B)), dptr(_vec(n_imp, self.B))), self._h)

    def get_disorder(self):
        out = np.empty((self.B, self.N))
        check(lib.dwhmc_get_disorder(self._h, dptr(out)), self._h)
        return out

    # ---- operators (1:1 with the reference)
    def init_static_H(self):
        check(lib.dwhmc_init_static_H(self._h), self._h)

    def update_H_BdG(self):
        check(lib.dwhmc_update_H(self._h), self._h)

    def diagonalize_H_BdG(self):
        check(lib.dwhmc_diagonalize(self._h), self._h)

    def compute_forces(self):
        check(lib.dwhmc_compute_forces(self._h), self._h)

    def compute_total_energy(self):
        out = np.empty(self.B)
        check(lib.dwhmc_total_energy(self._h, dptr(out)), self._h)
        return out

    def measure_observables(self):
        out = np.empty((self.B, _lib.NOBS))
        check(lib.dwhmc_measure_observables(self._h, dptr(out)), self._h)
        return out

    def measure_transport_and_spectra(self, eta: float, d_omega: float, omega_max: float):
        """measure_transport_and_spectra (src/Observables.jl:314-526) for every chain.  Returns a dict with
        superfluid_stiffness[B], dc_conductivity[B], omega_grid[n_w], optical_conductivity[B, n_w],
        dos_omega_grid[n_d], dos[B, n_d], dos_AN[B, n_d], A_k_w0[B, Lx, Ly]."""
        og = julia_range(eta, d_omega, omega_max)
        dg = julia_range(-omega_max, d_omega, omega_max)
        B = self.B
        scal = np.empty((B, 2)); sig = np.empty((B, len(og))); dos = np.empty((B, len(dg))); dan = np.empty((B, len(dg)))
        ak = np.empty((B, self.Ly, self.Lx))
        check(lib.dwhmc_measure_transport(self._h, float(eta), dptr(og), len(og), dptr(dg), len(dg), dptr(scal), dptr(sig),
                                          dptr(dos), dptr(dan), dptr(ak)), self._h)
        return dict(superfluid_stiffness=scal[:, 0].copy(), dc_conductivity=scal[:, 1].copy(), omega_grid=og,
                    optical_conductivity=sig, dos_omega_grid=dg, dos=dos, dos_AN=dan,
                    A_k_w0=np.ascontiguousarray(ak.transpose(0, 2, 1)))

    # ---- cache getters
    def get_H(self):
        out = np.empty((self.B, self.n, self.n), dtype=np.complex128)
        check(lib.dwhmc_get_H(self._h, dptr(out)), self._h)
        return out

    def get_eigenvalues(self):
        out = np.empty((self.B, self.n))
        check(lib.dwhmc_get_eigenvalues(self._h, dptr(out)), self._h)
        return out

    def get_eigenvectors(self):
        out = np.empty((self.B, self.n, self.n), dtype=np.complex128)
        check(lib.dwhmc_get_eigenvectors(self._h, dptr(out)), self._h)
        return out

    def get_forces(self):
        out = np.empty((self.B, 2, self.N), dtype=np.complex128)
        check(lib.dwhmc_get_forces(self._h, dptr(out)), self._h)
        return out

    def get_fermi(self):
        out = np.empty((self.B, self.n))
        check(lib.dwhmc_get_fermi(self._h, dptr(out)), self._h)
        return out

    # ---- trajectories
    def _steps(self, Nt, dt):
        nt = np.ascontiguousarray(np.broadcast_to(np.asarray(Nt, dtype=np.int32), (self.B,)))
        dtv = _vec(dt, self.B)
        return nt, dtv

    def trajectory(self, Nt, dt, pi0=None):
        """src/HMC.jl:77-124 for every chain; returns (H_old, H_new, dH).  Leaves a pending proposal."""
        nt, dtv = self._steps(Nt, dt)
        p = None if pi0 is None else self._field_in(pi0)
        Ho, Hn, dH = np.empty(self.B), np.empty(self.B), np.empty(self.B)
        check(lib.dwhmc_trajectory(self._h, iptr(nt), dptr(dtv), dptr(p), dptr(Ho), dptr(Hn), dptr(dH)), self._h)
        return Ho, Hn, dH

    def commit(self, accept):
        a = np.ascontiguousarray(np.asarray(accept).astype(np.int32))
        check(lib.dwhmc_commit(self._h, iptr(a)), self._h)

    def hmc_sweep(self, Nt, dt, pi0=None, uniforms=None):
        """hmc_sweep! (src/HMC.jl:71-144) for every chain; returns (accepted[B] bool, dH[B])."""
        nt, dtv = self._steps(Nt, dt)
        p = None if pi0 is None else self._field_in(pi0)
        u = None if uniforms is None else _vec(uniforms, self.B)
        acc = np.empty(self.B, dtype=np.int32)
        dH = np.empty(self.B)
        check(lib.dwhmc_hmc_sweep(self._h, iptr(nt), dptr(dtv), dptr(p), dptr(u), iptr(acc), dptr(dH)), self._h)
        return acc.astype(bool), dH

    def run_sweeps(self, n_sweeps: int, Nt, dt, observables: bool = False):
        """n_sweeps sweeps with on-device RNG and no host transfer inside.  Returns
        (n_accepted[B], last_dH[B], obs[n_sweeps, B, 9] or None)."""
        nt, dtv = self._steps(Nt, dt)
        nacc = np.empty(self.B, dtype=np.int32)
        dH = np.empty(self.B)
        obs = np.empty((n_sweeps, self.B, _lib.NOBS)) if observables else None
        check(lib.dwhmc_run_sweeps(self._h, n_sweeps, iptr(nt), dptr(dtv), iptr(nacc), dptr(dH), dptr(obs)), self._h)
        return nacc, dH, obs

    def band_halfwidth(self) -> int:
        """0 = dense eigensolver route; otherwise the half-bandwidth the band route works with."""
        out = C.c_int(0)
        check(lib.dwhmc_eigensolver_route(self._h, C.byref(out)), self._h)
        return int(out.value)

    # ---- instrumentation
    def set_profiling(self, level):
        """0 off; 1 per-stage CUDA-event timers; 2 also times every hemv launch (serialises the chain groups)."""
        check(lib.dwhmc_set_profiling(self._h, int(level)), self._h)

    def reset_timers(self):
        check(lib.dwhmc_reset_timers(self._h), self._h)

    def last_elapsed_ms(self) -> float:
        """Device time of the last run_sweeps (CUDA events on the library's stream)."""
        out = np.zeros(1)
        check(lib.dwhmc_last_elapsed_ms(self._h, dptr(out)), self._h)
        return float(out[0])

    def timers(self):
        out = np.zeros(8)
        check(lib.dwhmc_get_timers(self._h, dptr(out)), self._h)
        keys = ("assemble_ms", "tridiagonalize_ms", "stedc_ms", "backtransform_ms", "force_ms", "eigensolves",
                "launches", "hemv_ms")
        return dict(zip(keys, out))

    # ---- eigensolver stage entry points (parity tests)
    def debug_tridiagonalize(self):
        d = np.empty((self.B, self.n))
        e = np.empty((self.B, self.n - 1))
        check(lib.dwhmc_debug_tridiagonalize(self._h, dptr(d), dptr(e)), self._h)
        return d, e

    def debug_stedc(self, d, e):
        d = np.ascontiguousarray(d, dtype=np.float64).reshape(self.B, self.n)
        e = np.ascontiguousarray(e, dtype=np.float64).reshape(self.B, self.n - 1)
        w = np.empty((self.B, self.n))
        Z = np.empty((self.B, self.n, self.n))
        check(lib.dwhmc_debug_stedc(self._h, dptr(d), dptr(e), dptr(w), dptr(Z)), self._h)
        return w, Z

    def debug_heev(self, A):
        """A: (B, n, n) full Hermitian matrices -> (E (B, n), U (B, n, n) indexed [b, column, row])."""
        A = np.ascontiguousarray(np.asarray(A, dtype=np.complex128).reshape(self.B, self.n, self.n))
        E = np.empty((self.B, self.n))
        U = np.empty((self.B, self.n, self.n), dtype=np.complex128)
        check(lib.dwhmc_debug_heev(self._h, dptr(A), dptr(E), dptr(U)), self._h)
        return E, U
