"""CPU oracle for the DwaveHMC.jl molecular-dynamics force path.

TEST INFRASTRUCTURE ONLY.  This file is a NumPy/SciPy restatement of the
reference algorithm (reference paths are relative to /root/reference).  It may
be imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product path
(``libdwhmc.so`` and the ``dwhmc`` host package) never imports it and has no
CPU fallback.

PARITY UNPINNED: the reference ships no golden vectors, fixtures or seeded
tests for this path and Julia is not installed in this
image, so the reference itself cannot be executed.  The only numeric assertion
the reference holds on the path is ``scripts/bench_forces.jl:121-129`` (two
loop orders of the force contraction agree to 1e-10), restated below as
``bench_forces_orig`` / ``bench_forces_opt`` and checked in
``tests/test_oracle.py``.  Everything else is pinned by identities (finite
difference of the action, +-E symmetry, leapfrog reversibility ...).

Third-party arithmetic on the path (not vendored under /root/reference):
  * LAPACK ``zheevr`` through Julia ``eigen!(Hermitian(U,:U))``
    (src/Hamiltonian.jl:106; LinearAlgebra 1.11.0 -> libblastrampoline 5.11.0
    -> OpenBLAS_jll 0.3.27+1, Manifest.toml).  Restated here as
    ``scipy.linalg.eigh(H, lower=False, driver='evr')`` = the same routine.
  * LogExpFunctions 0.3.29 ``logistic`` (src/Observables.jl:27) and
    ``log1pexp`` (src/HMC.jl:25): restated below from their published
    piecewise definitions.

Layouts follow src/Types.jl: sites are 1-based in the reference; this file is
0-based everywhere, tables have shape (N, 4), fields have shape (N, 2) with
column 0 = +x bond, column 1 = +y bond.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla

# --------------------------------------------------------------------------
# LogExpFunctions 0.3.29 (Float64 branches)
# --------------------------------------------------------------------------
_LOGISTIC_LO = -744.4400719213812
_LOGISTIC_HI = 36.7368005696771


def logistic(x):
    """LogExpFunctions.logistic for Float64: e^x/(1+e^x), clamped to 0/1."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        e = np.exp(x)
        r = e / (1.0 + e)
    r = np.where(x < _LOGISTIC_LO, 0.0, np.where(x > _LOGISTIC_HI, 1.0, r))
    return r


def log1pexp(x):
    """LogExpFunctions.log1pexp for Float64 (piecewise, thresholds as upstream)."""
    x = np.asarray(x, dtype=np.float64)
    x0, x1, x2, x3 = -745.1332191019412, -36.7368005696771, 18.021826694558577, 33.23111882352963
    with np.errstate(over="ignore"):
        ex = np.exp(np.minimum(x, 700.0))
        out = np.where(
            x < x0, 0.0,
            np.where(x < x1, ex,
                     np.where(x < x2, np.log1p(ex),
                              np.where(x < x3, x + np.exp(-np.maximum(x, -700.0)), x))))
    return out


# --------------------------------------------------------------------------
# src/Types.jl
# --------------------------------------------------------------------------
@dataclass
class ModelParameters:
    """src/Types.jl:14-46 (struct) and :49-91 (constructor).  Tables 0-based."""
    Lx: int
    Ly: int
    t: float
    tp: float
    mu: float
    W: float
    n_imp: float
    beta: float
    J: float
    mass: float
    eta: float = 0.01
    d_omega: float = 0.002
    omega_max: float = 4.0
    N: int = field(init=False)
    nn_table: np.ndarray = field(init=False, repr=False)
    nnn_table: np.ndarray = field(init=False, repr=False)

    def __post_init__(self):
        self.N = self.Lx * self.Ly
        self.nn_table, self.nnn_table = neighbour_tables(self.Lx, self.Ly)


def neighbour_tables(Lx: int, Ly: int):
    """src/Types.jl:53-80.  site i = y*Lx + x (0-based); nn dirs {+x,+y,-x,-y};
    nnn dirs {+x+y, -x+y, -x-y, +x-y}; periodic boundaries."""
    N = Lx * Ly
    nn = np.zeros((N, 4), dtype=np.int64)
    nnn = np.zeros((N, 4), dtype=np.int64)

    def idx(x, y):
        return (y % Ly) * Lx + (x % Lx)

    for y in range(Ly):
        for x in range(Lx):
            i = idx(x, y)
            nn[i] = (idx(x + 1, y), idx(x, y + 1), idx(x - 1, y), idx(x, y - 1))
            nnn[i] = (idx(x + 1, y + 1), idx(x - 1, y + 1), idx(x - 1, y - 1), idx(x + 1, y - 1))
    return nn, nnn


@dataclass
class SimulationState:
    """src/Types.jl:101-116."""
    disorder_pot: np.ndarray   # float64 [N]
    Delta: np.ndarray          # complex128 [N, 2]
    pi: np.ndarray             # complex128 [N, 2]


def julia_round_half_even(x: float) -> int:
    """Julia round(Int, x) is ties-to-even (src/Types.jl:122)."""
    return int(np.rint(x))


def initialize_state(p: ModelParameters, rng: np.random.Generator) -> SimulationState:
    """src/Types.jl:118-134.  The reference is unseeded; the oracle takes an
    explicit NumPy generator.  Distributions: round(N*n_imp) distinct random
    sites get W; Re,Im(Delta) ~ U[-0.05, 0.05); pi = 0."""
    disorder = np.zeros(p.N, dtype=np.float64)
    n_imp_sites = julia_round_half_even(p.N * p.n_imp)
    imp = rng.permutation(p.N)[:n_imp_sites]
    disorder[imp] = p.W
    re = rng.random((p.N, 2))
    im = rng.random((p.N, 2))
    Delta = ((re - 0.5) + 1j * (im - 0.5)) * 0.1
    return SimulationState(disorder, Delta.astype(np.complex128), np.zeros((p.N, 2), np.complex128))


@dataclass
class ComputeCache:
    """src/Types.jl:145-180, hot-path members only."""
    H_base: np.ndarray
    E_n: np.ndarray
    U: np.ndarray
    forces: np.ndarray
    fermi_factors: np.ndarray
    Delta_backup: np.ndarray
    E_n_backup: np.ndarray
    U_backup: np.ndarray


def initialize_cache(p: ModelParameters) -> ComputeCache:
    """src/Types.jl:182-212 (transport/FFT members omitted: out of scope)."""
    dim = 2 * p.N
    z = lambda *s: np.zeros(s, dtype=np.complex128)
    return ComputeCache(z(dim, dim), np.zeros(dim), z(dim, dim), z(p.N, 2), np.zeros(dim),
                        z(p.N, 2), np.zeros(dim), z(dim, dim))


# --------------------------------------------------------------------------
# src/Hamiltonian.jl
# --------------------------------------------------------------------------
def init_static_H(cache: ComputeCache, p: ModelParameters, state: SimulationState) -> None:
    """src/Hamiltonian.jl:10-47.  Upper triangle only, assignments not sums."""
    N = p.N
    H = cache.H_base
    H[...] = 0.0
    for i in range(N):
        term = state.disorder_pot[i] - p.mu
        H[i, i] = term
        H[i + N, i + N] = -term
    for i in range(N):
        for d in range(4):
            j = p.nn_table[i, d]
            if j > i:
                H[i, j] = -p.t
                H[i + N, j + N] = p.t
        for d in range(4):
            j = p.nnn_table[i, d]
            if j > i:
                H[i, j] = -p.tp
                H[i + N, j + N] = p.tp


def update_H_BdG(cache: ComputeCache, p: ModelParameters, state: SimulationState) -> None:
    """src/Hamiltonian.jl:55-86.  Overwrite the 4N pairing entries (top-right block)."""
    N = p.N
    H = cache.H_base
    i = np.arange(N)
    jx = p.nn_table[:, 0]
    jy = p.nn_table[:, 1]
    vx = 0.5 * state.Delta[:, 0]
    vy = 0.5 * state.Delta[:, 1]
    # sequential order of the reference loop matters only if entries collide (L<3)
    for k in range(N):
        H[i[k], jx[k] + N] = vx[k]
        H[jx[k], i[k] + N] = vx[k]
        H[i[k], jy[k] + N] = vy[k]
        H[jy[k], i[k] + N] = vy[k]


def diagonalize_H_BdG(cache: ComputeCache, p: ModelParameters) -> None:
    """src/Hamiltonian.jl:96-114: eigen!(Hermitian(U,:U)) == LAPACK zheevr, upper
    triangle, all eigenpairs, ascending eigenvalues."""
    vals, vecs = sla.eigh(cache.H_base, lower=False, driver="evr", check_finite=False)
    cache.E_n[...] = vals
    cache.U[...] = vecs


def full_hermitian(cache: ComputeCache) -> np.ndarray:
    """The matrix Hermitian(H_base,:U) denotes (helper for tests)."""
    Hu = np.triu(cache.H_base)
    return Hu + np.triu(cache.H_base, 1).conj().T


# --------------------------------------------------------------------------
# src/Observables.jl:14-62
# --------------------------------------------------------------------------
def bond_correlators(cache: ComputeCache, p: ModelParameters):
    """P[i,dir] = -rho[i, j+N] - rho[j, i+N], rho = U diag(f) U^dagger
    (src/Observables.jl:24-53).  Also refreshes cache.fermi_factors."""
    N = p.N
    U, E = cache.U, cache.E_n
    f = logistic(-p.beta * E)
    cache.fermi_factors[...] = f
    Uf = U * f[None, :]
    P = np.empty((N, 2), dtype=np.complex128)
    ii = np.arange(N)
    for d in range(2):
        j = p.nn_table[:, d]
        rho1 = np.einsum("in,in->i", Uf[ii], U[j + N].conj())
        rho2 = np.einsum("in,in->i", Uf[j], U[ii + N].conj())
        P[:, d] = -rho1 - rho2
    return P


def compute_forces(cache: ComputeCache, p: ModelParameters, state: SimulationState) -> None:
    """src/Observables.jl:14-62: F = -(beta/2J) (Delta - J P)."""
    P = bond_correlators(cache, p)
    cache.forces[...] = -(p.beta / (2.0 * p.J)) * (state.Delta - p.J * P)


# --------------------------------------------------------------------------
# src/HMC.jl
# --------------------------------------------------------------------------
def fermion_energy(E_n: np.ndarray, beta: float) -> float:
    """src/HMC.jl:21-27: -sum_{E>0} [beta E + 2 log1pexp(-beta E)], sequential sum."""
    pos = E_n[E_n > 0]
    x = beta * pos
    terms = x + 2.0 * log1pexp(-x)
    acc = 0.0
    for v in terms:            # sequential accumulation as in the reference loop
        acc -= float(v)
    return acc


def compute_total_energy(cache: ComputeCache, p: ModelParameters, state: SimulationState) -> float:
    """src/HMC.jl:12-41."""
    E_f = fermion_energy(cache.E_n, p.beta)
    E_b = p.beta / (2.0 * p.J) * float(np.sum(np.abs(state.Delta) ** 2))
    E_k = 1.0 / (2.0 * p.mass) * float(np.sum(np.abs(state.pi) ** 2))
    return E_k + E_b + E_f


def refresh_momentum(state: SimulationState, p: ModelParameters, rng: np.random.Generator) -> None:
    """src/HMC.jl:51-61: randn!(pi) (Re,Im variance 1/2) then *= sqrt(2 m)."""
    n = state.pi.shape
    z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * math.sqrt(0.5)
    state.pi[...] = z * math.sqrt(2.0 * p.mass)


def hmc_sweep(cache: ComputeCache, p: ModelParameters, state: SimulationState, *, Nt: int, dt: float,
              pi0: np.ndarray | None = None, uniform=None, rng: np.random.Generator | None = None,
              return_energies: bool = False):
    """src/HMC.jl:71-144.  ``pi0`` injects the refreshed momentum (parity mode);
    ``uniform`` is a float or a zero-argument callable consumed lazily, exactly
    like ``rand()`` at :128 (only when dH >= 0)."""
    if pi0 is not None:
        state.pi[...] = pi0
    else:
        refresh_momentum(state, p, rng)
    H_old = compute_total_energy(cache, p, state)
    cache.Delta_backup[...] = state.Delta
    cache.E_n_backup[...] = cache.E_n
    cache.U_backup[...] = cache.U

    compute_forces(cache, p, state)
    state.pi += (0.5 * dt) * cache.forces
    coef_field = dt / (2.0 * p.mass)
    for step in range(1, Nt + 1):
        state.Delta += coef_field * state.pi
        update_H_BdG(cache, p, state)
        diagonalize_H_BdG(cache, p)
        compute_forces(cache, p, state)
        if step < Nt:
            state.pi += dt * cache.forces
    state.pi += (0.5 * dt) * cache.forces

    H_new = compute_total_energy(cache, p, state)
    dH = H_new - H_old
    accepted = False
    if dH < 0:
        accepted = True
    else:
        if callable(uniform):
            u = uniform()
        elif uniform is not None:
            u = float(uniform)
        else:
            u = rng.random()
        with np.errstate(over="ignore", invalid="ignore"):
            accepted = bool(u < np.exp(-dH))
    if not accepted:
        state.Delta[...] = cache.Delta_backup
        cache.E_n[...] = cache.E_n_backup
        cache.U[...] = cache.U_backup
        update_H_BdG(cache, p, state)
    if return_energies:
        return accepted, dH, H_old, H_new
    return accepted, dH


# --------------------------------------------------------------------------
# src/Observables.jl:70-222
# --------------------------------------------------------------------------
OBS_NAMES = ("total_energy", "D_amp", "D_local", "D_global", "S_D", "hole_conc",
             "D_diff", "D_pair", "D_localpair")


def measure_observables(cache: ComputeCache, p: ModelParameters, state: SimulationState) -> np.ndarray:
    """src/Observables.jl:88-222, returned in ObservablesResult field order (:70-80)."""
    N = p.N
    dx, dy = state.Delta[:, 0], state.Delta[:, 1]
    val_amp = float(np.sum(0.5 * (np.abs(dx) + np.abs(dy)))) / N
    val_local = float(np.sum(0.5 * np.abs(dx - dy))) / N
    sg = np.sum(0.5 * (dx - dy)) / N
    val_global = abs(sg)
    val_S = abs(sg) ** 2
    U, E = cache.U, cache.E_n
    pos = E > 0
    w = np.sum(np.abs(U[:N, :]) ** 2 - np.abs(U[N:, :]) ** 2, axis=0)
    val_hole = float(np.sum(w[pos] * np.tanh(0.5 * p.beta * E[pos]))) / N
    E_f = fermion_energy(E, p.beta)
    E_b = p.beta / (2.0 * p.J) * float(np.sum(np.abs(state.Delta) ** 2))
    total_energy = (E_f + E_b) / N
    P = bond_correlators(cache, p)
    Px, Py = P[:, 0], P[:, 1]
    val_diff = float(np.sum((np.abs(dx - p.J * Px) + np.abs(dy - p.J * Py)) / 2.0)) / N
    term = p.J * 0.5 * (Px - Py)
    val_pair = abs(np.sum(term) / N)
    val_localpair = float(np.sum(np.abs(term))) / N
    return np.array([total_energy, val_amp, val_local, val_global, val_S, val_hole,
                     val_diff, val_pair, val_localpair], dtype=np.float64)


# --------------------------------------------------------------------------
# src/Observables.jl:237-526  (transport and spectra)
# --------------------------------------------------------------------------
def julia_range(start: float, step: float, stop: float) -> np.ndarray:
    """collect(start:step:stop) for Float64 (length = floor((stop-start)/step + 1e-10) + 1; the
    reference's grids eta:d_omega:omega_max and -omega_max:d_omega:omega_max)."""
    n = int(math.floor((stop - start) / step + 1e-10)) + 1
    return start + step * np.arange(max(n, 0), dtype=np.float64)


def build_current_operator(p: ModelParameters) -> np.ndarray:
    """src/Observables.jl:237-283 as a dense 2N x 2N matrix: blockdiag(Jx, Jx),
    Jx[i, i+x] += i t, Jx[i+x, i] += -i t, same with t' for i+x+y and i+x-y (duplicates add, as sparse())."""
    N = p.N
    Jp = np.zeros((N, N), dtype=np.complex128)
    for i in range(N):
        for j, val in ((p.nn_table[i, 0], 1j * p.t), (p.nnn_table[i, 0], 1j * p.tp), (p.nnn_table[i, 3], 1j * p.tp)):
            Jp[i, j] += val
            Jp[j, i] += np.conj(val)
    J = np.zeros((2 * N, 2 * N), dtype=np.complex128)
    J[:N, :N] = Jp
    J[N:, N:] = Jp
    return J


def lorentzian(x, eta):
    return (1.0 / math.pi) * (eta / (x * x + eta * eta))


@dataclass
class SpectrumResult:
    """src/Observables.jl:293-311."""
    superfluid_stiffness: float
    dc_conductivity: float
    omega_grid: np.ndarray
    optical_conductivity: np.ndarray
    dos_omega_grid: np.ndarray
    dos: np.ndarray
    dos_AN: np.ndarray
    A_k_w0: np.ndarray


def measure_transport_and_spectra(cache: ComputeCache, p: ModelParameters) -> SpectrumResult:
    """src/Observables.jl:314-526.  Requires cache.fermi_factors current (the reference reads
    cache.fermi_factors as left by the last compute_forces!/measure_observables call)."""
    N, dim, beta = p.N, 2 * p.N, p.beta
    U, E, f = cache.U, cache.E_n, cache.fermi_factors
    Jx = build_current_operator(p)
    J_mn = U.conj().T @ (Jx @ U)                                     # :334-335
    J2 = np.abs(J_mn) ** 2
    # B. stiffness (:346-387)
    jx, jxpy, jxmy = p.nn_table[:, 0], p.nnn_table[:, 0], p.nnn_table[:, 3]
    Up, Un = U[:N], U[N:]
    val_dia = 0.0
    for n in range(dim):
        if E[n] > 0:
            u, v = Up[:, n], Un[:, n]
            w_n = 0.0
            for nb, hop in ((jx, p.t), (jxpy, p.tp), (jxmy, p.tp)):
                w_n += hop * 2.0 * float(np.sum((v * np.conj(v[nb]) - np.conj(u) * u[nb]).real))
            val_dia += w_n * math.tanh(0.5 * beta * E[n]) / N
    dE = E[None, :] - E[:, None]                                     # [n, m] = E[m] - E[n]
    df = f[:, None] - f[None, :]                                     # f[n] - f[m]
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(np.abs(dE) < 1e-8, (beta * f * (1.0 - f))[:, None] * np.ones_like(dE), df / dE)
    Lambda_xx = float(np.sum(ratio * J2)) / N
    stiffness = val_dia - Lambda_xx
    # C. conductivities (:398-425)
    eta = p.eta
    omega_grid = julia_range(p.eta, p.d_omega, p.omega_max)
    dc = float(np.sum((beta * f * (1.0 - f))[:, None] * J2 * lorentzian(dE, eta))) * (math.pi / N)
    sigma = np.zeros(len(omega_grid))
    wgt = np.where(np.abs(df) < 1e-12, 0.0, df) * J2
    sel = wgt != 0.0
    wsel, dsel = wgt[sel], dE[sel]
    for iw, w in enumerate(omega_grid):
        sigma[iw] = float(np.sum((wsel / w) * lorentzian(w - dsel, eta)))
    sigma *= math.pi / N
    # D. DOS, antinodal DOS, A(k, 0) (:431-519)
    dos_grid = julia_range(-p.omega_max, p.d_omega, p.omega_max)
    w_n = np.sum(np.abs(Up) ** 2, axis=0)
    xs = np.arange(N) % p.Lx + 1                                     # 1-based x, y as in mod1 / cld
    ys = np.arange(N) // p.Lx + 1
    sx = np.where(xs % 2 == 0, 1.0, -1.0)
    sy = np.where(ys % 2 == 0, 1.0, -1.0)
    w_AN = 0.5 * (np.abs(sx @ Up) ** 2 + np.abs(sy @ Up) ** 2) / N
    L = lorentzian(dos_grid[:, None] - E[None, :], eta)
    dos = (L @ w_n) / N
    dos_AN = L @ w_AN
    w0 = lorentzian(0.0 - E, eta)
    ak = np.zeros((p.Lx, p.Ly))
    for n in range(dim):
        if w0[n] > 1e-6:
            ur = Up[:, n].reshape(p.Ly, p.Lx).T                      # u_r[x, y], i = (y-1) Lx + x
            ak += np.abs(np.fft.fft2(ur)) ** 2 * w0[n]               # FFTW forward = exp(-2 pi i ...)
    ak /= N
    return SpectrumResult(stiffness, dc, omega_grid, sigma, dos_grid, dos, dos_AN, ak)


# --------------------------------------------------------------------------
# src/Simulation.jl:11-14
# --------------------------------------------------------------------------
def calc_optimal_dt(beta: float, J: float, mass: float, Nt: int) -> float:
    T = 2.0 * math.pi * math.sqrt(mass * J / beta)
    return T / (2 * Nt)


# --------------------------------------------------------------------------
# scripts/bench_forces.jl:36-110 -- the reference's only numeric KAT on the path
# --------------------------------------------------------------------------
def bench_forces_orig(U, f, nn2, J, beta_term, Delta):
    """scripts/bench_forces.jl:36-56 (bond-outer, eigenstate-inner)."""
    N = Delta.shape[0]
    F = np.zeros((N, 2), dtype=np.complex128)
    for i in range(N):
        for d in range(2):
            j = nn2[i, d]
            rho1 = np.sum(U[i, :] * f * np.conj(U[j + N, :]))
            rho2 = np.sum(U[j, :] * f * np.conj(U[i + N, :]))
            P = -rho1 - rho2
            F[i, d] = -beta_term * (Delta[i, d] - J * P)
    return F


def bench_forces_opt(U, f, nn2, J, beta_term, Delta):
    """scripts/bench_forces.jl:59-110 (eigenstate-outer loop order)."""
    N = Delta.shape[0]
    F = -beta_term * Delta.astype(np.complex128)
    c = beta_term * J
    for n in range(2 * N):
        fn = f[n]
        up = U[:N, n]
        dn = U[N:, n]
        for d in range(2):
            j = nn2[:, d]
            F[:, d] -= c * (up * fn * np.conj(dn[j]) + up[j] * fn * np.conj(dn))
    return F


# --------------------------------------------------------------------------
# helpers shared by tests / bench (not in the reference)
# --------------------------------------------------------------------------
def action(p: ModelParameters, disorder: np.ndarray, Delta: np.ndarray) -> float:
    """S = E_boson + E_fermion for the given field (used by the FD identity)."""
    st = SimulationState(disorder.copy(), Delta.copy(), np.zeros_like(Delta))
    c = initialize_cache(p)
    init_static_H(c, p, st)
    update_H_BdG(c, p, st)
    E = sla.eigvalsh(c.H_base, lower=False)
    return fermion_energy(E, p.beta) + p.beta / (2.0 * p.J) * float(np.sum(np.abs(Delta) ** 2))


def make_chain(p: ModelParameters, seed: int):
    """Seeded chain used for fixtures and parity tests: PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    st = initialize_state(p, rng)
    c = initialize_cache(p)
    init_static_H(c, p, st)
    update_H_BdG(c, p, st)
    diagonalize_H_BdG(c, p)
    return rng, st, c


def draw_momentum(p: ModelParameters, rng: np.random.Generator) -> np.ndarray:
    z = (rng.standard_normal((p.N, 2)) + 1j * rng.standard_normal((p.N, 2))) * math.sqrt(0.5)
    return z * math.sqrt(2.0 * p.mass)
