"""Single-chain operator API with the reference's names, argument order and mutation semantics
(src/DwaveHMC.jl:3-9 export list), backed by libdwhmc through a B = 1 ChainBatch.  The parity
tests are written against this module so that they read like calls into DwaveHMC.jl:

    p = ModelParameters(Lx, Ly, t, tp, mu, W, n_imp, beta, J, mass)
    state = initialize_state(p, rng); cache = initialize_cache(p)
    init_static_H(cache, p, state); update_H_BdG(cache, p, state); diagonalize_H_BdG(cache, p)
    accepted, dH = hmc_sweep(cache, p, state, Nt=6, dt=dt, rng=rng)

Differences forced by the host language: no `!` in names; 0-based neighbour tables are exposed as
``p.nn0`` while ``p.nn_table`` keeps the reference's 1-based Int64 values; the RNG is an explicit
NumPy generator (the reference uses Julia's unseeded task RNG)."""
from __future__ import annotations

import math
from collections import namedtuple
from dataclasses import dataclass, field

import numpy as np

from .batch import ChainBatch, calc_optimal_dt, neighbour_tables, OBS_NAMES  # noqa: F401

ObservablesResult = namedtuple("ObservablesResult", OBS_NAMES)   # src/Observables.jl:70-80
SpectrumResult = namedtuple("SpectrumResult", ("superfluid_stiffness", "dc_conductivity", "omega_grid",
                                               "optical_conductivity", "dos_omega_grid", "dos", "dos_AN",
                                               "A_k_w0"))                 # src/Observables.jl:293-311


@dataclass
class ModelParameters:
    """src/Types.jl:14-46; positional order of the constructor at :49-50."""
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
        self.omega_min = self.eta
        self.n_omega = int(math.floor((self.omega_max - self.omega_min) / self.d_omega)) + 1

    @property
    def nn0(self):
        return np.ascontiguousarray(self.nn_table - 1)


@dataclass
class SimulationState:
    """src/Types.jl:101-116: disorder_pot Float64[N]; Delta, pi ComplexF64[N, 2]."""
    disorder_pot: np.ndarray
    Delta: np.ndarray
    pi: np.ndarray


def initialize_state(p: ModelParameters, rng: np.random.Generator) -> SimulationState:
    """src/Types.jl:118-134 (same distributions; explicit generator)."""
    disorder = np.zeros(p.N)
    n_imp_sites = int(np.rint(p.N * p.n_imp))            # Julia round = ties to even
    disorder[rng.permutation(p.N)[:n_imp_sites]] = p.W
    re, im = rng.random((p.N, 2)), rng.random((p.N, 2))
    Delta = ((re - 0.5) + 1j * (im - 0.5)) * 0.1
    return SimulationState(disorder, Delta.astype(np.complex128), np.zeros((p.N, 2), np.complex128))


class ComputeCache:
    """src/Types.jl:145-180, hot-path members.  The arrays live on the GPU; the attributes below
    fetch them (H_base, E_n, U, forces, fermi_factors in the reference's orientation)."""

    def __init__(self, p: ModelParameters, device: int = 0):
        self.batch = ChainBatch(1, p.Lx, p.Ly, device=device, nn_table=p.nn_table, nnn_table=p.nnn_table)
        self._p = None

    def _sync_params(self, p: ModelParameters):
        key = (p.t, p.tp, p.mu, p.beta, p.J, p.mass)
        if key != self._p:
            self.batch.set_params(*key)
            self._p = key

    @property
    def H_base(self):
        return self.batch.get_H()[0].T.copy()

    @property
    def E_n(self):
        return self.batch.get_eigenvalues()[0]

    @property
    def U(self):
        return self.batch.get_eigenvectors()[0].T.copy()

    @property
    def forces(self):
        return self.batch.get_forces()[0].T.copy()

    @property
    def fermi_factors(self):
        return self.batch.get_fermi()[0]


def initialize_cache(p: ModelParameters, device: int = 0) -> ComputeCache:
    """src/Types.jl:182-212."""
    return ComputeCache(p, device)


def _push_field(cache, state):
    cache.batch.set_field(state.Delta[None])


def init_static_H(cache: ComputeCache, p: ModelParameters, state: SimulationState) -> None:
    """init_static_H!  src/Hamiltonian.jl:10-47."""
    cache._sync_params(p)
    cache.batch.set_disorder(state.disorder_pot[None])
    cache.batch.init_static_H()


def update_H_BdG(cache: ComputeCache, p: ModelParameters, state: SimulationState) -> None:
    """update_H_BdG!  src/Hamiltonian.jl:55-86."""
    _push_field(cache, state)
    cache.batch.update_H_BdG()


def diagonalize_H_BdG(cache: ComputeCache, p: ModelParameters) -> None:
    """diagonalize_H_BdG!  src/Hamiltonian.jl:96-114; raises EigenConvergenceError like LAPACKException."""
    cache.batch.diagonalize_H_BdG()


def compute_forces(cache: ComputeCache, p: ModelParameters, state: SimulationState) -> None:
    """compute_forces!  src/Observables.jl:14-62."""
    cache._sync_params(p)
    _push_field(cache, state)
    cache.batch.compute_forces()


def compute_total_energy(cache: ComputeCache, p: ModelParameters, state: SimulationState) -> float:
    """src/HMC.jl:12-41."""
    cache._sync_params(p)
    _push_field(cache, state)
    cache.batch.set_momentum(state.pi[None])
    return float(cache.batch.compute_total_energy()[0])


def refresh_momentum(state: SimulationState, p: ModelParameters, rng: np.random.Generator) -> None:
    """refresh_momentum!  src/HMC.jl:51-61."""
    z = (rng.standard_normal((p.N, 2)) + 1j * rng.standard_normal((p.N, 2))) * math.sqrt(0.5)
    state.pi[...] = z * math.sqrt(2.0 * p.mass)


def hmc_sweep(cache: ComputeCache, p: ModelParameters, state: SimulationState, *, Nt: int, dt: float,
              rng: np.random.Generator | None = None, pi0=None, uniform=None):
    """hmc_sweep!  src/HMC.jl:71-144.  Both keywords are required, as in the reference.  The
    uniform deviate is drawn lazily, only when dH >= 0 (:128)."""
    cache._sync_params(p)
    if pi0 is not None:
        state.pi[...] = pi0
    else:
        refresh_momentum(state, p, rng)
    _push_field(cache, state)
    _, _, dH = cache.batch.trajectory(Nt, dt, pi0=state.pi[None])
    dH = float(dH[0])
    if dH < 0:
        accepted = True
    else:
        u = uniform() if callable(uniform) else (float(uniform) if uniform is not None else rng.random())
        with np.errstate(over="ignore", invalid="ignore"):
            accepted = bool(u < np.exp(-dH))
    cache.batch.commit([int(accepted)])
    state.Delta[...] = cache.batch.get_field()[0].T
    state.pi[...] = cache.batch.get_momentum()[0].T
    return accepted, dH


def measure_observables(cache: ComputeCache, p: ModelParameters, state: SimulationState) -> ObservablesResult:
    """src/Observables.jl:88-222."""
    cache._sync_params(p)
    _push_field(cache, state)
    return ObservablesResult(*cache.batch.measure_observables()[0])


def build_current_operator(cache: ComputeCache, p: ModelParameters) -> None:
    """src/Observables.jl:237-283.  The current operator lives in the kernels (a 6-point gather with
    the handle's neighbour tables and the chain's t, t'); nothing to precompute on the host."""
    cache._sync_params(p)


def measure_transport_and_spectra(cache: ComputeCache, p: ModelParameters) -> SpectrumResult:
    """src/Observables.jl:314-526: uses cache.E_n, cache.U and the fermi_factors of the last
    compute_forces / measure_observables call, like the reference (:321)."""
    cache._sync_params(p)
    r = cache.batch.measure_transport_and_spectra(p.eta, p.d_omega, p.omega_max)
    return SpectrumResult(float(r["superfluid_stiffness"][0]), float(r["dc_conductivity"][0]), r["omega_grid"],
                          r["optical_conductivity"][0], r["dos_omega_grid"], r["dos"][0], r["dos_AN"][0], r["A_k_w0"][0])

