"""GPU parity tests: the CUDA path (through the C ABI, via the reference-shaped host API) against
the CPU oracle and the committed golden vectors.  Tolerances follow BASELINE.json's north star:
1e-10 relative in FP64 for energies, forces and dH (dH relative to |H_old|).
Eigenvectors are never compared column by column (phases / degenerate subspaces are arbitrary);
the residual, unitarity and every gauge-invariant quantity are."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import dwhmc_oracle as orc  # noqa: E402  (checker only)

RTOL = 1e-10
PHYS = dict(t=1.0, tp=-0.35, mu=-1.08, W=1.0, J=0.8, mass=1.0)
_ALL_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
GOLDEN = [p for p in _ALL_GOLDEN if not os.path.basename(p).startswith("transport_")]
TRANSPORT_GOLDEN = [p for p in _ALL_GOLDEN if os.path.basename(p).startswith("transport_")]


@pytest.fixture(scope="module")
def dw():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import dwhmc
    return dwhmc


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_vectors_single_chain(dw, path):
    """Reads like the reference's own call sequence (src/Simulation.jl:81-86, then hmc_sweep!)."""
    g = np.load(path)
    p = dw.ModelParameters(int(g["Lx"]), int(g["Ly"]), PHYS["t"], PHYS["tp"], PHYS["mu"], PHYS["W"],
                           float(g["n_imp"]), float(g["beta"]), PHYS["J"], PHYS["mass"])
    state = dw.SimulationState(g["disorder"].copy(), g["Delta0"].copy(), np.zeros_like(g["Delta0"]))
    cache = dw.initialize_cache(p)
    dw.init_static_H(cache, p, state)
    dw.update_H_BdG(cache, p, state)
    dw.diagonalize_H_BdG(cache, p)
    scale = np.max(np.abs(g["E0"]))
    assert np.max(np.abs(cache.E_n - g["E0"])) <= 1e-12 * scale
    dw.compute_forces(cache, p, state)
    assert rel(cache.forces, g["F0"]) <= RTOL
    obs = np.array(dw.measure_observables(cache, p, state))
    assert np.allclose(obs, g["obs0"], rtol=1e-9, atol=1e-11)
    for k in range(len(g["u"])):
        acc, dH = dw.hmc_sweep(cache, p, state, Nt=int(g["Nt"]), dt=float(g["dt"]), pi0=g["pi0"][k],
                               uniform=float(g["u"][k]))
        Hscale = max(abs(float(g["H_old"][k])), 1.0)
        assert abs(dH - g["dH"][k]) <= RTOL * Hscale
        assert acc == bool(g["accepted"][k])
        assert np.max(np.abs(state.Delta - g["Delta_end"][k])) <= 1e-10
        assert np.max(np.abs(state.pi - g["pi_end"][k])) <= 1e-9 * max(1.0, np.max(np.abs(g["pi_end"][k])))
        assert np.max(np.abs(cache.E_n - g["E_end"][k])) <= 1e-11 * scale
        obs = np.array(dw.measure_observables(cache, p, state))
        assert np.allclose(obs, g["obs_end"][k], rtol=1e-8, atol=1e-10)
    cache.batch.close()


def make_batch(dw, L, betas, n_imp, seed0, Nt=None):
    Lx, Ly = (L, L) if isinstance(L, int) else L
    B = len(betas)
    ps, sts, cs = [], [], []
    for b, beta in enumerate(betas):
        p = orc.ModelParameters(Lx, Ly, PHYS["t"], PHYS["tp"], PHYS["mu"], PHYS["W"], n_imp, float(beta), PHYS["J"],
                                PHYS["mass"])
        _, st, c = orc.make_chain(p, seed0 + b)
        ps.append(p); sts.append(st); cs.append(c)
    cb = dw.ChainBatch(B, Lx, Ly)
    cb.set_params(PHYS["t"], PHYS["tp"], PHYS["mu"], np.asarray(betas, float), PHYS["J"], PHYS["mass"])
    cb.set_disorder(np.stack([s.disorder_pot for s in sts]))
    cb.set_field(np.stack([s.Delta for s in sts]))
    cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
    return cb, ps, sts, cs


def test_H_base_bit_exact(dw):
    cb, ps, sts, cs = make_batch(dw, (6, 10), [2.0, 20.0], 0.05, 300)
    H = cb.get_H()
    for b in range(2):
        assert np.array_equal(H[b].T, cs[b].H_base)       # upper triangle, lower = 0, as the reference stores it
    cb.close()


def test_batched_sweeps_match_oracle_mixed_Nt_and_beta(dw):
    """Chains with different temperatures, disorder seeds and leapfrog step counts in one batch."""
    betas = [0.5, 5.0, 20.0, 100.0, 1000.0]
    cb, ps, sts, cs = make_batch(dw, 8, betas, 0.05, 400)
    B = len(betas)
    Nt = np.array([2, 6, 3, 5, 4], dtype=np.int32)
    dt = np.array([orc.calc_optimal_dt(p.beta, p.J, p.mass, int(k)) for p, k in zip(ps, Nt)]) * 0.5
    n_acc = 0
    for it in range(4):
        pi0 = np.stack([orc.draw_momentum(ps[b], np.random.default_rng(900 + 10 * it + b)) for b in range(B)])
        u = np.random.default_rng(990 + it).random(B)
        acc, dH = cb.hmc_sweep(Nt, dt, pi0=pi0, uniforms=u)
        D, Pi, E = cb.get_field(), cb.get_momentum(), cb.get_eigenvalues()
        for b in range(B):
            a_r, dH_r, Ho, Hn = orc.hmc_sweep(cs[b], ps[b], sts[b], Nt=int(Nt[b]), dt=float(dt[b]), pi0=pi0[b],
                                              uniform=float(u[b]), return_energies=True)
            assert abs(dH[b] - dH_r) <= RTOL * max(abs(Ho), 1.0), (it, b)
            assert bool(acc[b]) == a_r
            assert np.max(np.abs(D[b].T - sts[b].Delta)) <= 1e-10
            assert np.max(np.abs(Pi[b].T - sts[b].pi)) <= 1e-9 * max(1.0, np.max(np.abs(sts[b].pi)))
            assert np.max(np.abs(E[b] - cs[b].E_n)) <= 1e-11 * np.max(np.abs(cs[b].E_n))
            n_acc += int(a_r)
    assert 0 < n_acc < 4 * B            # both the accept and the restore branch were exercised
    O = cb.measure_observables()
    for b in range(B):
        assert np.allclose(O[b], orc.measure_observables(cs[b], ps[b], sts[b]), rtol=1e-8, atol=1e-10)
    cb.close()


def test_trajectory_commit_split_keeps_lazy_uniform(dw):
    cb, ps, sts, cs = make_batch(dw, 4, [5.0, 5.0], 0.0, 500)
    Nt, dt = 4, 0.05
    pi0 = np.stack([orc.draw_momentum(ps[b], np.random.default_rng(7 + b)) for b in range(2)])
    Ho, Hn, dH = cb.trajectory(Nt, dt, pi0=pi0)
    with pytest.raises(dw.DwhmcError):          # a second proposal before commit is a sequence error
        cb.trajectory(Nt, dt, pi0=pi0)
    D_prop = cb.get_field()
    cb.commit([1, 0])
    D = cb.get_field()
    assert np.array_equal(D[0], D_prop[0])                       # accepted: proposal kept
    assert np.array_equal(D[1].T, sts[1].Delta)                  # rejected: restored bit-exactly
    E = cb.get_eigenvalues()
    assert np.max(np.abs(E[1] - cs[1].E_n)) <= 1e-12 * np.max(np.abs(cs[1].E_n))
    for b in range(2):
        _, dH_r, Ho_r, Hn_r = orc.hmc_sweep(cs[b], ps[b], sts[b], Nt=Nt, dt=dt, pi0=pi0[b], uniform=0.0 if b == 0 else 1.0,
                                            return_energies=True)
        assert abs(Ho[b] - Ho_r) <= 1e-11 * max(abs(Ho_r), 1) and abs(dH[b] - dH_r) <= RTOL * max(abs(Ho_r), 1)
    with pytest.raises(dw.DwhmcError):
        cb.commit([1, 1])
    cb.close()


def test_force_is_gradient_of_action_config2_shape(dw):
    """BASELINE config 2: disordered L=16, 64 chains in one batch; dS/dRe(D) = -2 Re F,
    dS/dIm(D) = -2 Im F by central differences (h = 1e-5) on 16 random bond components."""
    B, L, h = 64, 16, 1e-5
    N = L * L
    rng = np.random.default_rng(16)
    betas = np.logspace(-1, 2, B)
    cb = dw.ChainBatch(B, L, L)
    cb.set_params(PHYS["t"], PHYS["tp"], PHYS["mu"], betas, PHYS["J"], PHYS["mass"])
    w = np.zeros((B, N))
    for b in range(B):
        w[b, rng.permutation(N)[:13]] = PHYS["W"]
    D0 = ((rng.random((B, 2, N)) - 0.5) + 1j * (rng.random((B, 2, N)) - 0.5)) * 0.1
    cb.set_disorder(w); cb.set_momentum(np.zeros((B, 2, N), complex))

    def action(D):
        cb.set_field(D); cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
        return cb.compute_total_energy()

    action(D0)
    cb.compute_forces()
    F = cb.get_forces()
    # chain 0 against the oracle as well
    p = orc.ModelParameters(L, L, PHYS["t"], PHYS["tp"], PHYS["mu"], PHYS["W"], 0.05, float(betas[0]), PHYS["J"], PHYS["mass"])
    st = orc.SimulationState(w[0], D0[0].T.copy(), np.zeros((N, 2), complex))
    c = orc.initialize_cache(p)
    orc.init_static_H(c, p, st); orc.update_H_BdG(c, p, st); orc.diagonalize_H_BdG(c, p); orc.compute_forces(c, p, st)
    assert rel(F[0].T, c.forces) <= RTOL
    for _ in range(16):
        d, i = int(rng.integers(2)), int(rng.integers(N))
        for part, comp in ((1.0, "real"), (1j, "imag")):
            Dp, Dm = D0.copy(), D0.copy()
            Dp[:, d, i] += part * h
            Dm[:, d, i] -= part * h
            fd = (action(Dp) - action(Dm)) / (2 * h)
            an = -2.0 * getattr(F[:, d, i], comp)
            assert np.all(np.abs(fd - an) <= 2e-6 * np.maximum(np.abs(an), 1.0)), (d, i, comp)
    cb.close()


def test_leapfrog_reversibility(dw):
    cb, ps, sts, cs = make_batch(dw, 8, [10.0, 50.0], 0.05, 600)
    D0 = cb.get_field()
    pi0 = np.stack([orc.draw_momentum(ps[b], np.random.default_rng(3 + b)) for b in range(2)])
    Nt, dt = 5, 0.02
    cb.trajectory(Nt, dt, pi0=pi0); cb.commit([1, 1])
    cb.trajectory(Nt, dt, pi0=-cb.get_momentum()); cb.commit([1, 1])
    assert np.max(np.abs(cb.get_field() - D0)) <= 1e-11
    assert np.max(np.abs(cb.get_momentum() + pi0.transpose(0, 2, 1))) <= 1e-10
    cb.close()


@pytest.mark.parametrize("L,B", [(16, 4), (24, 2)])
def test_eigensolver_properties_full_size(dw, L, B):
    """Size-independent properties at the benchmark sizes: residual, unitarity, +-E symmetry
    (particle-hole), trace, and agreement with LAPACK eigenvalues."""
    n = 2 * L * L
    cb, ps, sts, cs = make_batch(dw, L, np.logspace(0, 2, B), 0.05, 700)
    E, U = cb.get_eigenvalues(), cb.get_eigenvectors()
    for b in range(B):
        Hf = orc.full_hermitian(cs[b]); Ub = U[b].T
        nrm = np.max(np.abs(cs[b].E_n))
        assert np.max(np.abs(E[b] - cs[b].E_n)) <= 1e-12 * nrm
        assert np.max(np.abs(Hf @ Ub - Ub * E[b])) <= 1e-12 * nrm
        assert np.max(np.abs(Ub.conj().T @ Ub - np.eye(n))) <= 1e-12
        assert np.max(np.abs(E[b] + E[b][::-1])) <= 1e-12 * nrm
        assert abs(np.sum(E[b])) <= 1e-10 * nrm
        assert np.all(np.diff(E[b]) >= 0)
    cb.compute_forces()
    F = cb.get_forces()
    for b in range(B):
        orc.compute_forces(cs[b], ps[b], sts[b])
        assert rel(F[b].T, cs[b].forces) <= RTOL
    cb.close()


def test_degenerate_and_extreme_spectra(dw):
    """Clean lattice (massively degenerate), beta from 1e-3 to 1e5 (f(E) from flat to a step)."""
    betas = [1e-3, 1.0, 1e3, 1e5]
    cb, ps, sts, cs = make_batch(dw, 8, betas, 0.0, 800)
    cb.compute_forces()
    F = cb.get_forces()
    pi = np.zeros((4, 2, 64), complex)
    cb.set_momentum(pi)
    Eg = cb.compute_total_energy()
    for b in range(4):
        orc.compute_forces(cs[b], ps[b], sts[b])
        assert rel(F[b].T, cs[b].forces) <= RTOL
        Er = orc.compute_total_energy(cs[b], ps[b], sts[b])
        assert abs(Eg[b] - Er) <= 1e-12 * max(abs(Er), 1.0)
    # zero field: D = 0, the BdG matrix is block diagonal with exactly paired +-E
    cb.set_field(np.zeros((4, 2, 64), complex)); cb.update_H_BdG(); cb.diagonalize_H_BdG()
    E = cb.get_eigenvalues()
    assert np.max(np.abs(E + E[:, ::-1])) <= 1e-13 * np.max(np.abs(E))
    cb.close()


def test_particle_hole_shortcut_and_fallback(dw, monkeypatch):
    """The eigensolver back-transforms only the upper half of the spectrum and builds the partner
    columns (u, v) -> (-conj v, conj u) (tau_y H^* tau_y = -H).  (i) It must agree with the full
    back-transformation (DWHMC_PH=0); (ii) when more than one pair of levels sits at zero (here:
    mu = t' = 0, zero field: exact zero modes on the whole line kx + ky = pi) the chain falls back
    to the full path, and U is still a unitary eigenbasis with the oracle's correlators."""
    L, B = 8, 3
    N, n = L * L, 2 * L * L
    betas = np.array([2.0, 20.0, 200.0])
    rng = np.random.default_rng(77)
    w = np.zeros((B, N)); w[:, rng.permutation(N)[:3]] = 1.0
    delta = (rng.random((B, 2, N)) - 0.5 + 1j * (rng.random((B, 2, N)) - 0.5)) * 0.3
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("DWHMC_PH", mode)
        cb = dw.ChainBatch(B, L, L)
        cb.set_params(1.0, -0.35, -1.08, betas, 0.8, 1.0)
        cb.set_disorder(w); cb.set_field(delta)
        cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG(); cb.compute_forces()
        res[mode] = (cb.get_eigenvalues(), cb.get_eigenvectors(), cb.get_forces(), cb.measure_observables(),
                     cb.get_H())
        cb.close()
    monkeypatch.delenv("DWHMC_PH")
    E1, U1, F1, O1, H1 = res["1"]
    E0, U0, F0, O0, _ = res["0"]
    assert np.max(np.abs(E1 - E0)) <= 1e-13 * np.max(np.abs(E0))
    assert rel(F1, F0) <= 1e-12
    assert np.allclose(O1, O0, rtol=1e-11, atol=1e-13)
    for b in range(B):
        Hu = H1[b].T
        Hf = np.triu(Hu) + np.triu(Hu, 1).conj().T
        Ub = U1[b].T
        assert np.max(np.abs(Ub.conj().T @ Ub - np.eye(n))) <= 1e-12
        assert np.max(np.abs(Hf @ Ub - Ub * E1[b])) <= 1e-12 * np.max(np.abs(E1[b]))
        # partner structure of the shortcut: column k is C applied to column n-1-k
        assert np.max(np.abs(Ub[:N, :N] + Ub[N:, :N - 1:-1].conj())) <= 1e-15
    # fallback: massively degenerate zero modes
    ps, sts, cs = [], [], []
    for b, beta in enumerate(betas):
        p = orc.ModelParameters(L, L, 1.0, 0.0, 0.0, 1.0, 0.0, float(beta), 0.8, 1.0)
        _, st, c = orc.make_chain(p, 900 + b)
        st.Delta[...] = 0.0
        orc.update_H_BdG(c, p, st); orc.diagonalize_H_BdG(c, p); orc.compute_forces(c, p, st)
        ps.append(p); sts.append(st); cs.append(c)
    cb = dw.ChainBatch(B, L, L)
    cb.set_params(1.0, 0.0, 0.0, betas, 0.8, 1.0)
    cb.set_disorder(np.zeros((B, N))); cb.set_field(np.zeros((B, 2, N), complex))
    cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG(); cb.compute_forces()
    E, U, F = cb.get_eigenvalues(), cb.get_eigenvectors(), cb.get_forces()
    for b in range(B):
        Ub = U[b].T
        Hf = orc.full_hermitian(cs[b])
        assert np.sum(np.abs(E[b]) < 1e-12) >= 4          # the zero modes are there
        assert np.max(np.abs(Ub.conj().T @ Ub - np.eye(n))) <= 1e-12
        assert np.max(np.abs(Hf @ Ub - Ub * E[b])) <= 1e-12 * np.max(np.abs(E[b]))
        assert np.max(np.abs(F[b].T - cs[b].forces)) <= 1e-10 * max(np.max(np.abs(cs[b].forces)), 1.0)
    cb.close()


@pytest.mark.parametrize("L,B", [((6, 10), 3), ((10, 6), 2), (8, 150), ((5, 13), 2), (9, 2), (22, 2)])
def test_band_route_shapes(dw, monkeypatch, L, B):
    """Band route on rectangular lattices (the short ring is the fast index of the fold ordering either way) and on
    a batch larger than one cooperative launch of the chase kernel can hold (148 CTAs): spectrum and correlators
    equal the dense route's."""
    Lx, Ly = (L, L) if isinstance(L, int) else L
    N = Lx * Ly
    rng = np.random.default_rng(31)
    betas = np.linspace(2.0, 50.0, B)
    w = np.zeros((B, N)); w[:, rng.permutation(N)[:3]] = 1.0
    delta = (rng.random((B, 2, N)) - 0.5 + 1j * (rng.random((B, 2, N)) - 0.5)) * 0.2
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("DWHMC_BAND", mode)
        cb = dw.ChainBatch(B, Lx, Ly)
        cb.set_params(1.0, -0.35, -1.08, betas, 0.8, 1.0)
        cb.set_disorder(w); cb.set_field(delta)
        cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG(); cb.compute_forces()
        res[mode] = (cb.get_eigenvalues(), cb.get_forces(), cb.measure_observables())
        cb.close()
    monkeypatch.delenv("DWHMC_BAND")
    assert np.max(np.abs(res["1"][0] - res["0"][0])) <= 1e-12 * np.max(np.abs(res["0"][0]))
    assert rel(res["1"][1], res["0"][1]) <= 1e-11
    assert np.allclose(res["1"][2], res["0"][2], rtol=1e-10, atol=1e-12)


def test_band_route_matches_dense_route(dw, monkeypatch):
    """DWHMC_BAND=1: BdG matrix assembled straight into band storage (folded site order, half-bandwidth
    4L+4), bulge-chased to tridiagonal form, block-reflector back-transformation.  Same spectrum,
    forces, observables and trajectories as the dense route and the oracle."""
    L, B, Nt = 8, 3, 3
    betas = [2.0, 20.0, 200.0]
    monkeypatch.setenv("DWHMC_BAND", "1")
    cb, ps, sts, cs = make_batch(dw, L, betas, 0.05, 1300)
    monkeypatch.delenv("DWHMC_BAND")
    n = 2 * L * L
    E, U = cb.get_eigenvalues(), cb.get_eigenvectors()
    for b in range(B):
        Hf = orc.full_hermitian(cs[b]); Ub = U[b].T
        nrm = np.max(np.abs(cs[b].E_n))
        assert np.max(np.abs(E[b] - cs[b].E_n)) <= 1e-12 * nrm
        assert np.max(np.abs(Hf @ Ub - Ub * E[b])) <= 1e-12 * nrm
        assert np.max(np.abs(Ub.conj().T @ Ub - np.eye(n))) <= 1e-12
    d, e = cb.debug_tridiagonalize()
    import scipy.linalg as sl
    for b in range(B):
        wt = sl.eigh_tridiagonal(d[b], e[b], eigvals_only=True)
        assert np.max(np.abs(wt - cs[b].E_n)) <= 1e-12 * np.max(np.abs(cs[b].E_n))
    cb.compute_forces()
    F = cb.get_forces()
    dt = np.array([0.5 * orc.calc_optimal_dt(p.beta, p.J, p.mass, Nt) for p in ps])
    pi0 = np.stack([orc.draw_momentum(ps[b], np.random.default_rng(50 + b)) for b in range(B)])
    u = np.random.default_rng(51).random(B)
    acc, dH = cb.hmc_sweep(Nt, dt, pi0=pi0, uniforms=u)
    for b in range(B):
        orc.compute_forces(cs[b], ps[b], sts[b])
        assert rel(F[b].T, cs[b].forces) <= RTOL
        a_r, dH_r, Ho, _ = orc.hmc_sweep(cs[b], ps[b], sts[b], Nt=Nt, dt=float(dt[b]), pi0=pi0[b], uniform=float(u[b]),
                                         return_energies=True)
        assert abs(dH[b] - dH_r) <= RTOL * max(abs(Ho), 1.0)
        assert bool(acc[b]) == a_r
    cb.close()


@pytest.mark.parametrize("L,n_imp", [(8, 0.05), ((6, 10), 0.05), (8, 0.0)])
def test_transport_and_spectra_match_oracle(dw, L, n_imp):
    """measure_transport_and_spectra (src/Observables.jl:314-526) on the GPU against the oracle's restatement:
    stiffness, dc conductivity, Re sigma(omega), DOS, antinodal DOS and A(k, 0); eta = 8/N, d_omega = 0.2 eta as in
    scripts/batch_scan_T.jl:30-32.  Tolerance 1e-9 relative to the largest entry of each array."""
    betas = [2.0, 20.0, 200.0]
    cb, ps, sts, cs = make_batch(dw, L, betas, n_imp, 1500)
    N = ps[0].N
    eta = 8.0 / N
    cb.compute_forces()                                   # leaves fermi_factors, as in the reference's sweep loop
    r = cb.measure_transport_and_spectra(eta, 0.2 * eta, 4.0)
    for b in range(len(betas)):
        p = ps[b]
        p.eta, p.d_omega, p.omega_max = eta, 0.2 * eta, 4.0
        orc.compute_forces(cs[b], p, sts[b])
        ref = orc.measure_transport_and_spectra(cs[b], p)
        assert len(ref.omega_grid) == r["optical_conductivity"].shape[1]
        assert np.allclose(ref.omega_grid, r["omega_grid"], rtol=0, atol=1e-14)
        scale = max(abs(ref.superfluid_stiffness), 1e-3)
        assert abs(r["superfluid_stiffness"][b] - ref.superfluid_stiffness) <= 1e-9 * max(scale, 1.0)
        assert abs(r["dc_conductivity"][b] - ref.dc_conductivity) <= 1e-9 * max(abs(ref.dc_conductivity), 1e-6)
        for key, arr in (("optical_conductivity", ref.optical_conductivity), ("dos", ref.dos), ("dos_AN", ref.dos_AN),
                         ("A_k_w0", ref.A_k_w0)):
            got = r[key][b]
            assert got.shape == arr.shape, (key, got.shape, arr.shape)
            assert np.max(np.abs(got - arr)) <= 1e-9 * max(np.max(np.abs(arr)), 1e-12), key
    cb.close()


@pytest.mark.parametrize("path", TRANSPORT_GOLDEN, ids=[os.path.basename(p)[:-4] for p in TRANSPORT_GOLDEN])
def test_transport_golden_vectors(dw, path):
    """The committed transport / spectra vectors (tests/golden/make_golden.py) through the single-chain API."""
    g = np.load(path)
    p = dw.ModelParameters(int(g["Lx"]), int(g["Ly"]), PHYS["t"], PHYS["tp"], PHYS["mu"], PHYS["W"], float(g["n_imp"]),
                           float(g["beta"]), PHYS["J"], PHYS["mass"], eta=float(g["eta"]), d_omega=float(g["d_omega"]),
                           omega_max=float(g["omega_max"]))
    state = dw.SimulationState(g["disorder"].copy(), g["Delta0"].copy(), np.zeros_like(g["Delta0"]))
    cache = dw.initialize_cache(p)
    dw.init_static_H(cache, p, state); dw.update_H_BdG(cache, p, state); dw.diagonalize_H_BdG(cache, p)
    dw.measure_observables(cache, p, state)
    r = dw.measure_transport_and_spectra(cache, p)
    assert abs(r.superfluid_stiffness - float(g["stiffness"])) <= 1e-9
    assert abs(r.dc_conductivity - float(g["dc"])) <= 1e-9 * max(1.0, abs(float(g["dc"])))
    assert np.allclose(r.omega_grid, g["omega_grid"], rtol=0, atol=1e-14)
    for got, key in ((r.optical_conductivity, "sigma"), (r.dos, "dos"), (r.dos_AN, "dos_AN"), (r.A_k_w0, "A_k0")):
        assert np.max(np.abs(got - g[key])) <= 1e-9 * max(np.max(np.abs(g[key])), 1e-12), key
    cache.batch.close()


def test_transport_full_size_and_errors(dw):
    """BASELINE config 3 size (L = 24, n = 1152, 1436 frequencies): one chain against the oracle, the second chain
    through size-independent properties (sigma >= 0, DOS sum rule, A(k,0) Parseval); argument errors."""
    L = 24
    betas = [20.0, 1000.0]
    cb, ps, sts, cs = make_batch(dw, L, betas, 0.05, 1700)
    N = L * L
    eta = 8.0 / N
    cb.measure_observables()
    r = cb.measure_transport_and_spectra(eta, 0.2 * eta, 4.0)
    p = ps[0]
    p.eta, p.d_omega, p.omega_max = eta, 0.2 * eta, 4.0
    orc.measure_observables(cs[0], p, sts[0])
    ref = orc.measure_transport_and_spectra(cs[0], p)
    assert abs(r["superfluid_stiffness"][0] - ref.superfluid_stiffness) <= 1e-9
    assert abs(r["dc_conductivity"][0] - ref.dc_conductivity) <= 1e-9 * max(abs(ref.dc_conductivity), 1e-6)
    for key, arr in (("optical_conductivity", ref.optical_conductivity), ("dos", ref.dos), ("dos_AN", ref.dos_AN),
                     ("A_k_w0", ref.A_k_w0)):
        assert np.max(np.abs(r[key][0] - arr)) <= 1e-9 * max(np.max(np.abs(arr)), 1e-12), key
    E, U = cb.get_eigenvalues()[1], cb.get_eigenvectors()[1].T
    w_n = np.sum(np.abs(U[:N]) ** 2, axis=0)
    assert np.all(r["optical_conductivity"][1] >= -1e-12) and np.all(r["dos_AN"][1] >= 0)
    # DOS sum rule: the integral over the window equals the Lorentzian weight of every state inside it
    inside = (np.arctan((4.0 - E) / eta) - np.arctan((-4.0 - E) / eta)) / np.pi
    assert abs(np.sum(r["dos"][1]) * 0.2 * eta - np.sum(w_n * inside) / N) <= 2e-3
    w0 = orc.lorentzian(-E, eta)
    sel = w0 > 1e-6
    assert abs(np.sum(r["A_k_w0"][1]) - np.sum(w_n[sel] * w0[sel])) <= 1e-9 * max(np.sum(r["A_k_w0"][1]), 1.0)
    with pytest.raises(dw.DwhmcError):
        cb.measure_transport_and_spectra(0.0, 0.1, 4.0)                 # eta must be positive
    with pytest.raises(dw.DwhmcError):
        cb.init_state(1.0, 1.5)                                         # n_imp outside [0, 1]
    cb.trajectory(2, 0.01)
    with pytest.raises(dw.DwhmcError):
        cb.measure_transport_and_spectra(eta, 0.2 * eta, 4.0)           # proposal pending
    cb.commit(np.ones(2, int))
    cb.close()


def test_transport_single_chain_api(dw):
    """Reads like the reference: measure_observables, then measure_transport_and_spectra(cache, p)."""
    p = dw.ModelParameters(8, 8, 1.0, -0.35, -1.08, 1.0, 0.05, 20.0, 0.8, 1.0, eta=0.125, d_omega=0.025, omega_max=4.0)
    po = orc.ModelParameters(8, 8, 1.0, -0.35, -1.08, 1.0, 0.05, 20.0, 0.8, 1.0, eta=0.125, d_omega=0.025, omega_max=4.0)
    _, st, c = orc.make_chain(po, 1600)
    state = dw.SimulationState(st.disorder_pot.copy(), st.Delta.copy(), np.zeros_like(st.Delta))
    cache = dw.initialize_cache(p)
    dw.init_static_H(cache, p, state); dw.update_H_BdG(cache, p, state); dw.diagonalize_H_BdG(cache, p)
    dw.measure_observables(cache, p, state)
    dw.build_current_operator(cache, p)
    res = dw.measure_transport_and_spectra(cache, p)
    orc.measure_observables(c, po, st)
    ref = orc.measure_transport_and_spectra(c, po)
    assert abs(res.superfluid_stiffness - ref.superfluid_stiffness) <= 1e-9
    assert abs(res.dc_conductivity - ref.dc_conductivity) <= 1e-9 * max(abs(ref.dc_conductivity), 1e-6)
    assert np.max(np.abs(res.dos - ref.dos)) <= 1e-9 * np.max(np.abs(ref.dos))
    assert res.A_k_w0.shape == (8, 8)
    cache.batch.close()


def test_debug_stages(dw):
    import scipy.linalg as sl
    B, L = 2, 6
    n = 2 * L * L
    rng = np.random.default_rng(1)
    cb = dw.ChainBatch(B, L, L)
    A = rng.standard_normal((B, n, n)) + 1j * rng.standard_normal((B, n, n))
    A = A + A.conj().transpose(0, 2, 1)
    E, U = cb.debug_heev(A.transpose(0, 2, 1))
    for b in range(B):
        Ub = U[b].T
        assert np.max(np.abs(E[b] - np.linalg.eigvalsh(A[b]))) <= 1e-12 * np.max(np.abs(E[b]))
        assert np.max(np.abs(A[b] @ Ub - Ub * E[b])) <= 1e-12 * np.max(np.abs(E[b]))
    for name, d, e in (("wilkinson", np.abs(np.arange(n) - n // 2).astype(float), np.ones(n - 1)),
                       ("zeros", np.zeros(n), np.zeros(n - 1)),
                       ("split", rng.standard_normal(n), np.where(np.arange(n - 1) % 7 == 0, 0.0, 1.0)),
                       ("clustered", 1 + 1e-10 * rng.standard_normal(n), 1e-10 * rng.standard_normal(n - 1))):
        w, Z = cb.debug_stedc(np.tile(d, (B, 1)), np.tile(e, (B, 1)))
        wr = sl.eigh_tridiagonal(d, e, eigvals_only=True)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        for b in range(B):
            Zb = Z[b].T
            nrm = max(np.max(np.abs(wr)), 1e-300)
            assert np.max(np.abs(w[b] - wr)) <= 5e-14 * nrm, name
            assert np.max(np.abs(Zb.T @ Zb - np.eye(n))) <= 5e-14, name
            assert np.max(np.abs(T @ Zb - Zb * w[b])) <= 5e-14 * nrm, name
    cb.close()


def test_error_behaviour(dw):
    with pytest.raises(dw.DwhmcError):
        dw.ChainBatch(1, 2, 2)                       # L < 3: neighbours collide
    cb = dw.ChainBatch(1, 4, 4)
    with pytest.raises(dw.DwhmcError):
        cb.compute_total_energy()                    # parameters not set
    cb.set_params(1.0, -0.35, -1.08, 5.0, 0.8, 1.0)
    with pytest.raises(dw.DwhmcError):
        cb.hmc_sweep(0, 0.1)                         # Nt >= 1
    with pytest.raises(dw.DwhmcError):
        cb.commit([1])                               # nothing pending
    # NaN energy rejects (src/HMC.jl:128) and the restore branch runs
    cb.set_field(np.full((1, 2, 16), 0.01 + 0j)); cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
    pi0 = np.full((1, 2, 16), np.nan + 0j)
    try:
        acc, dH = cb.hmc_sweep(2, 0.1, pi0=pi0, uniforms=[0.0])
        assert not acc[0] and np.isnan(dH[0])
    except dw.EigenConvergenceError:
        pass                                         # LAPACKException twin is also acceptable for a NaN matrix
    cb.close()


def test_stale_cache_quirk(dw):
    """The reference lets callers run hmc_sweep! on a cache that was never diagonalised
    (scripts/benchmark_clean.jl:82-88): the first trajectory starts from E_n = 0, U = 0."""
    p = orc.ModelParameters(4, 4, PHYS["t"], PHYS["tp"], PHYS["mu"], PHYS["W"], 0.0, 5.0, PHYS["J"], PHYS["mass"])
    rng = np.random.Generator(np.random.PCG64(11))
    st = orc.initialize_state(p, rng)
    c = orc.initialize_cache(p)
    orc.init_static_H(c, p, st)
    pi0 = orc.draw_momentum(p, rng)
    cb = dw.ChainBatch(1, 4, 4)
    cb.set_params(p.t, p.tp, p.mu, p.beta, p.J, p.mass)
    cb.set_disorder(st.disorder_pot[None]); cb.set_field(st.Delta[None]); cb.init_static_H()
    acc, dH = cb.hmc_sweep(3, 0.1, pi0=pi0[None], uniforms=[0.5])
    a_r, dH_r = orc.hmc_sweep(c, p, st, Nt=3, dt=0.1, pi0=pi0, uniform=0.5)
    assert abs(dH[0] - dH_r) <= 1e-9 * max(abs(dH_r), 1.0) and bool(acc[0]) == a_r
    cb.close()


def test_device_rng_statistics(dw):
    """Throughput mode (Philox momenta / uniforms on device): <exp(-dH)> = 1 and momenta ~ N(0, m)."""
    B, L = 64, 4
    cb = dw.ChainBatch(B, L, L)
    cb.set_params(PHYS["t"], PHYS["tp"], PHYS["mu"], 5.0, PHYS["J"], 2.0)
    rng = np.random.default_rng(5)
    cb.set_disorder(np.zeros((B, 16)))
    cb.set_field(((rng.random((B, 2, 16)) - 0.5) + 1j * (rng.random((B, 2, 16)) - 0.5)) * 0.1)
    cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
    cb.seed(1234)
    dt = orc.calc_optimal_dt(5.0, 0.8, 2.0, 8)
    cb.run_sweeps(20, 8, dt)                        # thermalise
    vals, accs = [], 0
    for _ in range(10):
        nacc, dH, obs = cb.run_sweeps(1, 8, dt, observables=True)
        vals.append(np.exp(-dH)); accs += nacc.sum()
        assert obs.shape == (1, B, 9) and np.all(np.isfinite(obs))
    v = np.concatenate(vals)
    assert abs(v.mean() - 1.0) <= 5 * v.std() / np.sqrt(len(v)) + 0.02
    assert accs > 0.5 * len(v)
    cb.trajectory(1, 1e-9); pi = cb.get_momentum(); cb.commit(np.ones(B, int))
    x = np.concatenate([pi.real.ravel(), pi.imag.ravel()])
    assert abs(x.mean()) < 0.1 and abs(x.var() - 2.0) < 0.2      # Re, Im ~ N(0, m = 2)
    cb.close()


def test_device_initialize_state(dw):
    """initialize_state on the device (src/Types.jl:118-134): round(N n_imp) distinct impurity sites at W
    (Julia round = ties to even), Delta0 with Re, Im ~ U[-0.05, 0.05), pi = 0; chains differ; reseeding
    reproduces the draw."""
    B, L = 32, 10
    N = L * L
    cb = dw.ChainBatch(B, L, L)
    cb.set_params(PHYS["t"], PHYS["tp"], PHYS["mu"], 5.0, PHYS["J"], 1.0)
    n_imp = np.where(np.arange(B) % 2 == 0, 0.078, 0.025)       # 7.8 -> 8 sites, 2.5 -> 2 sites (ties to even)
    cb.seed(77)
    cb.init_state(3.0, n_imp)
    w, d, pi = cb.get_disorder(), cb.get_field(), cb.get_momentum()
    for b in range(B):
        assert set(np.unique(w[b])) <= {0.0, 3.0}
        assert np.count_nonzero(w[b]) == int(np.rint(N * n_imp[b]))
    assert np.all(pi == 0)
    x = np.concatenate([d.real.ravel(), d.imag.ravel()])
    assert x.min() >= -0.05 and x.max() < 0.05
    assert abs(x.mean()) < 2e-3 and abs(x.var() - 0.01 / 12) < 1e-4
    assert len({tuple(np.flatnonzero(w[b])) for b in range(0, B, 2)}) > B // 4          # chains differ
    hits = np.zeros(N)
    for b in range(B):
        hits += w[b] != 0
    assert hits.max() <= 8                                                           # no preferred site
    cb.seed(77)
    cb.init_state(3.0, n_imp)
    assert np.array_equal(cb.get_disorder(), w) and np.array_equal(cb.get_field(), d)
    cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
    assert np.all(np.isfinite(cb.measure_observables()))
    cb.close()


def test_batched_run_driver_matches_oracle_run(dw, tmp_path):
    """run_simulation (src/Simulation.jl:34-236) for two chains at once, host-RNG mode: the adaptive
    thermalisation decisions, accept flags, dH and the nine observables of every measured sweep equal
    a sequential oracle run that consumes the same generator; files have the reference's layout."""
    from dwhmc import simulation as sim
    L, n_therm, n_meas, Nt0, Ntm = 4, 10, 6, 4, 5
    betas, seeds = [5.0, 40.0], [11, 12]
    ps = [dw.ModelParameters(L, L, PHYS["t"], PHYS["tp"], PHYS["mu"], PHYS["W"], 0.1, b, PHYS["J"], PHYS["mass"]) for b in betas]
    dirs = [str(tmp_path / sim.scan_dir_T(1.0 / b)) for b in betas]
    for p in ps:
        p.eta, p.d_omega, p.omega_max = 0.5, 0.1, 4.0
    out = sim.run_simulation_batch(ps, dirs, n_therm=n_therm, n_measure=n_meas, Nt_therm_init=Nt0, Nt_measure=Ntm,
                                   measure_transport_freq=2, bin_size=2, seeds=seeds, rng_mode="host")
    for c, (beta, seed) in enumerate(zip(betas, seeds)):
        p = orc.ModelParameters(L, L, PHYS["t"], PHYS["tp"], PHYS["mu"], PHYS["W"], 0.1, beta, PHYS["J"], PHYS["mass"],
                                eta=0.5, d_omega=0.1, omega_max=4.0)
        rng, st, ca = orc.make_chain(p, seed)
        Nt, recent = Nt0, 0
        trans_rows = open(os.path.join(dirs[c], "transport.csv")).read().splitlines()
        assert trans_rows[0] == sim.TRANS_HEADER and len(trans_rows) == 1 + n_meas // 2
        bins = np.load(os.path.join(dirs[c], "spectra_bins.npz"))
        acc_spec, n_in_bin = None, 0
        dt = orc.calc_optimal_dt(p.beta, p.J, p.mass, Nt)
        for i in range(1, n_therm + 1):
            acc, _ = orc.hmc_sweep(ca, p, st, Nt=Nt, dt=dt, rng=rng)
            recent += acc
            if i % 5 == 0:
                Nt = sim.adapt_Nt(recent / 5, Nt); recent = 0
                dt = orc.calc_optimal_dt(p.beta, p.J, p.mass, Nt)
        assert out["Nt_therm_final"][c] == Nt
        dtm = orc.calc_optimal_dt(p.beta, p.J, p.mass, Ntm)
        rows = open(os.path.join(dirs[c], "observables.csv")).read().splitlines()
        assert rows[0] == sim.OBS_HEADER and len(rows) == n_meas + 1
        for i in range(1, n_meas + 1):
            acc, dH, Ho, _ = orc.hmc_sweep(ca, p, st, Nt=Ntm, dt=dtm, rng=rng, return_energies=True)
            obs = orc.measure_observables(ca, p, st)
            row = out["table"][i - 1, c]
            assert row[0] == i and bool(row[1]) == acc
            assert abs(row[2] - dH) <= 1e-9 * max(abs(Ho), 1.0)
            assert np.allclose(row[3:], obs, rtol=1e-7, atol=1e-9)
            assert rows[i] == sim.obs_csv_line(i, acc, row[2], row[3:]).rstrip("\n")
            if i % 2 == 0:                                   # measure_transport_freq = 2, bin_size = 2
                sp = orc.measure_transport_and_spectra(ca, p)
                k = i // 2
                it, stiff, dc = trans_rows[k].split(",")
                assert int(it) == i and abs(float(stiff) - sp.superfluid_stiffness) <= 2e-6
                assert abs(float(dc) - sp.dc_conductivity) <= 2e-6 * max(1.0, abs(sp.dc_conductivity))
                cur = [sp.optical_conductivity, sp.dos, sp.dos_AN, sp.A_k_w0]
                acc_spec = [a.copy() for a in cur] if n_in_bin == 0 else [a + b for a, b in zip(acc_spec, cur)]
                n_in_bin += 1
                if n_in_bin >= 2:
                    for key, a in zip(("opt_cond", "dos", "dos_AN", "A_k0"), acc_spec):
                        got = bins[f"sweep_{i}/{key}"]
                        assert np.max(np.abs(got - a / n_in_bin)) <= 1e-8 * max(np.max(np.abs(a)), 1e-12), key
                    assert int(bins[f"sweep_{i}/count"]) == 2
                    n_in_bin = 0
        assert np.allclose(bins["omega_grid"], orc.julia_range(p.eta, p.d_omega, p.omega_max))
        assert "Measurement Done." in open(os.path.join(dirs[c], "simulation.log")).read()


def test_scan_drivers_config4(dw, tmp_path):
    """BASELINE config 4 shapes at a small size: the beta scan (scripts/batch_scan_beta.jl) as one batch with
    transport every sweep, and the Nt-efficiency study (scripts/test_scan_Nt_efficiency.jl): acceptance rises with
    the number of leapfrog steps at fixed trajectory length and a 30-step trajectory is accepted almost always."""
    from dwhmc import simulation as sim
    betas = 10.0 ** np.linspace(-2, 5, 4)
    out = sim.batch_scan_beta(str(tmp_path), betas, Lx=6, Ly=6, n_therm=5, n_measure=4, Nt_therm=6, Nt_measure=6,
                              bin_size=2)
    assert out["table"].shape == (4, 4, 12) and len(out["transport"]) == 4
    for b in betas:
        d = tmp_path / sim.scan_dir_beta(b)
        assert len(open(d / "observables.csv").read().splitlines()) == 5
        assert len(open(d / "transport.csv").read().splitlines()) == 5
        z = np.load(d / "spectra_bins.npz")
        assert "sweep_2/dos" in z and "sweep_4/opt_cond" in z and int(z["sweep_4/count"]) == 2
        assert np.all(np.isfinite(z["sweep_4/A_k0"]))
    assert os.path.basename(str(tmp_path / sim.scan_dir_beta(0.01))) == "beta_0.01"
    Nts, dt, rate, eff = sim.scan_Nt_efficiency((2, 6, 30), Lx=6, Ly=6, n_warmup=10, n_measure=40, seed=3)
    assert np.allclose(Nts * dt, 2 * np.pi * np.sqrt(0.8 / 20.0))
    assert rate[2] >= 0.9 and rate[2] >= rate[0] - 0.05
    assert np.allclose(eff, rate / Nts)



def test_two_handles_concurrently(dw):
    """Two handles on one GPU driven from two host threads (handles are independent): the cooperative chase
    launches of the band route are chained per device, so they cannot deadlock half-resident; results equal a
    sequential run."""
    import threading
    L, B, Nt = 12, 40, 3
    N = L * L
    rng = np.random.default_rng(91)

    def make(seed):
        cb = dw.ChainBatch(B, L, L)
        cb.set_params(1.0, -0.35, -1.08, np.linspace(2, 40, B), 0.8, 1.0)
        w = np.zeros((B, N)); w[:, :7] = 1.0
        cb.set_disorder(w)
        r = np.random.default_rng(seed)
        cb.set_field((r.random((B, 2, N)) - 0.5 + 1j * (r.random((B, 2, N)) - 0.5)) * 0.1)
        cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG(); cb.seed(seed)
        return cb

    dt = np.full(B, 0.05)
    ref = []
    for seed in (1, 2):
        cb = make(seed)
        ref.append(cb.run_sweeps(3, Nt, dt)[1].copy())
        cb.close()
    cbs = [make(1), make(2)]
    out = [None, None]

    def work(i):
        out[i] = cbs[i].run_sweeps(3, Nt, dt)[1].copy()

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert all(not t.is_alive() for t in th), "concurrent handles hung"
    for i in range(2):
        assert np.allclose(out[i], ref[i], rtol=0, atol=1e-9 * np.max(np.abs(ref[i])) + 1e-9)
        cbs[i].close()


def test_chase_bitwise_reproducible_beside_another_handle(dw):
    """The band -> tridiagonal reduction of one handle gives the same bits on every run while a second handle keeps the
    SMs busy from another host thread.  (Regression: the corner message of the position-owning chase was once read
    from shared memory after the barrier behind which other warps rewrite it; about 1 run in 100 at this size then
    had one wrong diagonal entry.  tools/chase_race_check.py is the long version.)"""
    import threading
    L, B = 12, 12
    N = L * L

    def make(seed, nb):
        cb = dw.ChainBatch(nb, L, L)
        cb.set_params(1.0, -0.35, -1.08, np.linspace(2, 40, nb), 0.8, 1.0)
        w = np.zeros((nb, N)); w[:, :7] = 1.0
        cb.set_disorder(w)
        r = np.random.default_rng(seed)
        cb.set_field((r.random((nb, 2, N)) - 0.5 + 1j * (r.random((nb, 2, N)) - 0.5)) * 0.1)
        cb.init_static_H(); cb.update_H_BdG()
        return cb

    a = make(1, B)
    d0, e0 = (x.copy() for x in a.debug_tridiagonalize())
    b = make(2, 40)
    b.diagonalize_H_BdG(); b.seed(2)
    stop = threading.Event()

    def noise():
        dt = np.full(40, 0.05)
        while not stop.is_set():
            b.run_sweeps(1, 3, dt)

    th = threading.Thread(target=noise)
    th.start()
    try:
        bad = 0
        for _ in range(250):
            d, e = a.debug_tridiagonalize()
            bad += not (np.array_equal(d, d0) and np.array_equal(e, e0))
    finally:
        stop.set()
        th.join(timeout=120)
    assert not th.is_alive(), "the second handle hung"
    assert bad == 0, f"{bad} of 250 tridiagonalisations differ from the quiet run"
    a.close(); b.close()


def test_dense_route_against_oracle(dw, monkeypatch):
    """The dense eigensolver route (DWHMC_BAND=0: blocked tridiagonalisation + back-transformation; the route of
    lattices whose half-bandwidth exceeds 100, e.g. L = 32, see tests/test_gpu_round2.py for that size) against the
    oracle."""
    monkeypatch.setenv("DWHMC_BAND", "0")
    cb, ps, sts, cs = make_batch(dw, 12, [3.0, 30.0, 300.0], 0.05, 2100)
    monkeypatch.delenv("DWHMC_BAND")
    assert cb.band_halfwidth() == 0
    n = 2 * 144
    E, U = cb.get_eigenvalues(), cb.get_eigenvectors()
    cb.compute_forces()
    F = cb.get_forces()
    for b in range(3):
        Hf = orc.full_hermitian(cs[b]); Ub = U[b].T
        nrm = np.max(np.abs(cs[b].E_n))
        assert np.max(np.abs(E[b] - cs[b].E_n)) <= 1e-12 * nrm
        assert np.max(np.abs(Hf @ Ub - Ub * E[b])) <= 1e-12 * nrm
        assert np.max(np.abs(Ub.conj().T @ Ub - np.eye(n))) <= 1e-12
        orc.compute_forces(cs[b], ps[b], sts[b])
        assert rel(F[b].T, cs[b].forces) <= RTOL
    cb.close()


def test_default_route_is_band_where_a_chase_kernel_exists(dw):
    for L, expect in ((4, 0), (8, 36), ((6, 10), 28), (12, 52), (24, 100), (22, 100), (9, 44), ((5, 13), 28), (7, 0)):
        Lx, Ly = (L, L) if isinstance(L, int) else L
        cb = dw.ChainBatch(1, Lx, Ly)
        assert cb.band_halfwidth() == expect, (L, cb.band_halfwidth())
        cb.close()


CHASE_SWITCH_SHAPES = {
    # the position-owning kernel (the default) with short epochs: every position changes CTAs every 16 / 37 sweeps
    # through the band storage, at L = 8 with more tasks than CTAs, at L = 24, and on a rectangle
    "DWHMC_CHASE_Q=16": (["8", "150"], ["24", "3"], ["5x13", "2"]),
    "DWHMC_CHASE_Q=37": (["16", "40"], ["6x10", "3"]),
    # the sweep-owning kernel (band.cu), the band route's second kernel
    "DWHMC_CHASE=sweep": (["24", "2"], ["12", "5"], ["6x10", "3"], ["8", "150"], ["16", "64"], ["5x13", "2"]),
}


@pytest.mark.parametrize("switch", sorted(CHASE_SWITCH_SHAPES))
def test_chase_fallback_kernels(switch):
    """The bulge-chase kernels behind their (process-wide, read-once) switches against LAPACK, each in a process of its
    own: eigenvalues, residual and unitarity of two chains per shape.  The default is the position-owning kernel
    (band_systolic.cu; every other GPU test runs it, with one epoch for small batches and epochs of 5 b / 4 sweeps for
    large ones); here: short epochs and the sweep-owning kernel (band.cu), the band route's one fallback, incl. a
    rectangle whose half-bandwidth is rounded up (5 x 13: 24 -> 28)."""
    import re
    import subprocess
    import sys
    key, val = switch.split("=")
    env = dict(os.environ, **{key: val})
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for args in CHASE_SWITCH_SHAPES[switch]:
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "band_check.py")] + args, env=env,
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        rows = re.findall(r"chain \d+ E err (\S+) res (\S+) orth (\S+)", out.stdout)
        assert len(rows) == 2, out.stdout[-2000:]
        for err, res, orth in rows:
            assert float(err) <= 1e-12 and float(res) <= 1e-12 and float(orth) <= 1e-12, (switch, args, rows)
        assert re.search(r"route half-bandwidth (\d+)", out.stdout).group(1) != "0"
