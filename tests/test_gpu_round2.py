"""GPU parity tests added in round 2 ("close the parity holes"): the benchmark shape itself (L = 24,
64 chains, every CTA of the chase and of the back-transformation busy), BASELINE config 5 (L = 32), the reference's
physics check scripts/benchmark_clean.jl:112-123, chain averages against oracle chains, and the band route against
the ORACLE (not against the dense route) on every odd shape.  Same tolerances as tests/test_gpu_parity.py: 1e-10
relative for energies, forces and dH (dH relative to |H_old|); eigenvectors only through gauge-invariant quantities."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import dwhmc_oracle as orc  # noqa: E402  (checker only)

RTOL = 1e-10
PHYS = dict(t=1.0, tp=-0.35, mu=-1.08, W=1.0, J=0.8, mass=1.0)


@pytest.fixture(scope="module")
def dw():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import dwhmc
    return dwhmc


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def oracle_chain(Lx, Ly, beta, w, delta, J=PHYS["J"], n_imp=0.05):
    """Oracle chain with the given disorder w[N] and field delta[2, N] (the batch layout)."""
    p = orc.ModelParameters(Lx, Ly, PHYS["t"], PHYS["tp"], PHYS["mu"], PHYS["W"], n_imp, float(beta), J, PHYS["mass"])
    st = orc.SimulationState(np.array(w, float), np.array(delta).T.copy(), np.zeros((Lx * Ly, 2), complex))
    c = orc.initialize_cache(p)
    orc.init_static_H(c, p, st); orc.update_H_BdG(c, p, st); orc.diagonalize_H_BdG(c, p)
    return p, st, c


def test_bench_shape_L24_B64_two_sweeps(dw):
    """The shape bench.py times (src/HMC.jl:71-144 over 64 chains at L = 24, Nt = 6, temperatures of the whole
    scan): two hmc_sweep!s with injected momenta and uniforms; dH, accept / restore and the field after each sweep
    against the oracle on four chains spanning beta_max ... beta_min; size-independent checks on all 64."""
    L, B, Nt = 24, 64, 6
    N, n = L * L, 2 * L * L
    rng = np.random.default_rng(2402)
    betas = 1.0 / 10.0 ** np.linspace(-4, 3, B)                   # beta from 1e4 down to 1e-3
    w = np.zeros((B, N))
    for b in range(B):
        w[b, rng.permutation(N)[:int(np.rint(N * 0.05))]] = 1.0
    delta = (rng.random((B, 2, N)) - 0.5 + 1j * (rng.random((B, 2, N)) - 0.5)) * 0.1
    cb = dw.ChainBatch(B, L, L)
    assert cb.band_halfwidth() == 100                             # the band route, all 148 CTAs of the chase
    cb.set_params(PHYS["t"], PHYS["tp"], PHYS["mu"], betas, PHYS["J"], PHYS["mass"])
    cb.set_disorder(w); cb.set_field(delta)
    cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
    check = [0, 21, 42, 63]
    ora = {b: oracle_chain(L, L, betas[b], w[b], delta[b]) for b in check}
    E = cb.get_eigenvalues()
    for b in check:
        assert np.max(np.abs(E[b] - ora[b][2].E_n)) <= 1e-12 * np.max(np.abs(ora[b][2].E_n))
    assert np.max(np.abs(E + E[:, ::-1])) <= 1e-12 * np.max(np.abs(E))       # particle-hole, all 64 chains
    dt = np.array([orc.calc_optimal_dt(bt, PHYS["J"], PHYS["mass"], Nt) for bt in betas])
    n_acc = 0
    for it in range(2):
        pi0 = np.stack([orc.draw_momentum(ora[check[0]][0], np.random.default_rng(7000 + 100 * it + b)) for b in range(B)])
        u = np.random.default_rng(7900 + it).random(B)
        acc, dH = cb.hmc_sweep(Nt, dt, pi0=pi0, uniforms=u)
        D, E = cb.get_field(), cb.get_eigenvalues()
        assert np.all(np.isfinite(dH)) and np.all(np.isfinite(D.view(float)))
        for b in check:
            p, st, c = ora[b]
            a_r, dH_r, Ho, _ = orc.hmc_sweep(c, p, st, Nt=Nt, dt=float(dt[b]), pi0=pi0[b], uniform=float(u[b]),
                                             return_energies=True)
            assert abs(dH[b] - dH_r) <= RTOL * max(abs(Ho), 1.0), (it, b, dH[b], dH_r)
            assert bool(acc[b]) == a_r, (it, b)
            assert np.max(np.abs(D[b].T - st.Delta)) <= 1e-10, (it, b)
            assert np.max(np.abs(E[b] - c.E_n)) <= 1e-11 * np.max(np.abs(c.E_n)), (it, b)
            n_acc += int(a_r)
    O = cb.measure_observables()
    for b in check:
        p, st, c = ora[b]
        assert np.allclose(O[b], orc.measure_observables(c, p, st), rtol=1e-8, atol=1e-10), b
    # every chain: the cache holds a unitary eigenbasis of the matrix of its current field
    U = cb.get_eigenvectors()
    H = cb.get_H()
    E = cb.get_eigenvalues()
    for b in range(0, B, 9):
        Hu = H[b].T
        Hf = np.triu(Hu) + np.triu(Hu, 1).conj().T
        Ub = U[b].T
        nrm = np.max(np.abs(E[b]))
        assert np.max(np.abs(Hf @ Ub - Ub * E[b])) <= 1e-12 * nrm, b
        assert np.max(np.abs(Ub.conj().T @ Ub - np.eye(n))) <= 1e-12, b
    cb.close()


def test_config5_L32_eigen_force_observables_transport(dw):
    """BASELINE config 5 shape (L = 32, n = 2048; half-bandwidth 132 > 100, so the dense route): spectrum, residual,
    unitarity, forces, energy, the nine observables and transport / spectra against the oracle, two chains."""
    L, B = 32, 2
    N, n = L * L, 2 * L * L
    rng = np.random.default_rng(3201)
    betas = np.array([20.0, 1000.0])
    w = np.zeros((B, N))
    for b in range(B):
        w[b, rng.permutation(N)[:int(np.rint(N * 0.05))]] = 1.0
    delta = (rng.random((B, 2, N)) - 0.5 + 1j * (rng.random((B, 2, N)) - 0.5)) * 0.1
    cb = dw.ChainBatch(B, L, L)
    cb.set_params(PHYS["t"], PHYS["tp"], PHYS["mu"], betas, PHYS["J"], PHYS["mass"])
    cb.set_disorder(w); cb.set_field(delta)
    cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
    cb.compute_forces()
    E, U, F = cb.get_eigenvalues(), cb.get_eigenvectors(), cb.get_forces()
    cb.set_momentum(np.zeros((B, 2, N), complex))
    Etot = cb.compute_total_energy()
    O = cb.measure_observables()
    eta = 0.05
    tr = cb.measure_transport_and_spectra(eta, 0.05, 4.0)
    for b in range(B):
        p, st, c = oracle_chain(L, L, betas[b], w[b], delta[b])
        nrm = np.max(np.abs(c.E_n))
        Hf = orc.full_hermitian(c); Ub = U[b].T
        assert np.max(np.abs(E[b] - c.E_n)) <= 1e-12 * nrm
        assert np.max(np.abs(Hf @ Ub - Ub * E[b])) <= 1e-12 * nrm
        assert np.max(np.abs(Ub.conj().T @ Ub - np.eye(n))) <= 1e-12
        orc.compute_forces(c, p, st)
        assert rel(F[b].T, c.forces) <= RTOL
        Er = orc.compute_total_energy(c, p, st)
        assert abs(Etot[b] - Er) <= 1e-12 * max(abs(Er), 1.0)
        assert np.allclose(O[b], orc.measure_observables(c, p, st), rtol=1e-8, atol=1e-10)
        p.eta, p.d_omega, p.omega_max = eta, 0.05, 4.0
        ref = orc.measure_transport_and_spectra(c, p)
        assert abs(tr["superfluid_stiffness"][b] - ref.superfluid_stiffness) <= 1e-9
        assert abs(tr["dc_conductivity"][b] - ref.dc_conductivity) <= 1e-9 * max(abs(ref.dc_conductivity), 1e-6)
        for key, arr in (("optical_conductivity", ref.optical_conductivity), ("dos", ref.dos), ("dos_AN", ref.dos_AN),
                         ("A_k_w0", ref.A_k_w0)):
            assert np.max(np.abs(tr[key][b] - arr)) <= 1e-9 * max(np.max(np.abs(arr)), 1e-12), key
    cb.close()


def bcs_rhs(D, Lx, Ly, t, tp, mu, beta, J):
    """calc_BCS_RHS, scripts/benchmark_clean.jl:15-44: right-hand side of the d-wave gap equation on the L x L k grid."""
    kx = 2 * np.pi * np.arange(Lx)[None, :] / Lx
    ky = 2 * np.pi * np.arange(Ly)[:, None] / Ly
    eps = -2 * t * (np.cos(kx) + np.cos(ky)) - 4 * tp * np.cos(kx) * np.cos(ky) - mu
    gk = np.cos(kx) - np.cos(ky)
    Ek = np.sqrt(eps ** 2 + (D * gk) ** 2)
    return J / (Lx * Ly) * np.sum(gk ** 2 / (2 * Ek) * np.tanh(0.5 * beta * Ek)) * D


def test_clean_limit_bcs_gap_equation(dw):
    """The reference's physics check, scripts/benchmark_clean.jl:47-123: clean 10 x 10 lattice, beta = 180, J = 1.6,
    uniform d-wave start (Delta_x = 0.2, Delta_y = -0.2), 50 thermalisation sweeps at Nt = 20, 100 measured sweeps
    at Nt = 5; <|Delta_global|> has to reproduce itself through the BCS gap equation to 0.02 (:119).  Run here for
    eight independent chains at once with the on-device RNG; every chain has to pass, as the script's single one."""
    L, B = 10, 8
    N = L * L
    beta, J = 180.0, 1.6
    cb = dw.ChainBatch(B, L, L)
    cb.set_params(PHYS["t"], PHYS["tp"], PHYS["mu"], beta, J, PHYS["mass"])
    cb.set_disorder(np.zeros((B, N)))
    d0 = np.zeros((B, 2, N), complex); d0[:, 0, :] = 0.2; d0[:, 1, :] = -0.2
    cb.set_field(d0)
    cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
    cb.seed(20260118)
    dt_th = math.pi * math.sqrt(PHYS["mass"] * J / beta) / 20            # calc_optimal_dt, :7-10
    dt_me = math.pi * math.sqrt(PHYS["mass"] * J / beta) / 5
    cb.run_sweeps(50, 20, dt_th)
    nacc, _, obs = cb.run_sweeps(100, 5, dt_me, observables=True)        # obs [sweep, chain, 9]
    dglob = obs[:, :, 3]                                                  # Delta_global (ObservablesResult field 4, src/Observables.jl:70-80)
    assert np.all(nacc > 10)                                              # the chains move
    for b in range(B):
        m = float(np.mean(dglob[:, b]))
        assert 0.05 < m < 1.0
        assert abs(m - bcs_rhs(m, L, L, PHYS["t"], PHYS["tp"], PHYS["mu"], beta, J)) < 0.02, (b, m)
    cb.close()


def test_chain_averages_agree_with_oracle_chains(dw):
    """BASELINE north star: "observable averages must agree within statistical error bars".  16 GPU chains (device
    Philox momenta and uniforms) against 8 oracle chains (NumPy RNG) of the same disordered 8 x 8 model at beta = 5:
    the chain-to-chain scatter of the per-chain means gives the standard errors; all nine observables, the acceptance
    rate and <exp(-dH)> = 1 have to agree within 4 combined standard errors (seeds are fixed: deterministic)."""
    L, Nt, n_th, n_me = 8, 6, 20, 60
    N = L * L
    beta = 5.0
    rng = np.random.default_rng(88)
    w = np.zeros(N); w[rng.permutation(N)[:3]] = 1.0
    BG, BO = 16, 8
    dt = orc.calc_optimal_dt(beta, PHYS["J"], PHYS["mass"], Nt)
    d_start = (rng.random((BG + BO, 2, N)) - 0.5 + 1j * (rng.random((BG + BO, 2, N)) - 0.5)) * 0.1
    cb = dw.ChainBatch(BG, L, L)
    cb.set_params(PHYS["t"], PHYS["tp"], PHYS["mu"], beta, PHYS["J"], PHYS["mass"])
    cb.set_disorder(np.tile(w, (BG, 1))); cb.set_field(d_start[:BG])
    cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
    cb.seed(4242)
    cb.run_sweeps(n_th, Nt, dt)
    g_obs, g_acc, g_edh = [], np.zeros(BG), np.zeros(BG)
    for _ in range(n_me):
        acc, dH = cb.hmc_sweep(Nt, dt)
        g_obs.append(cb.measure_observables()); g_acc += acc; g_edh += np.exp(-dH)
    g_obs = np.mean(np.array(g_obs), axis=0)                              # [chain, 9]
    cb.close()
    o_obs, o_acc, o_edh = [], [], []
    for b in range(BO):
        p, st, c = oracle_chain(L, L, beta, w, d_start[BG + b])
        r = np.random.Generator(np.random.PCG64(600 + b))
        for _ in range(n_th):
            orc.hmc_sweep(c, p, st, Nt=Nt, dt=dt, rng=r)
        rows, na, ed = [], 0, 0.0
        for _ in range(n_me):
            a, dH = orc.hmc_sweep(c, p, st, Nt=Nt, dt=dt, rng=r)
            rows.append(orc.measure_observables(c, p, st)); na += int(a); ed += math.exp(-dH)
        o_obs.append(np.mean(rows, axis=0)); o_acc.append(na / n_me); o_edh.append(ed / n_me)
    o_obs = np.array(o_obs)

    def agree(g, o, what):
        mg, mo = np.mean(g), np.mean(o)
        se = math.sqrt(np.var(g, ddof=1) / len(g) + np.var(o, ddof=1) / len(o))
        assert abs(mg - mo) <= 4.0 * se + 1e-12, (what, mg, mo, se)

    for k in range(9):
        agree(g_obs[:, k], o_obs[:, k], f"observable {k}")
    agree(g_acc / n_me, np.array(o_acc), "acceptance")
    agree(g_edh / n_me, np.array(o_edh), "<exp(-dH)>")
    assert abs(np.mean(g_edh / n_me) - 1.0) <= 4.0 * np.std(g_edh / n_me, ddof=1) / math.sqrt(BG) + 1e-12


@pytest.mark.parametrize("L,B", [((6, 10), 3), ((10, 6), 2), (8, 150), ((5, 13), 2), (9, 2), (22, 2), (20, 3), (12, 5)])
def test_band_route_against_oracle(dw, monkeypatch, L, B):
    """Every shape of test_band_route_shapes (rectangles, odd sides, half-bandwidths rounded up to the next chase
    kernel, a batch larger than the grid) plus L = 12, 20, band route against the ORACLE: spectrum, forces, energy,
    observables and one trajectory."""
    Lx, Ly = (L, L) if isinstance(L, int) else L
    N = Lx * Ly
    rng = np.random.default_rng(31)
    betas = np.linspace(2.0, 50.0, B)
    w = np.zeros((B, N)); w[:, rng.permutation(N)[:3]] = 1.0
    delta = (rng.random((B, 2, N)) - 0.5 + 1j * (rng.random((B, 2, N)) - 0.5)) * 0.2
    monkeypatch.setenv("DWHMC_BAND", "1")
    cb = dw.ChainBatch(B, Lx, Ly)
    monkeypatch.delenv("DWHMC_BAND")
    assert cb.band_halfwidth() > 0
    cb.set_params(PHYS["t"], PHYS["tp"], PHYS["mu"], betas, PHYS["J"], PHYS["mass"])
    cb.set_disorder(w); cb.set_field(delta)
    cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG(); cb.compute_forces()
    E, F, O = cb.get_eigenvalues(), cb.get_forces(), cb.measure_observables()
    Nt = 3
    dt = np.array([0.5 * orc.calc_optimal_dt(bt, PHYS["J"], PHYS["mass"], Nt) for bt in betas])
    check = range(B) if B <= 5 else [0, B // 3, B - 1]
    ora = {b: oracle_chain(Lx, Ly, betas[b], w[b], delta[b]) for b in check}
    pi0 = np.stack([orc.draw_momentum(ora[0][0], np.random.default_rng(50 + b)) for b in range(B)])
    u = np.random.default_rng(51).random(B)
    acc, dH = cb.hmc_sweep(Nt, dt, pi0=pi0, uniforms=u)
    for b in check:
        p, st, c = ora[b]
        assert np.max(np.abs(E[b] - c.E_n)) <= 1e-12 * np.max(np.abs(c.E_n))
        orc.compute_forces(c, p, st)
        assert rel(F[b].T, c.forces) <= RTOL
        assert np.allclose(O[b], orc.measure_observables(c, p, st), rtol=1e-8, atol=1e-10)
        a_r, dH_r, Ho, _ = orc.hmc_sweep(c, p, st, Nt=Nt, dt=float(dt[b]), pi0=pi0[b], uniform=float(u[b]),
                                         return_energies=True)
        assert abs(dH[b] - dH_r) <= RTOL * max(abs(Ho), 1.0)
        assert bool(acc[b]) == a_r
    cb.close()


def test_eigensolve_is_bit_reproducible_at_bench_shape(dw):
    """Race check by determinism (compute-sanitizer is not available on the GPU pool): the bulge chase pipelines the
    sweeps of a chain over several CTAs through release/acquire counters and the back-transformation hands work items
    to CTAs by ticket, so the schedule differs from run to run -- the result must not.  Three eigensolves of the same
    64 matrices at L = 24 (all 148 CTAs busy in both kernels) have to agree bit for bit, eigenvectors included."""
    L, B = 24, 64
    N = L * L
    rng = np.random.default_rng(77)
    w = np.zeros((B, N))
    for b in range(B):
        w[b, rng.permutation(N)[:29]] = 1.0
    delta = (rng.random((B, 2, N)) - 0.5 + 1j * (rng.random((B, 2, N)) - 0.5)) * 0.1
    cb = dw.ChainBatch(B, L, L)
    cb.set_params(PHYS["t"], PHYS["tp"], PHYS["mu"], np.logspace(-1, 3, B), PHYS["J"], PHYS["mass"])
    cb.set_disorder(w); cb.set_field(delta)
    cb.init_static_H(); cb.update_H_BdG()
    ref = None
    for _ in range(3):
        cb.diagonalize_H_BdG()
        E = cb.get_eigenvalues()
        U = cb.get_eigenvectors()[::7].copy()          # every 7th chain: 10 x 21 MB
        if ref is None:
            ref = (E, U)
        else:
            assert np.array_equal(E, ref[0])
            assert np.array_equal(U.view(np.float64), ref[1].view(np.float64))
    cb.close()
