"""CPU tests of the oracle (oracle/dwhmc_oracle.py): golden vectors, the
reference's own force KAT (scripts/bench_forces.jl:121-129) and the identities
SURVEY.md section 4 lists."""
import glob
import os

import numpy as np
import pytest

import dwhmc_oracle as orc

PHYS = dict(t=1.0, tp=-0.35, mu=-1.08, W=1.0, J=0.8, mass=1.0)


def params(L, n_imp, beta, **kw):
    Lx, Ly = (L, L) if isinstance(L, int) else L
    d = dict(PHYS); d.update(kw)
    return orc.ModelParameters(Lx, Ly, d["t"], d["tp"], d["mu"], d["W"], n_imp, beta, d["J"], d["mass"])


def test_neighbour_tables_match_reference_convention():
    # src/Types.jl:60-80 with 0-based sites: i = y*Lx + x
    nn, nnn = orc.neighbour_tables(4, 3)
    assert nn.shape == (12, 4)
    assert tuple(nn[0]) == (1, 4, 3, 8)
    assert tuple(nnn[0]) == (5, 7, 11, 9)
    # +x then -x returns home
    assert np.all(nn[nn[:, 0], 2] == np.arange(12))
    assert np.all(nn[nn[:, 1], 3] == np.arange(12))


def test_logexp_functions():
    x = np.array([-800.0, -745.0, -100.0, -36.0, -1.0, 0.0, 1.0, 18.5, 34.0, 40.0])
    assert orc.logistic(x)[0] == 0.0 and orc.logistic(x)[-1] == 1.0
    ref = np.array([0 if v < -745.2 else np.log1p(np.exp(v)) if v < 30 else v for v in x])
    assert np.allclose(orc.log1pexp(x), ref, rtol=1e-15, atol=0)
    assert np.allclose(orc.logistic(x[2:8]), 1 / (1 + np.exp(-x[2:8])), rtol=1e-15)


TRANSPORT_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "transport_*.npz")))


@pytest.mark.parametrize("path", TRANSPORT_GOLDEN, ids=[os.path.basename(p)[:-4] for p in TRANSPORT_GOLDEN])
def test_transport_golden_reproduced_and_sum_rules(path):
    """The transport / spectra restatement reproduces its committed vectors, and obeys the checks the physics
    offers: the DOS integrates to the particle weight (Lorentzian tails aside), A(k, 0) sums to the zero-energy
    particle weight (Parseval), sigma(omega) >= 0 and the antinodal DOS is non-negative."""
    g = np.load(path)
    Lx, Ly = int(g["Lx"]), int(g["Ly"])
    p = orc.ModelParameters(Lx, Ly, 1.0, -0.35, -1.08, 1.0, float(g["n_imp"]), float(g["beta"]), 0.8, 1.0,
                            eta=float(g["eta"]), d_omega=float(g["d_omega"]), omega_max=float(g["omega_max"]))
    st = orc.SimulationState(g["disorder"].copy(), g["Delta0"].copy(), np.zeros_like(g["Delta0"]))
    c = orc.initialize_cache(p)
    orc.init_static_H(c, p, st); orc.update_H_BdG(c, p, st); orc.diagonalize_H_BdG(c, p)
    orc.measure_observables(c, p, st)
    r = orc.measure_transport_and_spectra(c, p)
    assert abs(r.superfluid_stiffness - float(g["stiffness"])) <= 1e-10
    assert abs(r.dc_conductivity - float(g["dc"])) <= 1e-10 * max(1.0, abs(float(g["dc"])))
    for got, key in ((r.optical_conductivity, "sigma"), (r.dos, "dos"), (r.dos_AN, "dos_AN"), (r.A_k_w0, "A_k0")):
        assert np.max(np.abs(got - g[key])) <= 1e-10 * max(np.max(np.abs(g[key])), 1e-12), key
    N = p.N
    w_n = np.sum(np.abs(c.U[:N]) ** 2, axis=0)
    inside = np.abs(c.E_n) < p.omega_max - 20 * p.eta
    integral = np.sum(r.dos) * p.d_omega
    assert abs(integral - np.sum(w_n) / N) <= 0.05 + np.sum(w_n[~inside]) / N
    w0 = orc.lorentzian(-c.E_n, p.eta)
    sel = w0 > 1e-6
    assert abs(np.sum(r.A_k_w0) - np.sum(w_n[sel] * w0[sel])) <= 1e-9 * max(np.sum(r.A_k_w0), 1.0)
    assert np.all(r.optical_conductivity >= -1e-12) and np.all(r.dos_AN >= 0)


@pytest.mark.parametrize("path", sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                                        if not os.path.basename(p).startswith("transport_")))
def test_golden_reproduced(path):
    g = np.load(path)
    p = params((int(g["Lx"]), int(g["Ly"])), float(g["n_imp"]), float(g["beta"]))
    rng, st, c = orc.make_chain(p, int(g["seed"]))
    assert np.array_equal(st.disorder_pot, g["disorder"])
    assert np.array_equal(st.Delta, g["Delta0"])
    assert np.allclose(c.E_n, g["E0"], rtol=0, atol=1e-13)
    orc.compute_forces(c, p, st)
    assert np.allclose(c.forces, g["F0"], rtol=1e-11, atol=1e-12)
    assert np.allclose(orc.measure_observables(c, p, st), g["obs0"], rtol=1e-11, atol=1e-13)
    for k in range(len(g["u"])):
        acc, dH, H_old, H_new = orc.hmc_sweep(c, p, st, Nt=int(g["Nt"]), dt=float(g["dt"]),
                                              pi0=g["pi0"][k], uniform=float(g["u"][k]),
                                              return_energies=True)
        assert acc == bool(g["accepted"][k])
        assert abs(H_old - g["H_old"][k]) <= 1e-11 * max(abs(H_old), 1)
        assert abs(dH - g["dH"][k]) <= 1e-10 * max(abs(H_old), 1)
        assert np.allclose(st.Delta, g["Delta_end"][k], rtol=1e-11, atol=1e-12)
        assert np.allclose(c.E_n, g["E_end"][k], rtol=0, atol=1e-12)


def test_bench_forces_kat_loop_orders_agree():
    # scripts/bench_forces.jl:8-33 data shapes (N=256 there; 64 here keeps it quick) and :121-129 check
    rng = np.random.Generator(np.random.PCG64(7))
    N = 64
    U = rng.random((2 * N, 2 * N)) + 1j * rng.random((2 * N, 2 * N))
    f = rng.random(2 * N)
    nn2 = np.stack([(np.arange(N) + 1) % N, (np.arange(N) - 1) % N], axis=1)
    Delta = rng.random((N, 2)) + 1j * rng.random((N, 2))
    F1 = orc.bench_forces_orig(U, f, nn2, 1.0, 0.5, Delta)
    F2 = orc.bench_forces_opt(U, f, nn2, 1.0, 0.5, Delta)
    assert np.max(np.abs(F1 - F2)) < 1e-10


def test_oracle_force_equals_bench_forces_formula():
    p = params(6, 0.05, 7.0)
    rng, st, c = orc.make_chain(p, 11)
    orc.compute_forces(c, p, st)
    F = orc.bench_forces_orig(c.U, c.fermi_factors, p.nn_table[:, :2], p.J, p.beta / (2 * p.J), st.Delta)
    assert np.allclose(F, c.forces, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("L,beta,n_imp", [(4, 3.0, 0.0), (6, 20.0, 0.05), (8, 50.0, 0.05)])
def test_force_is_minus_half_gradient_of_action(L, beta, n_imp):
    # SURVEY section 4 identity 1: dS/dRe(D) = -2 Re F, dS/dIm(D) = -2 Im F
    p = params(L, n_imp, beta)
    rng, st, c = orc.make_chain(p, 5)
    orc.compute_forces(c, p, st)
    h = 1e-5
    for (i, d) in [(0, 0), (3, 1), (p.N - 1, 0), (p.N // 2, 1)]:
        for comp, unit in ((0, 1.0), (1, 1j)):
            Dp = st.Delta.copy(); Dp[i, d] += h * unit
            Dm = st.Delta.copy(); Dm[i, d] -= h * unit
            g = (orc.action(p, st.disorder_pot, Dp) - orc.action(p, st.disorder_pot, Dm)) / (2 * h)
            F = c.forces[i, d]
            want = -2.0 * (F.real if comp == 0 else F.imag)
            assert abs(g - want) <= 2e-6 * max(1.0, abs(want))  # O(h^2 beta^2) truncation at h=1e-5


def test_spectrum_symmetry_and_pair_symmetry():
    p = params(6, 0.05, 12.0)
    rng, st, c = orc.make_chain(p, 3)
    E = c.E_n
    assert np.max(np.abs(E + E[::-1])) < 1e-12
    assert abs(np.trace(orc.full_hermitian(c))) < 1e-12
    f = orc.logistic(-p.beta * E)
    rho = (c.U * f) @ c.U.conj().T
    G = rho[: p.N, p.N:]
    assert np.max(np.abs(G - G.T)) < 1e-13
    G2 = -0.5 * (c.U[: p.N] * np.tanh(0.5 * p.beta * E)) @ c.U[p.N:].conj().T
    assert np.max(np.abs(G - G2)) < 1e-13


def test_leapfrog_reversible_and_dH_second_order():
    p = params(4, 0.0, 5.0)
    rng, st, c = orc.make_chain(p, 9)
    pi0 = orc.draw_momentum(p, rng)
    D0 = st.Delta.copy()
    dHs = []
    for Nt in (8, 16, 32):
        st.Delta[...] = D0
        orc.update_H_BdG(c, p, st); orc.diagonalize_H_BdG(c, p)
        dt = 0.4 / Nt
        acc, dH = orc.hmc_sweep(c, p, st, Nt=Nt, dt=dt, pi0=pi0, uniform=0.0)
        dHs.append(abs(dH))
        if Nt == 8:
            # reverse: flip momentum, integrate back
            assert acc
            pim = -st.pi.copy()
            acc2, dH2 = orc.hmc_sweep(c, p, st, Nt=Nt, dt=dt, pi0=pim, uniform=0.0)
            assert np.allclose(st.Delta, D0, atol=1e-11)
            assert abs(dH + dH2) < 1e-9
    assert dHs[1] < dHs[0] / 3.0 and dHs[2] < dHs[1] / 3.0


def test_stale_cache_quirk_runs():
    # scripts/benchmark_clean.jl:82-88 calls hmc_sweep! with E_n = 0, U = 0 (never diagonalised first)
    p = params(4, 0.0, 2.0)
    rng = np.random.Generator(np.random.PCG64(1))
    st = orc.initialize_state(p, rng)
    c = orc.initialize_cache(p)
    orc.init_static_H(c, p, st)
    acc, dH = orc.hmc_sweep(c, p, st, Nt=3, dt=0.1, pi0=orc.draw_momentum(p, rng), uniform=0.5)
    assert np.isfinite(dH)
