"""Generate the golden vectors under tests/golden/ from the CPU oracle.

The reference ships no fixtures and cannot run here
(no Julia), so these vectors are produced by ``oracle/dwhmc_oracle.py`` (LAPACK
zheevr through SciPy, the routine the reference calls).  Run:

    python tests/golden/make_golden.py

Seeds: NumPy PCG64(seed); the draw order is initialize_state (permutation,
Re Delta, Im Delta), then per sweep the momentum (Re, Im) and one uniform.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import dwhmc_oracle as orc  # noqa: E402

CASES = {
    # name: (L, n_imp, beta, Nt, n_sweeps, seed)
    "L4_clean_b5": (4, 0.0, 5.0, 4, 3, 1001),
    "L8_clean_b20": (8, 0.0, 20.0, 6, 3, 1002),       # BASELINE config 1 shape
    "L8_dis_b100": (8, 0.05, 100.0, 6, 2, 1003),
    "L6x10_dis_b2": ((6, 10), 0.05, 2.0, 5, 2, 1004),  # rectangular, n = 120
    "L16_dis_b20": (16, 0.05, 20.0, 6, 1, 2000),      # BASELINE config 2, chain 0
}

PHYS = dict(t=1.0, tp=-0.35, mu=-1.08, W=1.0, J=0.8, mass=1.0)  # scripts/batch_scan_T.jl:10-19


def run_case(L, n_imp, beta, Nt, n_sweeps, seed):
    Lx, Ly = (L, L) if isinstance(L, int) else L
    p = orc.ModelParameters(Lx, Ly, PHYS["t"], PHYS["tp"], PHYS["mu"], PHYS["W"], n_imp, beta,
                            PHYS["J"], PHYS["mass"])
    rng, st, c = orc.make_chain(p, seed)
    dt = orc.calc_optimal_dt(p.beta, p.J, p.mass, Nt)
    out = dict(Lx=Lx, Ly=Ly, n_imp=n_imp, beta=beta, Nt=Nt, dt=dt, seed=seed,
               disorder=st.disorder_pot.copy(), Delta0=st.Delta.copy(), E0=c.E_n.copy())
    orc.compute_forces(c, p, st)
    out["F0"] = c.forces.copy()
    out["obs0"] = orc.measure_observables(c, p, st)
    pis, us, accs, dHs, Holds, Hnews, Deltas, pis_end, obs, Es = [], [], [], [], [], [], [], [], [], []
    for _ in range(n_sweeps):
        pi0 = orc.draw_momentum(p, rng)
        u = rng.random()
        acc, dH, H_old, H_new = orc.hmc_sweep(c, p, st, Nt=Nt, dt=dt, pi0=pi0, uniform=u,
                                              return_energies=True)
        pis.append(pi0); us.append(u); accs.append(acc); dHs.append(dH)
        Holds.append(H_old); Hnews.append(H_new)
        Deltas.append(st.Delta.copy()); pis_end.append(st.pi.copy()); Es.append(c.E_n.copy())
        obs.append(orc.measure_observables(c, p, st))
    out.update(pi0=np.array(pis), u=np.array(us), accepted=np.array(accs), dH=np.array(dHs),
               H_old=np.array(Holds), H_new=np.array(Hnews), Delta_end=np.array(Deltas),
               pi_end=np.array(pis_end), E_end=np.array(Es), obs_end=np.array(obs))
    return out


TRANSPORT_CASES = {
    # name: (L, n_imp, beta, seed); eta = 8/N, d_omega = 0.2 eta, omega_max = 4 (scripts/batch_scan_T.jl:30-32)
    "transport_L8_dis_b20": (8, 0.05, 20.0, 3001),
    "transport_L6x10_dis_b200": ((6, 10), 0.05, 200.0, 3002),
}


def run_transport_case(L, n_imp, beta, seed):
    """measure_transport_and_spectra (src/Observables.jl:314-526) on the seeded initial state, after
    measure_observables (which leaves fermi_factors, as in the reference's sweep loop)."""
    Lx, Ly = (L, L) if isinstance(L, int) else L
    eta = 8.0 / (Lx * Ly)
    p = orc.ModelParameters(Lx, Ly, PHYS["t"], PHYS["tp"], PHYS["mu"], PHYS["W"], n_imp, beta, PHYS["J"],
                            PHYS["mass"], eta=eta, d_omega=0.2 * eta, omega_max=4.0)
    _, st, c = orc.make_chain(p, seed)
    orc.measure_observables(c, p, st)
    r = orc.measure_transport_and_spectra(c, p)
    return dict(Lx=Lx, Ly=Ly, n_imp=n_imp, beta=beta, seed=seed, eta=eta, d_omega=0.2 * eta, omega_max=4.0,
                disorder=st.disorder_pot.copy(), Delta0=st.Delta.copy(), stiffness=r.superfluid_stiffness,
                dc=r.dc_conductivity, omega_grid=r.omega_grid, sigma=r.optical_conductivity,
                dos_grid=r.dos_omega_grid, dos=r.dos, dos_AN=r.dos_AN, A_k0=r.A_k_w0)


def main():
    only_transport = "--transport-only" in sys.argv
    for name, args in CASES.items():
        if only_transport:
            break
        out = run_case(*args)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "dH", out["dH"], "acc", out["accepted"])
    for name, args in TRANSPORT_CASES.items():
        out = run_transport_case(*args)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "stiffness", out["stiffness"], "dc", out["dc"])


if __name__ == "__main__":
    main()
