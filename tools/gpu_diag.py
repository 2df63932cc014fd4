"""Stage-by-stage diagnostic of libdwhmc against the CPU oracle (run on the GPU box:
`python tools/gpu_diag.py [L ...]`).  Prints one line per check; never stops at the first failure."""
import os
import sys
import time
import traceback

import numpy as np
import scipy.linalg as sl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import dwhmc  # noqa: E402
import dwhmc_oracle as orc  # noqa: E402


def line(name, val, tol):
    ok = bool(np.isfinite(val) and val <= tol)
    print(f"  [{'ok' if ok else 'FAIL'}] {name}: {val:.3e} (tol {tol:.0e})", flush=True)
    return ok


def stage(fn):
    try:
        fn()
    except Exception:
        print("  [EXC] " + traceback.format_exc().replace("\n", "\n        "), flush=True)


def run(Lx, Ly, B=3, beta=20.0, n_imp=0.05):
    print(f"=== Lx={Lx} Ly={Ly} B={B} beta={beta}", flush=True)
    N, n = Lx * Ly, 2 * Lx * Ly
    rng = np.random.default_rng(Lx * 100 + Ly)
    cb = dwhmc.ChainBatch(B, Lx, Ly)
    ps, sts, cs = [], [], []
    for b in range(B):
        p = orc.ModelParameters(Lx, Ly, 1.0, -0.35, -1.08, 1.0, n_imp if b else 0.0, beta * (1 + b), 0.8, 1.0)
        _, st, c = orc.make_chain(p, 1000 + b)
        ps.append(p); sts.append(st); cs.append(c)
    cb.set_params([p.t for p in ps], [p.tp for p in ps], [p.mu for p in ps], [p.beta for p in ps],
                  [p.J for p in ps], [p.mass for p in ps])
    cb.set_disorder(np.stack([s.disorder_pot for s in sts]))
    cb.set_field(np.stack([s.Delta for s in sts]))

    def s_heev():
        A = rng.standard_normal((B, n, n)) + 1j * rng.standard_normal((B, n, n))
        A = A + A.conj().transpose(0, 2, 1)
        t0 = time.time()
        E, U = cb.debug_heev(A.transpose(0, 2, 1))
        dt = time.time() - t0
        err = res = orth = 0.0
        for b in range(B):
            Ub = U[b].T
            wr = np.linalg.eigvalsh(A[b])
            err = max(err, np.max(np.abs(E[b] - wr)) / np.max(np.abs(wr)))
            res = max(res, np.max(np.abs(A[b] @ Ub - Ub * E[b])) / np.max(np.abs(wr)))
            orth = max(orth, np.max(np.abs(Ub.conj().T @ Ub - np.eye(n))))
        print(f"  heev random: {dt*1e3:.1f} ms")
        line("heev random |E-Eref|/|E|", err, 1e-12)
        line("heev random residual", res, 1e-12)
        line("heev random orthogonality", orth, 1e-12)

    def s_stedc():
        d = rng.standard_normal((B, n)); e = rng.standard_normal((B, n - 1))
        w, Z = cb.debug_stedc(d, e)
        err = res = orth = 0.0
        for b in range(B):
            Zb = Z[b].T
            wr = sl.eigh_tridiagonal(d[b], e[b], eigvals_only=True)
            T = np.diag(d[b]) + np.diag(e[b], 1) + np.diag(e[b], -1)
            err = max(err, np.max(np.abs(w[b] - wr)))
            res = max(res, np.max(np.abs(T @ Zb - Zb * w[b])))
            orth = max(orth, np.max(np.abs(Zb.T @ Zb - np.eye(n))))
        line("stedc random |w-wref|", err, 1e-12)
        line("stedc random residual", res, 1e-12)
        line("stedc random orthogonality", orth, 1e-12)

    def s_H():
        cb.init_static_H(); cb.update_H_BdG()
        H = cb.get_H()
        err = max(np.max(np.abs(H[b].T - cs[b].H_base)) for b in range(B))
        line("H_base vs oracle", err, 0.0)

    def s_trid():
        d, e = cb.debug_tridiagonalize()
        err = 0.0
        for b in range(B):
            w = sl.eigh_tridiagonal(d[b], e[b], eigvals_only=True)
            err = max(err, np.max(np.abs(w - cs[b].E_n)))
        line("tridiagonal spectrum vs oracle E", err, 1e-12)

    def s_diag():
        cb.diagonalize_H_BdG()
        E = cb.get_eigenvalues(); U = cb.get_eigenvectors()
        err = res = orth = 0.0
        for b in range(B):
            Hf = orc.full_hermitian(cs[b]); Ub = U[b].T
            err = max(err, np.max(np.abs(E[b] - cs[b].E_n)))
            res = max(res, np.max(np.abs(Hf @ Ub - Ub * E[b])))
            orth = max(orth, np.max(np.abs(Ub.conj().T @ Ub - np.eye(n))))
        line("E_n vs oracle", err, 1e-12)
        line("BdG residual", res, 1e-12)
        line("BdG orthogonality", orth, 1e-12)

    def s_force():
        cb.compute_forces()
        F = cb.get_forces(); f = cb.get_fermi()
        errF = errf = 0.0
        for b in range(B):
            orc.compute_forces(cs[b], ps[b], sts[b])
            errF = max(errF, np.max(np.abs(F[b].T - cs[b].forces)) / np.max(np.abs(cs[b].forces)))
            errf = max(errf, np.max(np.abs(f[b] - cs[b].fermi_factors)))
        line("forces rel", errF, 1e-10)
        line("fermi factors", errf, 1e-13)
        pi = np.stack([orc.draw_momentum(ps[b], np.random.default_rng(5 + b)) for b in range(B)])
        cb.set_momentum(pi)
        Eg = cb.compute_total_energy()
        err = 0.0
        for b in range(B):
            sts[b].pi[...] = pi[b]
            Er = orc.compute_total_energy(cs[b], ps[b], sts[b])
            err = max(err, abs(Eg[b] - Er) / max(abs(Er), 1))
        line("total energy rel", err, 1e-12)
        O = cb.measure_observables()
        err = 0.0
        for b in range(B):
            Or = orc.measure_observables(cs[b], ps[b], sts[b])
            err = max(err, np.max(np.abs(O[b] - Or) / np.maximum(np.abs(Or), 1e-3)))
        line("observables rel", err, 1e-9)

    def s_sweep():
        Nt = np.array([3 + (b % 2) for b in range(B)], dtype=np.int32)
        dt = np.array([orc.calc_optimal_dt(ps[b].beta, ps[b].J, ps[b].mass, int(Nt[b])) for b in range(B)])
        for it in range(2):
            pi0 = np.stack([orc.draw_momentum(ps[b], np.random.default_rng(50 + 10 * it + b)) for b in range(B)])
            u = np.random.default_rng(70 + it).random(B)
            acc, dH = cb.hmc_sweep(Nt, dt, pi0=pi0, uniforms=u)
            errH = errD = 0.0; accok = True
            D = cb.get_field()
            for b in range(B):
                a_r, dH_r, Ho, Hn = orc.hmc_sweep(cs[b], ps[b], sts[b], Nt=int(Nt[b]), dt=float(dt[b]), pi0=pi0[b],
                                                  uniform=float(u[b]), return_energies=True)
                errH = max(errH, abs(dH[b] - dH_r) / max(abs(Ho), 1.0))
                errD = max(errD, np.max(np.abs(D[b].T - sts[b].Delta)))
                accok &= (bool(acc[b]) == a_r)
            print(f"  sweep {it}: dH = {dH}, acc = {acc}")
            line(f"sweep {it} dH vs oracle / |H|", errH, 1e-10)
            line(f"sweep {it} Delta after", errD, 1e-10)
            line(f"sweep {it} accept decisions differ", 0.0 if accok else 1.0, 0.0)
        E = cb.get_eigenvalues()
        err = max(np.max(np.abs(E[b] - cs[b].E_n)) for b in range(B))
        line("E_n after sweeps", err, 1e-10)

    for s in (s_heev, s_stedc, s_H, s_trid, s_diag, s_force, s_sweep):
        print(f" -- {s.__name__}", flush=True)
        stage(s)
    cb.close()


if __name__ == "__main__":
    print(dwhmc.version())
    sizes = [int(a) for a in sys.argv[1:]] or [4, 8]
    for L in sizes:
        run(L, L)
    if not sys.argv[1:]:
        run(6, 10)
