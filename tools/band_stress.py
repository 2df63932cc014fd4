"""Numerical stress of the default eigensolver route at full size: clean lattice (massive degeneracy), zero field,
large fields, strong disorder.  python tools/band_stress.py [L]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200"))
import dwhmc
L = int(sys.argv[1]) if len(sys.argv) > 1 else 24
N, n = L * L, 2 * L * L
rng = np.random.default_rng(5)
cases = {
    "clean, zero field": (np.zeros(N), np.zeros((2, N), complex)),
    "clean, uniform d-wave field": (np.zeros(N), np.stack([np.full(N, 0.3 + 0j), np.full(N, -0.3 + 0j)])),
    "clean, tiny random field": (np.zeros(N), (rng.random((2, N)) - 0.5 + 1j * (rng.random((2, N)) - 0.5)) * 1e-9),
    "strong disorder W=10, 30 %": (np.where(rng.random(N) < 0.3, 10.0, 0.0), (rng.random((2, N)) - 0.5 + 1j * (rng.random((2, N)) - 0.5)) * 0.1),
    "large field |Delta| ~ 5": (np.zeros(N), (rng.random((2, N)) - 0.5 + 1j * (rng.random((2, N)) - 0.5)) * 10.0),
}
B = len(cases)
cb = dwhmc.ChainBatch(B, L, L)
print("route: half-bandwidth", cb.band_halfwidth())
cb.set_params(1.0, -0.35, -1.08, 50.0, 0.8, 1.0)
cb.set_disorder(np.stack([c[0] for c in cases.values()]))
cb.set_field(np.stack([c[1] for c in cases.values()]))
cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
E, U, Hu = cb.get_eigenvalues(), cb.get_eigenvectors(), cb.get_H()
ok = True
for b, name in enumerate(cases):
    H = Hu[b].T; Hf = np.triu(H) + np.triu(H, 1).conj().T
    Ub = U[b].T
    wr = np.linalg.eigvalsh(Hf)
    nrm = max(np.max(np.abs(wr)), 1e-300)
    e1, e2, e3 = np.max(np.abs(E[b] - wr)) / nrm, np.max(np.abs(Hf @ Ub - Ub * E[b])) / nrm, np.max(np.abs(Ub.conj().T @ Ub - np.eye(n)))
    print(f"{name:32s} |E-E_lapack|/|E| {e1:.1e}  residual {e2:.1e}  unitarity {e3:.1e}")
    ok &= e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12
print("OK" if ok else "FAILED")
