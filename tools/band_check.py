"""Band-route check on the GPU box: python tools/band_check.py L B   (L: side, or LxxLy)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200"))
import dwhmc
Lx, Ly = (int(v) for v in sys.argv[1].split("x")) if "x" in sys.argv[1] else (int(sys.argv[1]),) * 2
B = int(sys.argv[2])
N, n = Lx * Ly, 2 * Lx * Ly
rng = np.random.default_rng(0)
cb = dwhmc.ChainBatch(B, Lx, Ly)
print("route half-bandwidth", cb.band_halfwidth())
cb.set_params(1.0, -0.35, -1.08, np.logspace(-1, 2, B), 0.8, 1.0)
w = np.zeros((B, N))
for b in range(B):
    w[b, rng.permutation(N)[:int(np.rint(N * 0.05))]] = 1.0
cb.set_disorder(w)
cb.set_field(((rng.random((B, 2, N)) - 0.5) + 1j * (rng.random((B, 2, N)) - 0.5)) * 0.1)
cb.init_static_H(); cb.update_H_BdG()
d, e = cb.debug_tridiagonalize()
Hu = cb.get_H()
import scipy.linalg as sl
for b in range(min(B, 2)):
    H = Hu[b].T; Hf = np.triu(H) + np.triu(H, 1).conj().T
    wr = np.linalg.eigvalsh(Hf)
    wt = sl.eigh_tridiagonal(d[b], e[b], eigvals_only=True)
    print("chain", b, "tridiagonal eigenvalue err", np.max(np.abs(wt - wr)))
t0 = time.time(); cb.diagonalize_H_BdG(); print("first diag", time.time() - t0)
t0 = time.time(); cb.diagonalize_H_BdG(); t1 = time.time() - t0
print(f"diagonalize: {t1*1e3:.1f} ms for {B} -> {B*(40/3)*n**3/t1/1e12:.2f} TFLOP/s algorithmic")
E = cb.get_eigenvalues(); U = cb.get_eigenvectors()
for b in range(min(B, 2)):
    H = Hu[b].T; Hf = np.triu(H) + np.triu(H, 1).conj().T
    Ub = U[b].T
    print("chain", b, "E err", np.max(np.abs(E[b] - np.linalg.eigvalsh(Hf))), "res", np.max(np.abs(Hf @ Ub - Ub * E[b])),
          "orth", np.max(np.abs(Ub.conj().T @ Ub - np.eye(n))))
cb.set_profiling(1); cb.reset_timers(); cb.diagonalize_H_BdG(); print(cb.timers())
