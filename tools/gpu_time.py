"""Per-stage timing of the trajectory on the GPU box: python tools/gpu_time.py L B [Nt] [sweeps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200"))
import dwhmc  # noqa: E402

L, B = int(sys.argv[1]), int(sys.argv[2])
Nt = int(sys.argv[3]) if len(sys.argv) > 3 else 6
sweeps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
N, n = L * L, 2 * L * L
rng = np.random.default_rng(0)
cb = dwhmc.ChainBatch(B, L, L)
beta = np.logspace(-1, 3, B) if B > 1 else np.array([20.0])
cb.set_params(1.0, -0.35, -1.08, beta, 0.8, 1.0)
w = np.zeros((B, N))
for b in range(B):
    w[b, rng.permutation(N)[:int(np.rint(N * 0.05))]] = 1.0
cb.set_disorder(w)
cb.set_field(((rng.random((B, 2, N)) - 0.5) + 1j * (rng.random((B, 2, N)) - 0.5)) * 0.1)
cb.init_static_H(); cb.update_H_BdG()
t0 = time.time(); cb.diagonalize_H_BdG(); print(f"first diagonalize (cold): {time.time()-t0:.3f} s")
t0 = time.time(); cb.diagonalize_H_BdG(); t1 = time.time() - t0
print(f"diagonalize: {t1*1e3:.1f} ms for {B} matrices of n={n} -> {t1/B*1e3:.3f} ms/matrix, "
      f"{B*(40/3)*n**3/t1/1e12:.2f} TFLOP/s algorithmic")
E = cb.get_eigenvalues(); U = cb.get_eigenvectors()[0].T
print("sym err", np.max(np.abs(E + E[:, ::-1])), "orth", np.max(np.abs(U.conj().T @ U - np.eye(n))))
dt = np.array([dwhmc.calc_optimal_dt(bb, 0.8, 1.0, Nt) for bb in beta])
cb.set_profiling(1); cb.reset_timers()
t0 = time.time(); nacc, dH, _ = cb.run_sweeps(sweeps, Nt, dt); t1 = time.time() - t0
tm = cb.timers()
print(f"profiled: {sweeps} sweeps x {B} chains, Nt={Nt}: {t1:.3f} s; timers {tm}")
cb.set_profiling(2); cb.reset_timers(); cb.diagonalize_H_BdG()
print("hemv-only ms per batched solve (serial groups):", cb.timers()["hemv_ms"])
cb.set_profiling(0); cb.reset_timers()
t0 = time.time(); nacc, dH, _ = cb.run_sweeps(sweeps, Nt, dt); t1 = time.time() - t0
print(f"unprofiled: {t1:.3f} s -> {sweeps*B/t1:.2f} traj/s, {sweeps*B*Nt/t1:.1f} eigensolves/s, "
      f"{sweeps*B*Nt*(40/3)*n**3/t1/1e12:.2f} TFLOP/s algorithmic; acc {nacc.mean()/sweeps:.2f}; dH {dH[:4]}")
print("launches", cb.timers()["launches"])
