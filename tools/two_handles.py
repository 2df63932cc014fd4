"""Experiment: K handles x (B/K) chains driven concurrently from K host threads vs one handle x B.
python tools/two_handles.py L B K [Nt] [sweeps] [stagger_ms]"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200"))
import dwhmc  # noqa: E402

L, B, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
Nt = int(sys.argv[4]) if len(sys.argv) > 4 else 6
sweeps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
stagger = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
N = L * L
rng = np.random.default_rng(0)


def make(Bk, k):
    cb = dwhmc.ChainBatch(Bk, L, L)
    beta = np.logspace(-1, 3, Bk)
    cb.set_params(1.0, -0.35, -1.08, beta, 0.8, 1.0)
    w = np.zeros((Bk, N))
    for b in range(Bk):
        w[b, rng.permutation(N)[:int(np.rint(N * 0.05))]] = 1.0
    cb.set_disorder(w)
    cb.set_field(((rng.random((Bk, 2, N)) - 0.5) + 1j * (rng.random((Bk, 2, N)) - 0.5)) * 0.1)
    cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
    dt = np.array([dwhmc.calc_optimal_dt(bb, 0.8, 1.0, Nt) for bb in beta])
    cb.seed(1234 + k)
    return cb, dt


cbs = [make(B // K, k) for k in range(K)]
for cb, dt in cbs:
    cb.run_sweeps(1, Nt, dt)     # warm-up


def work(k):
    cb, dt = cbs[k]
    if stagger > 0:
        time.sleep(k * stagger * 1e-3)
    cb.run_sweeps(sweeps, Nt, dt)


t0 = time.time()
th = [threading.Thread(target=work, args=(k,)) for k in range(K)]
for t in th:
    t.start()
for t in th:
    t.join()
t1 = time.time() - t0
print(f"K={K} handles x {B//K} chains, L={L}, Nt={Nt}, {sweeps} sweeps, stagger {stagger} ms: {t1:.3f} s -> "
      f"{sweeps*B/t1:.2f} traj/s")
