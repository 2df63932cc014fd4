"""One warm-up + one measured batched diagonalisation + force evaluation (profiling target):
python tools/prof_diag.py L B"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200"))
import dwhmc  # noqa: E402

L, B = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
N = L * L
rng = np.random.default_rng(0)
cb = dwhmc.ChainBatch(B, L, L)
cb.set_params(1.0, -0.35, -1.08, np.logspace(-1, 3, B), 0.8, 1.0)
w = np.zeros((B, N))
for b in range(B):
    w[b, rng.permutation(N)[:int(np.rint(N * 0.05))]] = 1.0
cb.set_disorder(w)
cb.set_field(((rng.random((B, 2, N)) - 0.5) + 1j * (rng.random((B, 2, N)) - 0.5)) * 0.1)
cb.init_static_H(); cb.update_H_BdG()
for _ in range(reps):
    cb.reset_timers()
    cb.diagonalize_H_BdG()
    cb.compute_forces()
print("launches per diagonalize+force:", cb.timers()["launches"])
