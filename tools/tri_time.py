"""Timing experiments on the tridiagonalisation stage only: python tools/tri_time.py L B"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200"))
import dwhmc
L, B = int(sys.argv[1]), int(sys.argv[2])
N = L * L
rng = np.random.default_rng(0)
cb = dwhmc.ChainBatch(B, L, L)
cb.set_params(1.0, -0.35, -1.08, np.logspace(-1, 2, B), 0.8, 1.0)
cb.set_disorder(np.zeros((B, N)))
cb.set_field(((rng.random((B, 2, N)) - 0.5) + 1j * (rng.random((B, 2, N)) - 0.5)) * 0.1)
cb.init_static_H(); cb.update_H_BdG()
cb.debug_tridiagonalize()
t0 = time.time()
for _ in range(3):
    cb.debug_tridiagonalize()
print(f"tridiagonalise (+assemble, d2h): {(time.time()-t0)/3*1e3:.1f} ms")
