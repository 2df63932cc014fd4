"""Small end-to-end run for compute-sanitizer (memcheck): L=6x10 lattice (n = 120, odd merge sizes), 3 chains."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200"))
import dwhmc  # noqa: E402

for (Lx, Ly, B) in ((6, 10, 3), (5, 13, 4), (16, 16, 4)):
    N = Lx * Ly
    rng = np.random.default_rng(1)
    cb = dwhmc.ChainBatch(B, Lx, Ly)
    cb.set_params(1.0, -0.35, -1.08, np.linspace(2, 50, B), 0.8, 1.0)
    cb.set_disorder((rng.random((B, N)) < 0.05) * 1.0)
    cb.set_field(((rng.random((B, 2, N)) - 0.5) + 1j * (rng.random((B, 2, N)) - 0.5)) * 0.1)
    cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
    nt = np.arange(B, dtype=np.int32) % 3 + 2
    acc, dH = cb.hmc_sweep(nt, 0.05)
    nacc, dH2, obs = cb.run_sweeps(2, nt, 0.05, observables=True)
    print(Lx, Ly, "dH", dH, "obs finite", np.isfinite(obs).all())
    cb.close()
print("done")
