"""Timing of measure_transport_and_spectra at the bench shape: python tools/transport_time.py L B"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200"))
import dwhmc
L, B = int(sys.argv[1]), int(sys.argv[2])
N = L * L
rng = np.random.default_rng(0)
cb = dwhmc.ChainBatch(B, L, L)
cb.set_params(1.0, -0.35, -1.08, 1.0 / 10.0 ** np.linspace(-4, 3, B), 0.8, 1.0)
w = np.zeros((B, N))
for b in range(B):
    w[b, rng.permutation(N)[:int(np.rint(N * 0.05))]] = 1.0
cb.set_disorder(w)
cb.set_field(((rng.random((B, 2, N)) - 0.5) + 1j * (rng.random((B, 2, N)) - 0.5)) * 0.1)
cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG(); cb.measure_observables()
eta = 8.0 / N
r = cb.measure_transport_and_spectra(eta, 0.2 * eta, 4.0)
t0 = time.time(); r = cb.measure_transport_and_spectra(eta, 0.2 * eta, 4.0); t1 = time.time() - t0
print(f"measure_transport_and_spectra: {t1*1e3:.1f} ms for {B} chains (n_omega {len(r['omega_grid'])}, n_dos {len(r['dos_omega_grid'])}); "
      f"stiffness[:3] {r['superfluid_stiffness'][:3]}")
