"""Race hunt for the bulge chase: handle A tridiagonalises the same matrices again and again while handle B runs sweeps
from another host thread; every (d, e) of A is compared bit for bit with a quiet run.
python tools/chase_race_check.py [reps] [L] [B]"""
import os, sys, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200"))
import dwhmc
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
L = int(sys.argv[2]) if len(sys.argv) > 2 else 12
B = int(sys.argv[3]) if len(sys.argv) > 3 else 12
N = L * L

def make(seed, B):
    cb = dwhmc.ChainBatch(B, L, L)
    cb.set_params(1.0, -0.35, -1.08, np.linspace(2, 40, B), 0.8, 1.0)
    w = np.zeros((B, N)); w[:, :7] = 1.0
    cb.set_disorder(w)
    r = np.random.default_rng(seed)
    cb.set_field((r.random((B, 2, N)) - 0.5 + 1j * (r.random((B, 2, N)) - 0.5)) * 0.1)
    cb.init_static_H(); cb.update_H_BdG()
    return cb

a = make(1, B)
d0, e0 = a.debug_tridiagonalize()
d0, e0 = d0.copy(), e0.copy()
for _ in range(5):
    d, e = a.debug_tridiagonalize()
    assert np.array_equal(d, d0) and np.array_equal(e, e0), "not reproducible even when quiet"
b = make(2, 40)
b.diagonalize_H_BdG(); b.seed(2)
stop = False
def noise():
    dt = np.full(40, 0.05)
    while not stop:
        b.run_sweeps(1, 3, dt)
th = threading.Thread(target=noise); th.start()
bad = 0
n = 2 * N
for rep in range(reps):
    d, e = a.debug_tridiagonalize()
    if not (np.array_equal(d, d0) and np.array_equal(e, e0)):
        bad += 1
        for c in range(B):
            dd = np.nonzero((d[c] != d0[c]) | np.append(e[c] != e0[c], False))[0]
            if len(dd):
                print(f"rep {rep} chain {c}: first differing index {dd[0]} of {n} (count {len(dd)}), |dd| max {np.abs(d[c]-d0[c]).max():.2e}, |de| max {np.abs(e[c]-e0[c]).max():.2e}")
stop = True; th.join()
print(f"{bad} of {reps} tridiagonalisations differ")
