"""Race hunt: two handles driven from two host threads, compared with sequential runs of the same handles.
python tools/two_handles_check.py [reps] [L] [B]"""
import os, sys, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200"))
import dwhmc
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
L = int(sys.argv[2]) if len(sys.argv) > 2 else 12
B = int(sys.argv[3]) if len(sys.argv) > 3 else 40
Nt, N = 3, L * L

def make(seed):
    cb = dwhmc.ChainBatch(B, L, L)
    cb.set_params(1.0, -0.35, -1.08, np.linspace(2, 40, B), 0.8, 1.0)
    w = np.zeros((B, N)); w[:, :7] = 1.0
    cb.set_disorder(w)
    r = np.random.default_rng(seed)
    cb.set_field((r.random((B, 2, N)) - 0.5 + 1j * (r.random((B, 2, N)) - 0.5)) * 0.1)
    cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG(); cb.seed(seed)
    return cb

dt = np.full(B, 0.05)
ref = []
for seed in (1, 2):
    cb = make(seed)
    ref.append(cb.run_sweeps(1, Nt, dt)[1].copy())
    cb.close()
bad = 0
for rep in range(reps):
    cbs = [make(1), make(2)]
    out = [None, None]
    def work(i):
        out[i] = cbs[i].run_sweeps(1, Nt, dt)[1].copy()
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th: t.start()
    for t in th: t.join()
    for i in range(2):
        d = np.abs(out[i] - ref[i])
        if d.max() > 0:
            bad += 1
            print(f"rep {rep} handle {i}: max |dH - ref| = {d.max():.3e} at chain {int(d.argmax())}, chains differing {int((d > 0).sum())}, ref {ref[i][int(d.argmax())]:.6f}")
        cbs[i].close()
print(f"{bad} mismatching (rep, handle) of {2 * reps}")
