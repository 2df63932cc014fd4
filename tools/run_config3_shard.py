"""BASELINE config 3 end to end on one GPU: the rank-0 shard (64 of the 512 chains: 32 T x 16 seeds, round-robin
over 8 ranks) of the L = 24 disordered temperature scan with the reference's control flow (adaptive thermalisation,
n_therm = 20, n_measure = 100, Nt_measure = 6, transport and spectra every sweep, bins of 10) and file layout.
Writes a per-chain summary CSV.  python tools/run_config3_shard.py OUT_DIR [n_measure] [world] [rank]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200"))
import dwhmc  # noqa: E402
from dwhmc.parallel import shard_chains  # noqa: E402

out = sys.argv[1]
n_measure = int(sys.argv[2]) if len(sys.argv) > 2 else 100
world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
rank = int(sys.argv[4]) if len(sys.argv) > 4 else 0
Ts = 10.0 ** np.linspace(-4, 3, 32)
n_seeds = 16
ids = shard_chains(len(Ts) * n_seeds, rank, world)
t0 = time.time()
res = dwhmc.batch_scan_T(out, Ts, n_seeds, Lx=24, Ly=24, n_therm=20, n_measure=n_measure, Nt_therm=20, Nt_measure=6,
                         measure_freq=1, bin_size=10, chain_ids=ids)
el = time.time() - t0
tab = res["table"]                                   # [sweep, chain, 12]
stiff = np.array([r[1] for r in res["transport"]])   # [sweep, chain]
dc = np.array([r[2] for r in res["transport"]])
rows = ["chain,T,seed,Nt_therm_final,acceptance,energy,energy_err,Delta_global,Delta_global_err,stiffness,stiffness_err,dc"]
for k, c in enumerate(ids):
    ip, sd = divmod(int(c), n_seeds)
    m = lambda a: (a.mean(), a.std(ddof=1) / np.sqrt(len(a)))
    e, ee = m(tab[:, k, 3]); g, ge = m(tab[:, k, 6]); s, se = m(stiff[:, k])
    rows.append(f"{c},{Ts[ip]:.6g},{sd},{res['Nt_therm_final'][k]},{res['acceptance'][k]:.3f},{e:.6f},{ee:.6f},{g:.6f},{ge:.6f},"
                f"{s:.6f},{se:.6f},{dc[:, k].mean():.6f}")
open(os.path.join(out, "summary.csv"), "w").write("\n".join(rows) + "\n")
sweeps = 20 + n_measure
print(f"config 3 shard: {len(ids)} chains, {sweeps} sweeps (+{n_measure} transport measurements) in {el:.1f} s "
      f"-> {len(ids) * sweeps / el:.1f} sweeps/s including adaptive thermalisation (Nt up to {res['Nt_therm_final'].max()}), "
      f"observables, transport, spectra and file output")
