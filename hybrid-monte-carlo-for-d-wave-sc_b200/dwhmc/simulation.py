"""Batched run driver: run_simulation (src/Simulation.jl:34-236) for many chains at once.

The reference runs one chain per call and the scan scripts loop over temperatures sequentially
(scripts/batch_scan_T.jl:54-74).  Here every chain of the scan advances in lock-step on the GPU
(one ChainBatch), with the reference's per-chain control flow kept on the host:
  * adaptive thermalisation (:104-130): every 5 sweeps, acceptance < 0.60 -> Nt += 2,
    > 0.95 and Nt > 4 -> Nt -= 1, dt = calc_optimal_dt(beta, J, m, Nt) -- per chain;
  * measurement loop (:151-228): hmc_sweep!, measure_observables, one CSV row per sweep, flushed;
  * every measure_transport_freq sweeps: measure_transport_and_spectra (:168-223), one transport.csv row
    ("%d,%.6f,%.6f"), spectra averaged over bin_size measurements;
  * files per chain directory: simulation.log (append), observables.csv (truncate; 12 columns,
    "%d,%d,%.5e" + 9 x ",%.6f"), transport.csv, spectra_bins.npz.
spectra_bins.npz stands in for the reference's spectra_bins.jld2 (JLD2 / HDF5 writers are not available to
this host package): same keys -- "omega_grid", "dos_omega_grid", "params" (the ModelParameters scalars) and per bin
"sweep_<i>/opt_cond", "/dos", "/dos_AN", "/A_k0", "/count"; INTEGRATION.md has the NPZ -> JLD2 snippet."""
from __future__ import annotations

import datetime
import math
import os
from typing import Sequence

import numpy as np

from .batch import ChainBatch, calc_optimal_dt
from .reference_api import ModelParameters, initialize_state

OBS_HEADER = ("Sweep,Accepted,dH,Energy,Delta_Amp,Delta_Loc,Delta_Glob,S_Delta,Hole_p,Delta_Diff,Delta_Pair,"
              "Delta_LocalPair")                                   # src/Simulation.jl:71
TRANS_HEADER = "Sweep,Superfluid_Stiffness,DC_Conductivity"       # src/Simulation.jl:73
THERM_WINDOW = 5                                                   # src/Simulation.jl:99


def obs_csv_line(sweep: int, accepted: bool, dH: float, obs: Sequence[float]) -> str:
    """@sprintf("%d,%d,%.5e,%.6f x 9\\n") of src/Simulation.jl:161-165 (Julia prints a two-digit exponent
    like C printf)."""
    return "%d,%d,%.5e," % (sweep, int(accepted), dH) + ",".join("%.6f" % v for v in obs) + "\n"


def adapt_Nt(rate: float, Nt: int) -> int:
    """src/Simulation.jl:116-120."""
    if rate < 0.60:
        return Nt + 2
    if rate > 0.95 and Nt > 4:
        return Nt - 1
    return Nt


def julia_float_str(x: float) -> str:
    """string(::Float64) as Julia prints it: shortest round-trip digits, fixed notation for
    1e-4 <= |x| < 1e6 (always with a decimal point), otherwise `d.ddde-5` style.  Used for the
    directory names T_$(round(T, sigdigits=3)) / beta_$(round(beta, digits=3))."""
    x = float(x)
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    if not math.isfinite(x):
        return "NaN" if x != x else ("Inf" if x > 0 else "-Inf")
    from decimal import Decimal
    sign, digits, exp = Decimal(repr(x)).as_tuple()
    digits = list(digits)
    while len(digits) > 1 and digits[-1] == 0:       # strip trailing zeros
        digits.pop(); exp += 1
    nd = len(digits)
    e10 = nd - 1 + exp                                # x = d.ddd * 10^e10
    ds = "".join(map(str, digits))
    neg = "-" if sign else ""
    if -5 < e10 < 6 and abs(x) >= 1e-4:
        if e10 >= nd - 1:
            return neg + ds + "0" * (e10 - nd + 1) + ".0"
        if e10 >= 0:
            return neg + ds[:e10 + 1] + "." + ds[e10 + 1:]
        return neg + "0." + "0" * (-e10 - 1) + ds
    return neg + ds[0] + "." + (ds[1:] or "0") + "e" + str(e10)


def round_sigdigits(x: float, sig: int) -> float:
    """round(x, sigdigits=sig)."""
    if x == 0 or not math.isfinite(x):
        return x
    return float(f"{x:.{sig - 1}e}")


def scan_dir_T(T: float) -> str:
    """scripts/batch_scan_T.jl:62."""
    return "T_" + julia_float_str(round_sigdigits(T, 3))


def scan_dir_beta(beta: float) -> str:
    """scripts/batch_scan_beta.jl: "beta_$(round(beta, digits=3))"."""
    return "beta_" + julia_float_str(round(beta, 3))


class _ChainFiles:
    def __init__(self, out_dir: str, verbose: bool):
        os.makedirs(out_dir, exist_ok=True)
        self.log = open(os.path.join(out_dir, "simulation.log"), "a")
        self.obs = open(os.path.join(out_dir, "observables.csv"), "w")
        self.trans = open(os.path.join(out_dir, "transport.csv"), "w")
        self.verbose = verbose
        self.obs.write(OBS_HEADER + "\n")
        self.trans.write(TRANS_HEADER + "\n")
        self.trans.flush()

    def tee(self, msg: str):
        ts = datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S")
        line = f"[{ts}] {msg}"
        self.log.write(line + "\n")
        self.log.flush()
        if self.verbose:
            print(line)

    def close(self):
        for f in (self.log, self.obs, self.trans):
            f.close()


class _SpectraBins:
    """The binning of src/Simulation.jl:179-221 for one chain; `store` mirrors the JLD2 groups."""

    def __init__(self, path: str, p: ModelParameters, omega_grid, dos_grid):
        self.path, self.count, self.acc = path, 0, None
        self.store = {"omega_grid": np.asarray(omega_grid), "dos_omega_grid": np.asarray(dos_grid),
                      "params": np.array([p.Lx, p.Ly, p.t, p.tp, p.mu, p.W, p.n_imp, p.beta, p.J, p.mass, p.eta,
                                          p.d_omega, p.omega_max])}
        np.savez(self.path, **self.store)                       # jldsave(...; params, omega_grid) at :89

    def add(self, sweep: int, bin_size: int, opt_cond, dos, dos_AN, ak0):
        new = [np.array(a, dtype=np.float64) for a in (opt_cond, dos, dos_AN, ak0)]
        if self.count == 0:
            self.acc, self.count = new, 1
        else:
            for a, b in zip(self.acc, new):
                a += b
            self.count += 1
        if self.count >= bin_size:
            for a in self.acc:
                a /= self.count
            for key, a in zip(("opt_cond", "dos", "dos_AN", "A_k0"), self.acc):
                self.store[f"sweep_{sweep}/{key}"] = a
            self.store[f"sweep_{sweep}/count"] = np.array(self.count)
            np.savez(self.path, **self.store)
            self.count = 0


def run_simulation_batch(params: Sequence[ModelParameters], out_dirs: Sequence[str], *, n_therm: int = 100,
                         n_measure: int = 500, Nt_therm_init: int = 10, Nt_measure: int = 5,
                         measure_transport_freq: int = 0, bin_size: int = 5, device: int = 0,
                         seeds: Sequence[int] | None = None, rng_mode: str = "device", verbose: bool = False):
    """run_simulation for len(params) chains sharing one lattice, in lock-step on one GPU.
    rng_mode "device": Philox momenta/uniforms on the GPU (throughput); "host": NumPy generators per
    chain, momenta and lazily-consumed uniforms injected (reproducible against the CPU oracle).
    Returns a dict with the per-sweep table [n_measure, B, 12] (CSV columns) and final Nt per chain."""
    B = len(params)
    assert B == len(out_dirs) and B > 0
    p0 = params[0]
    assert all((p.Lx, p.Ly) == (p0.Lx, p0.Ly) for p in params), "chains of one batch share the lattice"
    N = p0.N
    seeds = list(seeds) if seeds is not None else list(range(B))
    rngs = [np.random.Generator(np.random.PCG64(s)) for s in seeds]
    files = [_ChainFiles(d, verbose) for d in out_dirs]
    for f, p in zip(files, params):
        f.tee("Starting Simulation...")
        f.tee(f"System: {p.Lx}x{p.Ly}, β={p.beta}, n_imp={p.n_imp}, J={p.J}")
        f.tee(f"Config: Therm={n_therm}, Sweep={n_measure}, TransFreq={measure_transport_freq}, BinSize={bin_size}"
              f" (B200 batch of {B} chains)")
        f.tee("Initializing State...")
    states = [initialize_state(p, r) for p, r in zip(params, rngs)]
    cb = ChainBatch(B, p0.Lx, p0.Ly, device=device, nn_table=p0.nn_table, nnn_table=p0.nnn_table)
    try:
        beta = np.array([p.beta for p in params]); J = np.array([p.J for p in params])
        mass = np.array([p.mass for p in params])
        cb.set_params([p.t for p in params], [p.tp for p in params], [p.mu for p in params], beta, J, mass)
        cb.set_disorder(np.stack([s.disorder_pot for s in states]))
        cb.set_field(np.stack([s.Delta for s in states]))
        cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
        cb.seed(int(seeds[0]) * 7919 + 17)

        def sweep(Nt, dt):
            if rng_mode == "device":
                return cb.hmc_sweep(Nt, dt)
            pi0 = np.stack([(r.standard_normal((N, 2)) + 1j * r.standard_normal((N, 2))) * math.sqrt(0.5) *
                            math.sqrt(2.0 * m) for r, m in zip(rngs, mass)])
            _, _, dH = cb.trajectory(Nt, dt, pi0=pi0)
            acc = np.zeros(B, dtype=bool)
            for b in range(B):                       # rand() only when dH >= 0 (src/HMC.jl:128)
                acc[b] = dH[b] < 0 or (rngs[b].random() < math.exp(-dH[b]) if math.isfinite(dH[b]) else False)
            cb.commit(acc)
            return acc, dH

        # --- thermalisation with per-chain adaptive Nt
        Nt = np.full(B, Nt_therm_init, dtype=np.int32)
        dt = np.array([calc_optimal_dt(beta[b], J[b], mass[b], int(Nt[b])) for b in range(B)])
        for f, b in zip(files, range(B)):
            f.tee("--- Thermalization Start ---")
            f.tee(f"Init: Nt={Nt[b]}, dt={round(dt[b], 5)}")
        recent = np.zeros(B, dtype=int)
        for i in range(1, n_therm + 1):
            acc, _ = sweep(Nt, dt)
            recent += acc
            if i % THERM_WINDOW == 0:
                for b in range(B):
                    rate = recent[b] / THERM_WINDOW
                    old = int(Nt[b])
                    new = adapt_Nt(rate, old)
                    if new != old:
                        Nt[b] = new
                        dt[b] = calc_optimal_dt(beta[b], J[b], mass[b], new)
                        files[b].tee("Therm %d/%d. Rate=%.2f. Adjust Nt: %d -> %d, dt: %.4f" % (i, n_therm, rate, old, new, dt[b]))
                    elif i % 20 == 0:
                        files[b].tee("Therm %d/%d. Rate=%.2f. Nt=%d (Stable)" % (i, n_therm, rate, old))
                recent[:] = 0
        Nt_final = Nt.copy()
        for f in files:
            f.tee("Thermalization Done.")

        # --- measurement
        Ntm = np.full(B, Nt_measure, dtype=np.int32)
        dtm = np.array([calc_optimal_dt(beta[b], J[b], mass[b], Nt_measure) for b in range(B)])
        for f, b in zip(files, range(B)):
            f.tee("--- Measurement Start ---")
            f.tee(f"Settings: Nt={Nt_measure}, dt={round(dtm[b], 5)}")
        table = np.zeros((n_measure, B, 12))
        acc_total = np.zeros(B, dtype=int)
        transport_rows, bins = [], None
        if measure_transport_freq > 0:
            assert all((p.eta, p.d_omega, p.omega_max) == (p0.eta, p0.d_omega, p0.omega_max) for p in params), \
                "chains of one batch share the frequency grids"
        for i in range(1, n_measure + 1):
            acc, dH = sweep(Ntm, dtm)
            acc_total += acc
            obs = cb.measure_observables()
            for b in range(B):
                files[b].obs.write(obs_csv_line(i, acc[b], dH[b], obs[b]))
                files[b].obs.flush()
                table[i - 1, b] = (i, acc[b], dH[b], *obs[b])
            if measure_transport_freq > 0 and i % measure_transport_freq == 0:
                sp = cb.measure_transport_and_spectra(p0.eta, p0.d_omega, p0.omega_max)
                if bins is None:
                    bins = [_SpectraBins(os.path.join(d, "spectra_bins.npz"), p, sp["omega_grid"], sp["dos_omega_grid"])
                            for d, p in zip(out_dirs, params)]
                transport_rows.append((i, sp["superfluid_stiffness"].copy(), sp["dc_conductivity"].copy()))
                for b in range(B):
                    files[b].trans.write("%d,%.6f,%.6f\n" % (i, sp["superfluid_stiffness"][b], sp["dc_conductivity"][b]))
                    files[b].trans.flush()
                    bins[b].add(i, bin_size, sp["optical_conductivity"][b], sp["dos"][b], sp["dos_AN"][b], sp["A_k_w0"][b])
            for b in range(B):
                if i % 10 == 0:
                    files[b].tee("Meas %d/%d. Acc=%.2f. E=%.4f" % (i, n_measure, acc_total[b] / i, obs[b, 0]))
        for f in files:
            f.tee("Measurement Done.")
        return {"table": table, "Nt_therm_final": Nt_final, "acceptance": acc_total / max(n_measure, 1),
                "field": cb.get_field(), "transport": transport_rows}
    finally:
        cb.close()
        for f in files:
            f.close()


def run_simulation(p: ModelParameters, out_dir: str, *, n_therm: int = 100, n_measure: int = 500,
                   Nt_therm_init: int = 10, Nt_measure: int = 5, measure_transport_freq: int = 1, bin_size: int = 5,
                   verbose: bool = True, seed: int = 0, device: int = 0, rng_mode: str = "device"):
    """Single-chain form with the reference's keyword names and defaults (src/Simulation.jl:34-41)."""
    return run_simulation_batch([p], [out_dir], n_therm=n_therm, n_measure=n_measure, Nt_therm_init=Nt_therm_init,
                                Nt_measure=Nt_measure, measure_transport_freq=measure_transport_freq, bin_size=bin_size,
                                device=device, seeds=[seed], rng_mode=rng_mode, verbose=verbose)


def batch_scan_T(base_dir: str, Ts: Sequence[float], n_seeds: int = 1, *, Lx: int = 24, Ly: int = 24, t=1.0, tp=-0.35,
                 mu=-1.08, W=1.0, n_imp=0.05, J=0.8, mass=1.0, n_therm=20, n_measure=100, Nt_therm=20, Nt_measure=6,
                 measure_freq: int = 1, bin_size: int = 10, device: int = 0, chain_ids: Sequence[int] | None = None, **kw):
    """scripts/batch_scan_T.jl as one batch: chain c = (temperature c // n_seeds, seed c % n_seeds);
    directory T_<round(T, sigdigits=3)> (plus /seed_<k> when n_seeds > 1).  `chain_ids` restricts the
    call to this rank's shard (dwhmc.parallel.shard_chains)."""
    ids = list(range(len(Ts) * n_seeds)) if chain_ids is None else [int(c) for c in chain_ids]
    ps, dirs, seeds = [], [], []
    for c in ids:
        ip, k = divmod(c, n_seeds)
        T = float(Ts[ip])
        ps.append(ModelParameters(Lx, Ly, t, tp, mu, W, n_imp, 1.0 / T, J, mass, eta=8.0 / (Lx * Ly),
                                  d_omega=0.2 * 8.0 / (Lx * Ly), omega_max=4.0))
        d = os.path.join(base_dir, scan_dir_T(T))
        dirs.append(d if n_seeds == 1 else os.path.join(d, f"seed_{k}"))
        seeds.append(1_000_000 * 3 + 1000 * ip + k)
    return run_simulation_batch(ps, dirs, n_therm=n_therm, n_measure=n_measure, Nt_therm_init=Nt_therm,
                                Nt_measure=Nt_measure, measure_transport_freq=measure_freq, bin_size=bin_size,
                                device=device, seeds=seeds, **kw)


def batch_scan_beta(base_dir: str, betas: Sequence[float], *, Lx: int = 12, Ly: int = 12, t=1.0, tp=-0.35, mu=-1.08,
                    W=1.0, n_imp=0.0, J=0.8, mass=1.0, n_therm=20, n_measure=100, Nt_therm=20, Nt_measure=6,
                    measure_freq: int = 1, bin_size: int = 10, device: int = 0,
                    chain_ids: Sequence[int] | None = None, **kw):
    """scripts/batch_scan_beta.jl:52-71 as one batch: one chain per beta, directory beta_<round(beta, digits=3)>;
    eta = 8/N, d_omega = 0.2 eta, omega_max = 4 (:15-17)."""
    ids = list(range(len(betas))) if chain_ids is None else [int(c) for c in chain_ids]
    ps, dirs, seeds = [], [], []
    for c in ids:
        beta = float(betas[c])
        ps.append(ModelParameters(Lx, Ly, t, tp, mu, W, n_imp, beta, J, mass, eta=8.0 / (Lx * Ly),
                                  d_omega=0.2 * 8.0 / (Lx * Ly), omega_max=4.0))
        dirs.append(os.path.join(base_dir, scan_dir_beta(beta)))
        seeds.append(4_000_000 + 1000 * c)
    return run_simulation_batch(ps, dirs, n_therm=n_therm, n_measure=n_measure, Nt_therm_init=Nt_therm,
                                Nt_measure=Nt_measure, measure_transport_freq=measure_freq, bin_size=bin_size,
                                device=device, seeds=seeds, **kw)


def scan_Nt_efficiency(Nt_list: Sequence[int] = (2, 3, 4, 5, 6, 8, 10, 12, 15, 20, 30), *, Lx: int = 10, Ly: int = 10,
                       t=1.0, tp=-0.35, mu=-1.08, W=3.0, n_imp=0.078, beta=20.0, J=0.8, mass=1.0, n_warmup: int = 100,
                       n_measure: int = 100, seed: int = 0, device: int = 0):
    """scripts/test_scan_Nt_efficiency.jl:19-62: acceptance and Acc/Nt against the number of leapfrog steps at a
    fixed trajectory length L = 2 pi sqrt(m J / beta) (dt = L / Nt).  The reference runs the Nt values one after
    another from fresh states; here they are the chains of one batch (same disorder / Delta0 seed for all).
    Returns (Nt, dt, acceptance, efficiency) arrays."""
    Nts = np.asarray(list(Nt_list), dtype=np.int32)
    B = len(Nts)
    p = ModelParameters(Lx, Ly, t, tp, mu, W, n_imp, beta, J, mass)
    L_target = 2.0 * math.pi * math.sqrt(mass * J / beta)          # T_period / 2, T_period = 4 pi sqrt(m J / beta)
    dt = L_target / Nts
    st = initialize_state(p, np.random.Generator(np.random.PCG64(seed)))
    with ChainBatch(B, Lx, Ly, device=device, nn_table=p.nn_table, nnn_table=p.nnn_table) as cb:
        cb.set_params(t, tp, mu, beta, J, mass)
        cb.set_disorder(np.tile(st.disorder_pot, (B, 1)))
        cb.set_field(np.tile(st.Delta.T[None], (B, 1, 1)))
        cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
        cb.seed(seed * 31 + 5)
        cb.run_sweeps(n_warmup, Nts, dt)
        nacc, _, _ = cb.run_sweeps(n_measure, Nts, dt)
    rate = nacc / max(n_measure, 1)
    return Nts, dt, rate, rate / Nts

