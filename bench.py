#!/usr/bin/env python
"""bench.py -- HMC trajectories/s of the molecular-dynamics force path at L=24 (BASELINE.json).

One "step" = one hmc_sweep! (src/HMC.jl:71-144) over a batch of independent chains: momentum
refresh, H_old, Nt leapfrog steps (each: BdG assembly, 2N x 2N Hermitian eigendecomposition,
bond-correlator force, kick/drift), H_new, Metropolis, commit.  Workload = the per-GPU shard of
BASELINE config 3 (L=24 disordered T-scan, 32 T x 16 seeds = 512 chains over 8 GPUs = 64 chains
per GPU), Nt_measure = 6 and the physics of scripts/batch_scan_T.jl:10-36.  Weak scaling: every
rank owns 64 chains; there is no data-path collective (chains are independent), only an
end-of-run gather of the observables table.

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
  python bench.py --impl reference ...                         CPU oracle (the reference's algorithm,
                                                               LAPACK zheevr) on the host cores
Prints ONE JSON line (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200"))

PHYS = dict(t=1.0, tp=-0.35, mu=-1.08, W=1.0, n_imp=0.05, J=0.8, mass=1.0)   # scripts/batch_scan_T.jl:10-19
METRIC = "HMC trajectories/sec at L=24 (disordered T-scan shard, Nt=6)"
# Committed ncu evidence (profiles/): NOT measured by this script.  Only cited, with the file, in the keys whose
# name says so (`ncu_reference`); no `frac` in the JSON line is computed from these numbers.
NCU_REFERENCE = os.path.join("profiles", "r02_ncu_reference.json")


def temperatures(n_points=32):
    return 10.0 ** np.linspace(-4, 3, n_points)


def chain_setup(L, chain_ids, n_seeds=16):
    """(beta, disorder, Delta0) of the global chains `chain_ids` of the 32 T x 16 seed scan;
    seed = 3e6 + 1e3 * i_point + i_seed, NumPy PCG64."""
    N = L * L
    Ts = temperatures()
    betas, ws, Ds = [], [], []
    for c in chain_ids:
        ip, iseed = divmod(int(c), n_seeds)
        ip %= len(Ts)
        rng = np.random.Generator(np.random.PCG64(3_000_000 + 1000 * ip + iseed))
        w = np.zeros(N)
        w[rng.permutation(N)[:int(np.rint(N * PHYS["n_imp"]))]] = PHYS["W"]
        re, im = rng.random((N, 2)), rng.random((N, 2))
        betas.append(1.0 / Ts[ip]); ws.append(w); Ds.append((((re - 0.5) + 1j * (im - 0.5)) * 0.1).T)
    return np.array(betas), np.stack(ws), np.stack(Ds)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({nm for r in self.rows if len(r) >= 6 for nm, v in zip(names, r[2:6]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def fp64_peak_tflops(device, sustain_s=3.0):
    """No FP64 entry in MEASURED_PEAKS.json: measure cuBLAS DGEMM (torch.matmul fp64, 4096^3) live, the way the
    driver measures its bf16 peak: best of 10 single launches (burst) and back-to-back launches for `sustain_s`
    seconds (sustained: the honest denominator for a step that runs for seconds)."""
    import torch
    a = torch.randn(4096, 4096, dtype=torch.float64, device=device)
    b = torch.randn(4096, 4096, dtype=torch.float64, device=device)
    c = torch.empty_like(a)
    fl = 2 * 4096 ** 3
    for _ in range(2):
        torch.matmul(a, b, out=c)
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = max(best, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    reps = max(10, int(sustain_s * best * 1e12 / fl))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record(); torch.cuda.synchronize()
    sustained = reps * fl / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return best, sustained


def _oracle_trajectories(L, Nt, n_traj):
    """n_traj trajectories of one chain of the workload on the CPU oracle; returns elapsed seconds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dwhmc_oracle as orc
    beta, w, D0 = chain_setup(L, [8 * 16])          # a mid-scan temperature point
    p = orc.ModelParameters(L, L, PHYS["t"], PHYS["tp"], PHYS["mu"], PHYS["W"], PHYS["n_imp"], float(beta[0]), PHYS["J"],
                            PHYS["mass"])
    st = orc.SimulationState(w[0], D0[0].T.copy(), np.zeros((L * L, 2), complex))
    c = orc.initialize_cache(p)
    orc.init_static_H(c, p, st); orc.update_H_BdG(c, p, st); orc.diagonalize_H_BdG(c, p)
    rng = np.random.Generator(np.random.PCG64(1))
    dt = orc.calc_optimal_dt(p.beta, p.J, p.mass, Nt)
    t0 = time.perf_counter()
    for _ in range(n_traj):
        orc.hmc_sweep(c, p, st, Nt=Nt, dt=dt, rng=rng)
    return time.perf_counter() - t0


def cpu_oracle_sample(L, Nt, n_traj=1):
    """Time the CPU restatement (oracle = the reference's algorithm, LAPACK zheevr via SciPy) on a
    bounded sample of the workload, two ways: (i) as shipped -- one process, BLAS threads =
    all visible cores; (ii) throughput -- one single-threaded process per core, all at once.  Returns
    (best traj/s, cores, description)."""
    from threadpoolctl import threadpool_limits
    cores = len(os.sched_getaffinity(0))
    with threadpool_limits(limits=cores):            # torchrun exports OMP_NUM_THREADS=1; undo that here
        el = _oracle_trajectories(L, Nt, n_traj)
    shipped = n_traj / el
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker", "--L", str(L), "--nt", str(Nt)],
                              env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for _ in range(cores)]
    times = []
    for pr in procs:
        out, _ = pr.communicate()
        try:
            times.append(float(out.strip().splitlines()[-1]))
        except (ValueError, IndexError):
            pass
    multi = len(times) / max(times) if times else 0.0
    best = max(shipped, multi)
    cpu_oracle_sample.last = {"as_shipped_1proc_all_threads": shipped, "throughput_1thread_procs": multi}
    desc = (f"L={L}, Nt={Nt}, OpenBLAS zheevr: (i) 1 process x {cores} BLAS threads, {n_traj} trajectory(ies): "
            f"{shipped:.3f} traj/s; (ii) {len(times)} single-threaded processes x 1 trajectory each, concurrently: "
            f"{multi:.3f} traj/s; value = the faster")
    return best, cores, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L, Nt = args.L, args.nt
    if args.warmup > 0:
        _oracle_trajectories(min(L, 8), Nt, 1)       # warm the BLAS / page in SciPy; the sample itself is minutes-bounded
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, cores, desc = cpu_oracle_sample(L, Nt, 1)
        vals.append(v)
    el = time.perf_counter() - t0
    value = len(vals) / sum(1.0 / v for v in vals)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args),
           "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": cores, "kind": "port", "sample": desc},
           "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "Julia is not installed, so the reference cannot run; this arm times oracle/dwhmc_oracle.py, the "
                   "NumPy/SciPy restatement that calls the same LAPACK zheevr; each step = one bounded sample (see cpu_baseline.sample)",
           "wall_s": el}
    emit(out)


def workload_config(args):
    return {"workload": f"L={args.L} disordered T-scan shard (BASELINE config 3): {args.chains} chains/GPU "
                        f"(32 T x 16 seeds over 8 GPUs), n=2N={2 * args.L ** 2}, Nt={args.nt}, t'=-0.35 mu=-1.08 W=1 "
                        f"n_imp=0.05 J=0.8",
            "chains_per_gpu": args.chains, "L": args.L, "Nt": args.nt,
            "l2": "inputs larger than L2 (per-step working set > 5 GB vs 126 MB L2)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import dwhmc
    from dwhmc.parallel import gather_table, shard_chains

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L, B, Nt, K, W = args.L, args.chains, args.nt, args.steps, max(args.warmup, 3)
    N, n = L * L, 2 * L * L
    n_chains = B * world
    ids = shard_chains(n_chains, rank, world)
    beta, w, D0 = chain_setup(L, ids)
    cb = dwhmc.ChainBatch(B, L, L, device=local)
    cb.set_params(PHYS["t"], PHYS["tp"], PHYS["mu"], beta, PHYS["J"], PHYS["mass"])
    cb.set_disorder(w); cb.set_field(D0)
    cb.init_static_H(); cb.update_H_BdG(); cb.diagonalize_H_BdG()
    cb.seed(1234 + rank)
    dt = np.array([dwhmc.calc_optimal_dt(b, PHYS["J"], PHYS["mass"], Nt) for b in beta])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput: K sweeps, RNG on device, no host transfer inside
    cb.run_sweeps(W, Nt, dt)
    barrier()
    cb.reset_timers()
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        nacc, dH, _ = cb.run_sweeps(K, Nt, dt)
        ms_dev = cb.last_elapsed_ms()          # CUDA events on the library's stream
        barrier()
        wall = time.perf_counter() - t0
    launches = int(cb.timers()["launches"])
    ms_dev = max_over_ranks(ms_dev)
    value = n_chains * K / (ms_dev * 1e-3)

    # ---- end to end through the public batched API with host buffers
    rng = np.random.Generator(np.random.PCG64(99 + rank))
    pin = torch.empty((B, 2, N), dtype=torch.complex128).pin_memory()
    pi_host = pin.numpy()
    nt_arr = np.full(B, Nt, dtype=np.int32)
    h2d = pi_host.nbytes + 8 * B + 4 * B + 8 * B
    d2h = 4 * B + 8 * B + 8 * 9 * B
    sq = np.sqrt(PHYS["mass"])

    def e2e_step():
        pi_host[...] = (rng.standard_normal((B, 2, N)) + 1j * rng.standard_normal((B, 2, N))) * sq
        u = rng.random(B)
        acc, dHs = cb.hmc_sweep(nt_arr, dt, pi0=pi_host, uniforms=u)
        obs = cb.measure_observables()
        return acc, dHs, obs

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        acc, dHs, obs = e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = n_chains * K / e2e_s

    # ---- the same, with the per-sweep transport / spectra measurement every scan script runs (measure_freq = 1)
    eta = 8.0 / N                                      # scripts/batch_scan_T.jl:30-32
    tr = cb.measure_transport_and_spectra(eta, 0.2 * eta, 4.0)
    d2h_tr = sum(v.nbytes for k_, v in tr.items() if k_ not in ("omega_grid", "dos_omega_grid"))
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        acc, dHs, obs = e2e_step()
        tr = cb.measure_transport_and_spectra(eta, 0.2 * eta, 4.0)
    barrier()
    e2e_tr_s = max_over_ranks(time.perf_counter() - t0)
    e2e_tr_value = n_chains * K / e2e_tr_s

    # ---- stage breakdown for the roofline (separate profiled pass; event pairs per stage)
    cb.set_profiling(1); cb.reset_timers()
    cb.run_sweeps(1, Nt, dt)
    tm = cb.timers()
    cb.set_profiling(0)
    eig_ms = tm["tridiagonalize_ms"] + tm["stedc_ms"] + tm["backtransform_ms"]
    n_solves = tm["eigensolves"]                       # batched solves (each = B matrices)
    flops_per_solve = B * (40.0 / 3.0) * n ** 3        # 40/3 n^3 per eigendecomposition
    eig_tflops = n_solves * flops_per_solve / (eig_ms * 1e-3) / 1e12
    bw = cb.band_halfwidth()
    route = ("band (free dense->band stage: folded site order; position-owning bulge chase chase_sys_kernel, D&C dc_*_kernel + DMMA dc_gemm2_kernel, "
             "register-resident DMMA block reflectors band_apply2_kernel)" if bw else
             "dense (hetrd: hemv_reg_kernel + DMMA zgemm her2k, D&C, DMMA back-transformation)")
    stages = {("tridiagonalize (band route: position-owning bulge chase, chase_sys_kernel)" if bw else
               "tridiagonalize (dense route: hemv_reg_kernel + column steps + DMMA her2k)"): tm["tridiagonalize_ms"],
              "tridiagonal D&C": tm["stedc_ms"],
              "back-transformation": tm["backtransform_ms"]}
    dom_name = max(stages, key=stages.get)

    # ---- one chain at a time through the reference's own call sequence (scripts/test_simulation.jl:25-31)
    single = None
    if rank == 0:
        from dwhmc import reference_api as ra
        p1 = ra.ModelParameters(L, L, PHYS["t"], PHYS["tp"], PHYS["mu"], PHYS["W"], PHYS["n_imp"], float(beta[0]), PHYS["J"],
                                PHYS["mass"])
        st1 = ra.SimulationState(w[0].copy(), D0[0].T.copy(), np.zeros((N, 2), complex))
        c1 = ra.initialize_cache(p1, device=local)
        ra.init_static_H(c1, p1, st1); ra.update_H_BdG(c1, p1, st1); ra.diagonalize_H_BdG(c1, p1)
        r1 = np.random.Generator(np.random.PCG64(5))
        ra.hmc_sweep(c1, p1, st1, Nt=Nt, dt=float(dt[0]), rng=r1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n1 = 3
        for _ in range(n1):
            ra.hmc_sweep(c1, p1, st1, Nt=Nt, dt=float(dt[0]), rng=r1)
            ra.measure_observables(c1, p1, st1)
        single = n1 / (time.perf_counter() - t0)
        c1.batch.close()

    # ---- end-of-run gather of the observables table (the only collective of the run)
    table = gather_table(np.concatenate([dHs[:, None], acc[:, None].astype(float), obs], axis=1), ids, n_chains,
                         dist if world > 1 else None, dev if world > 1 else None)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        fp64_burst, fp64_sustained = fp64_peak_tflops(dev)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        cpu_v, cores, desc = cpu_oracle_sample(L, Nt, 2)
        cpu_modes = dict(cpu_oracle_sample.last)
        ncu_ref = None
        try:
            ncu_ref = json.load(open(os.path.join(ROOT, NCU_REFERENCE)))
        except (OSError, ValueError):
            pass
        dominant = {"kernel": dom_name, "stage_ms_per_batched_eigensolve": stages[dom_name] / n_solves,
                    "share_of_eigensolve": stages[dom_name] / eig_ms,
                    "bound": ("latency (a chain of barrier-separated phases per step; shared-memory pipe ~50 %, FP64 pipe ~25 %, L2 / HBM idle: "
                              "the blocks never leave the SM)" if bw else
                              "hbm (the trailing-matrix product A v of the one-stage reduction: 1 flop per byte)"),
                    "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None,
                    "note": "DRAM bytes cannot be measured outside a profiler; see ncu_reference for the committed capture"}
        if ncu_ref and bw and "chase_sys" in ncu_ref:
            c = ncu_ref["chase_sys"]
            gbs = c["dram_bytes"] / (c["duration_ms"] * 1e-3) / 1e9
            dominant["ncu_reference"] = dict(c, file=NCU_REFERENCE, achieved_gbs=gbs, frac_of_hbm_peak=gbs / hbm_peak,
                                             source="committed ncu --set full capture of this kernel, NOT measured in this run")
        out = {
            "metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args),
            "force_evals_per_s": value * (Nt + 1), "eigensolves_per_s": value * Nt,
            "acceptance": float(nacc.mean() / K), "wall_s_timed_region": wall,
            "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "ChainBatch.hmc_sweep(pi0, uniforms from pinned host) + measure_observables -> host"},
            "e2e_with_transport": {"value": e2e_tr_value, "unit": "trajectories/s", "h2d_bytes_per_step": h2d,
                                   "d2h_bytes_per_step": d2h + d2h_tr,
                                   "api": "as e2e, plus measure_transport_and_spectra (stiffness, sigma(omega), DOS, A(k, 0)) "
                                          "after every sweep, as every scan script of the reference does (measure_freq = 1)"},
            "single_chain": {"value": single, "unit": "trajectories/s", "chains": 1,
                             "api": "reference_api.hmc_sweep + measure_observables, one chain per handle: the call sequence of the "
                                    "unchanged reference scripts (scripts/test_simulation.jl:25-31)",
                             "cpu_1_process_all_threads": cpu_modes.get("as_shipped_1proc_all_threads"),
                             "note": "one chain cannot fill the GPU: small batches run the position-owning bulge chase "
                                     "(band_systolic.cu, 12 CTAs per chain at L = 24, one sweep per step time: ~10 ms per "
                                     "chase instead of ~31 with the sweep-owning kernel); batch chains (ChainBatch) for "
                                     "throughput"},
            "gpu_launches": launches,
            "clocks": clk.summary(),
            # the dense eigensolve (A3) sits on the FP64 tensor (DMMA) roofline, 40/3 n^3 flops per matrix
            "roofline": {"bound": "tensor", "kernel": "batched Hermitian eigensolve, " + route,
                         "half_bandwidth": bw, "achieved": eig_tflops, "peak": fp64_sustained, "unit": "TFLOP/s",
                         "frac": eig_tflops / fp64_sustained, "peak_burst": fp64_burst, "frac_of_burst_peak": eig_tflops / fp64_burst,
                         "traffic": None,
                         "peak_source": "measured live, cuBLAS DGEMM through torch.matmul fp64 4096^3: `peak` = back to back for "
                                        "3 s (sustained; the step runs for seconds), `peak_burst` = best of 10 single launches; "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "algorithmic_flops_per_launch": flops_per_solve,
                         "ms_per_batched_eigensolve": eig_ms / n_solves,
                         "share_of_step": eig_ms / (eig_ms + tm["assemble_ms"] + tm["force_ms"])},
            "roofline_dominant": dominant,
            # the DMMA kernel of the back-transformation on its own useful work (4 n^3 real flops per matrix with the
            # particle-hole half of the columns; the stage time includes the row un-permutation and the partner columns)
            "roofline_backtransform": {"bound": "tensor", "kernel": "band_apply2_kernel" if bw else "zgemm_dmma_kernel",
                                       "achieved": n_solves * B * 4.0 * n ** 3 / (tm["backtransform_ms"] * 1e-3) / 1e12,
                                       "peak": fp64_sustained, "unit": "TFLOP/s",
                                       "frac": n_solves * B * 4.0 * n ** 3 / (tm["backtransform_ms"] * 1e-3) / 1e12 / fp64_sustained},
            "shard_note": "weak scaling: rank r owns chains r, r+W, ... of the 512-chain scan (32 T x 16 seeds), so at N=1 the 64 "
                          "chains are the 4 lowest temperatures x 16 seeds and at N=8 every rank holds all 32 temperatures x 2 "
                          "seeds; the cost of a step does not depend on T, the acceptance at Nt=6 does (reported above)",
            "stage_ms_per_sweep": {k: v for k, v in tm.items() if k.endswith("_ms")},
            "cpu_baseline": {"value": cpu_v, "unit": "trajectories/s", "cores": cores, "kind": "port", "sample": desc},
            "gathered_table_shape": list(table.shape),
        }
        emit(out)
    cb.close()
    if world > 1:
        dist.destroy_process_group()


class _CleanStdout:
    """Everything libraries print to fd 1 while the bench runs (e.g. NCCL's version banner) goes to
    stderr; only the JSON line reaches stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


_JSON_LINES = []


def emit(obj):
    _JSON_LINES.append(json.dumps(obj))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=24)
    ap.add_argument("--chains", type=int, default=64, help="chains per GPU")
    ap.add_argument("--nt", type=int, default=6, help="leapfrog steps (Nt_measure, scripts/batch_scan_T.jl:33)")
    ap.add_argument("--cpu-worker", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_worker:
        print(_oracle_trajectories(args.L, args.nt, 1), flush=True)
        return
    with _CleanStdout():
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    for ln in _JSON_LINES:
        print(ln, flush=True)


if __name__ == "__main__":
    main()
