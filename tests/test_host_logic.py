"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/dwhmc.h
declares, it refuses to run without a CUDA device (no CPU fallback), the neighbour tables match the
oracle, and chain sharding / the end-of-run gather work at world_size 2 over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hybrid-monte-carlo-for-d-wave-sc_b200")


@pytest.fixture(scope="module")
def built():
    subprocess.check_call(["make", "-s", "-C", os.path.join(PKG, "csrc")])
    import dwhmc
    return dwhmc


def test_library_exports_every_header_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "dwhmc.h")).read()
    names = set(re.findall(r"\b(dwhmc_[a-zA-Z_]+)\s*\(", hdr))
    assert len(names) >= 30
    lib = ctypes.CDLL(built.LIB_PATH)
    for nm in sorted(names):
        assert hasattr(lib, nm), nm
    from dwhmc import _lib
    assert names == set(_lib.PROTOTYPES), names ^ set(_lib.PROTOTYPES)
    assert "sm_100a" in built.version()


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(built.DwhmcError) as ei:
        built.ChainBatch(1, 4, 4)
    assert ei.value.code == 4 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "dwhmc_oracle" not in src or f == "_never_", os.path.join(dirpath, f)


def test_neighbour_tables_match_oracle(built):
    import dwhmc_oracle as orc
    for Lx, Ly in ((4, 4), (6, 10), (3, 5)):
        nn, nnn = built.neighbour_tables(Lx, Ly)
        o_nn, o_nnn = orc.neighbour_tables(Lx, Ly)
        assert nn.dtype == np.int64 and nn.flags["F_CONTIGUOUS"]
        assert np.array_equal(nn - 1, o_nn) and np.array_equal(nnn - 1, o_nnn)
    assert built.calc_optimal_dt(20.0, 0.8, 1.0, 6) == orc.calc_optimal_dt(20.0, 0.8, 1.0, 6)


def test_shard_chains_partition(built):
    from dwhmc.parallel import chain_grid, shard_chains
    for n, w in ((512, 8), (7, 2), (3, 4)):
        ids = np.concatenate([shard_chains(n, r, w) for r in range(w)])
        assert sorted(ids) == list(range(n))
    pts, ip, iseed = chain_grid(np.logspace(-4, 3, 32), 16)
    assert len(pts) == 512 and ip[17] == 1 and iseed[17] == 1


WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from dwhmc.parallel import shard_chains, gather_table
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
n = 7
ids = shard_chains(n, dist.get_rank(), 2)
local = np.stack([np.full((3, 12), 100.0 * c) + np.arange(12) for c in ids])
tab = gather_table(local, ids, n, dist)
ref = np.stack([np.full((3, 12), 100.0 * c) + np.arange(12) for c in range(n)])
assert np.array_equal(tab, ref), tab
dist.destroy_process_group()
print("ok")
"""


def test_gather_world2_gloo(built, tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), PKG, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0 and b"ok" in out, out.decode()


def test_run_driver_formats_and_adaptive_Nt(built):
    """Host logic of the batched run driver: CSV row format (src/Simulation.jl:161-165), directory
    names (scripts/batch_scan_T.jl:62), adaptive-Nt rule (src/Simulation.jl:116-120)."""
    from dwhmc import simulation as sim
    line = sim.obs_csv_line(7, True, -1.23456789e-3, [0.1 * k for k in range(9)])
    assert line == "7,1,-1.23457e-03,0.000000,0.100000,0.200000,0.300000,0.400000,0.500000,0.600000,0.700000,0.800000\n"
    assert sim.OBS_HEADER.count(",") == 11 and line.count(",") == 11
    assert [sim.adapt_Nt(r, n) for r, n in ((0.4, 10), (0.6, 10), (0.96, 10), (1.0, 4), (1.0, 5))] == [12, 10, 9, 4, 4]
    for x, s in ((0.0001, "0.0001"), (0.00001, "1.0e-5"), (1000.0, "1000.0"), (0.000201, "0.000201"), (12.5, "12.5"),
                 (1.0e6, "1.0e6"), (123456.0, "123456.0"), (2.03e-5, "2.03e-5"), (0.5, "0.5"), (1e21, "1.0e21")):
        assert sim.julia_float_str(x) == s, (x, sim.julia_float_str(x))
    Ts = 10.0 ** np.linspace(-4, 3, 24)
    assert sim.scan_dir_T(Ts[0]) == "T_0.0001" and sim.scan_dir_T(Ts[-1]) == "T_1000.0"
    assert sim.scan_dir_T(Ts[1]) == "T_0.000202"
    assert sim.scan_dir_beta(0.01) == "beta_0.01" and sim.scan_dir_beta(100000.0) == "beta_100000.0"


def test_bench_reference_arm_contract(built):
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints exactly one JSON line on
    stdout with the contract's keys; small lattice so the CPU suite stays fast."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--L", "4", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "trajectories/s" and d["dtype"] == "f64" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]
