"""Chain sharding across the GPUs of one box.  Chains -- (T or beta point, disorder seed, chain)
triples, the iterations of the reference's sequential scan loop (scripts/batch_scan_T.jl:54-74) --
are independent Markov chains, so the run needs no data-path collective: every rank owns a
contiguous-by-stride slice of the global chain list and advances it on its own GPU.  The only
exchange is the end-of-run gather of the per-sweep table (Sweep/Accepted/dH + 9 observables, the
columns of the reference's observables.csv, src/Simulation.jl:71-73) to rank 0.

torch.distributed is plumbing here (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_chains(n_chains: int, rank: int, world: int) -> np.ndarray:
    """Global chain ids owned by `rank`: round-robin, so neighbouring temperatures (similar
    adaptive step counts) spread over devices."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(rank, n_chains, world, dtype=np.int64)


def chain_grid(points, n_seeds: int):
    """The scan as a flat chain list: chain id c -> (point index, seed index), point-major."""
    pts = np.asarray(points, dtype=np.float64)
    ip, iseed = np.divmod(np.arange(len(pts) * n_seeds), n_seeds)
    return pts[ip], ip, iseed


def gather_table(local: np.ndarray, ids: np.ndarray, n_chains: int, dist=None, device=None):
    """All ranks contribute local[len(ids), ...]; returns the [n_chains, ...] table on every rank
    (all_gather of padded blocks; a few hundred KB at the scan sizes).  With dist=None (single
    process) it just scatters into place."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    out = np.zeros((n_chains,) + local.shape[1:], dtype=np.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out[ids] = local
        return out
    import torch
    world = dist.get_world_size()
    per = (n_chains + world - 1) // world
    pad = np.zeros((per,) + local.shape[1:], dtype=np.float64)
    pad[:len(ids)] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    bufs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(bufs, t)
    for r in range(world):
        rid = shard_chains(n_chains, r, world)
        out[rid] = bufs[r][:len(rid)].cpu().numpy()
    return out
