"""Independent chains sharded over the GPUs of one box (one process per GPU).

The path partitions cleanly: chains share only the read-only Gram, prior and node types,
so every rank builds its own context and runs its block of chains with NO data-path
collective.  The only exchange is an all-gather of the fixed-capacity per-chain trace
blocks and counters (NCCL over NVLink on GPUs, gloo in the CPU tests).

Chain identity is GLOBAL: chain c gets the same seeds and therefore the same trajectory
whether the job runs on 1, 2, 4 or 8 GPUs (tests/test_dist_gloo.py, tests/test_gpu_*).
"""
from __future__ import annotations

import numpy as np

from .synth import chain_seeds

INT_COLUMNS = ("iter", "ChangedNode", "movetype", "additions", "deletions", "FN", "FP")


def shard_chains(n_chains_total: int, world_size: int, rank: int):
    """Contiguous block of global chain indices owned by ``rank``: (first, count)."""
    base, rem = divmod(n_chains_total, world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def pack_results(results, capacity: int):
    """ChainResult list -> (int32 [n, capacity, 7], float64 [n, capacity], int64 [n, 12])."""
    n = len(results)
    ints = np.zeros((n, capacity, len(INT_COLUMNS)), dtype=np.int32)
    gll = np.zeros((n, capacity), dtype=np.float64)
    meta = np.zeros((n, 12), dtype=np.int64)
    for i, r in enumerate(results):
        rows = len(r.trace["iter"])
        for k, name in enumerate(INT_COLUMNS):
            ints[i, :rows, k] = r.trace[name]
        gll[i, :rows] = r.trace["globalLL"]
        meta[i, 0] = rows
        meta[i, 1] = r.uniforms
        meta[i, 2] = r.valid_iters
        meta[i, 3:6] = r.proposed
        meta[i, 6:9] = r.reject
        meta[i, 9] = r.total_edges
        meta[i, 10] = r.n_nonpd
        meta[i, 11] = r.alg_bytes
    return ints, gll, meta


def unpack_results(ints, gll, meta):
    """Inverse of :func:`pack_results`: list of dicts (trace columns + counters)."""
    out = []
    for i in range(ints.shape[0]):
        rows = int(meta[i, 0])
        tr = {name: ints[i, :rows, k].copy() for k, name in enumerate(INT_COLUMNS)}
        tr["globalLL"] = gll[i, :rows].copy()
        out.append(dict(trace=tr, uniforms=int(meta[i, 1]), valid_iters=int(meta[i, 2]),
                        proposed=tuple(int(x) for x in meta[i, 3:6]),
                        reject=tuple(int(x) for x in meta[i, 6:9]), total_edges=int(meta[i, 9]),
                        n_nonpd=int(meta[i, 10]), alg_bytes=int(meta[i, 11])))
    return out


def all_gather_chain_blocks(ints, gll, meta, n_chains_total: int, group=None, device=None):
    """All-gather the per-rank blocks; every rank returns the arrays of ALL chains in global
    order.  Ranks may own different chain counts: blocks are padded to the largest."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    counts = [shard_chains(n_chains_total, world, r)[1] for r in range(world)]
    pad = max(counts)

    def gather(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        if t.shape[0] < pad:
            t = torch.cat([t, torch.zeros((pad - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype)])
        if device is not None:
            t = t.to(device)
        out = torch.empty((world * pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        out = out.cpu().numpy().reshape((world, pad) + tuple(t.shape[1:]))
        return np.concatenate([out[r, :counts[r]] for r in range(world)], axis=0)

    return gather(ints), gather(gll), gather(meta)


def run_sharded(ctx, n_chains_total: int, n_iter: int, output: int, rank: int, world_size: int,
                initial_network: int = 2, drop: int = 0, gather: bool = True, group=None,
                device=None):
    """Run this rank's block of chains on ``ctx`` and (optionally) all-gather the traces.

    Returns (results_of_all_chains | local results, local kernel ms, local ChainResult list)."""
    first, count = shard_chains(n_chains_total, world_size, rank)
    local, ms = [], 0.0
    if count > 0:
        seeds = chain_seeds(count, first_chain=first)
        local, ms = ctx.run(n_chains=count, n_iter=n_iter, output=output,
                            initial_network=initial_network, drop=drop, rng="wh", seeds=seeds)
    if not gather or world_size == 1:
        return local, ms, local
    cap = max(1, (n_iter + output - 1) // output)
    ints, gll, meta = pack_results(local, cap)
    gi, gg, gm = all_gather_chain_blocks(ints, gll, meta, n_chains_total, group=group, device=device)
    return unpack_results(gi, gg, gm), ms, local
