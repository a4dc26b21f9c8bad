"""Independent chains sharded over the GPUs of one box (one process per GPU).

The path partitions cleanly: chains share only the read-only Gram, prior and node types,
so every rank builds its own context and runs its block of chains with NO data-path
collective.  The only exchange is an all-gather of the fixed-capacity per-chain trace
blocks and counters (NCCL over NVLink on GPUs, gloo in the CPU tests).

Chain identity is GLOBAL: chain c gets the same seeds and therefore the same trajectory
whether the job runs on 1, 2, 4 or 8 GPUs (tests/test_dist_gloo.py, tests/test_gpu_*).

The one place with a real exchange step is the sufficient statistics of a matrix whose
sample axis is sharded over the GPUs (5,000 nodes x 1,000,000 samples = 40 GB): every rank
builds the partial centred Gram of its row blocks and the parts are all-gathered and added
in a FIXED block order (``sharded_sufficient_stats``), so the Gram -- and with it every
trajectory -- is bit-identical for 1, 2, 4 or 8 GPUs; an all-reduce would add in a
rank-count-dependent order.
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


def run_sharded_device(ctx, n_chains_total: int, n_iter: int, output: int, rank: int, world_size: int, device,
                       initial_network: int = 2, drop: int = 0, group=None, check: bool = True):
    """This rank's block of chains with the traces left in HBM, then ONE exchange step: the
    device-resident blocks are all-gathered over NCCL (NVLink) without touching the host.

    Returns a dict: ``ints`` int32 [n_chains_total, capacity, 7] and ``gll`` float64
    [n_chains_total, capacity] and ``n_rows`` int32 [n_chains_total] (CUDA tensors holding ALL
    chains in global order, identical on every rank), ``stats`` (this rank's bn_chain_stats),
    ``kernel_ms``, ``gather_ms`` (device time of the collectives) and ``gather_ok``: this rank's
    own block came back unchanged at its global position and the row counts agree."""
    import torch
    import torch.distributed as dist

    first, count = shard_chains(n_chains_total, world_size, rank)
    counts = [shard_chains(n_chains_total, world_size, r)[1] for r in range(world_size)]
    pad = max(counts)
    cap = max(1, (n_iter + output - 1) // output)
    # [7][pad][cap] so that the library's seven int columns are rows of ONE buffer
    ints = torch.zeros((len(INT_COLUMNS), pad, cap), dtype=torch.int32, device=device)
    gll = torch.zeros((pad, cap), dtype=torch.float64, device=device)
    n_rows = torch.zeros((pad,), dtype=torch.int32, device=device)
    stats, ms = None, 0.0
    if count > 0:
        if count != pad:   # the library lays the columns out for `count` chains: run into a view-sized buffer
            ints_run = torch.zeros((len(INT_COLUMNS), count, cap), dtype=torch.int32, device=device)
        else:
            ints_run = ints
        stats, ms = ctx.run_device(count, n_iter, output, ints_run.data_ptr(), gll.data_ptr(), n_rows.data_ptr(),
                                   initial_network=initial_network, drop=drop, rng="wh",
                                   seeds=chain_seeds(count, first_chain=first))
        if ints_run is not ints:
            ints[:, :count] = ints_run
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if world_size > 1:
        g_ints = torch.empty((world_size,) + tuple(ints.shape), dtype=ints.dtype, device=device)
        g_gll = torch.empty((world_size,) + tuple(gll.shape), dtype=gll.dtype, device=device)
        g_rows = torch.empty((world_size, pad), dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(g_ints, ints, group=group)
        dist.all_gather_into_tensor(g_gll, gll, group=group)
        dist.all_gather_into_tensor(g_rows, n_rows, group=group)
    else:
        g_ints, g_gll, g_rows = ints[None], gll[None], n_rows[None]
    e1.record()
    # global chain order (ranks own contiguous blocks; drop the padding)
    all_ints = torch.cat([g_ints[r, :, :counts[r]] for r in range(world_size)], dim=1).permute(1, 2, 0).contiguous()
    all_gll = torch.cat([g_gll[r, :counts[r]] for r in range(world_size)], dim=0)
    all_rows = torch.cat([g_rows[r, :counts[r]] for r in range(world_size)], dim=0)
    ok = True
    if check and count > 0:
        mine = ints[:, :count].permute(1, 2, 0)
        ok = bool(torch.equal(all_ints[first:first + count], mine) and torch.equal(all_gll[first:first + count], gll[:count])
                  and torch.equal(all_rows[first:first + count], n_rows[:count]))
    torch.cuda.synchronize(device)
    return dict(ints=all_ints, gll=all_gll, n_rows=all_rows, stats=stats, kernel_ms=ms,
                gather_ms=float(e0.elapsed_time(e1)), gather_ok=ok, first=first, count=count)


# ---------------------------------------------------------------------------
# row-sharded sufficient statistics
# ---------------------------------------------------------------------------
N_ROW_BLOCKS = 8  # fixed logical partition of the sample axis, whatever the GPU count


def row_blocks(n_samples: int, n_blocks: int = N_ROW_BLOCKS):
    """[(first_row, n_rows)] of the logical row blocks (multiples of 16 rows, last takes the rest)."""
    per = -(-n_samples // n_blocks)
    per = -(-per // 16) * 16
    out = []
    for b in range(n_blocks):
        lo = min(b * per, n_samples)
        hi = min(lo + per, n_samples)
        out.append((lo, hi - lo))
    return out


def blocks_of_rank(rank: int, world_size: int, n_blocks: int = N_ROW_BLOCKS):
    """Logical blocks owned by ``rank`` (contiguous; world_size must divide n_blocks)."""
    if n_blocks % world_size:
        raise ValueError(f"world_size {world_size} must divide the {n_blocks} row blocks")
    per = n_blocks // world_size
    return list(range(rank * per, (rank + 1) * per))


def _ordered_sum(parts):
    """parts[0] + parts[1] + ... in index order (elementwise, deterministic)."""
    acc = parts[0].clone()
    for b in range(1, parts.shape[0]):
        acc += parts[b]
    return acc


def sharded_sufficient_stats(local_blocks, n_samples: int, n_nodes: int, rank: int, world_size: int,
                             colsum_fn, gram_fn, device, group=None, n_blocks: int = N_ROW_BLOCKS):
    """Global column means and centred Gram from row blocks spread over the ranks.

    ``local_blocks``: this rank's blocks in logical order, whatever ``colsum_fn(block)`` ->
    tensor [P] and ``gram_fn(block, mean)`` -> tensor [P, P] understand (device tensors of
    shape (P, n_rows) on the GPU path; the CPU tests inject numpy-backed callables).
    Every rank returns the same (mean [P], centred Gram [P, P]) -- bit-identical for any
    world size because the parts are added in block order, not in rank-arrival order."""
    import torch
    import torch.distributed as dist

    mine = blocks_of_rank(rank, world_size, n_blocks)
    if len(local_blocks) != len(mine):
        raise ValueError(f"rank {rank} owns {len(mine)} row blocks, got {len(local_blocks)}")

    def gather(t):  # [n_local, ...] -> [n_blocks, ...] in logical block order
        if world_size == 1:
            return t
        out = torch.empty((n_blocks,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out

    sums = torch.stack([colsum_fn(b) for b in local_blocks]).to(device)
    mean = _ordered_sum(gather(sums)) / float(n_samples)
    parts = torch.stack([gram_fn(b, mean) for b in local_blocks]).to(device)
    gram = _ordered_sum(gather(parts))
    return mean, gram


def context_row_sharded(local_blocks, n_samples: int, n_nodes: int, graph_source, graph_target,
                        graph_node_type, rank: int, world_size: int, device, max_par=50, phi=1.0,
                        omega=6.9, group=None, n_blocks: int = N_ROW_BLOCKS):
    """Context whose sufficient statistics come from row blocks sharded over the GPUs.

    ``local_blocks``: CUDA float64 tensors of shape (n_nodes, n_rows_b) -- i.e. column-major
    n_rows_b x n_nodes, contiguous -- for the logical blocks ``blocks_of_rank(rank, world_size)``.
    Returns (Context, mean tensor, Gram tensor, ms spent in the Gram kernels of this rank)."""
    import torch

    from .api import Context, block_colsum_device, block_gram_device

    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(device).cuda_stream
    gram_ms = [0.0]

    def colsum_fn(blk):
        out = torch.empty(n_nodes, dtype=torch.float64, device=device)
        block_colsum_device(blk.data_ptr(), blk.stride(0), blk.shape[1], n_nodes, out.data_ptr(), dev_index, stream)
        return out

    def gram_fn(blk, mean):
        out = torch.empty((n_nodes, n_nodes), dtype=torch.float64, device=device)
        gram_ms[0] += block_gram_device(blk.data_ptr(), blk.stride(0), blk.shape[1], n_nodes, mean.data_ptr(),
                                        out.data_ptr(), dev_index, stream)
        return out

    mean, gram = sharded_sufficient_stats(local_blocks, n_samples, n_nodes, rank, world_size, colsum_fn, gram_fn,
                                          device, group=group, n_blocks=n_blocks)
    torch.cuda.synchronize(device)
    ctx = Context.from_stats_device(n_samples, n_nodes, mean.data_ptr(), gram.data_ptr(), graph_source,
                                    graph_target, graph_node_type, max_par=max_par, phi=phi, omega=omega,
                                    device=dev_index)
    return ctx, mean, gram, gram_ms[0]
