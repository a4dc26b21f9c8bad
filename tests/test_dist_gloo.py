"""N > 1 path on CPU: world_size-2 gloo run of the chain sharding + trace all-gather.
The per-rank compute is replaced by a deterministic stand-in keyed by the GLOBAL chain
index (the GPU tests cover the real kernels); what is tested is the host logic: block
partition, padding of uneven blocks, global ordering after the gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesnetworks_b200.dist import (all_gather_chain_blocks, pack_results, shard_chains,  # noqa: E402
                                     unpack_results)
from bayesnetworks_b200.synth import chain_seeds  # noqa: E402


class FakeResult:
    def __init__(self, chain, cap):
        rng = np.random.default_rng(1000 + chain)
        rows = int(rng.integers(1, cap + 1))
        self.trace = {k: rng.integers(0, 100, rows).astype(np.int32) for k in
                      ("iter", "ChangedNode", "movetype", "additions", "deletions", "FN", "FP")}
        self.trace["globalLL"] = rng.standard_normal(rows)
        seeds = chain_seeds(1, first_chain=chain)[0]
        self.uniforms, self.valid_iters = int(seeds[0]), int(seeds[1])
        self.proposed, self.reject = (0, chain, 2 * chain), (1, 2, 3)
        self.total_edges, self.n_nonpd, self.alg_bytes = chain, 0, 368 * chain


def _worker(rank, world, port, n_chains, cap, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_chains(n_chains, world, rank)
    local = [FakeResult(first + i, cap) for i in range(count)]
    ints, gll, meta = pack_results(local, cap)
    gi, gg, gm = all_gather_chain_blocks(ints, gll, meta, n_chains)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), gi=gi, gg=gg, gm=gm)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_chains_partition():
    for n, w in ((64, 1), (64, 2), (64, 8), (5, 2), (3, 4), (512, 8)):
        blocks = [shard_chains(n, w, r) for r in range(w)]
        assert sum(c for _, c in blocks) == n
        pos = 0
        for first, count in blocks:
            assert first == pos
            pos += count


@pytest.mark.parametrize("n_chains", [5, 8])
def test_all_gather_world2_gloo(tmp_path, n_chains):
    cap, world = 7, 2
    mp.spawn(_worker, args=(world, _free_port(), n_chains, cap, str(tmp_path)), nprocs=world, join=True)
    want_i, want_g, want_m = pack_results([FakeResult(c, cap) for c in range(n_chains)], cap)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["gi"], want_i) and np.array_equal(z["gg"], want_g)
        assert np.array_equal(z["gm"], want_m)
    res = unpack_results(want_i, want_g, want_m)
    assert len(res) == n_chains and res[3]["total_edges"] == 3
    assert np.array_equal(res[2]["trace"]["globalLL"], FakeResult(2, cap).trace["globalLL"])


# ---------------------------------------------------------------------------
# row-sharded sufficient statistics: fixed-order reduction over gloo
# ---------------------------------------------------------------------------
def _numpy_block_fns():
    import torch as _t

    def colsum_fn(blk):  # blk: (P, n_rows) float64 numpy
        return _t.from_numpy(blk.sum(axis=1))

    def gram_fn(blk, mean):
        xc = blk - mean.numpy()[:, None]
        return _t.from_numpy(xc @ xc.T)

    return colsum_fn, gram_fn


def _stats_data(n=1000, p=13):
    rng = np.random.default_rng(7)
    return np.ascontiguousarray((rng.standard_normal((p, n)) * 3.0 + rng.uniform(-50, 50, (p, 1))))


def _stats_worker(rank, world, port, out_dir):
    from bayesnetworks_b200.dist import blocks_of_rank, row_blocks, sharded_sufficient_stats
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X = _stats_data()
    p, n = X.shape
    rb = row_blocks(n)
    local = [np.ascontiguousarray(X[:, rb[b][0]:rb[b][0] + rb[b][1]]) for b in blocks_of_rank(rank, world)]
    colsum_fn, gram_fn = _numpy_block_fns()
    mean, gram = sharded_sufficient_stats(local, n, p, rank, world, colsum_fn, gram_fn, torch.device("cpu"))
    np.savez(os.path.join(out_dir, f"stats{rank}.npz"), mean=mean.numpy(), gram=gram.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_row_blocks_partition():
    from bayesnetworks_b200.dist import blocks_of_rank, row_blocks
    for n in (1000, 1_000_000, 17, 128, 100_003):
        rb = row_blocks(n)
        assert len(rb) == 8 and sum(c for _, c in rb) == n
        pos = 0
        for lo, cnt in rb:
            assert lo == pos and (lo % 16 == 0 or cnt == 0)
            pos += cnt
    for w in (1, 2, 4, 8):
        assert sorted(b for r in range(w) for b in blocks_of_rank(r, w)) == list(range(8))
    with pytest.raises(ValueError):
        blocks_of_rank(0, 3)


def test_sharded_stats_world2_gloo_matches_single_process(tmp_path):
    """The Gram of the row-sharded path does not depend on the number of ranks (bit-identical)."""
    from bayesnetworks_b200.dist import row_blocks, sharded_sufficient_stats
    world = 2
    mp.spawn(_stats_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    X = _stats_data()
    p, n = X.shape
    rb = row_blocks(n)
    blocks = [np.ascontiguousarray(X[:, lo:lo + cnt]) for lo, cnt in rb]
    colsum_fn, gram_fn = _numpy_block_fns()
    mean1, gram1 = sharded_sufficient_stats(blocks, n, p, 0, 1, colsum_fn, gram_fn, torch.device("cpu"))
    for r in range(world):
        z = np.load(tmp_path / f"stats{r}.npz")
        assert np.array_equal(z["mean"], mean1.numpy()) and np.array_equal(z["gram"], gram1.numpy())
    # and it is the centred cross-product matrix
    xc = X - X.mean(axis=1, keepdims=True)
    assert np.allclose(gram1.numpy(), xc @ xc.T, rtol=1e-12, atol=1e-9)
    assert np.allclose(mean1.numpy(), X.mean(axis=1), rtol=1e-14)
