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
