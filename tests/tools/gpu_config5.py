"""BASELINE.json configs[4] (Gram-build stress): synthetic Gaussian DAG, 5,000 nodes x 1,000,000 samples
(X = 40 GB FP64), sample axis sharded over the GPUs in 8 logical row blocks, chains sharded over the GPUs.
Run under torchrun (or plain python for one GPU).  Env: P, N, CHAINS (total), ITERS."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesnetworks_b200.dist import blocks_of_rank, context_row_sharded, row_blocks, shard_chains
from bayesnetworks_b200.synth import chain_seeds, make_dag, make_prior, simulate_torch

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
P, N = int(os.environ.get("P", 5000)), int(os.environ.get("N", 1000000))
chains, iters, MP = int(os.environ.get("CHAINS", 512)), int(os.environ.get("ITERS", 20000)), 8
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
dag = make_dag(P, seed=42); g = make_prior(dag, max_par=MP, seed=43); nt = g.node_type_codes()
rb = row_blocks(N)
# every logical block is generated on the device that owns it (X never crosses PCIe); block b uses seed 42 + b
mine = [simulate_torch(dag, rb[b][1], seed=42 + b, device=dev) for b in blocks_of_rank(rank, world)]
torch.cuda.synchronize()

def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

times = []
for rep in range(3):
    barrier(); t0 = time.perf_counter()
    ctx, mean, gram, ms = context_row_sharded(mine, N, P, g.source, g.target, nt, rank, world, dev, max_par=MP)
    barrier(); t1 = time.perf_counter()
    times.append((t1 - t0, ms))
    if rep < 2: ctx.close()
wall, kms = min(times)
chk = float(gram.double().sum().item()), float(torch.diagonal(gram).sum().item())
first, count = shard_chains(chains, world, rank)
barrier(); t0 = time.perf_counter()
res, cms = ctx.run(n_chains=count, n_iter=iters, output=100, rng="wh", seeds=chain_seeds(count, first_chain=first))
barrier(); t1 = time.perf_counter()
valid = torch.tensor([sum(r.valid_iters for r in res), cms], dtype=torch.float64, device=dev)
if world > 1:
    v2 = valid.clone(); dist.all_reduce(valid[:1], op=dist.ReduceOp.SUM); dist.all_reduce(v2[1:], op=dist.ReduceOp.MAX); valid[1] = v2[1]
ctx.close()
if rank == 0:
    flops = 2.0 * N * P * P
    print(json.dumps({"config": f"{P} nodes x {N} samples, {chains} chains x {iters} iters, {world} GPU(s), 8 logical row blocks",
                      "gram_wall_s_incl_allgather": wall, "gram_kernel_ms_rank0": kms,
                      "gram_tflops_full_count_boxwide": flops / wall / 1e12,
                      "gram_checksum_sum_trace": chk,
                      "chains_wall_s": t1 - t0, "chain_kernel_ms_max": float(valid[1].item()),
                      "proposals_per_sec": float(valid[0].item()) / (t1 - t0),
                      "iters_per_sec": chains * iters / (t1 - t0)}), flush=True)
if world > 1:
    dist.destroy_process_group()
