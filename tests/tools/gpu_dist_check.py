"""torchrun check of the multi-GPU path on real GPUs: chains sharded by global index, traces
all-gathered over NCCL; every rank must see the same trajectories a single GPU produces."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesnetworks_b200 import Context
from bayesnetworks_b200.dist import INT_COLUMNS, run_sharded
from bayesnetworks_b200.synth import chain_seeds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
z = np.load(os.path.join(os.path.dirname(__file__), "..", "golden", "network_p3sim8.npz"))
n_chains, n_iter, output = 2 * world + 1, 4000, 100    # uneven blocks on purpose
with Context.from_data(z["X"], z["source"], z["target"], z["node_type"], max_par=8, device=local) as ctx:
    allres, ms, localres = run_sharded(ctx, n_chains, n_iter, output, rank, world, device=torch.device("cuda", local))
    ok = len(allres) == n_chains
    if rank == 0:  # single-GPU truth for every chain
        truth, _ = ctx.run(n_chains=n_chains, n_iter=n_iter, output=output, rng="wh", seeds=chain_seeds(n_chains))
        for c in range(n_chains):
            for k in INT_COLUMNS:
                ok &= bool(np.array_equal(allres[c]["trace"][k], truth[c].trace[k]))
            ok &= bool(np.array_equal(allres[c]["trace"]["globalLL"], truth[c].trace["globalLL"]))
            ok &= allres[c]["uniforms"] == truth[c].uniforms
        g = np.load(os.path.join(os.path.dirname(__file__), "..", "golden", "golden_ref.npz"))
        ok &= bool(np.array_equal(allres[0]["trace"]["ChangedNode"], g["cfg2_ChangedNode"][:n_iter // output]))
# row-sharded Gram: every rank must end with the bits a single GPU produces from the same 8 blocks
from bayesnetworks_b200.dist import blocks_of_rank, context_row_sharded, row_blocks
dev = torch.device("cuda", local)
Xd = torch.from_numpy(np.ascontiguousarray(z["X"].T)).to(dev)
N, P = z["X"].shape
rb = row_blocks(N)
if 8 % world == 0:
    mine = [Xd[:, rb[b][0]:rb[b][0] + rb[b][1]].contiguous() for b in blocks_of_rank(rank, world)]
    cs, mean_s, gram_s, _ = context_row_sharded(mine, N, P, z["source"], z["target"], z["node_type"], rank, world,
                                                dev, max_par=8)
    r_sh = cs.run(n_iter=2000, output=100, rng="wh")[0][0]
    cs.close()
    if rank == 0:
        allb = [Xd[:, lo:lo + cnt].contiguous() for lo, cnt in rb]
        import torch.distributed as _d
        c1, mean_1, gram_1, _ = context_row_sharded(allb, N, P, z["source"], z["target"], z["node_type"], 0, 1, dev,
                                                    max_par=8)
        r_1 = c1.run(n_iter=2000, output=100, rng="wh")[0][0]
        c1.close()
        ok &= bool(torch.equal(gram_s, gram_1)) and bool(torch.equal(mean_s, mean_1))
        ok &= bool(np.array_equal(r_sh.trace["globalLL"], r_1.trace["globalLL"]))
# device-resident traces exchanged over NCCL without host staging (what bench.py times for N > 1)
from bayesnetworks_b200.dist import run_sharded_device
with Context.from_data(z["X"], z["source"], z["target"], z["node_type"], max_par=8, device=local) as ctx:
    out = run_sharded_device(ctx, n_chains, n_iter, output, rank, world, dev)
    ok &= bool(out["gather_ok"])
    for c in range(n_chains):   # every rank holds every chain's trace, equal to the host-staged gather
        rows = int(out["n_rows"][c].item())
        ok &= rows == len(allres[c]["trace"]["iter"])
        for k, name in enumerate(INT_COLUMNS):
            ok &= bool(np.array_equal(out["ints"][c, :rows, k].cpu().numpy(), allres[c]["trace"][name]))
        ok &= bool(np.array_equal(out["gll"][c, :rows].cpu().numpy(), allres[c]["trace"]["globalLL"]))
    gather_ms = out["gather_ms"]
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    msg = (f"dist check world={world} chains={n_chains} (uneven blocks): chains sharded by global index, host-staged "
           f"and device-resident NCCL all-gather, row-sharded Gram vs single GPU: {'OK' if t.item() == 1 else 'FAILED'}"
           f" (device gather {gather_ms:.3f} ms)")
    print(msg, flush=True)
    out_dir = os.path.join(os.path.dirname(__file__), "..", "..", "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"dist_check_n{world}.txt"), "w") as fh:
        fh.write(msg + "\n")
dist.destroy_process_group()
sys.exit(0 if t.item() == 1 else 1)
