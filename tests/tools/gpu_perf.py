"""Quick perf probe on the named config (not the bench): Gram ms, H2D path wall, chain kernel ms."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesnetworks_b200 import Context
from bayesnetworks_b200.synth import chain_seeds, make_dag, make_prior, simulate_torch
P = int(os.environ.get("P", 1000)); N = int(os.environ.get("N", 100000)); MP = int(os.environ.get("MP", 8))
chains = int(os.environ.get("CHAINS", 64)); iters = int(os.environ.get("ITERS", 100000))
dag = make_dag(P, seed=42); g = make_prior(dag, max_par=MP, seed=43); nt = g.node_type_codes()
X = simulate_torch(dag, N, seed=42, device="cuda"); torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    ctx = Context.from_device(X.data_ptr(), N, N, P, g.source, g.target, nt, max_par=MP)
    t1 = time.perf_counter()
    print(f"from_device wall {1e3*(t1-t0):.1f} ms, gram_ms {ctx.gram_ms:.3f}", flush=True)
    if rep == 0: ctx.close()
if os.environ.get("H2D", "1") == "1":
    Xh = torch.empty((P, N), dtype=torch.float64, pin_memory=True); Xh.copy_(X)
    for rep in range(2):
        t0 = time.perf_counter()
        c2 = Context.from_data(Xh.numpy().T, g.source, g.target, nt, max_par=MP)
        t1 = time.perf_counter(); c2.close()
        print(f"from_data (pinned) wall {1e3*(t1-t0):.1f} ms", flush=True)
seeds = chain_seeds(chains)
for it in [iters]:
    for rep in range(int(os.environ.get("REPS", 3))):   # the first launch sees the clocks ramping up
        t0 = time.perf_counter()
        res, ms = ctx.run(n_chains=chains, n_iter=it, output=100, rng="wh", seeds=seeds)
        t1 = time.perf_counter()
    vi = sum(r.valid_iters for r in res); win = sum(r.windows for r in res)
    print(f"run {chains} chains x {it}: kernel {ms:.1f} ms wall {1e3*(t1-t0):.1f} ms -> {chains*it/ms/1e3:.3f} M iters/s, "
          f"{1e3*ms/it:.3f} us/iter/chain, valid {vi/(chains*it):.3f}, iters/window {chains*it/win:.1f}, "
          f"edges {np.mean([r.total_edges for r in res]):.0f}, accepted {np.mean([sum(r.proposed)-sum(r.reject[1:]) for r in res]):.0f}", flush=True)
    kc = np.mean([r.kernel_cycles for r in res]) / it
    print(f"kernel cycles/iter/chain {kc:.0f} (max chain {np.max([r.kernel_cycles for r in res]) / it:.0f}); implied SM clock {np.max([r.kernel_cycles for r in res]) / ms / 1e3:.0f} MHz", flush=True)
    cyc = np.array([r.phase_cycles for r in res], dtype=np.float64).mean(0)
    names = ["refill", "records", "walk+repair", "commit", "acc_add", "acc_del", "del:collect", "del:bar0", "del:roundwork", "del:list", "del:anc", "del:roundbar"]
    print("cycles/iter/chain: " + ", ".join(f"{n} {c/it:.0f}" for n, c in zip(names, cyc)) + f"  total {cyc.sum()/it:.0f}; slots simulated/iter {np.mean([r.slots_simulated for r in res])/it:.2f}", flush=True)
