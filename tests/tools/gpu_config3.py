"""BASELINE.json configs[2]: synthetic Gaussian DAG, 100 nodes x 10,000 samples, one chain, N = 1e6, output 100."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesnetworks_b200 import Context
from bayesnetworks_b200.synth import chain_seeds, make_dag, make_prior, simulate_torch
dag = make_dag(100, seed=42)
for mp in (8, 50):
    g = make_prior(dag, max_par=mp, seed=43); nt = g.node_type_codes()
    X = simulate_torch(dag, 10000, seed=42, device="cuda"); torch.cuda.synchronize()
    for chains in (1, 64):
        with Context.from_device(X.data_ptr(), 10000, 10000, 100, g.source, g.target, nt, max_par=mp) as ctx:
            t0 = time.perf_counter()
            res, ms = ctx.run(n_chains=chains, n_iter=1000000, output=100, rng="wh", seeds=chain_seeds(chains))
            t1 = time.perf_counter()
        r = res[0]
        print(f"MaxPar {mp:2d}, {chains:2d} chain(s) x 1e6 iters: gram {ctx.gram_ms if False else 0:.0f} run wall {1e3*(t1-t0):.0f} ms (kernel {ms:.0f} ms) -> "
              f"{chains*1e6/(t1-t0):.3g} iters/s, {sum(x.valid_iters for x in res)/(t1-t0):.3g} proposals/s; rows {len(r.trace['iter'])}, "
              f"edges {r.total_edges}, accepted {sum(r.proposed) - sum(r.reject[1:])}", flush=True)
