"""BASELINE.json configs[0]/[1]: the shipped `network` dataset (81 nodes x 2,000 samples), bn_mcmc(N=50000),
MaxPar 50 (the R default) and 8, one chain and 64 chains; wall and kernel time on one GPU."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesnetworks_b200 import Context
z = np.load(os.path.join(os.path.dirname(__file__), "..", "golden", "network_p3sim8.npz"))
g = np.load(os.path.join(os.path.dirname(__file__), "..", "golden", "golden_ref.npz"))
for mp in (50, 8):
    for chains in (1, 64):
        t0 = time.perf_counter()
        with Context.from_data(z["X"], z["source"], z["target"], z["node_type"], max_par=mp) as ctx:
            t1 = time.perf_counter()
            res, ms = ctx.run(n_chains=chains, n_iter=50000, output=100, rng="rmt", seeds=[[1234, 0, 0]] * chains)
            t2 = time.perf_counter()
        ok = np.array_equal(res[0].trace["ChangedNode"], g["cfg1_ChangedNode"])
        print(f"MaxPar {mp:2d}, {chains:2d} chain(s) x 50000 iters: create {1e3*(t1-t0):.1f} ms, run wall {1e3*(t2-t1):.1f} ms "
              f"(kernel {ms:.1f} ms) -> {chains*50000/(t2-t1):.3g} iters/s; trace == reference golden: {ok}", flush=True)
