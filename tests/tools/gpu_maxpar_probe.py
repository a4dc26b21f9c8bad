"""Why is MaxPar > 8 slower?  Parent-count distribution of the final graphs and per-phase cycles."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesnetworks_b200 import Context
from bayesnetworks_b200.synth import chain_seeds, make_dag, make_prior, simulate_torch
dag = make_dag(100, seed=42)
X = simulate_torch(dag, 10000, seed=42, device="cuda"); torch.cuda.synchronize()
for mp in (8, 12, 50):
    g = make_prior(dag, max_par=mp, seed=43); nt = g.node_type_codes()
    with Context.from_device(X.data_ptr(), 10000, 10000, 100, g.source, g.target, nt, max_par=mp) as ctx:
        ctx.run(n_chains=4, n_iter=1000, output=100, rng="wh", seeds=chain_seeds(4))   # warm (module load)
        res, ms = ctx.run(n_chains=4, n_iter=200000, output=100, rng="wh", seeds=chain_seeds(4))
    npar = np.stack([r.final_npar for r in res])
    cyc = np.array([r.phase_cycles for r in res], dtype=np.float64).mean(0) / 200000
    print(f"MaxPar {mp}: kernel {ms:.1f} ms, final npar max {npar.max()}, hist {np.bincount(npar.ravel(), minlength=13)[:13].tolist()}, "
          f"cycles/iter {cyc.round(0).tolist()}", flush=True)
