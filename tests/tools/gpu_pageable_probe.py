"""Developer tool: end-to-end step from a PAGEABLE host matrix (what R owns) for a few staging-thread counts."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesnetworks_b200 import Context
from bayesnetworks_b200.synth import chain_seeds, make_dag, make_prior, simulate_numpy
P, N = 1000, 100000
dag = make_dag(P, seed=42); g = make_prior(dag, max_par=8, seed=43); nt = g.node_type_codes()
X = np.asfortranarray(simulate_numpy(dag, N, seed=42))   # ordinary (pageable) memory, column-major like R's
for rep in range(5):
    t0 = time.perf_counter()
    with Context.from_data(X, g.source, g.target, nt, max_par=8) as ctx:
        t1 = time.perf_counter()
        res, ms = ctx.run(n_chains=64, n_iter=100000, output=100, rng="wh", seeds=chain_seeds(64))
    t2 = time.perf_counter()
    print(f"threads {os.environ.get('BN_B200_STAGE_THREADS', 'default')}: create {1e3*(t1-t0):.1f} ms, step {1e3*(t2-t0):.1f} ms (chain kernel {ms:.1f})", flush=True)
