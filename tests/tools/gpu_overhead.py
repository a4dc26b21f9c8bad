"""Host-side overhead of one bench step: context creation, run wall vs kernel time, close, host-pointer path."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesnetworks_b200 import Context, set_default_stream
from bayesnetworks_b200.synth import chain_seeds, make_dag, make_prior, simulate_torch
P, N, MP, chains, iters = 1000, 100000, 8, 64, int(os.environ.get("ITERS", 100000))
dag = make_dag(P, seed=42); g = make_prior(dag, max_par=MP, seed=43); nt = g.node_type_codes()
X = simulate_torch(dag, N, seed=42, device="cuda"); torch.cuda.synchronize()
set_default_stream(torch.cuda.current_stream().cuda_stream)
seeds = chain_seeds(chains)
Xh = torch.empty((P, N), dtype=torch.float64, pin_memory=True); Xh.copy_(X); Xh_np = Xh.numpy().T
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(4):
    t0 = T(); ctx = Context.from_device(X.data_ptr(), N, N, P, g.source, g.target, nt, max_par=MP)
    t1 = T(); res, ms = ctx.run(n_chains=chains, n_iter=iters, output=100, rng="wh", seeds=seeds)
    gm = ctx.gram_ms; t2 = T(); ctx.close(); t3 = T()
    print(f"resident: create {1e3*(t1-t0):.2f} (gram {gm:.2f}) run {1e3*(t2-t1):.2f} (kernel {ms:.2f}) close {1e3*(t3-t2):.2f}  total {1e3*(t3-t0):.2f} ms", flush=True)
for rep in range(4):
    t0 = T(); ctx = Context.from_data(Xh_np, g.source, g.target, nt, max_par=MP)
    t1 = T(); res, ms = ctx.run(n_chains=chains, n_iter=iters, output=100, rng="wh", seeds=seeds)
    gm = ctx.gram_ms; t2 = T(); ctx.close(); t3 = T()
    print(f"host ptr: create {1e3*(t1-t0):.2f} (gram {gm:.2f}) run {1e3*(t2-t1):.2f} (kernel {ms:.2f}) close {1e3*(t3-t2):.2f}  total {1e3*(t3-t0):.2f} ms", flush=True)
