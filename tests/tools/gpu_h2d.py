import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesnetworks_b200 import Context
from bayesnetworks_b200.synth import make_dag, make_prior
P, N, MP = 1000, 100000, 8
dag = make_dag(P, seed=42); g = make_prior(dag, max_par=MP, seed=43); nt = g.node_type_codes()
Xh = torch.randn((P, N), dtype=torch.float64).pin_memory()
Xd = torch.empty((P, N), dtype=torch.float64, device="cuda")
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); Xd.copy_(Xh, non_blocking=True); torch.cuda.synchronize()
    print(f"torch H2D copy {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
for rep in range(8):
    t0 = time.perf_counter()
    c2 = Context.from_data(Xh.numpy().T, g.source, g.target, nt, max_par=MP)
    t1 = time.perf_counter(); gm = c2.gram_ms; c2.close(); t2 = time.perf_counter()
    print(f"from_data wall {1e3*(t1-t0):.1f} ms (gram {gm:.2f}) close {1e3*(t2-t1):.1f} ms", flush=True)
