"""Developer tool: first divergence between the two-CTA and the one-CTA chain kernels on a small case."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesnetworks_b200 import Context
from bayesnetworks_b200.synth import make_dag, make_prior, simulate_numpy

P = int(os.environ.get("P", 64)); MP = int(os.environ.get("MP", 3)); OMEGA = float(os.environ.get("OMEGA", 0.3))
ITERS = int(os.environ.get("ITERS", 20000)); NC = int(os.environ.get("NC", 1)); OUT = int(os.environ.get("OUT", 1))
dag = make_dag(P, seed=7 + P); g = make_prior(dag, max_par=MP, seed=8 + P); nt = g.node_type_codes()
X = simulate_numpy(dag, 300, seed=9 + P)
res = {}
with Context.from_data(X, g.source, g.target, nt, max_par=MP, omega=OMEGA) as ctx:
    for mode in ("0", os.environ.get("PIPE_MODE", "1")):
        os.environ["BN_B200_PIPE"] = mode
        res[mode] = ctx.run(n_chains=NC, n_iter=ITERS, output=OUT, rng="wh", seeds=[(11 + 3 * i, 22 + i, 33 + i) for i in range(NC)], log_moves=True)[0]
for ch in range(NC):
    a, b = res["0"][ch], res[os.environ.get("PIPE_MODE", "1")][ch]
    ma, mb = a.accepted_moves, b.accepted_moves
    n = min(len(ma), len(mb))
    d = np.nonzero((ma[:n] != mb[:n]).any(1))[0]
    print(f"chain {ch}: moves legacy {len(ma)} pipe {len(mb)}; uniforms {a.uniforms} {b.uniforms}; first differing move index {d[0] if len(d) else None}")
    if len(d):
        i = d[0]
        print(" legacy moves around:\n", ma[max(0, i - 3):i + 3], "\n pipe moves around:\n", mb[max(0, i - 3):i + 3])
    for k in ("iter", "ChangedNode", "movetype", "FN", "FP", "additions", "deletions"):
        ta, tb = a.trace[k], b.trace[k]
        m = min(len(ta), len(tb))
        dd = np.nonzero(ta[:m] != tb[:m])[0]
        if len(dd) or len(ta) != len(tb):
            j = dd[0] if len(dd) else m
            print(f"  column {k}: rows {len(ta)} vs {len(tb)}, first difference at row {j}: legacy {ta[max(0,j-2):j+3]} pipe {tb[max(0,j-2):j+3]} (iter legacy {a.trace['iter'][max(0,j-2):j+3]})")
