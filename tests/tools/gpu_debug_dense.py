"""Debug helper: first divergence between the CUDA chain and the oracle on a dense synthetic case."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesnetworks_b200 import Context
from bayesnetworks_b200.synth import make_dag, make_prior, simulate_numpy
from oracle.oracle import RNG_WH, Oracle
mp = int(os.environ.get("MP", 12)); omega = float(os.environ.get("OMEGA", 0.2)); P = int(os.environ.get("P", 40))
it = int(os.environ.get("ITERS", 20000))
dag = make_dag(P, seed=5); g = make_prior(dag, max_par=mp, seed=6); X = simulate_numpy(dag, 500, seed=7); nt = g.node_type_codes()
ref = Oracle().mcmc(X, g.source, g.target, nt, max_par=mp, omega=omega, n_iter=it, output=7, rng_kind=RNG_WH, seeds=(123, 456, 789))
with Context.from_data(X, g.source, g.target, nt, max_par=mp, omega=omega) as ctx:
    r = ctx.run(n_iter=it, output=7, rng="wh", seeds=(123, 456, 789), log_moves=True)[0][0]
a, b = r.accepted_moves, ref.accepted_moves()
n = min(len(a), len(b))
bad = np.nonzero((a[:n] != b[:n]).any(axis=1))[0]
print("moves", len(a), len(b), "first diff", bad[:1])
if bad.size:
    i = bad[0]
    print("gpu", a[max(0, i - 3):i + 2].tolist()); print("ref", b[max(0, i - 3):i + 2].tolist())
    # state of the child at that point
    par = {}
    for (itn, typ, c, j) in b[:i]:
        par.setdefault(c, [])
        if typ == 1: par[c].append(j)
        else: par[c].remove(j)
    for row in (a[i], b[i]):
        print("child", row[2], "parents before", par.get(row[2], []))
print("uniforms", r.uniforms, ref.uniforms, "nonpd", r.n_nonpd, ref.n_nonpd)
for k, rk in (("iter", "iter"), ("ChangedNode", "ChangedNode"), ("movetype", "movetype"), ("additions", "additions"),
              ("deletions", "deletions"), ("FN", "FN"), ("FP", "FP")):
    x, y = r.trace[k], getattr(ref, rk)
    if len(x) != len(y) or not np.array_equal(x, y):
        m = min(len(x), len(y)); d = np.nonzero(x[:m] != y[:m])[0]
        print(k, "len", len(x), len(y), "first diff row", d[:3], "gpu", x[d[:3]], "ref", y[d[:3]], "iter", r.trace["iter"][d[:3]])
print("gll maxdiff", np.abs(r.trace["globalLL"] - ref.globalLL[:len(r.trace["globalLL"])]).max())
print("proposed", r.proposed, ref.proposed, "reject", r.reject, ref.reject)
