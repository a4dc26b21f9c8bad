"""Child process of tests/test_gpu_pipeline.py: one case on both forms of the chain kernel, compared bit for bit.
Runs in its own process so that the (opt-in, experimental) two-CTA kernel can be given a time limit: a stall of
that kernel (early builds had a rare one, profiles/r02_two_cta_chain.md) must not take the test run with it.
usage: python pipe_case.py '<json spec>'  ->  prints OK or DIFF: <what>"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesnetworks_b200 import Context  # noqa: E402
from bayesnetworks_b200.synth import chain_seeds, make_dag, make_prior, simulate_numpy  # noqa: E402

COLS = ("iter", "ChangedNode", "movetype", "additions", "deletions", "FN", "FP", "globalLL")


def main():
    c = json.loads(sys.argv[1])
    P, MP = c["P"], c["max_par"]
    dag = make_dag(P, seed=c["seed"])
    g = make_prior(dag, max_par=MP, seed=c["seed"] + 1)
    nt = g.node_type_codes().copy()
    src, tgt = g.source, g.target
    if c.get("rare_children"):   # sources everywhere except four nodes: hundreds of uniforms per addition
        nt[:] = 1
        nt[5:9] = 0
        src, tgt = np.zeros(0, np.int32), np.zeros(0, np.int32)
    X = simulate_numpy(dag, c["N"], seed=c["seed"] + 2)
    kw = dict(n_chains=c["chains"], n_iter=c["n_iter"], output=c["output"], log_moves=True,
              initial_network=c.get("initial_network", 2), drop=c.get("drop", 0), tabulate=bool(c.get("tabulate")))
    rng = c.get("rng", "wh")
    if rng == "wh":
        kw.update(rng="wh", seeds=chain_seeds(c["chains"]))
    elif rng == "rmt":
        kw.update(rng="rmt", seeds=[(1234 + 7 * i, 0, 0) for i in range(c["chains"])])
    else:
        kw.update(rng="replay", replay=np.random.default_rng(5).random((c["chains"], c["n_iter"] * 12)))
    out = {}
    extra = {} if c.get("omega") is None else {"omega": c["omega"]}
    with Context.from_data(X, src, tgt, nt, max_par=MP, **extra) as ctx:
        for mode in ("1", "0"):
            os.environ["BN_B200_PIPE"] = mode
            out[mode] = ctx.run(**kw)[0]
    for ra, rb in zip(out["1"], out["0"]):
        for k in COLS:
            if not np.array_equal(ra.trace[k], rb.trace[k]):   # globalLL: same bits
                return f"DIFF: column {k}"
        if ra.uniforms != rb.uniforms or ra.valid_iters != rb.valid_iters or ra.n_nonpd != rb.n_nonpd:
            return "DIFF: counters"
        if list(ra.proposed) != list(rb.proposed) or list(ra.reject) != list(rb.reject):
            return "DIFF: proposed / reject"
        if not (np.array_equal(ra.final_parents, rb.final_parents) and np.array_equal(ra.final_npar, rb.final_npar)):
            return "DIFF: final graph"
        if not np.array_equal(ra.accepted_moves, rb.accepted_moves):
            return "DIFF: accepted moves"
        if c.get("tabulate") and not (np.array_equal(ra.edge_freq, rb.edge_freq) and np.array_equal(ra.npar_freq, rb.npar_freq)):
            return "DIFF: tabulation"
    if c.get("want_deletions") and sum(int(r.trace["deletions"][-1]) for r in out["1"]) == 0:
        return "DIFF: the case was meant to accept deletions"
    return "OK"


if __name__ == "__main__":
    print(main(), flush=True)
