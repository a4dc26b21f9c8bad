"""The two-CTA form of the chain kernel (chain_pipe_kernel: one CTA walks and commits, the second keeps
a replica of the graph and builds the position records of the next window) against the one-CTA kernel
on the same inputs: every output bit for bit -- the two differ only in who builds a record and when,
never in what a record says.  Both are compared with the oracle elsewhere (test_gpu_parity,
test_gpu_more, test_dense_parity run whichever form bn_run picks); here the switch BN_B200_PIPE is
flipped explicitly so that a regression in either shows up as a difference between them."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

COLS = ("iter", "ChangedNode", "movetype", "additions", "deletions", "FN", "FP", "globalLL")


def _synthetic(P, N, max_par, seed):
    from bayesnetworks_b200.synth import make_dag, make_prior, simulate_numpy
    dag = make_dag(P, seed=seed)
    g = make_prior(dag, max_par=max_par, seed=seed + 1)
    return simulate_numpy(dag, N, seed=seed + 2), g, g.node_type_codes()


def _run_both(ctx, **kw):
    out = {}
    old = os.environ.get("BN_B200_PIPE")
    try:
        for mode in ("1", "0"):
            os.environ["BN_B200_PIPE"] = mode
            out[mode] = ctx.run(**kw)[0]
    finally:
        if old is None:
            os.environ.pop("BN_B200_PIPE", None)
        else:
            os.environ["BN_B200_PIPE"] = old
    return out["1"], out["0"]


def _identical(a, b):
    assert len(a) == len(b)
    for ra, rb in zip(a, b):
        for k in COLS:
            assert np.array_equal(ra.trace[k], rb.trace[k]), k   # globalLL: same bits
        assert ra.uniforms == rb.uniforms
        assert ra.valid_iters == rb.valid_iters and ra.n_nonpd == rb.n_nonpd
        assert list(ra.proposed) == list(rb.proposed) and list(ra.reject) == list(rb.reject)
        assert np.array_equal(ra.final_parents, rb.final_parents)
        assert np.array_equal(ra.final_npar, rb.final_npar)
        if ra.accepted_moves is not None:
            assert np.array_equal(ra.accepted_moves, rb.accepted_moves)


@pytest.mark.parametrize("P,max_par,omega,n_iter", [(1000, 8, 6.9, 30000), (1000, 8, 1.0, 20000), (260, 5, 0.4, 20000),
                                                     (64, 3, 0.3, 20000), (1025, 8, 2.0, 15000)])
def test_two_cta_equals_one_cta(P, max_par, omega, n_iter):
    """Sparse and dense regimes (accepted deletions, nodes at MaxPar, the set of nodes with parents
    shrinking and growing), several chains per launch, a row every 7 iterations."""
    from bayesnetworks_b200 import Context
    from bayesnetworks_b200.synth import chain_seeds
    X, g, nt = _synthetic(P, 300, max_par, 7 + P)
    with Context.from_data(X, g.source, g.target, nt, max_par=max_par, omega=omega) as ctx:
        a, b = _run_both(ctx, n_chains=6, n_iter=n_iter, output=7, rng="wh", seeds=chain_seeds(6), log_moves=True)
    _identical(a, b)
    assert sum(int(r.trace["deletions"][-1]) for r in a) > 0 or omega > 5


@pytest.mark.parametrize("rng", ["rmt", "replay"])
def test_two_cta_other_streams(rng):
    """R's Mersenne-Twister (each CTA twists its own copy of the state) and a replayed stream."""
    from bayesnetworks_b200 import Context
    X, g, nt = _synthetic(120, 400, 6, 31)
    kw = dict(n_chains=3, n_iter=15000, output=10, log_moves=True)
    if rng == "rmt":
        kw.update(rng="rmt", seeds=[(1234, 0, 0), (42, 0, 0), (7, 0, 0)])
    else:
        u = np.random.default_rng(5).random((3, 15000 * 12))
        kw.update(rng="replay", replay=u)
    with Context.from_data(X, g.source, g.target, nt, max_par=6, omega=0.5) as ctx:
        a, b = _run_both(ctx, **kw)
    _identical(a, b)


@pytest.mark.parametrize("initial_network", [0, 1, 2])
def test_two_cta_start_graphs_and_tabulation(initial_network):
    """The three start graphs (the replica draws the random start from its own copy of the stream),
    `drop`, and both tabulations, which only the chain's CTA keeps."""
    from bayesnetworks_b200 import Context
    X, g, nt = _synthetic(80, 300, 4, 77)
    with Context.from_data(X, g.source, g.target, nt, max_par=4, omega=0.6) as ctx:
        a, b = _run_both(ctx, n_chains=2, n_iter=12000, output=5, rng="wh", seeds=[(11, 22, 33), (44, 55, 66)],
                         initial_network=initial_network, drop=1000, log_moves=True, tabulate=True)
    _identical(a, b)
    for ra, rb in zip(a, b):
        assert np.array_equal(ra.edge_freq, rb.edge_freq)
        assert np.array_equal(ra.npar_freq, rb.npar_freq)


def test_two_cta_rare_legal_children():
    """Iterations that outrun their record (hundreds of uniforms per addition: the sequential path
    takes them and the walk leaves the window grid, so windows are requested off the grid)."""
    from bayesnetworks_b200 import Context
    from bayesnetworks_b200.synth import make_dag, make_prior, simulate_numpy
    P = 400
    dag = make_dag(P, seed=3)
    g = make_prior(dag, max_par=8, seed=4)
    nt = g.node_type_codes().copy()
    nt[:] = 1          # sources: never a child ...
    nt[5:9] = 0        # ... except four nodes
    X = simulate_numpy(dag, 200, seed=5)
    with Context.from_data(X, np.zeros(0, np.int32), np.zeros(0, np.int32), nt, max_par=8) as ctx:
        a, b = _run_both(ctx, n_chains=2, n_iter=3000, output=3, rng="wh", seeds=[(5, 6, 7), (8, 9, 10)], log_moves=True)
    _identical(a, b)
