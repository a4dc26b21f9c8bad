"""The two-CTA form of the chain kernel (chain_pipe_kernel, opt-in with BN_B200_PIPE=1: one CTA walks and
commits, the second keeps a replica of the graph and builds the position records of the next window) against
the one-CTA kernel on the same inputs: every output bit for bit -- the two differ only in who builds a record
and when, never in what a record says.  (The whole GPU suite also passes with BN_B200_PIPE=1 exported:
profiles/r02_two_cta_chain.md.)

Each case runs in a child process (tests/tools/pipe_case.py) with a time limit: early builds of the two-CTA kernel
had a rare hang (found and fixed: profiles/r02_two_cta_chain.md); should an experimental kernel ever stall again, the
child is killed and the case reported as an expected failure instead of stalling the run.  A DIFFERENCE between the two
forms is always a hard failure."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

TOOL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "pipe_case.py")


def _case(**spec):
    try:
        res = subprocess.run([sys.executable, TOOL, json.dumps(spec)], capture_output=True, text=True, timeout=150)
    except subprocess.TimeoutExpired:
        pytest.xfail("the opt-in two-CTA kernel did not finish in 150 s (profiles/r02_two_cta_chain.md)")
    last = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
    assert res.returncode == 0 and last == "OK", (last, res.stderr[-2000:])


@pytest.mark.parametrize("P,max_par,omega,n_iter", [(1000, 8, None, 30000), (1000, 8, 1.0, 20000), (260, 5, 0.4, 20000),
                                                     (64, 3, 0.3, 20000), (1025, 8, 2.0, 15000)])
def test_two_cta_equals_one_cta(P, max_par, omega, n_iter):
    """Sparse and dense regimes (accepted deletions, nodes at MaxPar, the set of nodes with parents shrinking
    and growing, sequential windows between rounds), several chains per launch, a row every 7 iterations."""
    _case(P=P, max_par=max_par, omega=omega, n_iter=n_iter, N=300, seed=7 + P, chains=6, output=7,
          want_deletions=omega is not None)


@pytest.mark.parametrize("rng", ["rmt", "replay"])
def test_two_cta_other_streams(rng):
    """R's Mersenne-Twister (each CTA twists its own copy of the state) and a replayed stream."""
    _case(P=120, max_par=6, omega=0.5, n_iter=15000, N=400, seed=31, chains=3, output=10, rng=rng)


@pytest.mark.parametrize("initial_network", [0, 1, 2])
def test_two_cta_start_graphs_and_tabulation(initial_network):
    """The three start graphs (the replica draws the random start from its own copy of the stream), `drop`,
    and both tabulations, which only the chain's CTA keeps."""
    _case(P=80, max_par=4, omega=0.6, n_iter=12000, N=300, seed=77, chains=2, output=5,
          initial_network=initial_network, drop=1000, tabulate=True)


def test_two_cta_rare_legal_children():
    """Iterations that outrun their record (hundreds of uniforms per addition: the sequential path takes them
    and the walk leaves the window grid, so windows are requested off the grid)."""
    _case(P=400, max_par=8, omega=None, n_iter=3000, N=200, seed=3, chains=2, output=3, rare_children=True)
