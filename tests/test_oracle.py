"""The CPU oracle against the golden vectors generated from the UNMODIFIED reference
(tests/golden/make_golden.py) and against the reference's own golden trace (the legacy
xlsx).  CPU only; sized to run in well under a minute."""
import numpy as np
import pytest

from oracle.oracle import RNG_RMT, RNG_WH

COLS = ("iter", "ChangedNode", "movetype", "globalLL", "additions", "deletions", "FN", "FP")


def test_wichmann_hill_known_answers(oracle, golden):
    # SURVEY.md B.1, Bayes-networks/random4f.h seeds 10437/13568/30524
    want = [0.090953704848706129, 0.42809688836293813, 0.96060048711501489,
            0.43520064132698311, 0.027688522659340853, 0.39314819673496815]
    got = oracle.uniforms(1000, RNG_WH)
    assert np.array_equal(got[:6], np.array(want))
    assert np.array_equal(got, golden["wh_first1000"])


def test_r_mersenne_twister_known_answers(oracle, golden):
    # universally known R answers: set.seed(s); runif(3)
    known = {1234: (0.1137034, 0.6222994, 0.6092747), 42: (0.9148060, 0.9370754, 0.2861395),
             1: (0.2655087, 0.3721239, 0.5728534), 123: (0.2875775, 0.7883051, 0.4089769)}
    for seed, want in known.items():
        got = oracle.uniforms(1000, RNG_RMT, (seed,))
        np.testing.assert_allclose(got[:3], want, atol=5e-8)
        assert np.array_equal(got, golden[f"rmt_seed{seed}_first1000"])
        assert got.min() > 0 and got.max() < 1


def test_invert_pds_matches_reference(oracle, golden):
    off = 0
    for n in golden["pds_dims"]:
        A = golden["pds_in"][off:off + n * n].reshape(n, n)
        want = golden["pds_out"][off:off + n * n].reshape(n, n)
        off += n * n
        rc, inv = oracle.invert_pds(A)
        assert rc == 0
        assert np.array_equal(inv, want)
        np.testing.assert_allclose(inv @ A, np.eye(n), atol=1e-9)
    rc, inv = oracle.invert_pds(np.array([[1.0, 2.0], [2.0, 1.0]]))
    assert rc == int(golden["pds_nonpd_rc"]) == 14
    assert np.array_equal(inv, np.eye(2))  # the reference carries on with the identity


def test_sufficient_statistics_bit_exact(oracle, dataset, golden):
    sum_x, sum_xx = oracle.gram(dataset["X"])
    assert np.array_equal(sum_x, golden["sumX"])
    assert np.array_equal(sum_xx, golden["sumXX"])
    # SURVEY.md B.2 spot values
    assert sum_x[0] == 1000 and sum_xx[0, 0] == 1000
    assert sum_xx[0, 21] == 589.68000000000029
    assert sum_xx[21, 22] == 863.51720000000012
    assert sum_xx[80, 80] == 19.557199999999966


def test_scores_bit_exact(oracle, dataset, golden):
    from conftest import prior_lists
    P = dataset["X"].shape[1]
    par, npar = prior_lists(dataset["source"], dataset["target"], P, 50)
    got = oracle.score_graph(dataset["X"], par, npar, pad_dim=0)
    assert np.array_equal(got, golden["prior_scores"])
    # SURVEY.md B.3
    assert got[0] == 34.06543249794975 and got[21] == 961.45456890719174
    assert float(golden["prior_globalLL"]) == 16791.445774684958
    acc = 0.0
    for v in got:
        acc += v
    assert acc == float(golden["prior_globalLL"])
    got0 = oracle.score_graph(dataset["X"], np.full((P, 50), -1, np.int32), np.zeros(P, np.int32))
    assert np.array_equal(got0, golden["null_scores"])
    # the identity-padded 51-dim inversion of the reference gives the same bits
    padded = oracle.score_graph(dataset["X"], par, npar, pad_dim=51)
    assert np.array_equal(padded, got)


@pytest.mark.parametrize("name,kind,seeds", [("cfg1", RNG_RMT, (1234,)),
                                              ("cfg2", RNG_WH, (10437, 13568, 30524))])
def test_chain_bit_exact(oracle, dataset, golden, name, kind, seeds):
    r = oracle.mcmc(dataset["X"], dataset["source"], dataset["target"], dataset["node_type"],
                    max_par=50, n_iter=50000, output=100, rng_kind=kind, seeds=seeds)
    for k in COLS:
        assert np.array_equal(getattr(r, k), golden[f"{name}_{k}"]), k
    assert r.uniforms == int(golden[f"{name}_uniforms"])
    assert np.array_equal(r.accepted_moves(), golden[f"{name}_accepted_moves"])
    assert np.array_equal(np.asarray(r.edges(), np.int32), golden[f"{name}_final_edges"])


def test_chain_every_iteration_bit_exact(oracle, dataset, golden):
    r = oracle.mcmc(dataset["X"], dataset["source"], dataset["target"], dataset["node_type"],
                    max_par=8, n_iter=4000, output=1, rng_kind=RNG_WH)
    for k in COLS:
        assert np.array_equal(getattr(r, k), golden[f"every_wh_{k}"]), k
    r = oracle.mcmc(dataset["X"], dataset["source"], dataset["target"], dataset["node_type"],
                    max_par=8, n_iter=2000, output=1, rng_kind=RNG_RMT, seeds=(99,),
                    initial_network=0)
    for k in COLS:
        assert np.array_equal(getattr(r, k), golden[f"every_init0_{k}"]), k


def test_readme_example_summary(golden):
    # README.md:41-74: plateau ~1.68e4, additions ~234, deletions ~200, FN ~10-12, FP ~1-3
    assert golden["cfg1_additions"][-1] == 234 and golden["cfg1_deletions"][-1] == 200
    assert 16700 < golden["cfg1_globalLL"][-1] < 16850
    assert golden["cfg1_FN"][-1] in (10, 11, 12) and golden["cfg1_FP"][-1] in (1, 2, 3)


def test_legacy_mode_reproduces_reference_xlsx(oracle, dataset, legacy_xlsx):
    """`Bayes-networks/iterations - null start.xlsx`: 1,100 rows of a 110,000-iteration run of
    the legacy program (float X, Wichmann-Hill, validity check disabled)."""
    Xf = dataset["X"].astype(np.float32).astype(np.float64)
    r = oracle.mcmc(Xf, dataset["source"], dataset["target"], dataset["node_type"], max_par=50,
                    n_iter=110000, output=100, rng_kind=RNG_WH, legacy=True, log_moves=False)
    g = legacy_xlsx
    assert len(r.iter) == 1100 == len(g["iter"])
    for mine, col in ((r.iter, "iter"), (r.ChangedNode, "chngd"), (r.legacy["Npar"], "Npar"),
                      (r.movetype, "type"), (r.legacy["Edges"], "Edges"), (r.FP, "FP"), (r.FN, "FN"),
                      (r.legacy["Agree"], "Agree"), (r.additions, "Add"), (r.deletions, "Delete")):
        assert np.array_equal(mine, g[col].astype(np.int64)), col
    # the file holds printf("%11.4f") / ("%9.4f") values
    assert np.max(np.abs(r.globalLL - g["lnL"])) <= 5.0001e-5
    assert np.max(np.abs(r.legacy["lnPrior"] - g["lnPrior"])) <= 5.0001e-5
    assert r.additions[-1] == 464 and r.deletions[-1] == 428
