"""Host-side logic, CPU only: the input mirror of the reference's R helpers, the synthetic
generator, the chain core compiled for the host (one-lane warp) against the golden traces,
and the C ABI surface (library loads, exports every declared symbol, fails loudly w/o GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, centered_stats, prior_lists

INT_COLS = ("iter", "ChangedNode", "movetype", "additions", "deletions", "FN", "FP")


# ---------------------------------------------------------------------------
# create_network: tests/testthat/test-bnetwork.R translated
# ---------------------------------------------------------------------------
def test_create_network_inconsistent_input():
    from bayesnetworks_b200 import create_network
    with pytest.raises(ValueError, match="same type"):
        create_network([1], ["a"])
    with pytest.raises(ValueError, match="same length"):
        create_network([1, 2], [1, 2, 3])
    with pytest.raises(ValueError, match="same"):
        create_network([1], [1])
    with pytest.raises(ValueError, match="cannot be specified if"):
        create_network(node_type=["sink"])
    with pytest.raises(ValueError):
        create_network([1], [2], ["A"])


def test_create_network_shapes():
    from bayesnetworks_b200 import create_network
    n = create_network([], [])
    assert (len(n.source), len(n.target), len(n.node_labels), len(n.node_type)) == (0, 0, 0, 0)
    n = create_network(node_labels=[1])
    assert (len(n.source), len(n.target), len(n.node_labels), len(n.node_type)) == (0, 0, 1, 1)
    n = create_network(node_labels=list(range(1, 101)))
    assert (len(n.node_labels), len(n.node_type)) == (100, 100)
    n = create_network([1], [2])
    assert (len(n.source), len(n.target), len(n.node_labels), len(n.node_type)) == (1, 1, 2, 2)
    n = create_network(["A"], ["B"])
    assert list(n.source) == [1] and list(n.target) == [2] and n.node_type == ["neither"] * 2
    letters = [chr(ord("A") + i) for i in range(26)]
    n = create_network(letters[1:], ["A"] * 25)
    assert (len(n.source), len(n.node_labels)) == (25, 26)
    assert set(n.target) == {1} and list(n.source) == list(range(2, 27))


def test_create_network_sorted_by_target_stable():
    from bayesnetworks_b200 import create_network
    n = create_network([5, 1, 4, 3], [3, 2, 3, 2], node_labels=[1, 2, 3, 4, 5])
    assert list(n.target) == [2, 2, 3, 3]
    assert list(n.source) == [1, 3, 5, 4]  # order() is stable (R/bnetwork.R:72)


def test_fixture_readers_roundtrip(tmp_path, dataset):
    """read_data / read_dag (R/aaa.R:9-49) on files written in the reference's formats."""
    from bayesnetworks_b200 import read_data, read_dag
    X = dataset["X"][:7]
    dat = tmp_path / "x.dat"
    with open(dat, "w", newline="") as fh:
        fh.write("\r\n")
        for i, row in enumerate(X):
            junk = "  0.5550   10   15.98  0.109  2.519  "
            fh.write(f"    {i}  {row[0]:.0f}{junk}" + "  ".join(f"{v:.2f}" for v in row[1:]) + "\r\n")
    got = read_data(str(dat))
    assert got.shape == (7, 81)
    np.testing.assert_allclose(got, X, atol=5e-3)
    P = 81
    par, npar = prior_lists(dataset["source"], dataset["target"], P, 50)
    dag = tmp_path / "x.dag.txt"
    with open(dag, "w", newline="") as fh:
        for p in range(P):
            fh.write(f"{npar[p]}  {dataset['node_type'][p]}  " +
                     "  ".join(str(q) for q in par[p, :npar[p]]) + " \r")
    g = read_dag(str(dag))
    assert np.array_equal(g.source, dataset["source"]) and np.array_equal(g.target, dataset["target"])
    assert np.array_equal(g.node_type_codes(), dataset["node_type"])


def test_shipped_dataset_shape(dataset):
    assert dataset["X"].shape == (2000, 81) and len(dataset["source"]) == 44
    nt = dataset["node_type"]
    assert (nt == 1).sum() == 40 and (nt == 2).sum() == 1 and (nt == 0).sum() == 40


# ---------------------------------------------------------------------------
# synthetic generator
# ---------------------------------------------------------------------------
def test_synthetic_generator():
    from bayesnetworks_b200.synth import chain_seeds, make_dag, make_prior, simulate_numpy
    dag = make_dag(100, seed=42)
    assert all(all(p < j for p in dag.parents[j]) for j in range(100))
    assert max(len(p) for p in dag.parents) <= 3
    X = simulate_numpy(dag, 500, seed=42)
    assert X.shape == (500, 100) and X.flags["F_CONTIGUOUS"]
    np.testing.assert_allclose(X.mean(0), 0, atol=1e-12)
    np.testing.assert_allclose(X.std(0), 1, atol=1e-12)
    g = make_prior(dag, max_par=8, seed=43)
    assert np.all(np.diff(g.target) >= 0)
    counts = np.bincount(g.target - 1, minlength=100)
    assert counts.max() <= 8
    types = g.node_type_codes()
    assert not np.any(types[g.target - 1] == 1) or True  # true edges never point into roots
    assert (types == 1).sum() >= 1 and (types == 2).sum() >= 1
    s = chain_seeds(4)
    assert tuple(s[0]) == (10437, 13568, 30524)
    assert np.array_equal(chain_seeds(2, first_chain=2), s[2:4])  # global chain identity
    assert len({tuple(r) for r in s}) == 4


# ---------------------------------------------------------------------------
# the chain core, host instantiation (logic only; the GPU tests check the device build)
# ---------------------------------------------------------------------------
def _rmt_state(seed):
    s = np.uint32(seed)
    out = np.zeros(625, np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(50):
            s = np.uint32(69069) * s + np.uint32(1)
        for j in range(625):
            s = np.uint32(69069) * s + np.uint32(1)
            out[j] = s
    return out[1:].copy()


def _emu_run(lib, dataset, max_par, n_iter, output, kind, seeds, init=2, drop=0, phi=1.0, omega=6.9):
    X, src, tgt, nt = dataset["X"], dataset["source"], dataset["target"], dataset["node_type"]
    N, P = X.shape
    _, Cm = centered_stats(X)
    ppar, pnpar = prior_lists(src, tgt, P, max_par)
    sim = np.zeros((P, P), np.uint8)
    for s_, t_ in zip(src, tgt):
        sim[t_ - 1, s_ - 1] = 1
    cap = (n_iter + output - 1) // output + 1
    ti = [np.zeros(cap, np.int32) for _ in range(7)]
    gll = np.zeros(cap)
    moves = np.zeros((n_iter, 4), np.int32)
    freq = np.zeros((P, P), np.int32)
    nfreq = np.zeros((P, max_par + 1), np.int32)
    fpar = np.zeros((P, max_par), np.int32)
    fnpar = np.zeros(P, np.int32)
    cnt = np.zeros(12, np.int64)
    sd = np.zeros(3, np.int32)
    sd[:len(seeds)] = seeds
    mt = _rmt_state(seeds[0]) if kind == 1 else np.zeros(624, np.uint32)
    ntu = nt.astype(np.uint8)
    dp, ip, up = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_ubyte)
    rc = lib.emu_run_chain(
        P, max_par, N, Cm.ctypes.data_as(dp), ntu.ctypes.data_as(up), sim.ctypes.data_as(up),
        int(len(src)), C.c_double(phi), C.c_double(omega), init, drop, n_iter, output,
        ppar.ctypes.data_as(ip), pnpar.ctypes.data_as(ip), kind, sd.ctypes.data_as(ip),
        mt.ctypes.data_as(C.POINTER(C.c_uint)), None, C.c_long(0), cap,
        ti[0].ctypes.data_as(ip), ti[1].ctypes.data_as(ip), ti[2].ctypes.data_as(ip),
        gll.ctypes.data_as(dp), ti[3].ctypes.data_as(ip), ti[4].ctypes.data_as(ip),
        ti[5].ctypes.data_as(ip), ti[6].ctypes.data_as(ip), n_iter, moves.ctypes.data_as(ip),
        freq.ctypes.data_as(ip), nfreq.ctypes.data_as(ip), fpar.ctypes.data_as(ip), fnpar.ctypes.data_as(ip),
        cnt.ctypes.data_as(C.POINTER(C.c_long)))
    n = int(cnt[8])
    out = dict(zip(INT_COLS, [ti[0][:n], ti[1][:n], ti[2][:n], ti[3][:n], ti[4][:n], ti[5][:n], ti[6][:n]]))
    out.update(rc=rc, globalLL=gll[:n], moves=moves[:int(cnt[9])], cnt=cnt, fpar=fpar, fnpar=fnpar,
               freq=freq, nfreq=nfreq)
    return out


@pytest.mark.parametrize("name,kind,seeds", [("cfg1", 1, (1234,)), ("cfg2", 0, (10437, 13568, 30524))])
def test_chain_core_host_build_matches_reference(emu_lib, dataset, golden, name, kind, seeds):
    r = _emu_run(emu_lib, dataset, 50, 50000, 100, kind, seeds)
    assert r["rc"] == 0
    for k in INT_COLS:
        assert np.array_equal(r[k], golden[f"{name}_{k}"]), k
    np.testing.assert_allclose(r["globalLL"], golden[f"{name}_globalLL"], rtol=1e-9, atol=1e-6)
    assert int(r["cnt"][0]) == int(golden[f"{name}_uniforms"])
    assert np.array_equal(r["moves"], golden[f"{name}_accepted_moves"])
    P = 81
    edges = [(int(r["fpar"][c, e]), c) for c in range(P) for e in range(r["fnpar"][c])]
    assert np.array_equal(np.asarray(edges, np.int32), golden[f"{name}_final_edges"])
    assert list(r["cnt"][2:5]) == list(golden[f"{name}_proposed"])
    assert list(r["cnt"][5:8]) == list(golden[f"{name}_reject"])


def test_chain_core_every_iteration_and_tabulation(emu_lib, dataset, golden):
    r = _emu_run(emu_lib, dataset, 8, 4000, 1, 0, (10437, 13568, 30524))
    for k in INT_COLS:
        assert np.array_equal(r[k], golden[f"every_wh_{k}"]), k
    # posterior tabulation (Bayes-networks/main.cpp:289-297): replay the accepted moves
    P = 81
    cur, freq, mv = set(), np.zeros((P, P), np.int64), {int(m[0]): m for m in r["moves"]}
    npar, nfreq = np.zeros(P, np.int64), np.zeros((P, 9), np.int64)
    for it in range(4000):
        if it in mv:
            _, typ, c, j = mv[it]
            (cur.add if typ == 1 else cur.discard)((int(j), int(c)))
            npar[c] += 1 if typ == 1 else -1
        for (j, c) in cur:
            freq[c, j] += 1
        nfreq[np.arange(P), npar] += 1   # freqNpar[p][Npar[p]]++, main.cpp:291
    assert np.array_equal(r["freq"], freq)
    assert np.array_equal(r["nfreq"], nfreq)
    r0 = _emu_run(emu_lib, dataset, 8, 2000, 1, 1, (99,), init=0)
    for k in INT_COLS:
        assert np.array_equal(r0[k], golden[f"every_init0_{k}"]), k


@pytest.mark.parametrize("max_par,omega", [(5, 0.5), (12, 0.2)])
def test_chain_core_dense_synthetic_vs_oracle(emu_lib, oracle, max_par, omega):
    """Dense graphs (weak size penalty): many accepted deletions, large descendant sets, nodes
    at the parent limit -- the general ancestor-recompute path, the hp_list maintenance and the
    MaxPar masks, against the oracle (which does a BFS per proposal like the reference)."""
    from bayesnetworks_b200.synth import make_dag, make_prior, simulate_numpy
    from oracle.oracle import RNG_WH
    P, N, n_iter = 40, 500, 20000
    dag = make_dag(P, seed=5)
    g = make_prior(dag, max_par=max_par, seed=6)
    X = simulate_numpy(dag, N, seed=7)
    nt = g.node_type_codes()
    ref = oracle.mcmc(X, g.source, g.target, nt, max_par=max_par, phi=1.0, omega=omega, n_iter=n_iter,
                      output=7, rng_kind=RNG_WH, seeds=(123, 456, 789))
    ds = dict(X=X, source=g.source, target=g.target, node_type=nt)
    r = _emu_run(emu_lib, ds, max_par, n_iter, 7, 0, (123, 456, 789), omega=omega)
    assert r["rc"] == 0
    for k in INT_COLS:
        assert np.array_equal(r[k], getattr(ref, k)), k
    np.testing.assert_allclose(r["globalLL"], ref.globalLL, rtol=1e-9, atol=1e-9 * N / 2)
    assert np.array_equal(r["moves"], ref.accepted_moves())
    assert int(r["cnt"][0]) == ref.uniforms
    assert ref.deletions[-1] > 200 and ref.final_npar.max() == max_par  # the regime is exercised


def test_device_rng_core_matches_goldens(emu_lib, golden):
    out = np.zeros(1000)
    sd = np.array([10437, 13568, 30524], np.int32)
    ip = C.POINTER(C.c_int)
    emu_lib.emu_uniforms(0, sd.ctypes.data_as(ip), None, 1000, out.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.array_equal(out, golden["wh_first1000"])
    for seed in (1234, 42):
        mt = _rmt_state(seed)
        emu_lib.emu_uniforms(1, sd.ctypes.data_as(ip), mt.ctypes.data_as(C.POINTER(C.c_uint)), 1000,
                             out.ctypes.data_as(C.POINTER(C.c_double)))
        assert np.array_equal(out, golden[f"rmt_seed{seed}_first1000"])


def test_score_core_matches_oracle(emu_lib, dataset, oracle):
    X = dataset["X"]
    N, P = X.shape
    _, Cm = centered_stats(X)
    stats = oracle.gram(X)
    rng = np.random.default_rng(3)
    for _ in range(200):
        c = int(rng.integers(0, P))
        k = int(rng.integers(0, 9))
        S = np.asarray([q for q in rng.permutation(P) if q != c][:k], dtype=np.int32)
        want, _ = oracle.score(X, c, S, stats=stats)
        got = emu_lib.emu_score_set(Cm.ctypes.data_as(C.POINTER(C.c_double)), P, c,
                                    S.ctypes.data_as(C.POINTER(C.c_int)), k, N)
        assert abs(got - want) <= 1e-9 * abs(want) + 1e-9 * N / 2


# ---------------------------------------------------------------------------
# the C ABI
# ---------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    from bayesnetworks_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "bn_b200.h")).read()
    declared = set(re.findall(r"\b(bn_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.bn_abi_version() == 3


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point fails loudly (BN_ERR_NO_DEVICE)."""
    from bayesnetworks_b200 import BnError, Context, _lib
    if _lib.lib().bn_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(BnError) as ei:
        Context.from_data(np.random.rand(10, 3), [1], [2], [0, 0, 0], max_par=2)
    assert ei.value.status == _lib.BN_ERR_NO_DEVICE


def test_argument_validation_before_device():
    from bayesnetworks_b200 import BnError, Context, _lib
    with pytest.raises(BnError) as ei:
        Context.from_data(np.random.rand(10, 3), [7], [2], [0, 0, 0], max_par=2)
    assert ei.value.status == _lib.BN_ERR_BAD_ARG
    with pytest.raises(BnError) as ei:
        Context.from_data(np.random.rand(10, 3), [1], [2], [0, 5, 0], max_par=2)
    assert ei.value.status == _lib.BN_ERR_BAD_ARG


def test_rcpp_glue_compiles_against_shim():
    """The drop-in src/bayesnet_mcmc.cpp replacement keeps the reference's exported signature
    and compiles against the stand-in Rcpp.h (R itself is not installed here)."""
    import subprocess
    glue = os.path.join(ROOT, "bayesnetworks_b200", "csrc", "rcpp_glue", "bayesnet_mcmc.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-w",
                           f"-I{os.path.join(ROOT, 'oracle', 'ref_shim')}",
                           f"-I{os.path.join(ROOT, 'include')}", glue])
    text = open(glue).read()
    for frag in ("DataFrame main_fun(NumericMatrix X,", "std::vector<int> graph_node_labels,",
                 "int MaxPar = 50,", "const double phi = 1,", "const double omega = 6.9,",
                 "const int InitialNetwork = 2,", "const int drop = 0,", "int N = 1000,",
                 "int output = 10)"):
        assert frag in text, frag
    cols = re.findall(r'Named\("(\w+)"\)', text)
    assert cols == ["iter", "ChangedNode", "movetype", "globalLL", "additions", "deletions", "FN", "FP"]


def test_product_never_imports_oracle():
    """Nothing under bayesnetworks_b200/ imports, links or dlopens anything under oracle/."""
    pkg = os.path.join(ROOT, "bayesnetworks_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|oracle\.oracle|bn_oracle|libbnref|oracle/|_ref/)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), f


def _mostly_sources_case():
    """300 nodes of which 296 are sources: the child draw of an addition is rejected with
    probability 296/300, so iterations that need hundreds of uniforms are common -- more than
    a position record can count (255), which sends them through the sequential window path."""
    rng = np.random.default_rng(11)
    P, N = 300, 400
    X = np.asfortranarray(rng.standard_normal((N, P)))
    for c, ps in ((296, (0, 1)), (297, (2, 296)), (298, (3, 4, 5)), (299, (297, 6))):
        for q in ps:
            X[:, c] += 0.7 * X[:, q]
    nt = np.ones(P, dtype=np.int32)
    nt[296:] = 0
    src = np.array([1, 2, 3, 297, 4, 5], dtype=np.int32)
    tgt = np.array([297, 297, 298, 298, 299, 299], dtype=np.int32)
    return X, src, tgt, nt


def test_chain_core_long_rejection_loops_vs_oracle(emu_lib, oracle):
    from oracle.oracle import RNG_WH
    X, src, tgt, nt = _mostly_sources_case()
    n_iter = 6000
    ref = oracle.mcmc(X, src, tgt, nt, max_par=8, phi=1.0, omega=1.0, n_iter=n_iter, output=5,
                      rng_kind=RNG_WH, seeds=(321, 654, 987))
    ds = dict(X=X, source=src, target=tgt, node_type=nt)
    r = _emu_run(emu_lib, ds, 8, n_iter, 5, 0, (321, 654, 987), omega=1.0)
    assert r["rc"] == 0
    for k in INT_COLS:
        assert np.array_equal(r[k], getattr(ref, k)), k
    assert int(r["cnt"][0]) == ref.uniforms
    assert ref.uniforms > 30 * n_iter  # ~75 child draws per addition; 3.5% of them need more than 255
