"""Golden fixtures for the DENSE regime at the benchmark's own shape, from the reference itself.

Run in the build container (needs /root/reference and oracle/_ref built by ``make -C oracle``):

    python tests/golden/make_golden_dense.py

Every case is a synthetic linear-Gaussian DAG (bayesnetworks_b200.synth, seeded) with
P in {896, 897, 1000, 1024, 1025} nodes -- the sizes around the 1,024-node boundary of the
ancestor bitsets (8 chunks of 128 bits; 897..1,024 nodes take a dedicated code path in the
CUDA kernel, 1,025 the generic one) -- MaxPar 8, 300 rows, and prior weights that make the
chain accept thousands of additions AND deletions, saturate nodes at MaxPar and propose
many cyclic additions (invalid iterations, stale `valid` deletions).

For each case the UNMODIFIED reference sources (oracle/_ref/libbnref.so, i.e.
src/bayesnet_mcmc.cpp:27-72 + src/network.h) produce the 8 trace columns and the uniform
count; the C restatement (oracle/bn_oracle.c) is asserted identical on all of them here and
supplies what the reference does not export (accepted-move log, final edge lists).

Writes tests/golden/dense_ref.npz (committed).  The GPU box has no /root/reference: the
tests read only this file and regenerate X from the seeds (a checksum of X is stored).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from bayesnetworks_b200.synth import make_dag, make_prior, simulate_numpy  # noqa: E402
from oracle.oracle import RNG_WH, Oracle, Ref  # noqa: E402

COLS = ("iter", "ChangedNode", "movetype", "globalLL", "additions", "deletions", "FN", "FP")
N_ROWS = 300
MAX_PAR = 8
N_ITER = 20000
OUTPUT = 100


def dense_cases():
    """(name, P, phi, omega, seeds).  'mid': ~2,500 edges, thousands of accepted deletions;
    'sat': ~6,000 edges, most nodes at MaxPar, a third of the iterations invalid."""
    cases = []
    for P in (896, 897, 1000, 1024, 1025):
        cases.append((f"p{P}_mid", P, 0.0, 0.0, (101 + P, 202 + P, 303 + P)))
        cases.append((f"p{P}_sat", P, 0.0, -1.0, (11 + P, 22 + P, 33 + P)))
    # eight chains at the benchmark's node count (one launch of 8 chains on the GPU)
    for c in range(8):
        cases.append((f"p1000_chain{c}", 1000, 0.0, 0.0, (5000 + 7 * c, 6000 + 11 * c, 7000 + 13 * c)))
    return cases


def dense_inputs(P):
    """The dataset of a case: depends on P only."""
    dag = make_dag(P, seed=P)
    g = make_prior(dag, max_par=MAX_PAR, seed=P + 1)
    X = simulate_numpy(dag, N_ROWS, seed=P + 2)
    return X, g, g.node_type_codes()


def x_digest(X):
    return hashlib.sha1(np.ascontiguousarray(X).tobytes()).hexdigest()


def moves_digest(mv):
    """sha1 of the accepted-move log as int32 rows (iter, movetype, child, parent)."""
    return hashlib.sha1(np.ascontiguousarray(mv, dtype=np.int32).tobytes()).hexdigest()


def main():
    R, O = Ref(), Oracle()
    out = {}
    inputs = {}
    for name, P, phi, omega, seeds in dense_cases():
        if P not in inputs:
            inputs[P] = dense_inputs(P)
            out[f"x_sha1_p{P}"] = np.frombuffer(x_digest(inputs[P][0]).encode(), dtype=np.uint8)
        X, g, nt = inputs[P]
        r = R.main_fun(X, g.source, g.target, nt, MaxPar=MAX_PAR, phi=phi, omega=omega, N=N_ITER, output=OUTPUT,
                       rng_kind=RNG_WH, seeds=seeds)
        o = O.mcmc(X, g.source, g.target, nt, max_par=MAX_PAR, phi=phi, omega=omega, n_iter=N_ITER, output=OUTPUT,
                   rng_kind=RNG_WH, seeds=seeds)
        for k in COLS:
            assert np.array_equal(getattr(o, k), getattr(r, k)), (name, k)   # bit-equal, globalLL included
        assert o.uniforms == r.uniforms, name
        assert o.n_nonpd == 0, name
        for k in COLS:
            out[f"{name}_{k}"] = getattr(r, k)
        out[f"{name}_uniforms"] = np.int64(r.uniforms)
        mv = o.accepted_moves()
        out[f"{name}_moves_sha1"] = np.frombuffer(moves_digest(mv).encode(), dtype=np.uint8)
        out[f"{name}_n_moves"] = np.int64(len(mv))
        if "chain" not in name or name.endswith("chain0"):  # full log where a mismatch needs locating
            out[f"{name}_moves_iter"] = mv[:, 0].astype(np.int32)
            out[f"{name}_moves_type"] = mv[:, 1].astype(np.int8)
            out[f"{name}_moves_child"] = mv[:, 2].astype(np.int16)
            out[f"{name}_moves_parent"] = mv[:, 3].astype(np.int16)
        out[f"{name}_final_edges"] = np.asarray(o.edges(), dtype=np.int16)
        out[f"{name}_reject"] = np.asarray(o.reject, dtype=np.int32)
        out[f"{name}_proposed"] = np.asarray(o.proposed, dtype=np.int32)
        print(f"{name}: additions {r.additions[-1]} deletions {r.deletions[-1]} invalid {o.reject[0]} "
              f"edges {len(o.edges())} at MaxPar {(o.final_npar == MAX_PAR).sum()} uniforms {r.uniforms}")
    path = os.path.join(HERE, "dense_ref.npz")
    np.savez_compressed(path, **out)
    print("written", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
