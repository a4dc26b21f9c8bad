"""Synthetic linear-Gaussian DAG workloads of the named shapes.

Same construction on the host (numpy, for parity tests against the oracle) and
on the device (torch CUDA generator, so the big configurations never cross
PCIe):

* topological order = node index; node j draws ``min(j, U{0..3})`` distinct
  parents uniformly from ``0..j-1``; weights ``+-U(0.5, 1.5)``;
  ``x_j = sum w x_pa + eps``, ``eps ~ N(0,1)``; every column is standardised
  as it is produced (mean 0, sd 1) to keep conditioning sane at depth.
* prior ("external") network = the true DAG with 10 % of its edges removed and
  10 % spurious edges added (seed 43), spurious edges respect the order, the
  source/sink types and ``max_par``.
* node types: the first 10 % of the roots are sources, the last 5 % of the
  leaves are sinks.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .network import Network, create_network


@dataclass
class SyntheticDag:
    n_nodes: int
    parents: list            # parents[j] = list of parent indices (< j)
    weights: list            # weights[j] aligned with parents[j]
    node_type: np.ndarray    # 0 neither / 1 source / 2 sink

    def edges(self):
        return [(p, j) for j in range(self.n_nodes) for p in self.parents[j]]


def make_dag(n_nodes: int, seed: int = 42, max_true_parents: int = 3) -> SyntheticDag:
    rng = np.random.default_rng(seed)
    parents, weights = [], []
    for j in range(n_nodes):
        d = min(j, int(rng.integers(0, max_true_parents + 1)))
        pa = sorted(rng.choice(j, size=d, replace=False).tolist()) if d else []
        w = (rng.uniform(0.5, 1.5, size=d) * rng.choice([-1.0, 1.0], size=d)).tolist()
        parents.append(pa)
        weights.append(w)
    has_child = np.zeros(n_nodes, bool)
    for j in range(n_nodes):
        for p in parents[j]:
            has_child[p] = True
    roots = [j for j in range(n_nodes) if not parents[j]]
    leaves = [j for j in range(n_nodes) if not has_child[j] and parents[j]]
    node_type = np.zeros(n_nodes, dtype=np.int32)
    for j in roots[: max(1, len(roots) // 10)]:
        node_type[j] = 1
    for j in leaves[len(leaves) - max(1, len(leaves) // 20):]:
        node_type[j] = 2
    return SyntheticDag(n_nodes, parents, weights, node_type)


def make_prior(dag: SyntheticDag, max_par: int, seed: int = 43, drop_frac: float = 0.1,
               add_frac: float = 0.1) -> Network:
    """The external network handed to ``bn_mcmc`` as ``graph``."""
    rng = np.random.default_rng(seed)
    edges = dag.edges()
    keep = [e for e in edges if rng.random() >= drop_frac]
    have = set(keep)
    npar = np.zeros(dag.n_nodes, dtype=np.int64)
    for _, c in keep:
        npar[c] += 1
    n_add = int(round(add_frac * len(edges)))
    tries = 0
    while n_add > 0 and tries < 100 * (n_add + 1):
        tries += 1
        a, b = sorted(rng.choice(dag.n_nodes, size=2, replace=False).tolist())
        if (a, b) in have or dag.node_type[a] == 2 or dag.node_type[b] == 1 or npar[b] >= max_par:
            continue
        have.add((a, b))
        keep.append((a, b))
        npar[b] += 1
        n_add -= 1
    types = [("neither", "source", "sink")[t] for t in dag.node_type]
    return create_network(source=[e[0] for e in keep], target=[e[1] for e in keep],
                          node_labels=list(range(dag.n_nodes)), node_type=types)


def simulate_numpy(dag: SyntheticDag, n_samples: int, seed: int = 42) -> np.ndarray:
    """(n_samples, n_nodes) float64, Fortran order (R's column-major layout)."""
    rng = np.random.default_rng(seed)
    X = np.empty((n_samples, dag.n_nodes), order="F")
    for j in range(dag.n_nodes):
        col = rng.standard_normal(n_samples)
        for p, w in zip(dag.parents[j], dag.weights[j]):
            col += w * X[:, p]
        col -= col.mean()
        col /= col.std()
        X[:, j] = col
    return X


def simulate_torch(dag: SyntheticDag, n_samples: int, seed: int = 42, device="cuda"):
    """Device-resident X as a (n_nodes, n_samples) float64 tensor: row p is column p of the
    R matrix, i.e. the memory is column-major n_samples x n_nodes with ld = n_samples."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    X = torch.empty((dag.n_nodes, n_samples), dtype=torch.float64, device=device)
    for j in range(dag.n_nodes):
        col = torch.randn(n_samples, dtype=torch.float64, device=device, generator=gen)
        for p, w in zip(dag.parents[j], dag.weights[j]):
            col.add_(X[p], alpha=w)
        col.sub_(col.mean())
        col.div_(col.std(unbiased=False))
        X[j] = col
    return X


def chain_seeds(n_chains: int, first_chain: int = 0) -> np.ndarray:
    """Wichmann-Hill seed triples by GLOBAL chain index (independent of the GPU count):
    chain 0 = the reference's 10437/13568/30524 (random4f.h:19-21), chain c > 0 from
    SplitMix64(1234 + c) mapped into [1,30268] x [1,30306] x [1,30322].  Mirrors the
    default of bn_run (csrc/bn_api.cu)."""
    mask = (1 << 64) - 1

    def splitmix(state):
        state = (state + 0x9E3779B97F4A7C15) & mask
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        return state, z ^ (z >> 31)

    out = np.zeros((n_chains, 3), dtype=np.int32)
    for i in range(n_chains):
        c = first_chain + i
        if c == 0:
            out[i] = (10437, 13568, 30524)
            continue
        st = 1234 + c
        st, a = splitmix(st)
        st, b = splitmix(st)
        st, d = splitmix(st)
        out[i] = (1 + a % 30268, 1 + b % 30306, 1 + d % 30322)
    return out
