"""Input side of the path: the ``bayesnetworks_network`` object and fixture readers.

Host-side mirror of the reference's R helpers (pure input shaping, no compute):

* :func:`create_network`  -- ``R/bnetwork.R:34-80``  (validation, 1-based index
  encoding against ``node_labels``, edges sorted by target with a stable order)
* :func:`read_data`       -- ``R/aaa.R:9-14``        (drop columns 1,3..7 of the ``.dat``)
* :func:`read_dag`        -- ``R/aaa.R:27-49`` / ``data-raw/network.R:9-25``
  (``Npar nodetype parents...`` records, 0-based parents, CR line ends)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

NODE_TYPES = ("neither", "source", "sink")
_NODE_TYPE_CODE = {"neither": 0, "source": 1, "sink": 2}


@dataclass
class Network:
    """``bayesnetworks_network``: 1-based ``source``/``target`` positions in ``node_labels``."""

    source: np.ndarray
    target: np.ndarray
    node_labels: list
    node_type: list = field(default_factory=list)

    @property
    def n_nodes(self) -> int:
        return len(self.node_labels)

    def node_type_codes(self) -> np.ndarray:
        """``c(neither=0, source=1, sink=2)[graph$node_type]`` (``R/bn_mcmc.R:15-17``)."""
        return np.asarray([_NODE_TYPE_CODE[t] for t in self.node_type], dtype=np.int32)


def create_network(source: Sequence = (), target: Sequence = (),
                   node_labels: Optional[Sequence] = None,
                   node_type: Optional[Sequence[str]] = None) -> Network:
    """Validate and index-encode an edge list (``R/bnetwork.R:34-80``).

    Error messages follow the reference so its testthat cases
    (``tests/testthat/test-bnetwork.R``) translate one to one.
    """
    source = list(source)
    target = list(target)

    def _kind(xs):
        if not xs:
            return None
        return "character" if isinstance(xs[0], str) else "numeric"

    if _kind(source) is not None and _kind(target) is not None and _kind(source) != _kind(target):
        raise ValueError("`source` and `sink` must be the same type.")
    if len(source) != len(target):
        raise ValueError("`source` and `sink` must be the same length.")
    if any(s == t for s, t in zip(source, target)):
        raise ValueError("`target` and `source` cannot be the same for an egde.")
    if node_labels is None:
        if node_type is not None:
            raise ValueError("`node_type` cannot be specified if `node_labels` is left unspecified.")
        node_labels = sorted(set(source) | set(target))
    node_labels = list(node_labels)
    if node_type is None:
        node_type = ["neither"] * len(node_labels)
    node_type = list(node_type)
    if len(node_type) != len(node_labels):
        raise ValueError("`node_type` must be the same length as `node_labels`.")
    for t in node_type:
        if t not in _NODE_TYPE_CODE:
            raise ValueError("`node_type` must be one of \"neither\", \"source\" or \"sink\".")
    pos = {}
    for i, lab in enumerate(node_labels):
        pos.setdefault(lab, i + 1)  # match(): first occurrence, 1-based
    if not all(x in pos for x in set(source) | set(target)):
        raise ValueError("All nodes in `source` and `target` must be specified in `node_labels`")
    src = np.asarray([pos[s] for s in source], dtype=np.int32)
    tgt = np.asarray([pos[t] for t in target], dtype=np.int32)
    order = np.argsort(tgt, kind="stable")  # R's order() is stable
    return Network(source=src[order], target=tgt[order], node_labels=node_labels,
                   node_type=node_type)


def read_data(path: str) -> np.ndarray:
    """``.dat`` -> (n_samples, n_nodes) float64 matrix; keeps column 2 and 8.. (``R/aaa.R:11``)."""
    rows = []
    with open(path, "r", newline="") as fh:
        for line in fh.read().replace("\r", "\n").split("\n"):
            parts = line.split()
            if parts:
                rows.append(parts)
    full = np.asarray(rows, dtype=np.float64)
    keep = [1] + list(range(7, full.shape[1]))
    return np.ascontiguousarray(full[:, keep])


def read_dag(path: str) -> Network:
    """``.dag.txt`` -> :class:`Network` (record = ``Npar nodetype parent...``, 0-based parents).

    Follows ``data-raw/network.R:9-25``: edges are emitted node by node in file
    order (``src = par[node, 1..Npar]``, ``target = node``), labels are ``0..P-1``.
    """
    with open(path, "r", newline="") as fh:
        recs = [ln.split() for ln in fh.read().replace("\r", "\n").split("\n") if ln.strip()]
    src, tgt, types = [], [], []
    for node, rec in enumerate(recs):
        npar, ntype = int(rec[0]), int(rec[1])
        pars = [int(x) for x in rec[2:2 + npar]]
        if len(pars) != npar:
            raise ValueError(f"dag record {node}: expected {npar} parents, found {len(pars)}")
        types.append(NODE_TYPES[ntype])
        for p in pars:
            src.append(p)
            tgt.append(node)
    return create_network(source=src, target=tgt, node_labels=list(range(len(recs))),
                          node_type=types)
