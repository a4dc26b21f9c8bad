"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle.

``Oracle``  wraps oracle/_build/libbn_oracle.so (the plain-C restatement, bn_oracle.c).
``Ref``     wraps oracle/_ref/libbnref.so (the reference's own sources compiled
            unmodified against ref_shim/Rcpp.h; present only where it was built).

Nothing under bayesnetworks_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libbn_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libbnref.so")
LEGACY_BIN = os.path.join(HERE, "_ref", "legacy_main")
REFERENCE_ROOT = "/root/reference"

RNG_WH, RNG_RMT, RNG_REPLAY = 0, 1, 2
WH_DEFAULT_SEEDS = (10437, 13568, 30524)  # Bayes-networks/random4f.h:19-21


def build(ref: Optional[bool] = None) -> None:
    """Compile the C restatement and, where /root/reference exists, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref is None:
        ref = os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", f"REF={REFERENCE_ROOT}"])


def have_ref() -> bool:
    return os.path.exists(REF_SO)


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


class _Rng(C.Structure):
    _fields_ = [("kind", C.c_int), ("ix", C.c_int), ("iy", C.c_int), ("iz", C.c_int),
                ("mt", C.c_uint32 * 624), ("mti", C.c_int),
                ("replay", _dp), ("replay_len", C.c_long), ("draws", C.c_long)]


class _Trace(C.Structure):
    _fields_ = [("capacity", C.c_int), ("n_rows", C.c_int),
                ("iter", _ip), ("changed_node", _ip), ("movetype", _ip),
                ("global_ll", _dp), ("additions", _ip), ("deletions", _ip),
                ("fn", _ip), ("fp", _ip),
                ("npar_changed", _ip), ("log_prior", _dp), ("hr", _dp),
                ("total_edges", _ip), ("agree", _ip)]


class _MoveLog(C.Structure):
    _fields_ = [("capacity", C.c_long), ("n", C.c_long), ("iter", _ip),
                ("movetype", C.POINTER(C.c_byte)), ("child", _ip), ("parent", _ip),
                ("valid", C.POINTER(C.c_byte)), ("accepted", C.POINTER(C.c_byte))]


class _Counters(C.Structure):
    _fields_ = [("uniforms", C.c_long), ("proposed", C.c_int * 3), ("reject", C.c_int * 3),
                ("n_nonpd", C.c_int), ("total_edges_member", C.c_int),
                ("fp_member", C.c_int), ("fn_member", C.c_int)]


class _Args(C.Structure):
    _fields_ = [("X", _dp), ("N", C.c_int), ("P", C.c_int),
                ("src", _ip), ("tgt", _ip), ("n_edges", C.c_int), ("node_type", _ip),
                ("max_par", C.c_int), ("phi", C.c_double), ("omega", C.c_double),
                ("initial_network", C.c_int), ("drop", C.c_int), ("n_iter", C.c_int),
                ("output", C.c_int), ("legacy", C.c_int), ("pad_dim", C.c_int)]


@dataclass
class McmcResult:
    """Trace columns (src/network.h:353-364) plus everything the tests compare."""
    iter: np.ndarray
    ChangedNode: np.ndarray
    movetype: np.ndarray
    globalLL: np.ndarray
    additions: np.ndarray
    deletions: np.ndarray
    FN: np.ndarray
    FP: np.ndarray
    uniforms: int = 0
    proposed: tuple = ()
    reject: tuple = ()
    n_nonpd: int = 0
    final_parents: Optional[np.ndarray] = None   # [P, max_par], -1 padded
    final_npar: Optional[np.ndarray] = None
    moves: Optional[dict] = None                  # per-iteration move log
    legacy: Optional[dict] = None                 # legacy extra columns

    def edges(self):
        """(parent, child) pairs, 0-based, per-child list order."""
        out = []
        for c in range(len(self.final_npar)):
            for e in range(int(self.final_npar[c])):
                out.append((int(self.final_parents[c, e]), c))
        return out

    def accepted_moves(self):
        m = self.moves
        sel = m["accepted"] == 1
        return np.stack([m["iter"][sel], m["movetype"][sel], m["child"][sel], m["parent"][sel]], 1)


def _colmajor(X):
    return np.asfortranarray(np.asarray(X, dtype=np.float64))


class Oracle:
    """The plain-C restatement."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.bno_rng_uniform.restype = C.c_double
        L.bno_score.restype = C.c_double
        L.bno_mcmc.restype = C.c_int
        L.bno_invert_pds.restype = C.c_int
        L.bno_cholesky_decomp.restype = C.c_int

    # -- RNG ---------------------------------------------------------------
    def _rng(self, kind, seeds=None, replay=None):
        r = _Rng()
        if kind == RNG_WH:
            s = seeds if seeds is not None else WH_DEFAULT_SEEDS
            self.lib.bno_rng_init_wh(C.byref(r), int(s[0]), int(s[1]), int(s[2]))
        elif kind == RNG_RMT:
            self.lib.bno_rng_init_rmt(C.byref(r), C.c_uint32(int(seeds[0]) if np.ndim(seeds) else int(seeds)))
        else:
            self.lib.bno_rng_init_replay(C.byref(r), _d(replay), C.c_long(len(replay)))
        return r

    def rmt_state_after(self, seed, n_draws, state=None):
        """R's Mersenne-Twister state as .Random.seed[2:626] (position + 624 words, int32) after
        ``set.seed(seed)`` -- or starting from ``state`` (625 ints) -- and ``n_draws`` uniforms."""
        r = self._rng(RNG_RMT, (seed,) if state is None else (0,))
        if state is not None:
            st = (np.asarray(state).astype(np.int64) & 0xFFFFFFFF).astype(np.uint32)
            r.mti = int(st[0])
            for i in range(624):
                r.mt[i] = int(st[1 + i])
        self.lib.bno_rng_skip(C.byref(r), C.c_long(int(n_draws)))
        out = np.empty(625, dtype=np.uint32)
        out[0] = r.mti
        out[1:] = np.frombuffer(r.mt, dtype=np.uint32)
        return out.view(np.int32)

    def uniforms(self, n, kind=RNG_WH, seeds=None):
        r = self._rng(kind, seeds)
        return np.array([self.lib.bno_rng_uniform(C.byref(r)) for _ in range(n)])

    # -- linear algebra ------------------------------------------------------
    def invert_pds(self, A):
        A = np.ascontiguousarray(A, dtype=np.float64)
        n = A.shape[0]
        out = np.empty_like(A)
        rc = self.lib.bno_invert_pds(_d(A), n, _d(out))
        return rc, out

    def cholesky(self, A):
        A = np.ascontiguousarray(A, dtype=np.float64)
        n = A.shape[0]
        out = np.empty_like(A)
        rc = self.lib.bno_cholesky_decomp(_d(A), n, _d(out))
        return rc, out

    # -- sufficient statistics / scores -------------------------------------
    def gram(self, X):
        Xf = _colmajor(X)
        N, P = Xf.shape
        sumX = np.empty(P)
        sumXX = np.empty((P, P), order="F")
        self.lib.bno_gram(_d(Xf), N, P, _d(sumX), _d(sumXX))
        return sumX, np.ascontiguousarray(sumXX)

    def score(self, X, p, parents, pad_dim=0, stats=None):
        Xf = _colmajor(X)
        N, P = Xf.shape
        sumX, sumXX = stats if stats is not None else self.gram(Xf)
        sumXXf = np.asfortranarray(sumXX)
        par = np.asarray(parents, dtype=np.int32)
        err = C.c_int(0)
        s = self.lib.bno_score(_d(Xf), N, P, _d(sumX), _d(sumXXf), int(p), _i(par), len(par),
                               int(pad_dim), C.byref(err))
        return s, err.value

    def score_graph(self, X, parents, npar, pad_dim=0):
        """Scores of every node; ``parents`` is [P, max_par] int32."""
        Xf = _colmajor(X)
        N, P = Xf.shape
        parents = np.ascontiguousarray(parents, dtype=np.int32)
        npar = np.ascontiguousarray(npar, dtype=np.int32)
        out = np.empty(P)
        self.lib.bno_score_graph(_d(Xf), N, P, _i(parents), _i(npar), parents.shape[1],
                                 int(pad_dim), _d(out))
        return out

    # -- the chain -----------------------------------------------------------
    def mcmc(self, X, src_1b, tgt_1b, node_type, max_par=50, phi=1.0, omega=6.9,
             initial_network=2, drop=0, n_iter=1000, output=100, rng_kind=RNG_WH,
             seeds=None, replay=None, legacy=False, pad_dim=0, log_moves=True) -> McmcResult:
        Xf = _colmajor(X)
        N, P = Xf.shape
        src = np.ascontiguousarray(src_1b, dtype=np.int32)
        tgt = np.ascontiguousarray(tgt_1b, dtype=np.int32)
        nt = np.ascontiguousarray(node_type, dtype=np.int32)
        a = _Args(_d(Xf), N, P, _i(src), _i(tgt), len(src), _i(nt), int(max_par), float(phi),
                  float(omega), int(initial_network), int(drop), int(n_iter), int(output),
                  int(bool(legacy)), int(pad_dim))
        cap = (n_iter + output - 1) // output + 1
        ints = {k: np.zeros(cap, dtype=np.int32) for k in
                ("iter", "changed", "movetype", "additions", "deletions", "fn", "fp",
                 "npar_changed", "total_edges", "agree")}
        dbl = {k: np.zeros(cap) for k in ("gll", "log_prior", "hr")}
        tr = _Trace(cap, 0, _i(ints["iter"]), _i(ints["changed"]), _i(ints["movetype"]),
                    _d(dbl["gll"]), _i(ints["additions"]), _i(ints["deletions"]), _i(ints["fn"]),
                    _i(ints["fp"]), _i(ints["npar_changed"]), _d(dbl["log_prior"]), _d(dbl["hr"]),
                    _i(ints["total_edges"]), _i(ints["agree"]))
        mv = None
        mv_arrays = None
        if log_moves:
            mv_arrays = dict(iter=np.zeros(n_iter, np.int32), movetype=np.zeros(n_iter, np.int8),
                             child=np.zeros(n_iter, np.int32), parent=np.zeros(n_iter, np.int32),
                             valid=np.zeros(n_iter, np.int8), accepted=np.zeros(n_iter, np.int8))
            bp = C.POINTER(C.c_byte)
            mv = _MoveLog(n_iter, 0, _i(mv_arrays["iter"]),
                          mv_arrays["movetype"].ctypes.data_as(bp), _i(mv_arrays["child"]),
                          _i(mv_arrays["parent"]), mv_arrays["valid"].ctypes.data_as(bp),
                          mv_arrays["accepted"].ctypes.data_as(bp))
        cnt = _Counters()
        fpar = np.full((P, max_par), -1, dtype=np.int32)
        fnpar = np.zeros(P, dtype=np.int32)
        rng = self._rng(rng_kind, seeds, replay)
        rc = self.lib.bno_mcmc(C.byref(a), C.byref(rng), C.byref(tr),
                               C.byref(mv) if mv is not None else None, C.byref(cnt),
                               _i(fpar), _i(fnpar))
        if rc != 0:
            raise RuntimeError(f"bno_mcmc failed rc={rc}")
        n = tr.n_rows
        for c in range(P):
            fpar[c, fnpar[c]:] = -1
        res = McmcResult(iter=ints["iter"][:n].copy(), ChangedNode=ints["changed"][:n].copy(),
                         movetype=ints["movetype"][:n].copy(), globalLL=dbl["gll"][:n].copy(),
                         additions=ints["additions"][:n].copy(),
                         deletions=ints["deletions"][:n].copy(), FN=ints["fn"][:n].copy(),
                         FP=ints["fp"][:n].copy(), uniforms=int(cnt.uniforms),
                         proposed=tuple(cnt.proposed), reject=tuple(cnt.reject),
                         n_nonpd=int(cnt.n_nonpd), final_parents=fpar, final_npar=fnpar)
        if mv_arrays is not None:
            m = int(mv.n)
            res.moves = {k: v[:m].copy() for k, v in mv_arrays.items()}
        if legacy:
            res.legacy = dict(Npar=ints["npar_changed"][:n].copy(), lnPrior=dbl["log_prior"][:n].copy(),
                              HR=dbl["hr"][:n].copy(), Edges=ints["total_edges"][:n].copy(),
                              Agree=ints["agree"][:n].copy())
        return res


class Ref:
    """The reference's own sources, compiled unmodified (oracle/_ref/libbnref.so)."""

    def __init__(self):
        if not have_ref():
            raise FileNotFoundError(f"{REF_SO} not built (needs /root/reference; run oracle.build())")
        self.lib = C.CDLL(REF_SO)
        self.lib.ref_main_fun.restype = C.c_int

    def main_fun(self, X, src_1b, tgt_1b, node_type, MaxPar=50, phi=1.0, omega=6.9,
                 InitialNetwork=2, drop=0, N=1000, output=10, rng_kind=RNG_RMT, seeds=(1234,),
                 replay=None) -> McmcResult:
        Xf = _colmajor(X)
        n, P = Xf.shape
        src = np.ascontiguousarray(src_1b, dtype=np.int32)
        tgt = np.ascontiguousarray(tgt_1b, dtype=np.int32)
        nt = np.ascontiguousarray(node_type, dtype=np.int32)
        sd = np.zeros(3, dtype=np.int32)
        sd[:len(seeds)] = seeds
        cap = (N + output - 1) // output + 1
        ints = {k: np.zeros(cap, dtype=np.int32) for k in
                ("iter", "changed", "movetype", "additions", "deletions", "fn", "fp")}
        gll = np.zeros(cap)
        draws = C.c_long(0)
        diag = C.c_long(0)
        rp = np.ascontiguousarray(replay, dtype=np.float64) if replay is not None else None
        rows = self.lib.ref_main_fun(
            _d(Xf), n, P, _i(src), _i(tgt), len(src), _i(nt), int(MaxPar), C.c_double(phi),
            C.c_double(omega), int(InitialNetwork), int(drop), int(N), int(output), int(rng_kind),
            _i(sd), _d(rp), C.c_long(0 if rp is None else len(rp)), cap, _i(ints["iter"]),
            _i(ints["changed"]), _i(ints["movetype"]), _d(gll), _i(ints["additions"]),
            _i(ints["deletions"]), _i(ints["fn"]), _i(ints["fp"]), C.byref(draws), C.byref(diag))
        if rows < 0:
            raise RuntimeError(f"ref_main_fun failed rc={rows}")
        r = min(rows, cap)
        return McmcResult(iter=ints["iter"][:r].copy(), ChangedNode=ints["changed"][:r].copy(),
                          movetype=ints["movetype"][:r].copy(), globalLL=gll[:r].copy(),
                          additions=ints["additions"][:r].copy(),
                          deletions=ints["deletions"][:r].copy(), FN=ints["fn"][:r].copy(),
                          FP=ints["fp"][:r].copy(), uniforms=int(draws.value))

    def scores(self, X, src_1b, tgt_1b, node_type, MaxPar=50):
        Xf = _colmajor(X)
        n, P = Xf.shape
        src = np.ascontiguousarray(src_1b, dtype=np.int32)
        tgt = np.ascontiguousarray(tgt_1b, dtype=np.int32)
        nt = np.ascontiguousarray(node_type, dtype=np.int32)
        out = np.empty(P)
        gll = C.c_double(0)
        lp = C.c_double(0)
        rc = self.lib.ref_scores(_d(Xf), n, P, _i(src), _i(tgt), len(src), _i(nt), int(MaxPar),
                                 _d(out), C.byref(gll), C.byref(lp))
        if rc != 0:
            raise RuntimeError("ref_scores failed")
        return out, gll.value, lp.value

    def gram(self, X):
        Xf = _colmajor(X)
        n, P = Xf.shape
        sumX = np.empty(P)
        sumXX = np.empty((P, P), order="F")
        if self.lib.ref_gram(_d(Xf), n, P, _d(sumX), _d(sumXX)) != 0:
            raise RuntimeError("ref_gram failed")
        return sumX, np.ascontiguousarray(sumXX)

    def invert_pds(self, A):
        A = np.ascontiguousarray(A, dtype=np.float64).copy()
        n = A.shape[0]
        out = np.empty_like(A)
        rc = self.lib.ref_invert_pds(_d(A), n, _d(out))
        return rc, out
