// One MCMC chain = one warp.  Replaces the loop of src/bayesnet_mcmc.cpp:45-70
// and network::{propose_addition, propose_deletion, CheckValidity, checker,
// LogPrior, logger} (src/network.h:254-364,415-437).
//
// The reference's chain is strictly sequential, but a rejected proposal leaves
// the graph untouched and rejections are the common case (~99% after burn-in),
// so the warp executes a WINDOW of up to 32 consecutive iterations
// speculatively:
//   phase A  (warp-uniform, integer only): replay the reference's draw order
//            for each iteration assuming all earlier ones in the window were
//            rejected -- move type, rejection-sampled (child, parent), the
//            acyclicity test, the stale `valid` / TotalEdges / FN / FP members
//            (SURVEY.md Appendix A), the acceptance uniform.
//   phase B  (one lane per iteration): score the proposed parent set with a
//            k-dim Cholesky from the centred Gram (score_core.cuh).
//   phase C  Hastings ratio with the reference's expression, first accepted
//            iteration wins; everything after it is discarded and the uniform
//            stream position rewinds to just after it.
// The committed result is exactly the sequential one.
//
// Acyclicity (pathExists, src/network.h:366-413, a BFS per proposal) is an O(1)
// bit test against per-node ancestor bitsets, maintained on accepted moves.
#pragma once

#include "bn_common.cuh"
#include "rng_core.cuh"
#include "score_core.cuh"

namespace bn {

constexpr int WIN = 32;  // speculative window (iterations), one lane each

struct ChainParams {  // read-only, shared by all chains of a run
  int P, max_par, W, n_samples;
  const double* C;      // centred Gram [P][ldc]
  int64_t ldc;
  const uint8_t* node_type;  // [P] 0 neither / 1 source / 2 sink
  const uint8_t* sim_edge;   // [parent + child*P] prior adjacency, src/network.h:138-146
  int n_sim_edges;
  double phi, omega;
  int initial_network, drop, n_iter, output_every;
  int trace_capacity, moves_capacity;
  // prior graph (InitialNetwork == 0 start), [P][max_par] + [P]
  const int* prior_par;
  const int* prior_npar;
};

struct ChainMem {  // per-chain global memory
  int* par;            // [P][max_par] ordered parent lists (edges[child])
  int* npar;           // [P]
  int* born;           // [P][max_par] first counted iteration of the edge (tabulation)
  double* base;        // [P] score of each node under the current graph
  uint32_t* anc;       // [P][W] ancestor bitsets
  int* anc_cnt;        // [P] popcount of anc rows (topological key)
  uint32_t* haspar;    // [W] nodes with >= 1 parent
  unsigned long long* sortbuf;  // [pow2 >= P] scratch for the ancestor rebuild
  // outputs
  int* t_iter; int* t_changed; int* t_movetype; double* t_gll;
  int* t_add; int* t_del; int* t_fn; int* t_fp;
  int* moves;          // [moves_capacity][4]
  int* edge_freq;      // [parent + child*P] or null
};

struct ChainScalars {  // lives in registers (warp-uniform)
  int64_t iter;        // next iteration index
  int64_t read_pos;    // committed uniform stream position
  int valid;           // stale `valid` flag, src/bayesnet_mcmc.cpp:40
  int te_m, fp_m, fn_m;  // members left by the last LogPrior(), src/network.h:262-275
  int te_true, agree_true, n_haspar;
  int proposed[3], reject[3];
  int n_rows, n_moves, n_nonpd;
  int64_t valid_iters;
  int gll_ok; double gll;
  int64_t alg_bytes;   // sum over scored proposals of 8*(k'+1)(k'+2)/2 + 8
  int win;             // current window size
  int windows;
  int status;
};

struct WindowSlots {  // shared memory on the device
  int child[WIN], parent[WIN], pos[WIN], kk[WIN];
  int te_m[WIN], fp_m[WIN], fn_m[WIN];
  int64_t pos_after[WIN];
  double u_acc[WIN], new_score[WIN];
  signed char type[WIN], valid[WIN], do_check[WIN], accept[WIN], nonpd[WIN];
};

BN_HD bool test_bit(const uint32_t* row, int b) { return (row[b >> 5] >> (b & 31)) & 1u; }

// LogPrior value from the integer counts, evaluated like src/network.h:277:
//   - phi * dist - omega * TotalEdges
BN_HD double prior_value(double phi, double omega, int dist, int total_edges) {
  return sub_rn(mul_rn(-phi, (double)dist), mul_rn(omega, (double)total_edges));
}

// index of the k-th (0-based) set bit of a W-word bitset, by the whole warp
BN_HD int select_kth(const uint32_t* bits, int W, int k) {
  const int l = Warp::lane();
  int before = 0;
  for (int w0 = 0; w0 < W; w0 += Warp::NL) {
    const int w = w0 + l;
    uint32_t word = (w < W) ? bits[w] : 0u;
    const int cnt = popc32(word);
    const int incl = Warp::incl_scan(cnt);
    const int total = Warp::shfl(incl, Warp::NL - 1);
    if (k < before + total) {
      const uint32_t m = Warp::ballot(before + incl > k);
      const int src = ffs32(m) - 1;
      int r = k - before - (incl - cnt);
      int res = -1;
      if (l == src) {
        for (int i = 0; i < r; i++) word &= word - 1;
        res = w * 32 + ffs32(word) - 1;
      }
      return Warp::shfl(res, src);
    }
    before += total;
  }
  return -1;
}

// ---------------------------------------------------------------------------
// Ancestor bitsets
// ---------------------------------------------------------------------------

// after adding parent j to child c: every node in {c} u desc(c) gains anc[j] u {j}
BN_HD void anc_after_add(const ChainParams& p, ChainMem& m, int j, int c) {
  const int l = Warp::lane(), W = p.W, P = p.P;
  const uint32_t* aj = m.anc + (int64_t)j * W;
  for (int d0 = 0; d0 < P; d0 += Warp::NL) {
    const int d = d0 + l;
    const int flag = (d < P) && (d == c || test_bit(m.anc + (int64_t)d * W, c));
    uint32_t mask = Warp::ballot(flag);
    while (mask) {
      const int b = ffs32(mask) - 1;
      mask &= mask - 1;
      uint32_t* ad = m.anc + (int64_t)(d0 + b) * W;
      int cnt = 0;
      for (int w = l; w < W; w += Warp::NL) {
        uint32_t v = ad[w] | aj[w];
        if (w == (j >> 5)) v |= 1u << (j & 31);
        ad[w] = v;
        cnt += popc32(v);
      }
      cnt = Warp::sum(cnt);
      if (l == 0) m.anc_cnt[d0 + b] = cnt;
    }
  }
  Warp::sync();
}

// after removing a parent of child c: recompute anc for {c} u desc(c) in a
// topological order.  |anc(x)| < |anc(d)| whenever x is an ancestor of d, so
// sorting by the OLD ancestor counts gives such an order.
BN_HD void anc_after_delete(const ChainParams& p, ChainMem& m, int c) {
  const int l = Warp::lane(), W = p.W, P = p.P;
  int n = 0;
  for (int d0 = 0; d0 < P; d0 += Warp::NL) {
    const int d = d0 + l;
    const int flag = (d < P) && (d == c || test_bit(m.anc + (int64_t)d * W, c));
    const uint32_t mask = Warp::ballot(flag);
    if (flag) {
      const int off = popc32(mask & ((1u << l) - 1u));
      m.sortbuf[n + off] = ((unsigned long long)(uint32_t)m.anc_cnt[d] << 32) | (uint32_t)d;
    }
    n += popc32(mask);
  }
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int i = n + l; i < n2; i += Warp::NL) m.sortbuf[i] = ~0ull;
  Warp::sync();
  // bitonic sort, ascending
  for (int k = 2; k <= n2; k <<= 1) {
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int i = l; i < n2; i += Warp::NL) {
        const int ixj = i ^ jj;
        if (ixj > i) {
          const unsigned long long a = m.sortbuf[i], b = m.sortbuf[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { m.sortbuf[i] = b; m.sortbuf[ixj] = a; }
        }
      }
      Warp::sync();
    }
  }
  for (int idx = 0; idx < n; idx++) {
    const int d = (int)(uint32_t)(m.sortbuf[idx] & 0xffffffffull);
    uint32_t* ad = m.anc + (int64_t)d * W;
    const int* pd = m.par + (int64_t)d * p.max_par;
    const int kd = m.npar[d];
    int cnt = 0;
    for (int w = l; w < W; w += Warp::NL) {
      uint32_t v = 0;
      for (int e = 0; e < kd; e++) {
        const int q = pd[e];
        v |= m.anc[(int64_t)q * W + w];
        if (w == (q >> 5)) v |= 1u << (q & 31);
      }
      ad[w] = v;
      cnt += popc32(v);
    }
    cnt = Warp::sum(cnt);
    if (l == 0) m.anc_cnt[d] = cnt;
    Warp::sync();
  }
}

// full build (chain start from a non-empty graph): Jacobi sweeps to the fixpoint
BN_HD void anc_build_all(const ChainParams& p, ChainMem& m) {
  const int l = Warp::lane(), W = p.W, P = p.P;
  for (int64_t i = l; i < (int64_t)P * W; i += Warp::NL) m.anc[i] = 0u;
  Warp::sync();
  for (int round = 0; round <= P; round++) {
    int changed = 0;
    for (int d = 0; d < P; d++) {
      const int kd = m.npar[d];
      if (kd == 0) continue;
      uint32_t* ad = m.anc + (int64_t)d * W;
      const int* pd = m.par + (int64_t)d * p.max_par;
      for (int w = l; w < W; w += Warp::NL) {
        uint32_t v = 0;
        for (int e = 0; e < kd; e++) {
          const int q = pd[e];
          v |= m.anc[(int64_t)q * W + w];
          if (w == (q >> 5)) v |= 1u << (q & 31);
        }
        if (v != ad[w]) { ad[w] = v; changed = 1; }
      }
      Warp::sync();
    }
    if (Warp::ballot(changed) == 0u) break;
  }
  for (int d = 0; d < P; d++) {
    int cnt = 0;
    for (int w = l; w < W; w += Warp::NL) cnt += popc32(m.anc[(int64_t)d * W + w]);
    cnt = Warp::sum(cnt);
    if (l == 0) m.anc_cnt[d] = cnt;
  }
  Warp::sync();
}

// ---------------------------------------------------------------------------
// Chain start: network::network graph part, src/network.h:115-122,138-170
// ---------------------------------------------------------------------------
template <int KMAX>
BN_HD void chain_init(const ChainParams& p, ChainMem& m, ChainScalars& s) {
  const int l = Warp::lane(), P = p.P, MP = p.max_par;
  for (int64_t i = l; i < (int64_t)P * MP; i += Warp::NL) {
    m.par[i] = (p.initial_network == 0) ? p.prior_par[i] : -1;
    m.born[i] = p.drop;  // edges of the start graph are counted from iteration `drop`
  }
  for (int i = l; i < P; i += Warp::NL) m.npar[i] = (p.initial_network == 0) ? p.prior_npar[i] : 0;
  for (int w = l; w < p.W; w += Warp::NL) m.haspar[w] = 0u;
  Warp::sync();
  if (l == 0) {
    int te = 0, ag = 0, nh = 0;
    for (int c = 0; c < P; c++) {
      const int k = m.npar[c];
      if (k) { m.haspar[c >> 5] |= 1u << (c & 31); nh++; }
      for (int e = 0; e < k; e++) {
        te++;
        if (p.sim_edge[(int64_t)m.par[(int64_t)c * MP + e] + (int64_t)c * P]) ag++;
      }
    }
    s.te_true = te; s.agree_true = ag; s.n_haspar = nh;
  }
  s.te_true = Warp::shfl(s.te_true, 0);
  s.agree_true = Warp::shfl(s.agree_true, 0);
  s.n_haspar = Warp::shfl(s.n_haspar, 0);
  Warp::sync();
  if (s.te_true > 0) anc_build_all(p, m);
  else {
    for (int64_t i = l; i < (int64_t)P * p.W; i += Warp::NL) m.anc[i] = 0u;
    for (int i = l; i < P; i += Warp::NL) m.anc_cnt[i] = 0;
  }
  // base scores
  s.n_nonpd = 0;
  {
    double L[KMAX * (KMAX + 1) / 2], z[KMAX];
    int S[KMAX];
    for (int c = l; c < P; c += Warp::NL) {
      const int k = m.npar[c];
      for (int e = 0; e < k; e++) S[e] = m.par[(int64_t)c * MP + e];
      int npd = 0;
      m.base[c] = score_set(p.C, p.ldc, c, S, k, p.n_samples, L, z, &npd);
    }
  }
  Warp::sync();
  s.iter = 0; s.read_pos = 0;
  s.valid = 1;                       // src/bayesnet_mcmc.cpp:40
  s.te_m = 0; s.fp_m = 0; s.fn_m = 0;  // members start at 0 (src/network.h:49-51,64)
  for (int t = 0; t < 3; t++) { s.proposed[t] = 0; s.reject[t] = 0; }
  s.n_rows = 0; s.n_moves = 0; s.valid_iters = 0; s.alg_bytes = 0;
  s.gll_ok = 0; s.gll = 0.0;
  s.win = 4; s.windows = 0; s.status = 0;
}

// globalLL = sum_p score(p) of the kept graph (LogLikelihood(1), src/network.h:239-247).
// Lane l sums p = l, l+32, ... ascending, then a fixed-order tree: deterministic.
BN_HD double sum_base(const ChainParams& p, const ChainMem& m) {
  double acc = 0.0;
  for (int c = Warp::lane(); c < p.P; c += Warp::NL) acc += m.base[c];
  return Warp::sum(acc);
}

// ---------------------------------------------------------------------------
// Phase A: replay the draw order of `nslots` iterations (warp-uniform).
// Returns the number of slots filled; sets *overflow when a single iteration
// outran the ring.
// ---------------------------------------------------------------------------
BN_HD int phase_a(const ChainParams& p, const ChainMem& m, const ChainScalars& s,
                  const RngStream& rng, WindowSlots& ws, int want, int* overflow) {
  const int P = p.P, MP = p.max_par;
  int64_t pos = s.read_pos;
  int valid = s.valid, te_m = s.te_m, fp_m = s.fp_m, fn_m = s.fn_m;
  const int fp_true = s.te_true - s.agree_true;
  const int fn_true = p.n_sim_edges - s.agree_true;
  const int64_t hi = rng.gen_hi;
  int n = 0;
  *overflow = 0;
#define BN_U(dst)                                        \
  do {                                                   \
    if (pos >= hi) { ovf = 1; dst = 0.75; }              \
    else dst = rng.ubuf[pos & (RNG_CAP - 1)];            \
    pos++;                                               \
  } while (0)
  while (n < want) {
    int ovf = 0;
    double u;
    BN_U(u);  // u_move, src/bayesnet_mcmc.cpp:48
    int type, c = 0, j = 0, e = -1;
    if (u > 0.5 || te_m < 3) {
      // propose_addition, src/network.h:281-306
      for (;;) {
        BN_U(u);
        c = (int)(P * u);
        if (ovf || (p.node_type[c] != 1 && m.npar[c] < MP)) break;
      }
      const int kc = ovf ? 0 : m.npar[c];
      const int* pc = m.par + (int64_t)c * MP;
      for (;;) {
        BN_U(u);
        j = (int)(P * u);
        if (ovf) break;
        int ok = (p.node_type[j] != 2 && j != c);
        for (int q = 0; q < kc; q++) if (pc[q] == j) ok = 0;
        if (ok) break;
      }
      type = 1;
      te_m = s.te_true; fp_m = fp_true; fn_m = fn_true;  // OldLogPrior = LogPrior(), :302
      // CheckValidity -> pathExists (src/network.h:366-432): is c an ancestor of j?
      if (!ovf) valid = !(j == c || test_bit(m.anc + (int64_t)j * p.W, c));
    } else {
      // propose_deletion, src/network.h:308-328
      BN_U(u);  // drawn and discarded (:309)
      BN_U(u);
      const int idx = (int)(s.n_haspar * u);
      BN_U(u);
      if (!ovf) {
        c = select_kth(m.haspar, p.W, idx);
        e = (int)(m.npar[c] * u);
        j = m.par[(int64_t)c * MP + e];
      }
      type = 2;
      te_m = s.te_true; fp_m = fp_true; fn_m = fn_true;  // :323
      // `valid` keeps the previous iteration's value (src/bayesnet_mcmc.cpp:50-52)
    }
    double ua = 0.0;
    if (valid && !ovf) {
      // checker(): NewLogPrior = LogPrior() on the proposed graph, src/network.h:333
      const int ag = p.sim_edge[(int64_t)j + (int64_t)c * P] ? 1 : 0;
      const int te_new = s.te_true + (type == 1 ? 1 : -1);
      const int ag_new = s.agree_true + (type == 1 ? ag : -ag);
      te_m = te_new; fp_m = te_new - ag_new; fn_m = p.n_sim_edges - ag_new;
      BN_U(ua);  // acceptance uniform, :335
    }
    if (ovf) { *overflow = (n == 0); break; }
    if (Warp::lane() == 0) {
      ws.child[n] = c; ws.parent[n] = j; ws.pos[n] = e;
      ws.type[n] = (signed char)type; ws.valid[n] = (signed char)valid;
      ws.te_m[n] = te_m; ws.fp_m[n] = fp_m; ws.fn_m[n] = fn_m;
      ws.pos_after[n] = pos; ws.u_acc[n] = ua;
    }
    n++;
  }
#undef BN_U
  Warp::sync();
  return n;
}

// ---------------------------------------------------------------------------
// Phase B + C for one slot (one lane): score the proposed set and decide.
// ---------------------------------------------------------------------------
template <int KMAX>
BN_HD void phase_bc(const ChainParams& p, const ChainMem& m, const ChainScalars& s,
                    WindowSlots& ws, int i) {
  if (!ws.valid[i]) { ws.accept[i] = 0; ws.nonpd[i] = 0; return; }
  const int c = ws.child[i], j = ws.parent[i], MP = p.max_par;
  const int* pc = m.par + (int64_t)c * MP;
  const int k = m.npar[c];
  double L[KMAX * (KMAX + 1) / 2], z[KMAX];
  int S[KMAX];
  int kk = 0;
  if (ws.type[i] == 1) {
    for (int e = 0; e < k; e++) S[kk++] = pc[e];
    S[kk++] = j;  // push_back, src/network.h:303
  } else {
    const int del = ws.pos[i];
    for (int e = 0; e < k; e++) if (e != del) S[kk++] = pc[e];  // erase keeps the order, :325
  }
  int npd = 0;
  const double nw = score_set(p.C, p.ldc, c, S, kk, p.n_samples, L, z, &npd);
  ws.new_score[i] = nw;
  ws.kk[i] = kk;
  ws.nonpd[i] = (signed char)npd;
  // HR = exp(NewLogLike - OldLogLike + NewLogPrior - OldLogPrior), src/network.h:334
  const int fp_true = s.te_true - s.agree_true, fn_true = p.n_sim_edges - s.agree_true;
  const double old_prior = prior_value(p.phi, p.omega, fp_true + fn_true, s.te_true);
  const double new_prior = prior_value(p.phi, p.omega, ws.fp_m[i] + ws.fn_m[i], ws.te_m[i]);
  const double arg = sub_rn(add_rn(sub_rn(nw, m.base[c]), new_prior), old_prior);
  const double HR = exp(arg);
  ws.accept[i] = (ws.u_acc[i] > HR) ? 0 : 1;  // reject iff runif > HR (NaN accepts), :335
}

// ---------------------------------------------------------------------------
// Commit: counters, trace rows, and the accepted move if any (warp-uniform).
// ---------------------------------------------------------------------------
BN_HD void write_row(const ChainParams& p, ChainMem& m, ChainScalars& s, int64_t it,
                     const WindowSlots& ws, int i) {
  if (!s.gll_ok) { s.gll = sum_base(p, m); s.gll_ok = 1; }
  if (s.n_rows < p.trace_capacity) {
    if (Warp::lane() == 0) {
      const int r = s.n_rows;
      m.t_iter[r] = (int)it;
      m.t_changed[r] = ws.child[i];
      m.t_movetype[r] = ws.type[i];
      m.t_gll[r] = s.gll;
      m.t_add[r] = s.proposed[1] - s.reject[1];
      m.t_del[r] = s.proposed[2] - s.reject[2];
      m.t_fn[r] = ws.fn_m[i];
      m.t_fp[r] = ws.fp_m[i];
    }
    s.n_rows++;
  }
}

BN_HD void apply_move(const ChainParams& p, ChainMem& m, ChainScalars& s, int64_t it,
                      const WindowSlots& ws, int i) {
  const int c = ws.child[i], j = ws.parent[i], MP = p.max_par, l = Warp::lane();
  int* pc = m.par + (int64_t)c * MP;
  int* bc = m.born + (int64_t)c * MP;
  const int k = m.npar[c];
  const int ag = p.sim_edge[(int64_t)j + (int64_t)c * p.P] ? 1 : 0;
  const int64_t first_counted = (it > p.drop) ? it : p.drop;  // Tabulate(): main.cpp:392
  Warp::sync();
  if (ws.type[i] == 1) {
    if (l == 0) {
      pc[k] = j; bc[k] = (int)first_counted; m.npar[c] = k + 1;
      if (k == 0) m.haspar[c >> 5] |= 1u << (c & 31);
      m.base[c] = ws.new_score[i];
    }
    if (k == 0) s.n_haspar++;
    s.te_true++; s.agree_true += ag;
    Warp::sync();
    anc_after_add(p, m, j, c);
  } else {
    const int del = ws.pos[i];
    if (l == 0) {
      if (m.edge_freq) {
        const int64_t cnt = first_counted - bc[del];
        if (cnt > 0) m.edge_freq[(int64_t)j + (int64_t)c * p.P] += (int)cnt;
      }
      for (int e = del; e + 1 < k; e++) { pc[e] = pc[e + 1]; bc[e] = bc[e + 1]; }
      pc[k - 1] = -1;
      m.npar[c] = k - 1;
      if (k == 1) m.haspar[c >> 5] &= ~(1u << (c & 31));
      m.base[c] = ws.new_score[i];
    }
    if (k == 1) s.n_haspar--;
    s.te_true--; s.agree_true -= ag;
    Warp::sync();
    anc_after_delete(p, m, c);
  }
  if (s.n_moves < p.moves_capacity) {
    if (l == 0) {
      int* mv = m.moves + (int64_t)s.n_moves * 4;
      mv[0] = (int)it; mv[1] = ws.type[i]; mv[2] = c; mv[3] = j;
    }
  }
  s.n_moves++;
  s.gll_ok = 0;
  Warp::sync();
}

// Commit slots [0, ncommit): all but possibly the last are rejections/invalid.
BN_HD void commit(const ChainParams& p, ChainMem& m, ChainScalars& s, const WindowSlots& ws,
                  int ncommit) {
  for (int i = 0; i < ncommit; i++) {
    const int64_t it = s.iter + i;
    const int type = ws.type[i];
    if (ws.valid[i]) {
      s.valid_iters++;
      s.alg_bytes += 4 * (int64_t)(ws.kk[i] + 1) * (ws.kk[i] + 2) + 8;
      if (it >= p.drop) s.proposed[type]++;  // src/network.h:331
      if (ws.nonpd[i]) s.n_nonpd++;
      if (ws.accept[i]) {
        apply_move(p, m, s, it, ws, i);
      } else if (it >= p.drop) {
        s.reject[type]++;  // src/bayesnet_mcmc.cpp:58
      }
      if (it % p.output_every == 0) write_row(p, m, s, it, ws, i);  // :63-65
    } else {
      s.reject[0]++;  // notValid(), src/network.h:434-437 (not guarded by drop)
    }
  }
  const int last = ncommit - 1;
  s.valid = ws.valid[last];
  s.te_m = ws.te_m[last]; s.fp_m = ws.fp_m[last]; s.fn_m = ws.fn_m[last];
  s.read_pos = ws.pos_after[last];
  s.iter += ncommit;
  // lanes are not in lockstep: nobody may still be reading the slots when lane 0
  // starts writing the next window's
  Warp::sync();
}

// ---------------------------------------------------------------------------
// The whole chain.
// ---------------------------------------------------------------------------
template <int KMAX>
BN_HD void run_chain(const ChainParams& p, ChainMem& m, ChainScalars& s, RngStream& rng,
                     WindowSlots& ws) {
  const int l = Warp::lane();
  chain_init<KMAX>(p, m, s);
  while (s.iter < p.n_iter) {
    rng_top_up(rng, s.read_pos);
    int want = s.win;
    if ((int64_t)want > p.n_iter - s.iter) want = (int)(p.n_iter - s.iter);
    int overflow = 0;
    const int n = phase_a(p, m, s, rng, ws, want, &overflow);
    if (n == 0) {
      s.status = 4;  // BN_ERR_NO_LEGAL_PROPOSAL
      break;
    }
    for (int i = l; i < n; i += Warp::NL) phase_bc<KMAX>(p, m, s, ws, i);
    Warp::sync();
    int first = -1;
    for (int i = 0; i < n; i++) if (ws.valid[i] && ws.accept[i]) { first = i; break; }
    const int ncommit = (first >= 0) ? first + 1 : n;
    commit(p, m, s, ws, ncommit);
    s.windows++;
    if (first >= 0) {
      int w = 2 * (first + 1);
      s.win = w < 2 ? 2 : (w > WIN ? WIN : w);
    } else {
      int w = 2 * s.win;
      s.win = w > WIN ? WIN : w;
    }
  }
  // flush the posterior tabulation of the surviving edges
  if (m.edge_freq) {
    Warp::sync();
    for (int c = l; c < p.P; c += Warp::NL) {
      const int k = m.npar[c];
      for (int e = 0; e < k; e++) {
        const int64_t cnt = (int64_t)p.n_iter - m.born[(int64_t)c * p.max_par + e];
        if (cnt > 0) m.edge_freq[(int64_t)m.par[(int64_t)c * p.max_par + e] + (int64_t)c * p.P] += (int)cnt;
      }
    }
  }
  Warp::sync();
}

}  // namespace bn
