/* TEST INFRASTRUCTURE ONLY -- see bn_oracle.h.
 *
 * Plain-C restatement of the reference's CPU algorithm.  It is deliberately
 * scalar and keeps the reference's operation ORDER (summation order, the
 * descending k loop of the Cholesky, left-to-right evaluation of HR, the
 * per-iteration uniform draw order and every quirk listed in SURVEY.md
 * Appendix A) so that, compiled with the same compiler flags as oracle/_ref
 * (-O2, no FMA contraction), it reproduces the reference bit for bit.
 */
#include "bn_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ========================================================================= */
/* Uniform generators                                                         */
/* ========================================================================= */

void bno_rng_init_wh(bno_rng* r, int ix, int iy, int iz) {
  memset(r, 0, sizeof(*r));
  r->kind = BNO_RNG_WH;
  r->ix = ix; r->iy = iy; r->iz = iz;
}

/* Wichmann-Hill AS183 as written in Bayes-networks/random4f.h:27-40: three
 * Schrage-style LCG updates, a conditional +modulus, then the FP64 combine
 * ix/30269.0 + iy/30307.0 + iz/30323.0 (left to right) minus its floor. */
static double wh_next(bno_rng* r) {
  int qx = (int)floor(r->ix / 177.0);
  int qy = (int)floor(r->iy / 176.0);
  int qz = (int)floor(r->iz / 178.0);
  r->ix = 171 * (r->ix - 177 * qx) - 2 * qx;
  r->iy = 172 * (r->iy - 176 * qy) - 35 * qy;
  r->iz = 170 * (r->iz - 178 * qz) - 63 * qz;
  if (r->ix < 0) r->ix += 30269;
  if (r->iy < 0) r->iy += 30307;
  if (r->iz < 0) r->iz += 30323;
  double v = r->ix / 30269.0 + r->iy / 30307.0 + r->iz / 30323.0;
  int whole = (int)floor(v);
  return v - whole;
}

/* R's set.seed(seed) for the default Mersenne-Twister kind (R sources,
 * src/main/RNG.c: RNG_Init + FixupSeeds): the integer seed is scrambled by an
 * LCG (69069*s+1) 50 times, then fills the 625-word i_seed[]; word 0 is the
 * position counter (forced to 624 = "regenerate"), words 1..624 are mt[]. */
void bno_rng_init_rmt(bno_rng* r, uint32_t seed) {
  memset(r, 0, sizeof(*r));
  r->kind = BNO_RNG_RMT;
  for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
  uint32_t first = 0;
  for (int j = 0; j < 625; j++) {
    seed = 69069u * seed + 1u;
    if (j == 0) first = seed; else r->mt[j - 1] = seed;
  }
  (void)first;
  r->mti = 624;
}

/* MT19937 generation + R's tempering, scaling to [0,1) and open-interval
 * fix-up (RNG.c: MT_genrand, fixup). */
static double rmt_next(bno_rng* r) {
  static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
  uint32_t y;
  if (r->mti >= 624) {
    int kk;
    for (kk = 0; kk < 624 - 397; kk++) {
      y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
      r->mt[kk] = r->mt[kk + 397] ^ (y >> 1) ^ mag01[y & 0x1u];
    }
    for (; kk < 623; kk++) {
      y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
      r->mt[kk] = r->mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 0x1u];
    }
    y = (r->mt[623] & 0x80000000u) | (r->mt[0] & 0x7fffffffu);
    r->mt[623] = r->mt[396] ^ (y >> 1) ^ mag01[y & 0x1u];
    r->mti = 0;
  }
  y = r->mt[r->mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  double v = (double)y * 2.3283064365386963e-10; /* [0,1) */
  const double i2_32m1 = 2.328306437080797e-10;
  if (v <= 0.0) return 0.5 * i2_32m1;
  if ((1.0 - v) <= 0.0) return 1.0 - 0.5 * i2_32m1;
  return v;
}

void bno_rng_init_replay(bno_rng* r, const double* u, long n) {
  memset(r, 0, sizeof(*r));
  r->kind = BNO_RNG_REPLAY;
  r->replay = u;
  r->replay_len = n;
}

double bno_rng_uniform(bno_rng* r) {
  double u;
  switch (r->kind) {
    case BNO_RNG_WH: u = wh_next(r); break;
    case BNO_RNG_RMT: u = rmt_next(r); break;
    default:
      u = (r->draws < r->replay_len) ? r->replay[r->draws] : 0.5;
      break;
  }
  r->draws++;
  return u;
}

void bno_rng_skip(bno_rng* r, long n) {
  for (long i = 0; i < n; i++) (void)bno_rng_uniform(r);
}

double bno_rng_uniform_cb(void* r) { return bno_rng_uniform((bno_rng*)r); }

/* ========================================================================= */
/* Cholesky / PDS inverse (src/cholesky22.h)                                   */
/* ========================================================================= */

/* src/cholesky22.h:25-66.  x, c: row-major n*n.  Reads the upper triangle of
 * x, writes the lower-triangular factor into c (upper part of c keeps x).
 * The k loop runs DOWN from i-1 to 0.  Returns 4 when a pivot is <= 0. */
int bno_cholesky_decomp(const double* x, int n, double* c) {
  if (!x) return 1;
  if (n < 1) return 2;
  if (!c) return 3;
  for (int i = 0; i < n * n; i++) c[i] = x[i];
  for (int i = 0; i < n; i++) {
    for (int j = i; j < n; j++) {
      double acc = x[i * n + j];
      for (int k = i - 1; k >= 0; k--) acc -= c[i * n + k] * c[j * n + k];
      if (j == i) {
        if (acc <= 0.0) return 4;
        c[i * n + i] = sqrt(acc);
      } else {
        c[j * n + i] = acc / c[i * n + i];
      }
    }
  }
  return 0;
}

/* src/cholesky22.h:92-170 (and the flat wrapper :202-242): c <- I, factor,
 * then for every row of c a forward substitution (k ascending) followed by a
 * back substitution (k descending).  A failed factorisation returns rc+10 and
 * leaves c = I (the reference then carries on with that). */
int bno_invert_pds(const double* x, int n, double* c) {
  if (!x) return 1;
  if (n < 1) return 2;
  if (!c) return 1;
  double* s = (double*)malloc(sizeof(double) * (size_t)n * (size_t)n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) c[i * n + j] = (i == j) ? 1.0 : 0.0;
  int rc = bno_cholesky_decomp(x, n, s);
  if (rc != 0) {
    free(s);
    return rc + 10;
  }
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) {
      double acc = c[i * n + j];
      for (int k = 0; k < j; k++) acc -= s[j * n + k] * c[i * n + k];
      c[i * n + j] = acc / s[j * n + j];
    }
    for (int j = n - 1; j >= 0; j--) {
      double acc = c[i * n + j];
      for (int k = n - 1; k > j; k--) acc -= s[k * n + j] * c[i * n + k];
      c[i * n + j] = acc / s[j * n + j];
    }
  }
  free(s);
  return 0;
}

/* ========================================================================= */
/* Sufficient statistics (src/network.h:124-136)                               */
/* ========================================================================= */

/* sample index outermost, then p1, then p2; both triangles accumulated. */
void bno_gram(const double* X, int N, int P, double* sumX, double* sumXX) {
  for (int p = 0; p < P; p++) sumX[p] = 0.0;
  for (long i = 0; i < (long)P * P; i++) sumXX[i] = 0.0;
  for (int n = 0; n < N; n++) {
    for (int p1 = 0; p1 < P; p1++) {
      double a = X[(size_t)n + (size_t)p1 * N];
      sumX[p1] += a;
      for (int p2 = 0; p2 < P; p2++)
        sumXX[(size_t)p1 + (size_t)p2 * P] += a * X[(size_t)n + (size_t)p2 * N];
    }
  }
}

/* Bayes-networks/main.cpp:39,88-94: X is float there, so every product is
 * rounded to binary32 before it is added to the FP64 accumulator. */
static void gram_legacy(const double* X, int N, int P, double* sumX, double* sumXX) {
  for (int p = 0; p < P; p++) sumX[p] = 0.0;
  for (long i = 0; i < (long)P * P; i++) sumXX[i] = 0.0;
  for (int n = 0; n < N; n++) {
    for (int p1 = 0; p1 < P; p1++) {
      float a = (float)X[(size_t)n + (size_t)p1 * N];
      sumX[p1] += a;
      for (int p2 = 0; p2 < P; p2++) {
        float b = (float)X[(size_t)n + (size_t)p2 * N];
        volatile float prod = a * b;
        sumXX[(size_t)p1 + (size_t)p2 * P] += prod;
      }
    }
  }
}

/* ========================================================================= */
/* Node score (src/network.h:183-237)                                          */
/* ========================================================================= */

double bno_score(const double* X, int N, int P, const double* sumX,
                 const double* sumXX, int p, const int* parents, int npar,
                 int pad_dim, int* err) {
  int dim = (pad_dim > npar + 1) ? pad_dim : npar + 1;
  double* SXX = (double*)calloc((size_t)dim * dim, sizeof(double));
  double* SXXinv = (double*)malloc(sizeof(double) * (size_t)dim * dim);
  double* SXY = (double*)calloc((size_t)dim, sizeof(double));
  double* beta = (double*)calloc((size_t)dim, sizeof(double));

  double SY = sumX[p];
  double SYY = sumXX[(size_t)p + (size_t)p * P];
  SXX[0] = N;
  SXY[0] = sumX[p];
  for (int a = 0; a < npar; a++) {
    int p1 = parents[a];
    SXY[a + 1] = sumXX[(size_t)p + (size_t)p1 * P];
    SXX[(a + 1) * dim + 0] = sumX[p1];
    SXX[0 * dim + (a + 1)] = SXX[(a + 1) * dim + 0];
    for (int b = 0; b < npar; b++) {
      int p2 = parents[b];
      SXX[(a + 1) * dim + (b + 1)] = sumXX[(size_t)p1 + (size_t)p2 * P];
    }
  }
  for (int q = npar + 1; q < dim; q++) SXX[q * dim + q] = 1.0;

  int rc = bno_invert_pds(SXX, dim, SXXinv);
  if (err) *err = rc;

  for (int a = 0; a < npar + 1; a++)
    for (int b = 0; b < npar + 1; b++) beta[a] += SXY[b] * SXXinv[a * dim + b];

  double resid2 = 0.0;
  for (int n = 0; n < N; n++) {
    double EX = beta[0];
    for (int a = 0; a < npar; a++)
      EX += beta[a + 1] * X[(size_t)n + (size_t)parents[a] * N];
    resid2 += pow(X[(size_t)n + (size_t)p * N] - EX, 2);
  }
  resid2 /= N - npar - 1;
  SYY -= SY * SY / N;
  SYY /= N - 1;
  double lnLR = -(N / 2.0) * log(resid2 / SYY);

  free(SXX); free(SXXinv); free(SXY); free(beta);
  return lnLR;
}

void bno_score_graph(const double* X, int N, int P, const int* parents,
                     const int* npar, int max_par, int pad_dim, double* out) {
  double* sumX = (double*)malloc(sizeof(double) * (size_t)P);
  double* sumXX = (double*)malloc(sizeof(double) * (size_t)P * P);
  bno_gram(X, N, P, sumX, sumXX);
  for (int p = 0; p < P; p++)
    out[p] = bno_score(X, N, P, sumX, sumXX, p, parents + (size_t)p * max_par,
                       npar[p], pad_dim, NULL);
  free(sumX); free(sumXX);
}

/* ========================================================================= */
/* The chain (src/bayesnet_mcmc.cpp:27-72 + src/network.h)                     */
/* ========================================================================= */

typedef struct {
  const bno_mcmc_args* a;
  int N, P, max_par;
  double* sumX; double* sumXX;
  int* par;       /* [P][max_par] ordered parent lists (edges[child]) */
  int* npar;
  unsigned char* sim_edge; /* [parent + child*P] */
  int n_sim_edges;
  /* members left by the last LogPrior() call (src/network.h:262-275) */
  int total_edges, n_agree, fp, fn;
  int n_nonpd;
} chain_t;

static double chain_score(chain_t* c, int p) {
  int err = 0;
  double s = bno_score(c->a->X_colmajor, c->N, c->P, c->sumX, c->sumXX, p,
                       c->par + (size_t)p * c->max_par, c->npar[p],
                       c->a->pad_dim, &err);
  if (err) c->n_nonpd++;
  return s;
}

/* src/network.h:254-279: full recount, overwrites the members. */
static double chain_log_prior(chain_t* c) {
  c->total_edges = 0;
  c->n_agree = 0;
  for (int p = 0; p < c->P; p++)
    for (int e = 0; e < c->npar[p]; e++) {
      c->total_edges++;
      if (c->sim_edge[(size_t)c->par[(size_t)p * c->max_par + e] + (size_t)p * c->P])
        c->n_agree++;
    }
  c->fp = c->total_edges - c->n_agree;
  c->fn = c->n_sim_edges - c->n_agree;
  int dist = c->fp + c->fn;
  return -c->a->phi * dist - c->a->omega * c->total_edges;
}

/* src/network.h:366-413: BFS from `from` along parent links; true when
 * `target` is reachable (the new edge from->target would close a cycle). */
static int chain_path_exists(chain_t* c, int from, int target) {
  if (from == target) return 1;
  int P = c->P;
  unsigned char* seen = (unsigned char*)calloc((size_t)P, 1);
  int* queue = (int*)malloc(sizeof(int) * (size_t)P);
  int head = 0, tail = 0, found = 0;
  seen[from] = 1;
  queue[tail++] = from;
  while (head < tail && !found) {
    int s = queue[head++];
    for (int e = 0; e < c->npar[s]; e++) {
      int q = c->par[(size_t)s * c->max_par + e];
      if (q == target) { found = 1; break; }
      if (!seen[q]) { seen[q] = 1; queue[tail++] = q; }
    }
  }
  free(seen); free(queue);
  return found;
}

int bno_mcmc(const bno_mcmc_args* a, bno_rng* rng, bno_trace* trace,
             bno_movelog* moves, bno_counters* counters, int* final_parents,
             int* final_npar) {
  chain_t ch;
  memset(&ch, 0, sizeof(ch));
  ch.a = a; ch.N = a->N; ch.P = a->P; ch.max_par = a->max_par;
  int P = a->P, MP = a->max_par;
  /* InitialNetwork: 1 = random start, 2 = empty graph, anything else keeps the supplied graph
   * (src/network.h:148-170) */
  const int init = (a->initial_network == 1 || a->initial_network == 2) ? a->initial_network : 0;

  ch.sumX = (double*)malloc(sizeof(double) * (size_t)P);
  ch.sumXX = (double*)malloc(sizeof(double) * (size_t)P * P);
  if (a->legacy) gram_legacy(a->X_colmajor, a->N, P, ch.sumX, ch.sumXX);
  else bno_gram(a->X_colmajor, a->N, P, ch.sumX, ch.sumXX);

  /* src/network.h:115-122: parent lists from the 1-based edge list (edges[tgt-1].push_back(src-1));
   * src/network.h:138-146: prior adjacency simEdge(parent, child) = 1, NsimEdges counts list entries.
   * The supplied graph may exceed MaxPar parents per node as long as the chain does not start
   * from it (with InitialNetwork == 2 it only feeds simEdge / NsimEdges, :164-169). */
  ch.par = (int*)malloc(sizeof(int) * (size_t)P * MP);
  ch.npar = (int*)calloc((size_t)P, sizeof(int));
  for (long i = 0; i < (long)P * MP; i++) ch.par[i] = -1;
  ch.sim_edge = (unsigned char*)calloc((size_t)P * P, 1);
  for (int e = 0; e < a->n_edges; e++) {
    int child = a->edge_tgt_1b[e] - 1, parent = a->edge_src_1b[e] - 1;
    ch.sim_edge[(size_t)parent + (size_t)child * P] = 1;
    ch.n_sim_edges++;
    if (init == 0) {
      if (ch.npar[child] >= MP) return -2;
      ch.par[(size_t)child * MP + ch.npar[child]++] = parent;
    }
  }
  if (init == 1) {
    /* Random start.  The reference's own version (src/network.h:148-163) writes edges[p][s] into
     * vectors sized by the prior graph (out of bounds) and avoids neither duplicate parents nor
     * cycles: undefined behaviour, NOT a parity target.  This is the defined variant the CUDA path
     * implements (include/bn_b200.h): same draw order from the chain's stream before iteration 0
     * -- for every non-source node Npar = int(MaxPar * u), then each parent as int(P * u) -- but a
     * candidate that is the node itself, a sink, already a parent or cycle-closing is re-drawn,
     * and a slot that finds no parent in 100 draws ends the node's list. */
    for (int p = 0; p < P; p++) {
      if (a->node_type[p] == 1) continue;
      int want = (int)(MP * bno_rng_uniform(rng));
      for (int s = 0; s < want; s++) {
        int found = -1;
        for (int tries = 0; tries < 100 && found < 0; tries++) {
          int src = (int)(P * bno_rng_uniform(rng));
          int ok = (src != p && a->node_type[src] != 2 && !chain_path_exists(&ch, src, p));
          for (int e = 0; e < s; e++) if (ch.par[(size_t)p * MP + e] == src) ok = 0;
          if (ok) found = src;
        }
        if (found < 0) break;
        ch.par[(size_t)p * MP + s] = found;
        ch.npar[p] = s + 1;
      }
    }
  }

  int proposed[3] = {0, 0, 0}, reject[3] = {0, 0, 0};
  int valid = 1;  /* src/bayesnet_mcmc.cpp:40 */
  int movetype = 0, changed = 0;
  int* curr_outputs = (int*)malloc(sizeof(int) * (size_t)P);
  double old_ll = 0, old_prior = 0, new_ll = 0, new_prior = 0;
  if (trace) trace->n_rows = 0;
  if (moves) moves->n = 0;

  for (int i = 0; i < a->n_iter; i++) {
    /* save_graph(): we keep a single-edge undo record instead of the deep copy */
    int undo_child = -1, undo_pos = -1, undo_parent = -1;
    int prop_parent = -1;

    double u_move = bno_rng_uniform(rng);
    int do_add = a->legacy ? (u_move < 0.5 || ch.total_edges < 3)   /* main.cpp:357 */
                           : (u_move > 0.5 || ch.total_edges < 3);  /* bayesnet_mcmc.cpp:48 */
    if (do_add) {
      /* src/network.h:281-306 */
      int newoutput = -1, newinput = -1, found = 0;
      while (!found) {
        newoutput = (int)(P * bno_rng_uniform(rng));
        if (a->node_type[newoutput] != 1 && ch.npar[newoutput] < MP) found = 1;
      }
      found = 0;
      while (!found) {
        newinput = (int)(P * bno_rng_uniform(rng));
        if (a->node_type[newinput] != 2 && newinput != newoutput) found = 1;
        for (int pp = 0; pp < ch.npar[newoutput]; pp++)
          if (newinput == ch.par[(size_t)newoutput * MP + pp]) found = 0;
      }
      changed = newoutput;
      old_ll = chain_score(&ch, changed);
      old_prior = chain_log_prior(&ch);
      ch.par[(size_t)newoutput * MP + ch.npar[newoutput]] = newinput;
      ch.npar[newoutput]++;
      movetype = 1;
      undo_child = newoutput; undo_pos = ch.npar[newoutput] - 1; undo_parent = newinput;
      prop_parent = newinput;
      /* bayesnet_mcmc.cpp:50 -> network.h:415-432; legacy: check disabled (main.cpp:350-351,359) */
      if (!a->legacy) valid = !chain_path_exists(&ch, newinput, newoutput);
    } else {
      /* src/network.h:308-328: the first uniform is drawn and discarded. */
      (void)bno_rng_uniform(rng);
      int cnt = 0;
      for (int p = 0; p < P; p++)
        if (ch.npar[p]) curr_outputs[cnt++] = p;
      int deloutput = curr_outputs[(int)(cnt * bno_rng_uniform(rng))];
      int deledge = (int)(ch.npar[deloutput] * bno_rng_uniform(rng));
      int delinput = ch.par[(size_t)deloutput * MP + deledge];
      changed = deloutput;
      old_ll = chain_score(&ch, changed);
      old_prior = chain_log_prior(&ch);
      for (int e = deledge; e + 1 < ch.npar[deloutput]; e++)
        ch.par[(size_t)deloutput * MP + e] = ch.par[(size_t)deloutput * MP + e + 1];
      ch.npar[deloutput]--;
      movetype = 2;
      undo_child = deloutput; undo_pos = deledge; undo_parent = delinput;
      prop_parent = delinput;
      /* `valid` keeps its previous value (bayesnet_mcmc.cpp:40,50,52). */
    }

    int proposed_type = movetype;
    int accepted = 0;
    if (valid) {
      /* checker(): src/network.h:330-336 */
      if (i >= a->drop) proposed[movetype]++;
      new_ll = chain_score(&ch, changed);
      new_prior = chain_log_prior(&ch);
      double HR = exp(new_ll - old_ll + new_prior - old_prior);
      int rejected = bno_rng_uniform(rng) > HR;
      if (rejected) {
        if (i >= a->drop) reject[movetype]++;
      } else {
        accepted = 1;
      }
      if (rejected) {
        /* restore_graph() */
        if (movetype == 1) {
          ch.npar[undo_child]--;
        } else {
          for (int e = ch.npar[undo_child]; e > undo_pos; e--)
            ch.par[(size_t)undo_child * MP + e] = ch.par[(size_t)undo_child * MP + e - 1];
          ch.par[(size_t)undo_child * MP + undo_pos] = undo_parent;
          ch.npar[undo_child]++;
        }
      }
      if (i % a->output == 0 && trace && trace->n_rows < trace->capacity) {
        /* logger(): src/network.h:338-351 -- globalLL of the KEPT graph,
         * FN/FP members as left by the last LogPrior() (proposed graph). */
        double gll = 0.0;
        for (int p = 0; p < P; p++) gll += chain_score(&ch, p);
        int r = trace->n_rows++;
        trace->iter[r] = i;
        trace->changed_node[r] = changed;
        trace->movetype[r] = movetype;
        trace->global_ll[r] = gll;
        trace->additions[r] = proposed[1] - reject[1];
        trace->deletions[r] = proposed[2] - reject[2];
        trace->fn[r] = ch.fn;
        trace->fp[r] = ch.fp;
        if (trace->npar_changed) trace->npar_changed[r] = ch.npar[changed];
        if (trace->log_prior) trace->log_prior[r] = new_prior;
        if (trace->hr) trace->hr[r] = HR;
        if (trace->total_edges) trace->total_edges[r] = ch.total_edges;
        if (trace->agree) trace->agree[r] = ch.n_agree;
      }
    } else {
      /* restore_graph(); notValid(): src/network.h:434-437 */
      if (movetype == 1) {
        ch.npar[undo_child]--;
      } else {
        for (int e = ch.npar[undo_child]; e > undo_pos; e--)
          ch.par[(size_t)undo_child * MP + e] = ch.par[(size_t)undo_child * MP + e - 1];
        ch.par[(size_t)undo_child * MP + undo_pos] = undo_parent;
        ch.npar[undo_child]++;
      }
      movetype = 0;
      reject[0]++;
    }
    if (moves && moves->n < moves->capacity) {
      long m = moves->n++;
      moves->iter[m] = i;
      moves->movetype[m] = (signed char)proposed_type;
      moves->child[m] = changed;
      moves->parent[m] = prop_parent;
      moves->valid[m] = (signed char)(valid ? 1 : 0);
      moves->accepted[m] = (signed char)accepted;
    }
    if (a->legacy && (i + 1) > a->drop) {
      /* Tabulate(): main.cpp:289-297,392 recounts TotalEdges on the kept graph */
      int te = 0;
      for (int p = 0; p < P; p++) te += ch.npar[p];
      ch.total_edges = te;
    }
  }

  if (counters) {
    counters->uniforms = rng->draws;
    for (int t = 0; t < 3; t++) { counters->proposed[t] = proposed[t]; counters->reject[t] = reject[t]; }
    counters->n_nonpd = ch.n_nonpd;
    counters->total_edges_member = ch.total_edges;
    counters->fp_member = ch.fp; counters->fn_member = ch.fn;
  }
  if (final_parents) memcpy(final_parents, ch.par, sizeof(int) * (size_t)P * MP);
  if (final_npar) memcpy(final_npar, ch.npar, sizeof(int) * (size_t)P);

  free(curr_outputs);
  free(ch.sumX); free(ch.sumXX); free(ch.par); free(ch.npar); free(ch.sim_edge);
  return 0;
}
