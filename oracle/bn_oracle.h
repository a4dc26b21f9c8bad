/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the structure-MCMC scoring hot path of
 * USCbiostats/bayesnetworks, used as the parity oracle for the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this; libbn_b200.so never links it.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Pinning: the restatement is checked against
 *   (1) the reference's own sources compiled unmodified (oracle/_ref, see
 *       oracle/Makefile) on the shipped dataset, and
 *   (2) the golden vectors in tests/golden/ generated from (1), and
 *   (3) the legacy 1,100-row trace `Bayes-networks/iterations - null start.xlsx`
 *       (via the legacy-semantics mode).
 * The only third-party arithmetic is R's unif_rand (Mersenne-Twister + set.seed
 * scrambling, R >= 3.1 per DESCRIPTION:25, src/main/RNG.c in the R sources,
 * not under /root/reference): restated from the published algorithm and pinned
 * by the universally known answers set.seed(1234); runif(3) etc.
 */
#ifndef BN_ORACLE_H
#define BN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- uniform generators ------------------------------------------------- */
enum { BNO_RNG_WH = 0, BNO_RNG_RMT = 1, BNO_RNG_REPLAY = 2 };

typedef struct bno_rng {
  int kind;
  /* Wichmann-Hill (Bayes-networks/random4f.h:17-49) */
  int ix, iy, iz;
  /* R Mersenne-Twister (R RNG.c: MT_sgenrand/MT_genrand/fixup) */
  uint32_t mt[624];
  int mti;
  /* replay */
  const double* replay;
  long replay_len;
  /* bookkeeping */
  long draws;
} bno_rng;

void bno_rng_init_wh(bno_rng* r, int ix, int iy, int iz);
void bno_rng_init_rmt(bno_rng* r, uint32_t seed);          /* == set.seed(seed) */
void bno_rng_init_replay(bno_rng* r, const double* u, long n);
double bno_rng_uniform(bno_rng* r);                          /* == R::runif(0,1) / RandomUniform() */
double bno_rng_uniform_cb(void* r);                          /* void* adaptor */
void bno_rng_skip(bno_rng* r, long n);                       /* draw and discard n uniforms */

/* ---- dense linear algebra (src/cholesky22.h) ---------------------------- */
int bno_cholesky_decomp(const double* x, int n, double* c);  /* :25-66, row-major n*n */
int bno_invert_pds(const double* x, int n, double* c);       /* :92-170 / :202-242 */

/* ---- sufficient statistics (src/network.h:124-136) ---------------------- */
void bno_gram(const double* X_colmajor, int N, int P, double* sumX, double* sumXX_colmajor);

/* ---- node score (src/network.h:183-237) --------------------------------- */
/* dim = MaxPar+1 reproduces the reference's identity-padded inversion;
 * dim = npar+1 gives the identical result in O(npar^3) (padding rows only add
 * exact zeros).  Pass pad_dim <= 0 for npar+1.  *err receives InvertPDS's rc. */
double bno_score(const double* X_colmajor, int N, int P,
                 const double* sumX, const double* sumXX_colmajor,
                 int p, const int* parents, int npar, int pad_dim, int* err);

/* ---- the MCMC driver (src/bayesnet_mcmc.cpp:27-72 + src/network.h) ------- */
typedef struct bno_trace {
  int capacity;      /* rows allocated by the caller */
  int n_rows;        /* rows written */
  int* iter;
  int* changed_node;
  int* movetype;
  double* global_ll;
  int* additions;
  int* deletions;
  int* fn;
  int* fp;
  /* legacy-only extras (Bayes-networks/main.cpp:378-381); may be NULL */
  int* npar_changed;
  double* log_prior;
  double* hr;
  int* total_edges;
  int* agree;
} bno_trace;

/* per-iteration move log (optional): kind 0 invalid, 1 add, 2 delete */
typedef struct bno_movelog {
  long capacity;
  long n;
  int* iter;
  signed char* movetype;  /* 0 invalid / 1 add / 2 delete (as proposed) */
  int* child;
  int* parent;
  signed char* valid;
  signed char* accepted;
} bno_movelog;

typedef struct bno_counters {
  long uniforms;
  int proposed[3];
  int reject[3];
  int n_nonpd;
  int total_edges_member;
  int fp_member, fn_member;
} bno_counters;

typedef struct bno_mcmc_args {
  const double* X_colmajor; int N; int P;
  const int* edge_src_1b; const int* edge_tgt_1b; int n_edges;
  const int* node_type;            /* 0 neither, 1 source, 2 sink */
  int max_par; double phi; double omega;
  int initial_network;             /* 0 given graph, 2 empty (1 is UB in the reference) */
  int drop; int n_iter; int output;
  int legacy;                      /* 0: Rcpp semantics; 1: Bayes-networks/main.cpp semantics */
  int pad_dim;                     /* see bno_score */
} bno_mcmc_args;

/* final_parents: [P * max_par] ints, final_npar: [P].  Returns 0 on success. */
int bno_mcmc(const bno_mcmc_args* a, bno_rng* rng, bno_trace* trace,
             bno_movelog* moves, bno_counters* counters,
             int* final_parents, int* final_npar);

/* scores of every node for a given graph (LogLikelihood(1) terms) */
void bno_score_graph(const double* X_colmajor, int N, int P,
                     const int* parents, const int* npar, int max_par,
                     int pad_dim, double* out_scores);

#ifdef __cplusplus
}
#endif
#endif
