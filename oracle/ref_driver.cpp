// TEST INFRASTRUCTURE ONLY -- never linked into libbn_b200.so.
//
// Driver that compiles the reference's OWN sources where they lie
// (/root/reference/src/{bayesnet_mcmc.cpp,network.h,cholesky22.h}, found via
// -I, not copied) against oracle/ref_shim/Rcpp.h and exposes them through a
// small C surface for ctypes:
//   ref_main_fun   -> the unmodified main_fun() (src/bayesnet_mcmc.cpp:27-72)
//   ref_scores     -> network::score(p) for every node of a given graph
//   ref_gram       -> the constructor's sumX / sumXX (src/network.h:124-136)
//   ref_invert_pds -> InvertPDS (src/cholesky22.h:202-242)
// The uniform source behind R::runif is selectable (bn_oracle.c generators).
// Output: oracle/_ref/libbnref.so (git-ignored).
#define private public   // tests need score(), sumX, sumXX, counters
#include "bayesnet_mcmc.cpp"
#undef private

#include "bn_oracle.h"

namespace Rcpp { ShimErrStream Rcerr; }
long bn_shim_rprintf_calls = 0;
namespace R {
shim_unif_fn shim_unif = nullptr;
void* shim_unif_state = nullptr;
long shim_unif_draws = 0;
}

static std::vector<int> to_vec(const int* p, int n) { return std::vector<int>(p, p + n); }

static void install_rng(bno_rng* rng) {
  R::shim_unif = bno_rng_uniform_cb;
  R::shim_unif_state = rng;
  R::shim_unif_draws = 0;
}

extern "C" {

// rng_kind: BNO_RNG_WH (seeds[0..2]), BNO_RNG_RMT (seeds[0] = set.seed value),
// BNO_RNG_REPLAY (replay buffer).  Trace arrays have `capacity` rows.
// Returns the number of rows, or -1 on exception.
int ref_main_fun(const double* X, int N, int P, const int* src_1b, const int* tgt_1b,
                 int n_edges, const int* node_type, int MaxPar, double phi, double omega,
                 int InitialNetwork, int drop, int n_iter, int output, int rng_kind,
                 const int* seeds, const double* replay, long replay_len, int capacity,
                 int* iter, int* changed, int* movetype, double* global_ll, int* additions,
                 int* deletions, int* fn, int* fp, long* uniforms_drawn, long* diag_messages) {
  bno_rng rng;
  if (rng_kind == BNO_RNG_WH) bno_rng_init_wh(&rng, seeds[0], seeds[1], seeds[2]);
  else if (rng_kind == BNO_RNG_RMT) bno_rng_init_rmt(&rng, (uint32_t)seeds[0]);
  else bno_rng_init_replay(&rng, replay, replay_len);
  install_rng(&rng);
  Rcpp::Rcerr.n_messages = 0;
  bn_shim_rprintf_calls = 0;
  try {
    Rcpp::NumericMatrix Xm(N, P, X);
    std::vector<int> labels(P);
    for (int i = 0; i < P; i++) labels[i] = i;
    Rcpp::DataFrame df = main_fun(Xm, to_vec(src_1b, n_edges), to_vec(tgt_1b, n_edges), labels,
                                  to_vec(node_type, P), MaxPar, phi, omega, InitialNetwork, drop,
                                  n_iter, output);
    const char* names[8] = {"iter", "ChangedNode", "movetype", "globalLL",
                            "additions", "deletions", "FN", "FP"};
    // the column order is part of the contract (src/network.h:353-364)
    for (int c = 0; c < 8; c++)
      if (df.columns.size() != 8 || df.columns[c].name != names[c]) return -2;
    int rows = (int)df.columns[0].values.size();
    int n = rows < capacity ? rows : capacity;
    for (int r = 0; r < n; r++) {
      iter[r] = (int)df.columns[0].values[r];
      changed[r] = (int)df.columns[1].values[r];
      movetype[r] = (int)df.columns[2].values[r];
      global_ll[r] = df.columns[3].values[r];
      additions[r] = (int)df.columns[4].values[r];
      deletions[r] = (int)df.columns[5].values[r];
      fn[r] = (int)df.columns[6].values[r];
      fp[r] = (int)df.columns[7].values[r];
    }
    if (uniforms_drawn) *uniforms_drawn = R::shim_unif_draws;
    if (diag_messages) *diag_messages = Rcpp::Rcerr.n_messages + bn_shim_rprintf_calls;
    return rows;
  } catch (...) {
    return -1;
  }
}

// score(p) for every node p of the graph given by the edge list (InitialNetwork=0).
int ref_scores(const double* X, int N, int P, const int* src_1b, const int* tgt_1b, int n_edges,
               const int* node_type, int MaxPar, double* out_scores, double* out_global_ll,
               double* out_log_prior) {
  bno_rng rng;
  bno_rng_init_wh(&rng, 10437, 13568, 30524);
  install_rng(&rng);
  try {
    Rcpp::NumericMatrix Xm(N, P, X);
    network net(Xm, 0, MaxPar, 1.0, 6.9, to_vec(src_1b, n_edges), to_vec(tgt_1b, n_edges),
                to_vec(node_type, P));
    for (int p = 0; p < P; p++) out_scores[p] = net.score(p);
    if (out_global_ll) *out_global_ll = net.LogLikelihood(1);
    if (out_log_prior) *out_log_prior = net.LogPrior();
    return 0;
  } catch (...) {
    return -1;
  }
}

int ref_gram(const double* X, int N, int P, double* sumX, double* sumXX_colmajor) {
  try {
    Rcpp::NumericMatrix Xm(N, P, X);
    std::vector<int> none, types(P, 0);
    network net(Xm, 2, 1, 1.0, 6.9, none, none, types);
    for (int p = 0; p < P; p++) sumX[p] = net.sumX[p];
    for (int a = 0; a < P; a++)
      for (int b = 0; b < P; b++) sumXX_colmajor[(size_t)a + (size_t)b * P] = net.sumXX(a, b);
    return 0;
  } catch (...) {
    return -1;
  }
}

int ref_invert_pds(double* x_rowmajor, int n, double* c_rowmajor) {
  return InvertPDS(x_rowmajor, n, c_rowmajor);
}

}  // extern "C"
