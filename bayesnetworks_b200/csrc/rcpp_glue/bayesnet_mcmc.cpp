// Drop-in replacement for the reference's src/bayesnet_mcmc.cpp: the SAME exported
// signature and defaults (src/bayesnet_mcmc.cpp:27-38), the SAME eight result columns in
// the same order (src/network.h:353-364) -- but instead of running the CPU loop of
// src/bayesnet_mcmc.cpp:40-71 it calls the B200 library through the C ABI
// (include/bn_b200.h).  RcppExports.cpp / R/RcppExports.R / R/bn_mcmc.R stay untouched;
// src/Makevars (next to this file) is added, src/network.h and src/cholesky22.h go.
//
// R is not installed in the build image: this file is compiled there against the test suite's
// stand-in Rcpp.h and RUN on the GPU box through tests/tools/glue_driver.cpp
// (tests/test_rcpp_glue.py: set.seed(1234) state in -> the README trace, .Random.seed out).
//
// Uniform stream.  The reference draws every uniform from R's global generator through
// R::runif, bracketed by the RNGScope of src/RcppExports.cpp:13 (GetRNGstate on entry,
// PutRNGstate on exit).  The device cannot call back into R, so R's generator itself runs on
// the device (BN_RNG_RMT: MT19937, R's scaling and open-interval fix-up) and its STATE crosses
// the boundary: .Random.seed[2:626] goes in, the state after exactly the uniforms the chain
// consumed comes back and is installed as R's state.  `set.seed(s); bn_mcmc(...)` therefore
// gives the reference's trajectory, and the next runif() in the session returns what it would
// have returned after the reference's run.  The mapping is one to one because main_fun is the
// only consumer of R's stream between entry and exit (the reference draws nothing else either).
// If R's generator is not the Mersenne-Twister (RNGkind() changed), the chain is seeded from one
// unif_rand() draw instead -- reproducible under set.seed, but not the CPU package's trajectory.
// Define BN_B200_RNG_WH to use the legacy program's Wichmann-Hill stream
// (Bayes-networks/random4f.h seeds) instead.
#include <Rcpp.h>
#include <string>
#include <vector>

#include "bn_b200.h"
using namespace Rcpp;

namespace {
// .Random.seed = c(kind code, position, 624 state words) for the Mersenne-Twister (kind %% 100 == 3)
bool fetch_r_mt_state(int* st, int* kind_code) {
  PutRNGstate();  // the scope's GetRNGstate() loaded the state; make the variable current
  Environment g = Environment::global_env();
  if (!g.exists(".Random.seed")) return false;
  IntegerVector rs = g[".Random.seed"];
  if (rs.size() != 626 || rs[0] % 100 != 3) return false;
  *kind_code = rs[0];
  for (int i = 0; i < 625; i++) st[i] = rs[1 + i];
  return true;
}

void store_r_mt_state(const int* st, int kind_code) {
  IntegerVector rs(626);
  rs[0] = kind_code;
  for (int i = 0; i < 625; i++) rs[1 + i] = st[i];
  Environment::global_env().assign(".Random.seed", rs);
  GetRNGstate();  // reload the generator: the scope's PutRNGstate() on exit writes this state back
}
}  // namespace

// [[Rcpp::export]]
DataFrame main_fun(NumericMatrix X,
                   std::vector<int> graph_source,
                   std::vector<int> graph_target,
                   std::vector<int> graph_node_labels,
                   std::vector<int> graph_node_type,
                   int MaxPar = 50,
                   const double phi = 1,
                   const double omega = 6.9,
                   const int InitialNetwork = 2,
                   const int drop = 0,
                   int N = 1000,
                   int output = 10) {
  const int n_samples = X.nrow(), n_nodes = X.ncol();
  const int capacity = output > 0 ? (N + output - 1) / output + 1 : 1;
  IntegerVector iter(capacity), changed(capacity), movetype(capacity), additions(capacity),
      deletions(capacity), fn(capacity), fp(capacity);
  NumericVector globalLL(capacity);

  int rng_kind = BN_RNG_RMT;
  int seed_store[3] = {0, 0, 0};
  const int* seeds = seed_store;
  std::vector<int> mt_in(625), mt_out(625);
  const int *state_in = nullptr;
  int *state_out = nullptr, kind_code = 0;
#ifdef BN_B200_RNG_WH
  rng_kind = BN_RNG_WH;
  seeds = nullptr;  // 10437 / 13568 / 30524
#else
  if (fetch_r_mt_state(mt_in.data(), &kind_code)) {
    state_in = mt_in.data();
    state_out = mt_out.data();
  } else {
    seed_store[0] = (int)(R::runif(0, 1) * 2147483647.0);
  }
#endif

  const int rows = bn_main_fun(&X(0, 0), n_samples, n_nodes, graph_source.data(), graph_target.data(),
                               (int)graph_source.size(), graph_node_labels.data(), graph_node_type.data(),
                               MaxPar, phi, omega, InitialNetwork, drop, N, output, rng_kind, seeds, capacity,
                               &iter[0], &changed[0], &movetype[0], &globalLL[0], &additions[0], &deletions[0],
                               &fn[0], &fp[0], state_in, state_out);
  if (rows < 0) Rcpp::stop(std::string("bayesnetworks (B200): ") + bn_last_error());
  if (state_out) store_r_mt_state(state_out, kind_code);

  IntegerVector o_iter(rows), o_changed(rows), o_movetype(rows), o_add(rows), o_del(rows), o_fn(rows), o_fp(rows);
  NumericVector o_gll(rows);
  for (int r = 0; r < rows; r++) {
    o_iter[r] = iter[r]; o_changed[r] = changed[r]; o_movetype[r] = movetype[r]; o_gll[r] = globalLL[r];
    o_add[r] = additions[r]; o_del[r] = deletions[r]; o_fn[r] = fn[r]; o_fp[r] = fp[r];
  }
  return DataFrame::create(
    Named("iter")        = o_iter,
    Named("ChangedNode") = o_changed,
    Named("movetype")    = o_movetype,
    Named("globalLL")    = o_gll,
    Named("additions")   = o_add,
    Named("deletions")   = o_del,
    Named("FN")          = o_fn,
    Named("FP")          = o_fp
  );
}
