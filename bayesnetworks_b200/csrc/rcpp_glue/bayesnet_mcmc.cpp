// Drop-in replacement for the reference's src/bayesnet_mcmc.cpp: the SAME exported
// signature and defaults (src/bayesnet_mcmc.cpp:27-38), the SAME eight result columns in
// the same order (src/network.h:353-364) -- but instead of running the CPU loop of
// src/bayesnet_mcmc.cpp:40-71 it calls the B200 library through the C ABI
// (include/bn_b200.h).  RcppExports.cpp / R/RcppExports.R / R/bn_mcmc.R stay untouched.
//
// R is not installed in the build image, so this file is only compile-checked there, against
// the test suite's stand-in Rcpp.h (tests/test_host_logic.py); under real R it needs
// src/Makevars:  PKG_LIBS = -L<dir> -lbn_b200   (see INTEGRATION.md).
//
// Uniform stream: R::runif is replaced by a device stream.  By default the chain is
// seeded from R's RNG (one unif_rand() draw -> an R-MT seed), so set.seed() keeps runs
// reproducible; define BN_B200_RNG_WH to use the reference program's Wichmann-Hill
// stream (Bayes-networks/random4f.h seeds) instead.
#include <Rcpp.h>
#include <string>
#include <vector>

#include "bn_b200.h"
using namespace Rcpp;

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

#ifdef BN_B200_RNG_WH
  const int rng_kind = BN_RNG_WH;
  const int* seeds = nullptr;  // 10437 / 13568 / 30524
#else
  const int rng_kind = BN_RNG_RMT;
  int seed_store[3] = {(int)(R::runif(0, 1) * 2147483647.0), 0, 0};
  const int* seeds = seed_store;
#endif

  const int rows = bn_main_fun(&X(0, 0), n_samples, n_nodes, graph_source.data(), graph_target.data(),
                               (int)graph_source.size(), graph_node_labels.data(), graph_node_type.data(),
                               MaxPar, phi, omega, InitialNetwork, drop, N, output, rng_kind, seeds, capacity,
                               &iter[0], &changed[0], &movetype[0], &globalLL[0], &additions[0], &deletions[0],
                               &fn[0], &fp[0]);
  if (rows < 0) Rcpp::stop(std::string("bayesnetworks (B200): ") + bn_last_error());

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
