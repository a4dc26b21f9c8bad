// TEST INFRASTRUCTURE ONLY.  Runs the PRODUCT's Rcpp glue (bayesnetworks_b200/csrc/rcpp_glue/
// bayesnet_mcmc.cpp, the drop-in for the reference's src/bayesnet_mcmc.cpp) on a machine without
// R: the file is compiled unmodified against the stand-in Rcpp.h (oracle/ref_shim) and linked
// with libbn_b200.so by the link line of the glue's src/Makevars; this driver plays R -- it
// installs .Random.seed, calls main_fun() with the reference's argument types and hands back
// the DataFrame columns and the .Random.seed the call leaves behind.
#include <Rcpp.h>

Rcpp::ShimErrStream Rcpp::Rcerr;
long bn_shim_rprintf_calls = 0;
long bn_shim_rngstate_calls = 0;
namespace R {
static double no_unif(void*) { return 0.5; }
shim_unif_fn shim_unif = no_unif;
void* shim_unif_state = nullptr;
long shim_unif_draws = 0;
}  // namespace R

#include "bayesnet_mcmc.cpp"  // the glue, as a maintainer would drop it into src/

extern "C" int glue_main_fun(const double* X, int n, int p, const int* src, const int* tgt, int n_edges,
                             const int* labels, const int* types, int MaxPar, double phi, double omega,
                             int InitialNetwork, int drop, int N, int output,
                             const int* random_seed_in /* 626 or NULL */, int* random_seed_out /* 626 */,
                             int capacity, int* iter, int* changed, int* movetype, double* gll, int* add,
                             int* del, int* fn, int* fp, long* r_unif_draws, char* err, int err_len) {
  using namespace Rcpp;
  shim_globals().clear();
  R::shim_unif_draws = 0;
  if (random_seed_in) {
    IntegerVector rs(626);
    for (int i = 0; i < 626; i++) rs[i] = random_seed_in[i];
    shim_globals()[".Random.seed"] = rs;
  }
  NumericMatrix Xm(n, p, X);
  std::vector<int> vs(src, src + n_edges), vt(tgt, tgt + n_edges), vl(labels, labels + p), vn(types, types + p);
  try {
    DataFrame df = main_fun(Xm, vs, vt, vl, vn, MaxPar, phi, omega, InitialNetwork, drop, N, output);
    static const char* names[8] = {"iter", "ChangedNode", "movetype", "globalLL", "additions", "deletions", "FN", "FP"};
    if (df.columns.size() != 8) return -100;
    for (int c = 0; c < 8; c++)
      if (df.columns[c].name != names[c]) return -101;  // the reference's column order
    const int rows = (int)df.columns[0].values.size();
    if (rows > capacity) return -102;
    int* ints[8] = {iter, changed, movetype, nullptr, add, del, fn, fp};
    for (int c = 0; c < 8; c++)
      for (int r = 0; r < rows; r++) {
        if (c == 3) gll[r] = df.columns[c].values[r];
        else ints[c][r] = (int)df.columns[c].values[r];
      }
    if (random_seed_out && shim_globals().count(".Random.seed")) {
      const IntegerVector& rs = shim_globals()[".Random.seed"];
      for (int i = 0; i < 626 && i < rs.size(); i++) random_seed_out[i] = rs[i];
    }
    if (r_unif_draws) *r_unif_draws = R::shim_unif_draws;
    return rows;
  } catch (const std::exception& e) {
    if (err && err_len > 0) { strncpy(err, e.what(), (size_t)err_len - 1); err[err_len - 1] = 0; }
    return -1;
  }
}
