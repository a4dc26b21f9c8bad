// TEST INFRASTRUCTURE ONLY -- never linked into libbn_b200.so.
//
// Stand-in for <Rcpp.h> so that the reference's own sources
//   /root/reference/src/bayesnet_mcmc.cpp, src/network.h, src/cholesky22.h
// compile UNMODIFIED with plain g++ (R and Rcpp are not installed in this
// image).  It provides exactly the surface those three files touch
// (SURVEY.md Appendix C): value-semantics Vector / Matrix containers with
// column-major operator(), clone(), DataFrame::create(Named(..) = ..),
// Rcerr, Rprintf and a pluggable R::runif.
//
// This header is original code written for this repo; it contains no
// reference source.
#ifndef BN_B200_ORACLE_RCPP_SHIM_H
#define BN_B200_ORACLE_RCPP_SHIM_H

#include <vector>
#include <list>
#include <string>
#include <cstring>
#include <cmath>
#include <cstdio>
#include <cstdarg>
#include <iostream>
#include <initializer_list>
#include <utility>
#include <stdexcept>
#include <map>

namespace Rcpp {

template <typename T>
class ShimVector {
  std::vector<T> v_;
 public:
  ShimVector() {}
  // size constructors (zero filled, like Rcpp)
  ShimVector(int n) : v_(n < 0 ? 0 : (size_t)n, T(0)) {}
  ShimVector(unsigned int n) : v_((size_t)n, T(0)) {}
  ShimVector(long n) : v_(n < 0 ? 0 : (size_t)n, T(0)) {}
  ShimVector(unsigned long n) : v_((size_t)n, T(0)) {}
  ShimVector(std::initializer_list<T> il) : v_(il) {}
  T& operator[](long i) { return v_[(size_t)i]; }
  const T& operator[](long i) const { return v_[(size_t)i]; }
  void fill(T x) { for (auto& e : v_) e = x; }
  void push_back(T x) { v_.push_back(x); }
  long size() const { return (long)v_.size(); }
  const std::vector<T>& std_vector() const { return v_; }
};

template <typename T>
class ShimMatrix {
  int nr_ = 0, nc_ = 0;
  std::vector<T> v_;
 public:
  ShimMatrix() {}
  ShimMatrix(int nr, int nc) : nr_(nr), nc_(nc), v_((size_t)nr * (size_t)nc, T(0)) {}
  // borrow-by-copy from a column-major buffer (used by the test driver only)
  ShimMatrix(int nr, int nc, const T* colmajor)
      : nr_(nr), nc_(nc), v_(colmajor, colmajor + (size_t)nr * (size_t)nc) {}
  int nrow() const { return nr_; }
  int ncol() const { return nc_; }
  T& operator()(int i, int j) { return v_[(size_t)i + (size_t)j * (size_t)nr_]; }
  const T& operator()(int i, int j) const { return v_[(size_t)i + (size_t)j * (size_t)nr_]; }
};

typedef ShimVector<double> NumericVector;
typedef ShimVector<int> IntegerVector;
typedef ShimMatrix<double> NumericMatrix;
typedef ShimMatrix<int> IntegerMatrix;

template <typename C>
inline C clone(const C& x) { return x; }

// DataFrame::create(Named("a") = vec, ...): columns are stored as doubles,
// the driver converts back (all int columns are exactly representable).
struct ShimColumn {
  std::string name;
  std::vector<double> values;
  bool is_int = false;
};

class Named {
  std::string name_;
 public:
  explicit Named(const char* n) : name_(n) {}
  ShimColumn operator=(const IntegerVector& v) const {
    ShimColumn c; c.name = name_; c.is_int = true;
    for (int x : v.std_vector()) c.values.push_back((double)x);
    return c;
  }
  ShimColumn operator=(const NumericVector& v) const {
    ShimColumn c; c.name = name_; c.values = v.std_vector();
    return c;
  }
};

class DataFrame {
 public:
  std::vector<ShimColumn> columns;
  template <typename... Cols>
  static DataFrame create(Cols&&... cols) {
    DataFrame df;
    (df.columns.push_back(std::forward<Cols>(cols)), ...);
    return df;
  }
  const ShimColumn* find(const char* name) const {
    for (auto& c : columns) if (c.name == name) return &c;
    return nullptr;
  }
};

// diagnostics go to a sink by default (the reference prints on non-PD
// matrices and on >100 proposal tries); the driver counts them.
struct ShimErrStream {
  long n_messages = 0;
  template <typename T> ShimErrStream& operator<<(const T&) { n_messages++; return *this; }
};
extern ShimErrStream Rcerr;

// R's global environment, as far as the B200 glue (bayesnetworks_b200/csrc/rcpp_glue) touches it:
// integer vectors by name (.Random.seed).  The reference's own sources do not use this.
inline std::map<std::string, IntegerVector>& shim_globals() {
  static std::map<std::string, IntegerVector> g;
  return g;
}
class Environment {
 public:
  static Environment global_env() { return Environment(); }
  bool exists(const std::string& name) const { return shim_globals().count(name) > 0; }
  IntegerVector operator[](const std::string& name) const { return shim_globals()[name]; }
  void assign(const std::string& name, const IntegerVector& v) const { shim_globals()[name] = v; }
};

// Rcpp::stop -> C++ exception (BEGIN_RCPP/END_RCPP turn it into an R error)
inline void stop(const std::string& msg) { throw std::runtime_error(msg); }

}  // namespace Rcpp

extern long bn_shim_rprintf_calls;
inline void Rprintf(const char*, ...) { bn_shim_rprintf_calls++; }

// R_ext/Random.h: the stand-in keeps R's generator state in the .Random.seed variable only, so
// loading / storing it are counted no-ops.
extern long bn_shim_rngstate_calls;
inline void GetRNGstate() { bn_shim_rngstate_calls++; }
inline void PutRNGstate() { bn_shim_rngstate_calls++; }

// Pluggable uniform source: the driver installs the generator.
namespace R {
typedef double (*shim_unif_fn)(void*);
extern shim_unif_fn shim_unif;
extern void* shim_unif_state;
extern long shim_unif_draws;
inline double runif(double a, double b) {
  // R's runif(a,b): a + (b-a)*unif_rand(), redrawn while outside (0,1).
  if (a == b) return a;
  double u;
  do { u = shim_unif(shim_unif_state); shim_unif_draws++; } while (u <= 0 || u >= 1);
  return a + (b - a) * u;
}
}  // namespace R

#endif
