// Gaussian local score of one node given an ordered parent list, from the
// centred Gram matrix.  Replaces network::score (src/network.h:183-237) and
// the InvertPDS call it makes (src/cholesky22.h:92-170):
//
//   reference:  beta = (X_S'X_S)^-1 X_S'y via a (MaxPar+1)-dim inverse, then a
//               pass over all N samples for the residual sum of squares
//   here:       RSS = C_cc - b' A^-1 b  with A = C[S,S], b = C[S,c] from the
//               centred cross-product matrix C (the intercept row/column of
//               the reference's SXX is absorbed by the centring), through a
//               k-dim Cholesky factor and one triangular solve -- no pass
//               over the data, O(k^3/6) instead of O(51^3 + N k).
//
//   score = -(N/2) * log( [RSS/(N-k-1)] / [C_cc/(N-1)] )      (src/network.h:232-236)
#pragma once

#include "bn_common.cuh"

namespace bn {

// L: k*(k+1)/2 scratch (row-packed lower triangle), z: k scratch.
// Returns -inf and sets *nonpd when a pivot is <= 0 (the reference only
// prints in that case, src/network.h:213-215; see DESIGN.md "deviations").
BN_HD double score_set(const double* __restrict__ C, int64_t ldc, int c,
                       const int* S, int k, int n_samples,
                       double* L, double* z, int* nonpd) {
  const double Ccc = ld_shared_ro(C + (int64_t)c * ldc + c);
  // gather first (independent loads pipeline), factor in place afterwards
  for (int i = 0; i < k; i++) {
    const double* row = C + (int64_t)S[i] * ldc;
    double* Li = L + (i * (i + 1)) / 2;
    for (int m = 0; m <= i; m++) Li[m] = ld_shared_ro(row + S[m]);
    z[i] = ld_shared_ro(row + c);
  }
  double rss = Ccc;
  for (int i = 0; i < k; i++) {
    double* Li = L + (i * (i + 1)) / 2;
    for (int m = 0; m < i; m++) {
      const double* Lm = L + (m * (m + 1)) / 2;
      double acc = Li[m];
      for (int t = 0; t < m; t++) acc -= Li[t] * Lm[t];
      Li[m] = acc / Lm[m];
    }
    double d = Li[i];
    for (int t = 0; t < i; t++) d -= Li[t] * Li[t];
    if (!(d > 0.0)) {
      if (nonpd) *nonpd = 1;
      return -INFINITY;
    }
    d = sqrt(d);
    Li[i] = d;
    double acc = z[i];
    for (int t = 0; t < i; t++) acc -= Li[t] * z[t];
    acc = acc / d;
    z[i] = acc;
    rss -= acc * acc;
  }
  const double resid2 = rss / (double)(n_samples - k - 1);
  const double syy = Ccc / (double)(n_samples - 1);
  return -((double)n_samples / 2.0) * log(resid2 / syy);
}

BN_HD double rsqrt_f64(double d) {
#if defined(__CUDA_ARCH__)
  return rsqrt(d);
#else
  return 1.0 / sqrt(d);
#endif
}

// Register-resident variant for small parent limits (K <= 8): every index is a
// compile-time constant after unrolling, so the factor lives in registers, the gathers
// issue back to back (one L2 round trip instead of one per row) and the divisions
// become multiplications by the reciprocal pivot.
template <int K>
BN_HD double score_set_small(const double* __restrict__ C, int64_t ldc, int c, const int (&S)[K], int k,
                             int n_samples, int* nonpd) {
  double L[K * (K + 1) / 2], z[K], rinv[K];
  const double Ccc = ld_shared_ro(C + (int64_t)c * ldc + c);
#pragma unroll
  for (int i = 0; i < K; i++) {
    const bool on = i < k;
    const double* row = C + (int64_t)(on ? S[i] : c) * ldc;
#pragma unroll
    for (int m = 0; m <= i; m++) L[i * (i + 1) / 2 + m] = on ? ld_shared_ro(row + S[m]) : 0.0;
    z[i] = on ? ld_shared_ro(row + c) : 0.0;
  }
  double rss = Ccc;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < K; i++) {
    if (i < k) {
#pragma unroll
      for (int m = 0; m < i; m++) {
        double acc = L[i * (i + 1) / 2 + m];
#pragma unroll
        for (int t = 0; t < m; t++) acc -= L[i * (i + 1) / 2 + t] * L[m * (m + 1) / 2 + t];
        L[i * (i + 1) / 2 + m] = acc * rinv[m];
      }
      double d = L[i * (i + 1) / 2 + i];
#pragma unroll
      for (int t = 0; t < i; t++) d -= L[i * (i + 1) / 2 + t] * L[i * (i + 1) / 2 + t];
      if (!(d > 0.0)) { bad = true; d = 1.0; }
      const double r = rsqrt_f64(d);
      rinv[i] = r;
      double acc = z[i];
#pragma unroll
      for (int t = 0; t < i; t++) acc -= L[i * (i + 1) / 2 + t] * z[t];
      acc *= r;
      z[i] = acc;
      rss -= acc * acc;
    }
  }
  if (bad) {
    if (nonpd) *nonpd = 1;
    return -INFINITY;
  }
  const double resid2 = rss / (double)(n_samples - k - 1);
  const double syy = Ccc / (double)(n_samples - 1);
  return -((double)n_samples / 2.0) * log(resid2 / syy);
}

// One shared copy of the K = 8 scorer on the device: the fully unrolled gather + Cholesky is
// ~2,500 instructions, and the chain kernel reaches it from four places (chain start, the
// sequential window path, the records of the chain's warp and of its helper warps).  Inlined
// four times it made up 40 % of a 300 KB kernel that misses the instruction cache on every
// rarely taken path; as a call it is one copy.  All arguments travel in registers.
struct Parents8 { int s[8]; };
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
inline
#endif
double score_set8(const double* C, int64_t ldc, int c, Parents8 S, int k, int n_samples) {
  int npd = 0;
  return score_set_small<8>(C, ldc, c, S.s, k, n_samples, &npd);  // -inf <=> not positive definite
}

// Parent sets of more than 8 nodes (possible when MaxPar > 8): the factor does not fit in
// registers.  A function of its own so that its local arrays (16 KB at KMAX = 64) are a frame
// that exists during the call only, not part of every thread's frame in the chain kernel.
// type 1: parents pc[0..k) plus j; type 2: pc[0..k) without slot del; type 0: pc[0..k) as is.
template <int KMAX>
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
inline
#endif
double score_set_big(const double* C, int64_t ldc, int c, const int* pc, int k, int type, int j, int del,
                     int n_samples, int* kk_out, int* nonpd) {
  double L[KMAX * (KMAX + 1) / 2], z[KMAX];
  int S[KMAX];
  int kk = 0;
  for (int e = 0; e < k; e++)
    if (type != 2 || e != del) S[kk++] = pc[e];
  if (type == 1) S[kk++] = j;
  *kk_out = kk;
  return score_set(C, ldc, c, S, kk, n_samples, L, z, nonpd);
}

}  // namespace bn
