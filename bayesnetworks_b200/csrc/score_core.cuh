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
//               over the data.
//
//   score = -(N/2) * log( [RSS/(N-k-1)] / [C_cc/(N-1)] )      (src/network.h:232-236)
//
// Three forms: score_set (generic, scratch arrays: score_nodes / fallbacks), score_set8
// (MaxPar <= 8 chain kernels: the factor of the proposed set in registers) and the per-node
// factor cache of the MaxPar > 8 chain kernels (O(k^2) per proposal, second half of this file).
#pragma once

#include "bn_common.cuh"

namespace bn {

// A residual sum of squares at or below this fraction of C_cc is rounding noise of the Gram route
// (the parent set reproduces the child exactly): the set is treated like a parent Gram that is not
// positive definite -- score -inf, counted in n_nonpd -- instead of letting a zero or negative RSS
// put a NaN into the node's score (which would accept every later proposal at that node,
// `runif > NaN` being false).  The reference's data pass gives a tiny positive RSS there and accepts.
constexpr double RSS_FLOOR = 1e-13;

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
  if (!(rss > Ccc * RSS_FLOOR)) {
    if (nonpd) *nonpd = 1;
    return -INFINITY;
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

#if defined(__CUDACC__)
#define BN_NOINLINE static __host__ __device__ __noinline__
#else
#define BN_NOINLINE inline
#endif

struct Parents8 { int s[8]; };

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
  if (bad || !(rss > Ccc * RSS_FLOOR)) {
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
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
inline
#endif
double score_set8(const double* C, int64_t ldc, int c, Parents8 S, int k, int n_samples) {
  int npd = 0;
  return score_set_small<8>(C, ldc, c, S.s, k, n_samples, &npd);  // -inf <=> not positive definite
}

// ---------------------------------------------------------------------------
// Per-node factor cache.
//
// The chain scores thousands of single-edge changes against the SAME parent set before one is
// accepted, so each node keeps the Cholesky factor of its current (ordered) parent set:
//   A = C[S,S] = L L',  z = L^-1 C[S,c],  rss = C_cc - z'z
// and a proposal is O(k^2) arithmetic on it instead of an O(k^3) factorisation plus a
// (k+1)(k+2)/2-entry gather:
//   addition of j   one more row of the factor (the same arithmetic a fresh factorisation of
//                   S + [j] in push_back order, src/network.h:303, performs for its last row):
//                   w = L^-1 C[S,j], d = C_jj - w'w, e = C_jc - w'z, rss' = rss - e^2/d
//   deletion of S_e rss' = rss + (y'z)^2 / (y'y) with y = column e of L^-1 (the classical
//                   drop-one-regressor identity, beta_e^2 / (A^-1)_ee)
// An accepted addition appends the candidate row; an accepted deletion re-factorises the node.
// Replaces the (MaxPar+1)-dim InvertPDS + pass over all samples of network::score
// (src/network.h:183-237, src/cholesky22.h:92-170) for ANY MaxPar <= 64.
//
//   score = -(N/2) log( [rss/(N-k-1)] / [C_cc/(N-1)] ) = -(N/2) log( rss * (1/C_cc) * ratio[k] )
// with ratio[k] = (N-1)/(N-k-1) tabulated per run (ScoreConsts).
//
// Layout of one node's block (fac_stride(mp) doubles, mp = fac_mp(MaxPar); rows padded to an even
// length so that every row starts 16-byte aligned):
//   fac_row(i) ..      L[i][0..i-1], then 1 / L[i][i]
//   fac_zoff(mp) ..    z[0..mp-1]
//   fac_tail(mp)       rss (NaN = the current parent Gram is not positive definite), 1 / C_cc
// Candidate row of an addition record (row_stride(mp) doubles): w[0..k-1], then at row_tail(mp):
// d, e, rss', unused.
// ---------------------------------------------------------------------------
BN_HD constexpr int fac_mp(int max_par) { return max_par < 8 ? 8 : max_par; }  // the layout's MaxPar (>= 8)
BN_HD constexpr int fac_row(int i) { return 2 * (i >> 1) * ((i >> 1) + 1) + (i & 1) * (2 * (i >> 1) + 2); }
BN_HD constexpr int fac_zoff(int mp) { return fac_row(mp); }
BN_HD constexpr int fac_tail(int mp) { return fac_row(mp) + ((mp + 1) & ~1); }
BN_HD constexpr int fac_stride(int mp) { return (fac_tail(mp) + 2 + 3) & ~3; }
BN_HD constexpr int row_tail(int mp) { return (mp + 1) & ~1; }
BN_HD constexpr int row_stride(int mp) { return row_tail(mp) + 4; }

struct ScoreConsts {
  double half_n;        // N / 2.0 (src/network.h:235, int N)
  const double* ratio;  // [max_par + 2]: (N - 1) / (N - k - 1)
};
BN_HD double score_from_rss(double rss, double icc, int kk, const ScoreConsts& sc) {
  return -sc.half_n * log(rss * icc * sc.ratio[kk]);
}

struct alignas(16) D2 { double x, y; };
BN_HD D2 ld2_l2(const double* p) {  // 16-byte load that bypasses L1 (the block is written by other warps)
#if defined(__CUDA_ARCH__)
  const double2 v = __ldcg((const double2*)p);
  D2 r; r.x = v.x; r.y = v.y; return r;
#else
  D2 r; r.x = p[0]; r.y = p[1]; return r;
#endif
}
BN_HD double ld1_l2(const double* p) { return ld_shared_ro(p); }

// Factorise node c's block for the ordered parent list S[0..k).  One lane; the block lives in
// global memory and is read back by the same thread (any MaxPar).  Returns the node's score, or
// -inf (and a NaN rss mark) when a pivot is <= 0.
BN_NOINLINE double factor_node(const double* C, int64_t ldc, int c, const int* S, int k, ScoreConsts sc,
                               double* F, int mp) {
  const double Ccc = ld_shared_ro(C + (int64_t)c * ldc + c);
  const double icc = 1.0 / Ccc;
  double* z = F + fac_zoff(mp);
  double rss = Ccc;
  for (int i = 0; i < k; i++) {
    const double* row = C + (int64_t)S[i] * ldc;
    double* Li = F + fac_row(i);
    for (int m = 0; m < i; m++) Li[m] = ld_shared_ro(row + S[m]);   // gather first: the loads pipeline
    double d = ld_shared_ro(row + S[i]);
    double acc = ld_shared_ro(row + c);
    for (int m = 0; m < i; m++) {
      const double* Lm = F + fac_row(m);
      double a = Li[m];
      for (int t = 0; t < m; t++) a -= Li[t] * Lm[t];
      a *= Lm[m];  // 1 / L[m][m]
      Li[m] = a;
      d -= a * a;
      acc -= a * z[m];
    }
    if (!(d > 0.0)) {
      F[fac_tail(mp)] = NAN; F[fac_tail(mp) + 1] = icc;
      return -INFINITY;
    }
    const double r = rsqrt_f64(d);
    Li[i] = r;
    acc *= r;
    z[i] = acc;
    rss -= acc * acc;
  }
  if (!(rss > Ccc * RSS_FLOOR)) {
    F[fac_tail(mp)] = NAN; F[fac_tail(mp) + 1] = icc;
    return -INFINITY;
  }
  F[fac_tail(mp)] = rss; F[fac_tail(mp) + 1] = icc;
  return score_from_rss(rss, icc, k, sc);
}

// The same for MaxPar <= 8 with the factor in registers (every index a compile-time constant),
// written out as 16-byte pairs at the end.
BN_NOINLINE double factor_node8(const double* C, int64_t ldc, int c, Parents8 S, int k, ScoreConsts sc, double* F,
                                int mp) {
  constexpr int K = 8;
  double L[fac_row(K)], z[K];
  const double Ccc = ld_shared_ro(C + (int64_t)c * ldc + c);
#pragma unroll
  for (int q = 0; q < fac_row(K); q++) L[q] = 0.0;
#pragma unroll
  for (int i = 0; i < K; i++) {
    const bool on = i < k;
    const double* row = C + (int64_t)(on ? S.s[i] : c) * ldc;
#pragma unroll
    for (int m = 0; m <= i; m++) L[fac_row(i) + m] = on ? ld_shared_ro(row + S.s[m]) : 0.0;
    z[i] = on ? ld_shared_ro(row + c) : 0.0;
  }
  double rss = Ccc;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < K; i++) {
    if (i < k) {
      double d = L[fac_row(i) + i], acc = z[i];
#pragma unroll
      for (int m = 0; m < i; m++) {
        double a = L[fac_row(i) + m];
#pragma unroll
        for (int t = 0; t < m; t++) a -= L[fac_row(i) + t] * L[fac_row(m) + t];
        a *= L[fac_row(m) + m];
        L[fac_row(i) + m] = a;
        d -= a * a;
        acc -= a * z[m];
      }
      if (!(d > 0.0)) { bad = true; d = 1.0; }
      const double r = rsqrt_f64(d);
      L[fac_row(i) + i] = r;
      acc *= r;
      z[i] = acc;
      rss -= acc * acc;
    }
  }
  const double icc = 1.0 / Ccc;
  if (!(rss > Ccc * RSS_FLOOR)) bad = true;
  const int nrow = fac_row(k);
#pragma unroll
  for (int q = 0; q < fac_row(K) / 2; q++)
    if (2 * q < nrow) { D2 v; v.x = L[2 * q]; v.y = L[2 * q + 1]; *(D2*)(F + 2 * q) = v; }
#pragma unroll
  for (int q = 0; q < K / 2; q++)
    if (2 * q < k) { D2 v; v.x = z[2 * q]; v.y = z[2 * q + 1]; *(D2*)(F + fac_zoff(mp) + 2 * q) = v; }
  D2 t; t.x = bad ? NAN : rss; t.y = icc;
  *(D2*)(F + fac_tail(mp)) = t;
  return bad ? -INFINITY : score_from_rss(rss, icc, k, sc);
}

// Accepted deletion of the parent in slot e: the factor of the reduced (order-preserving) list by
// Givens rotations instead of a fresh O(k^3) factorisation.  Dropping row e of L leaves a
// (k-1) x k matrix with one super-diagonal from row e on; rotating the column pairs (t, t+1),
// t = e..k-2, makes it lower triangular again (the last column becomes zero).  z takes the same
// rotations; its last component leaves the regression: rss' = rss + zeta^2.  Rows below e are
// untouched.  One lane, in place in the node's block; O((k - e)^2).  Returns the new score.
// The block must hold a valid factor (rss not NaN).
template <int KMAX>
BN_NOINLINE double factor_downdate(double* F, int k, int e, ScoreConsts sc, int mp) {
  double cs[KMAX], sn[KMAX], x[KMAX + 1];
  double* z = F + fac_zoff(mp);
  for (int i = e; i + 1 < k; i++) {  // new row i = old row i + 1
    const int r = i + 1;
    const double* Lr = F + fac_row(r);
    for (int t = 0; t < r; t++) x[t] = Lr[t];
    x[r] = 1.0 / Lr[r];  // the diagonal slot holds the reciprocal pivot
    for (int t = e; t < i; t++) {  // the rotations defined by the rows above
      const double a = x[t], b = x[t + 1];
      x[t] = cs[t] * a + sn[t] * b;
      x[t + 1] = cs[t] * b - sn[t] * a;
    }
    const double a = x[i], b = x[i + 1];  // b: the old pivot, > 0
    const double rinv = rsqrt_f64(a * a + b * b);
    cs[i] = a * rinv; sn[i] = b * rinv;
    double* Li = F + fac_row(i);
    for (int t = 0; t < i; t++) Li[t] = x[t];
    Li[i] = rinv;
  }
  double zt = z[e];
  for (int t = e; t + 1 < k; t++) {
    const double a = zt, b = z[t + 1];
    z[t] = cs[t] * a + sn[t] * b;
    zt = cs[t] * b - sn[t] * a;
  }
  const double rss = F[fac_tail(mp)] + zt * zt, icc = F[fac_tail(mp) + 1];
  F[fac_tail(mp)] = rss;
  return score_from_rss(rss, icc, k - 1, sc);
}

// flags of the proposal scorers
enum { SCORE_NPD = 1,      // addition: the enlarged parent Gram is not positive definite (score -inf)
       SCORE_NOFACTOR = 2  // deletion: the current block is marked not positive definite; the caller
                           // scores the reduced set from scratch
};

// Generic MaxPar: w[] / y[] are the only arrays; the block is read row by row, either straight from
// L2 (LOCAL = false) or from a private copy (LOCAL = true, see score_move_staged).
// type 1: score of S + [j] (rowout receives the candidate row); type 2: score of S without slot del.
template <int KMAX, bool LOCAL>
BN_HD double score_move_rows(const double* C, int64_t ldc, const double* diag, int c, const int* S, int k,
                             int type, int j, int del, ScoreConsts sc, const double* F, int mp, int mp_row,
                             double* rowout, int* flags) {
  auto LD2 = [](const double* q) -> D2 { if (LOCAL) return *(const D2*)q; return ld2_l2(q); };
  auto LD1 = [](const double* q) -> double { if (LOCAL) return *q; return ld1_l2(q); };
  double w[KMAX];
  const double rss = LD1(F + fac_tail(mp)), icc = LD1(F + fac_tail(mp) + 1);
  const double* z = F + fac_zoff(mp);
  *flags = 0;
  if (type == 1) {
    double dj = ld_shared_ro(diag + j), ej = ld_shared_ro(C + (int64_t)c * ldc + j);
    for (int i = 0; i < k; i++) {
      const double* Li = F + fac_row(i);
      double a = ld_shared_ro(C + (int64_t)S[i] * ldc + j);
      int t = 0;
      for (; t + 1 < i; t += 2) {
        const D2 l = LD2(Li + t);
        a -= l.x * w[t];
        a -= l.y * w[t + 1];
      }
      const D2 l = LD2(Li + t);  // (L[i][i-1], 1/L[i][i]) or (1/L[i][i], pad)
      if (t < i) { a -= l.x * w[t]; a *= l.y; } else a *= l.x;
      w[i] = a;
      dj -= a * a;
      ej -= a * LD1(z + i);
    }
    bool bad = !(rss == rss) || !(dj > 0.0);
    double rss_new = bad ? NAN : rss - ej * ej / dj;
    if (!bad && !(rss_new * icc > RSS_FLOOR)) { bad = true; rss_new = NAN; }
    if (rowout) {
      for (int t = 0; t < k; t++) rowout[t] = w[t];
      rowout[row_tail(mp_row)] = dj; rowout[row_tail(mp_row) + 1] = ej; rowout[row_tail(mp_row) + 2] = rss_new;
    }
    if (bad) { *flags = SCORE_NPD; return -INFINITY; }
    return score_from_rss(rss_new, icc, k + 1, sc);
  }
  if (!(rss == rss)) { *flags = SCORE_NOFACTOR; return 0.0; }
  double yy, yz;
  {
    const double re = LD1(F + fac_row(del) + del);
    w[del] = re; yy = re * re; yz = re * LD1(z + del);
  }
  for (int i = del + 1; i < k; i++) {
    const double* Li = F + fac_row(i);
    double a = 0.0;
    for (int t = del; t < i; t++) a -= LD1(Li + t) * w[t];
    a *= LD1(Li + i);
    w[i] = a;
    yy += a * a;
    yz += a * LD1(z + i);
  }
  return score_from_rss(rss + yz * yz / yy, icc, k - 1, sc);
}

template <int KMAX>
BN_NOINLINE double score_move_stream(const double* C, int64_t ldc, const double* diag, int c, const int* S, int k,
                                     int type, int j, int del, ScoreConsts sc, const double* F, int mp,
                                     double* rowout, int* flags) {
  return score_move_rows<KMAX, false>(C, ldc, diag, c, S, k, type, j, del, sc, F, mp, mp, rowout, flags);
}

// MaxPar <= 8: the whole block travels in registers (25 16-byte loads issued back to back: one
// L2 round trip) and the arithmetic is straight-line: rows beyond k are zero, so the unused part
// of the triangle contributes nothing and needs no branches.  Both forms are evaluated and the
// move type selects (a warp holds additions and deletions side by side anyway).
BN_NOINLINE double score_move8(const double* C, int64_t ldc, const double* diag, int c, Parents8 S, int k, int type,
                               int j, int del, ScoreConsts sc, const double* F, int mp, double* rowout, int* flags) {
  constexpr int K = 8;
  double L[fac_row(K)], z[K];
  const int nrow = fac_row(k);  // doubles of the rows in use (even)
#pragma unroll
  for (int q = 0; q < fac_row(K) / 2; q++) {
    D2 v; v.x = 0.0; v.y = 0.0;
    if (2 * q < nrow) v = ld2_l2(F + 2 * q);
    L[2 * q] = v.x; L[2 * q + 1] = v.y;
  }
#pragma unroll
  for (int q = 0; q < K / 2; q++) {
    D2 v; v.x = 0.0; v.y = 0.0;
    if (2 * q < k) v = ld2_l2(F + fac_zoff(mp) + 2 * q);
    // (the slot behind z[k-1] was never written: it may hold anything, NaN included)
    z[2 * q] = v.x; z[2 * q + 1] = (2 * q + 1 < k) ? v.y : 0.0;
  }
  const D2 tail = ld2_l2(F + fac_tail(mp));
  const double rss = tail.x, icc = tail.y;
  const bool add = type == 1;
  double g[K], dj = 1.0, ej = 0.0;
#pragma unroll
  for (int i = 0; i < K; i++) g[i] = 0.0;
  if (add) {
    dj = ld_shared_ro(diag + j);
    ej = ld_shared_ro(C + (int64_t)c * ldc + j);
#pragma unroll
    for (int i = 0; i < K; i++)
      if (i < k) g[i] = ld_shared_ro(C + (int64_t)S.s[i] * ldc + j);
  }
  // addition: w = L^-1 C[S,j]
  double w[K];
#pragma unroll
  for (int i = 0; i < K; i++) {
    double a = g[i];
#pragma unroll
    for (int t = 0; t < i; t++) a -= L[fac_row(i) + t] * w[t];
    a *= L[fac_row(i) + i];
    w[i] = a;
    dj -= a * a;
    ej -= a * z[i];
  }
  // deletion: y = column `del` of L^-1 (y[t] = 0 for t < del falls out of the recurrence)
  double y[K], yy = 0.0, yz = 0.0;
#pragma unroll
  for (int i = 0; i < K; i++) {
    double a = 0.0;
#pragma unroll
    for (int t = 0; t < i; t++) a -= L[fac_row(i) + t] * y[t];
    a = (i == del) ? L[fac_row(i) + i] : a * L[fac_row(i) + i];
    y[i] = a;
    yy += a * a;
    yz += a * z[i];
  }
  const bool nofac = !(rss == rss);
  const double num = add ? ej * ej : yz * yz, den = add ? dj : yy;
  const double q = num / den;
  const double rss_new = add ? rss - q : rss + q;
  const bool bad = add && (nofac || !(dj > 0.0) || !(rss_new * icc > RSS_FLOOR));
  if (add && rowout) {
#pragma unroll
    for (int t = 0; t < K / 2; t++) { D2 v; v.x = w[2 * t]; v.y = w[2 * t + 1]; *(D2*)(rowout + 2 * t) = v; }
    D2 a; a.x = dj; a.y = ej;
    *(D2*)(rowout + row_tail(mp)) = a;
    D2 b; b.x = bad ? NAN : rss_new; b.y = 0.0;
    *(D2*)(rowout + row_tail(mp) + 2) = b;
  }
  *flags = bad ? SCORE_NPD : ((!add && nofac) ? SCORE_NOFACTOR : 0);
  const double s = score_from_rss(rss_new, icc, add ? k + 1 : k - 1, sc);
  return bad ? -INFINITY : s;
}

// scoring from scratch (no factor): the fallback of SCORE_NOFACTOR
template <int KMAX>
BN_NOINLINE double score_scratch(const double* C, int64_t ldc, int c, const int* pc, int k, int del, int n_samples,
                                 int* nonpd) {
  double L[KMAX * (KMAX + 1) / 2], z[KMAX];
  int S2[KMAX];
  int kk = 0;
  for (int e = 0; e < k; e++) if (e != del) S2[kk++] = pc[e];
  return score_set(C, ldc, c, S2, kk, n_samples, L, z, nonpd);
}

}  // namespace bn
