// Host-side interface of the Gram build (gram.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bn {

constexpr int GRAM_MEAN_MAX_CHUNKS = 64;

struct GramPlan {
  int n_tiles;          // ceil(P / 128)
  int n_pairs;          // upper-triangular tile pairs
  int stages_total;     // ceil(N / 32) pipeline stages along the sample axis
  int n_splits;         // split-K factor
  int64_t items;        // CTAs = n_splits * n_pairs
  int64_t workspace_bytes;
  int64_t ld_centered;  // leading dimension of the centred copy (multiple of 16)
};

GramPlan gram_plan(int n_samples, int P, int n_sms);

// column sums and means of an n_samples x P block (d_scratch_part: P * GRAM_MEAN_MAX_CHUNKS doubles)
void gram_column_sums(const double* dX, int64_t ldx, int n_samples, int P, double* d_scratch_part,
                      double* d_colsum, double* d_mean, cudaStream_t stream);

// Enqueues means -> centring -> DMMA Gram -> reduction on `stream`.  With mean_given the
// means stage is skipped and d_mean is read as supplied (row block of a sharded matrix).
// Returns nullptr on success or a static error string.
const char* gram_build(const double* dX, int64_t ldx, int n_samples, int P, double* dXc,
                       int64_t ld_centered, double* d_partial, const GramPlan& pl, double* d_colsum,
                       double* d_mean, double* d_C, int64_t ldc, double* d_scratch_part,
                       int* d_error_flag, cudaStream_t stream, int64_t* launches, bool mean_given = false);

// X in (pinned) host memory, column-major n_samples x P: the H2D copy is pipelined with the
// build, one block of 128 columns at a time; block_arrived: pl.n_tiles events.  Same bits.
const char* gram_build_from_host(const double* hX, int n_samples, int P, double* dXc, int64_t ld_centered,
                                 double* d_partial, const GramPlan& pl, double* d_colsum, double* d_mean,
                                 double* d_C, int64_t ldc, double* d_scratch_part, int* d_error_flag,
                                 cudaStream_t stream, cudaStream_t copy_stream, cudaEvent_t* block_arrived,
                                 int64_t* launches);

}  // namespace bn
