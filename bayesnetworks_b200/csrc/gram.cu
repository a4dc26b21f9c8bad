// K1 -- sufficient statistics on the FP64 tensor pipe.
//
// Replaces the constructor's scalar triple loop (src/network.h:124-136:
// sumX[p1] += X(n,p1); sumXX(p1,p2) += X(n,p1)*X(n,p2), O(N P^2)).
//
//   1. column_mean_kernel   mean[p] = sum_n X(n,p) / N               (HBM bound)
//   2. center_kernel        Xc[p][n] = X(n,p) - mean[p]  into a context-owned,
//                           16-sample padded buffer                    (HBM bound)
//   3. gram_dmma_kernel     C = Xc Xc'  -- the dense contraction.  Upper-
//                           triangular 128x128 output tiles x split-K; operands
//                           staged by TMA (cp.async.bulk.tensor, 128B swizzle)
//                           through a 3-stage mbarrier ring, multiplied with
//                           mma.sync m8n8k4 f64 (SASS DMMA.8x8x4 -- the only
//                           FP64 tensor instruction sm_100a has; tcgen05 has no
//                           f64 kind), accumulators in registers.
//   4. gram_reduce_kernel   fixed-order sum of the split-K partials + mirror
//                           to the lower triangle (deterministic: no atomics).
//
// Centring first keeps the score's RSS = C_cc - b'A^-1 b free of the
// mean^2/variance cancellation an uncentred Gram would have.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "gram.h"

namespace bn {

// ---------------------------------------------------------------------------
// 1. column means
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) column_sum_partial_kernel(
    const double* __restrict__ X, int64_t ldx, int n_samples, int chunks, double* __restrict__ part) {
  const int p = blockIdx.x, ch = blockIdx.y;
  const int64_t per = ((int64_t)n_samples + chunks - 1) / chunks;
  const int64_t lo = (int64_t)ch * per;
  int64_t hi = lo + per;
  if (hi > n_samples) hi = n_samples;
  const double* col = X + (int64_t)p * ldx;
  double acc = 0.0;
  for (int64_t n = lo + threadIdx.x; n < hi; n += 256) acc += col[n];
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[(int64_t)p * chunks + ch] = sh[0];
}

__global__ void column_mean_finish_kernel(const double* __restrict__ part, int chunks, int P,
                                          int n_samples, double* __restrict__ colsum,
                                          double* __restrict__ mean) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double acc = 0.0;
  for (int c = 0; c < chunks; c++) acc += part[(int64_t)p * chunks + c];
  colsum[p] = acc;
  mean[p] = acc / (double)n_samples;
}

// ---------------------------------------------------------------------------
// 2. centring into the padded buffer (works in place when Xc == X, ldc == ldx)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) center_kernel(const double* X, int64_t ldx, int n_samples,
                                                     const double* __restrict__ mean, double* Xc,
                                                     int64_t ldc) {
  const int p = blockIdx.y;
  const double mu = mean[p];
  const double* src = X + (int64_t)p * ldx;
  double* dst = Xc + (int64_t)p * ldc;
  for (int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x; n < ldc; n += (int64_t)gridDim.x * 256)
    dst[n] = (n < n_samples) ? (src[n] - mu) : 0.0;
}

// ---------------------------------------------------------------------------
// 3. the DMMA Gram kernel
// ---------------------------------------------------------------------------
constexpr int TILE = 128;          // output tile edge (variables)
constexpr int BOX_K = 16;          // samples per TMA box row: 16 * 8 B = 128 B = swizzle span
constexpr int BOXES = 2;           // boxes per pipeline stage -> 32 samples
constexpr int STAGE_K = BOX_K * BOXES;
constexpr int NSTAGE = 3;
constexpr int BOX_BYTES = TILE * BOX_K * 8;            // 16 KB
constexpr int OPERAND_BYTES = BOX_BYTES * BOXES;       // 32 KB per operand per stage
constexpr int STAGE_BYTES = 2 * OPERAND_BYTES;         // 64 KB
constexpr int GRAM_CONSUMER_WARPS = 8;
constexpr int GRAM_THREADS = (GRAM_CONSUMER_WARPS + 1) * 32;
constexpr int GRAM_SMEM = NSTAGE * STAGE_BYTES + 1024 /*align*/ + 64 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a pipeline bug must not hang the GPU box
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 24); spin++) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int x /*sample*/, int y /*variable*/) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// Work item = (k-split s, tile pair t).  Items with the same s are adjacent in
// blockIdx so that co-resident CTAs stream the same sample range (L2 reuse of
// the row panels).
__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_dmma_kernel(const __grid_constant__ CUtensorMap tmap, int n_tiles, int n_pairs,
                 int stages_total, int n_splits, double* __restrict__ partial, int* __restrict__ error_flag,
                 int tj_only) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + NSTAGE * STAGE_BYTES);
  uint64_t* empty = full + NSTAGE;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tj_only < 0: the grid covers every (split, pair).  tj_only >= 0: only the pairs whose
  // second tile is tj_only (those that become computable when that block of columns has
  // arrived from the host); the partial tile lands in the same slot either way.
  int split, ti, tj;
  if (tj_only < 0) {
    split = blockIdx.x / n_pairs;
    int pair = blockIdx.x - split * n_pairs;
    // unrank the upper-triangular pair index: rows ti have (n_tiles - ti) entries
    ti = 0;
    while (pair >= n_tiles - ti) { pair -= n_tiles - ti; ti++; }
    tj = ti + pair;
  } else {
    split = blockIdx.x / (tj_only + 1);
    ti = blockIdx.x - split * (tj_only + 1);
    tj = tj_only;
  }
  const int item = split * n_pairs + ti * n_tiles - (ti * (ti - 1)) / 2 + (tj - ti);
  const bool diag = (ti == tj);

  const int per = (stages_total + n_splits - 1) / n_splits;
  const int st_lo = split * per;
  int st_hi = st_lo + per;
  if (st_hi > stages_total) st_hi = stages_total;
  const int n_st = st_hi > st_lo ? st_hi - st_lo : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], GRAM_CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == GRAM_CONSUMER_WARPS) {
    // ===== TMA producer (one elected lane) =====
    if (lane == 0) {
      const uint32_t bytes = diag ? OPERAND_BYTES : STAGE_BYTES;
      for (int k = 0; k < n_st; k++) {
        const int s = k % NSTAGE;
        if (k >= NSTAGE) {
          if (!mbar_wait(&empty[s], ((k / NSTAGE) - 1) & 1)) { atomicExch(error_flag, 1); break; }
        }
        uint8_t* stage = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full[s], bytes);
        const int x0 = (st_lo + k) * STAGE_K;
        for (int b = 0; b < BOXES; b++) {
          tma_load_2d(stage + b * BOX_BYTES, &tmap, &full[s], x0 + b * BOX_K, ti * TILE);
          if (!diag)
            tma_load_2d(stage + OPERAND_BYTES + b * BOX_BYTES, &tmap, &full[s], x0 + b * BOX_K, tj * TILE);
        }
      }
    }
    return;
  }

  // ===== consumers: 8 warps as 2 (rows) x 4 (cols); warp tile 64 x 32 =====
  const int wr = warp >> 2, wc = warp & 3;
  const int g = lane >> 2, t = lane & 3;
  const int rho = ((g & 1) << 2) | (g >> 1);  // MMA row/col g  <->  tile row rho(g): keeps the
                                              // 8 lanes of a quarter-warp on 8 distinct 16 B columns
  // byte offset of this lane's 16 B chunk inside an 8-row group, for k-half h
  uint32_t lane_off[2];
#pragma unroll
  for (int h = 0; h < 2; h++) lane_off[h] = (uint32_t)(rho * 128 + (((t + 4 * h) ^ rho) << 4));

  double acc[8][4][2];
#pragma unroll
  for (int mi = 0; mi < 8; mi++)
#pragma unroll
    for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  const uint32_t smem_base = smem_u32(smem);
  bool ok = true;
  for (int k = 0; k < n_st && ok; k++) {
    const int s = k % NSTAGE;
    if (!mbar_wait(&full[s], (k / NSTAGE) & 1)) { atomicExch(error_flag, 2); ok = false; break; }
    const uint32_t a_base = smem_base + s * STAGE_BYTES + (wr * 64) * 128;
    const uint32_t b_base = smem_base + s * STAGE_BYTES + (diag ? 0 : OPERAND_BYTES) + (wc * 32) * 128;
#pragma unroll
    for (int b = 0; b < BOXES; b++) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        double2 af[8], bf[4];
#pragma unroll
        for (int mi = 0; mi < 8; mi++) af[mi] = lds128(a_base + b * BOX_BYTES + mi * 8 * 128 + lane_off[h]);
#pragma unroll
        for (int ni = 0; ni < 4; ni++) bf[ni] = lds128(b_base + b * BOX_BYTES + ni * 8 * 128 + lane_off[h]);
        // two k4 steps: samples {2t} and {2t+1} of this half (any fixed assignment of
        // samples to the MMA's k index is valid as long as A and B agree)
#pragma unroll
        for (int mi = 0; mi < 8; mi++)
#pragma unroll
          for (int ni = 0; ni < 4; ni++) {
            dmma_884(acc[mi][ni][0], acc[mi][ni][1], af[mi].x, bf[ni].x);
            dmma_884(acc[mi][ni][0], acc[mi][ni][1], af[mi].y, bf[ni].y);
          }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // epilogue: partial tile [128][128] for this work item
  double* out = partial + (int64_t)item * TILE * TILE;
#pragma unroll
  for (int mi = 0; mi < 8; mi++) {
    const int r = wr * 64 + mi * 8 + rho;
#pragma unroll
    for (int ni = 0; ni < 4; ni++) {
      const int c0 = wc * 32 + ni * 8 + t;  // MMA column 2t   <-> tile column rho(2t)   = t
      out[r * TILE + c0] = acc[mi][ni][0];
      out[r * TILE + c0 + 4] = acc[mi][ni][1];  // MMA column 2t+1 <-> rho(2t+1) = t + 4
    }
  }
}

// ---------------------------------------------------------------------------
// 4. split-K reduction (fixed order) + symmetric fill
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_reduce_kernel(const double* __restrict__ partial, int n_tiles,
                                                          int n_pairs, int n_splits, int P,
                                                          double* __restrict__ C, int64_t ldc) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (int64_t)P * P) return;
  const int i = (int)(idx / P), j = (int)(idx % P);
  const int a = i < j ? i : j, b = i < j ? j : i;  // element (a,b), a <= b, lives in tile (ta <= tb)
  const int ta = a / TILE, tb = b / TILE;
  // rank of (ta,tb) in the upper-triangular enumeration
  const int pair = ta * n_tiles - (ta * (ta - 1)) / 2 + (tb - ta);
  const int64_t off = (int64_t)(a % TILE) * TILE + (b % TILE);
  double acc = 0.0;
  for (int s = 0; s < n_splits; s++)
    acc += partial[((int64_t)s * n_pairs + pair) * TILE * TILE + off];
  C[(int64_t)i * ldc + j] = acc;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);

static encode_fn_t get_encode_fn() {
  static encode_fn_t fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (encode_fn_t)p;
  }
  return fn;
}

GramPlan gram_plan(int n_samples, int P, int n_sms) {
  GramPlan pl;
  pl.n_tiles = (P + TILE - 1) / TILE;
  pl.n_pairs = pl.n_tiles * (pl.n_tiles + 1) / 2;
  pl.stages_total = (n_samples + STAGE_K - 1) / STAGE_K;
  // split K so the grid is close to a whole number of waves (1 CTA per SM) while each
  // CTA still streams >= 64 stages (2048 samples); keep the partial workspace bounded.
  int best = 1;
  double best_score = -1.0;
  const int max_splits_by_k = pl.stages_total / 64 > 0 ? pl.stages_total / 64 : 1;
  const int64_t ws_cap = (int64_t)3 << 30;  // bytes
  for (int s = 1; s <= max_splits_by_k && s <= 4096; s++) {
    const int64_t items = (int64_t)s * pl.n_pairs;
    if (items * TILE * TILE * 8 > ws_cap && s > 1) break;
    const int64_t waves = (items + n_sms - 1) / n_sms;
    const double eff = (double)items / (double)(waves * n_sms);
    // prefer fuller waves; among equals prefer more waves up to ~8 (tail amortisation)
    const double score = eff + 1e-3 * (double)(waves < 8 ? waves : 8);
    if (score > best_score + 1e-12) { best_score = score; best = s; }
  }
  pl.n_splits = best;
  pl.items = (int64_t)best * pl.n_pairs;
  pl.workspace_bytes = pl.items * TILE * TILE * 8;
  pl.ld_centered = ((int64_t)n_samples + 15) / 16 * 16;
  return pl;
}

void gram_column_sums(const double* dX, int64_t ldx, int n_samples, int P, double* d_scratch_part,
                      double* d_colsum, double* d_mean, cudaStream_t stream) {
  int chunks = (int)((n_samples + 65535) / 65536);
  if (chunks < 1) chunks = 1;
  if (chunks > GRAM_MEAN_MAX_CHUNKS) chunks = GRAM_MEAN_MAX_CHUNKS;
  column_sum_partial_kernel<<<dim3(P, chunks), 256, 0, stream>>>(dX, ldx, n_samples, chunks, d_scratch_part);
  column_mean_finish_kernel<<<(P + 127) / 128, 128, 0, stream>>>(d_scratch_part, chunks, P, n_samples,
                                                                  d_colsum, d_mean);
}

const char* gram_build(const double* dX, int64_t ldx, int n_samples, int P, double* dXc, int64_t ld_centered,
                       double* d_partial, const GramPlan& pl, double* d_colsum, double* d_mean,
                       double* d_C, int64_t ldc, double* d_scratch_part, int* d_error_flag,
                       cudaStream_t stream, int64_t* launches, bool mean_given) {
  // 1. means (skipped for a row block of a sharded matrix: the caller supplies the global means)
  if (!mean_given) {
    gram_column_sums(dX, ldx, n_samples, P, d_scratch_part, d_colsum, d_mean, stream);
    if (launches) *launches += 2;
  }
  // 2. centre
  int gx = (int)((ld_centered + 255) / 256);
  if (gx > 1024) gx = 1024;
  center_kernel<<<dim3(gx, P), 256, 0, stream>>>(dX, ldx, n_samples, d_mean, dXc, ld_centered);
  // 3. DMMA Gram
  encode_fn_t enc = get_encode_fn();
  if (!enc) return "cuTensorMapEncodeTiled entry point not available";
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {(cuuint64_t)n_samples, (cuuint64_t)P};
  cuuint64_t gstride[1] = {(cuuint64_t)ld_centered * 8};
  cuuint32_t box[2] = {BOX_K, TILE};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)dXc, gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed";
  if (cudaFuncSetAttribute(gram_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM) !=
      cudaSuccess)
    return "cudaFuncSetAttribute(gram_dmma_kernel) failed";
  cudaMemsetAsync(d_error_flag, 0, sizeof(int), stream);
  gram_dmma_kernel<<<(unsigned)pl.items, GRAM_THREADS, GRAM_SMEM, stream>>>(
      tmap, pl.n_tiles, pl.n_pairs, pl.stages_total, pl.n_splits, d_partial, d_error_flag, -1);
  // 4. reduce
  const int64_t total = (int64_t)P * P;
  gram_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_partial, pl.n_tiles, pl.n_pairs,
                                                                          pl.n_splits, P, d_C, ldc);
  if (launches) *launches += 3;
  return nullptr;
}

// The same build with X still in (pinned) host memory: the matrix crosses PCIe one block of
// TILE columns at a time on `copy_stream`, and as soon as block b is there `stream` runs its
// column sums, centres it in place and multiplies the tile pairs (i <= b, b).  Only the pairs of
// the last block (2 / (n_tiles + 1) of the work) are left when the copy ends.  Every partial
// tile is the same launch-independent computation, so the result has the bits of gram_build.
// ---------------------------------------------------------------------------
// Pageable host matrices (what R owns).  cudaMemcpyAsync from pageable memory is staged by the
// driver through one thread (~8 GB/s); here a few host threads copy 8 MB chunks into a ring of
// pinned buffers and every chunk goes to the device at the PCIe rate while the next one is being
// staged.  Marshalling only: bytes are moved, nothing is computed on the host.
// ---------------------------------------------------------------------------
namespace {
constexpr size_t STAGE_CHUNK = (size_t)8 << 20;
constexpr int STAGE_BUFS = 4;
constexpr int STAGE_THREADS_MAX = 16;
// host threads that fill the pinned ring: 8 by default (the copy of 800 MB is bound by the host's memcpy rate
// with 4: 62.8 ms per end-to-end step with 4 threads, 55.4 with 8, 52.3 from pinned memory); BN_B200_STAGE_THREADS overrides
static int stage_threads() {
  static int n = 0;
  if (n == 0) {
    const char* e = getenv("BN_B200_STAGE_THREADS");
    if (e) {
      n = atoi(e);
    } else {
      const int hw = (int)std::thread::hardware_concurrency();  // (half of the cores, at most 8)
      n = hw >= 16 ? 8 : (hw >= 4 ? hw / 2 : 2);
    }
    if (n < 1) n = 1;
    if (n > STAGE_THREADS_MAX) n = STAGE_THREADS_MAX;
  }
  return n;
}

class StagePool {
 public:
  static StagePool& get() { static StagePool p; return p; }
  std::mutex use;  // one staged copy at a time
  bool ready() {
    if (state_ == 0) {
      state_ = -1;
      bool ok = true;
      for (int b = 0; b < STAGE_BUFS && ok; b++) {
        ok = cudaHostAlloc(&buf_[b], STAGE_CHUNK, cudaHostAllocDefault) == cudaSuccess &&
             cudaEventCreateWithFlags(&free_[b], cudaEventDisableTiming) == cudaSuccess;
      }
      if (ok) {
        for (int t = 0; t < stage_threads(); t++) workers_.emplace_back([this, t] { work(t); });
        state_ = 1;
      } else {
        cudaGetLastError();
      }
    }
    return state_ == 1;
  }
  void* buffer(int b) { return buf_[b]; }
  cudaEvent_t& event(int b) { return free_[b]; }
  // dst[0..bytes) = src[0..bytes) by all workers; returns when done
  void copy(void* dst, const void* src, size_t bytes) {
    std::unique_lock<std::mutex> lk(mu_);
    dst_ = (char*)dst; src_ = (const char*)src; bytes_ = bytes; pending_ = stage_threads(); gen_++;
    cv_.notify_all();
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  StagePool() {}
  ~StagePool() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      quit_ = true;
      cv_.notify_all();
    }
    for (auto& w : workers_) w.join();
  }
  void work(int t) {
    uint64_t seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return quit_ || gen_ != seen; });
      if (quit_) return;
      seen = gen_;
      const size_t per = (bytes_ / stage_threads() + 63) & ~(size_t)63;
      const size_t lo = per * t < bytes_ ? per * t : bytes_;
      const size_t hi = lo + per < bytes_ ? lo + per : bytes_;
      char* d = dst_; const char* s = src_;
      lk.unlock();
      if (hi > lo) memcpy(d + lo, s + lo, hi - lo);
      lk.lock();
      if (--pending_ == 0) done_.notify_all();
    }
  }
  int state_ = 0;
  void* buf_[STAGE_BUFS] = {};
  cudaEvent_t free_[STAGE_BUFS] = {};
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  char* dst_ = nullptr; const char* src_ = nullptr;
  size_t bytes_ = 0; int pending_ = 0; uint64_t gen_ = 0; bool quit_ = false;
};

bool host_is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}

// columns [c0, c0 + nc) of the column-major host matrix -> device rows of stride ld_centered
cudaError_t staged_block_copy(StagePool& sp, int& ring, const double* hX, int n_samples, int c0, int nc,
                              double* dst, int64_t ld_centered, cudaStream_t copy_stream) {
  const size_t col_bytes = (size_t)n_samples * 8;
  int cols_per = (int)(STAGE_CHUNK / col_bytes);
  if (cols_per < 1) return cudaErrorInvalidValue;  // (a single column above 8 MB: the caller uses the plain path)
  for (int c = 0; c < nc; c += cols_per) {
    const int n = (nc - c < cols_per) ? nc - c : cols_per;
    const int b = ring++ % STAGE_BUFS;
    cudaError_t e = cudaEventSynchronize(sp.event(b));  // the previous copy out of this buffer is done
    if (e != cudaSuccess) return e;
    sp.copy(sp.buffer(b), hX + (int64_t)(c0 + c) * n_samples, col_bytes * n);
    e = cudaMemcpy2DAsync(dst + (int64_t)c * ld_centered, (size_t)ld_centered * 8, sp.buffer(b), col_bytes, col_bytes,
                          (size_t)n, cudaMemcpyHostToDevice, copy_stream);
    if (e != cudaSuccess) return e;
    cudaEventRecord(sp.event(b), copy_stream);
  }
  return cudaSuccess;
}
}  // namespace

const char* gram_build_from_host(const double* hX, int n_samples, int P, double* dXc, int64_t ld_centered,
                                 double* d_partial, const GramPlan& pl, double* d_colsum, double* d_mean,
                                 double* d_C, int64_t ldc, double* d_scratch_part, int* d_error_flag,
                                 cudaStream_t stream, cudaStream_t copy_stream, cudaEvent_t* block_arrived,
                                 int64_t* launches) {
  encode_fn_t enc = get_encode_fn();
  if (!enc) return "cuTensorMapEncodeTiled entry point not available";
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {(cuuint64_t)n_samples, (cuuint64_t)P};
  cuuint64_t gstride[1] = {(cuuint64_t)ld_centered * 8};
  cuuint32_t box[2] = {BOX_K, TILE};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)dXc, gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed";
  if (cudaFuncSetAttribute(gram_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM) !=
      cudaSuccess)
    return "cudaFuncSetAttribute(gram_dmma_kernel) failed";
  cudaMemsetAsync(d_error_flag, 0, sizeof(int), stream);
  int chunks = (int)((n_samples + 65535) / 65536);
  if (chunks < 1) chunks = 1;
  if (chunks > GRAM_MEAN_MAX_CHUNKS) chunks = GRAM_MEAN_MAX_CHUNKS;
  int gx = (int)((ld_centered + 255) / 256);
  if (gx > 1024) gx = 1024;
  // pageable source of some size: staged through pinned buffers by a few host threads
  StagePool& sp = StagePool::get();
  std::unique_lock<std::mutex> stage_lock(sp.use, std::defer_lock);
  bool staged = (size_t)n_samples * P * 8 >= ((size_t)32 << 20) && (size_t)n_samples * 8 <= STAGE_CHUNK &&
                host_is_pageable(hX);
  if (staged) {
    stage_lock.lock();
    staged = sp.ready();
    if (!staged) stage_lock.unlock();
  }
  int ring = 0;
  for (int b = 0; b < pl.n_tiles; b++) {
    const int c0 = b * TILE;
    const int nc = (P - c0 < TILE) ? P - c0 : TILE;
    double* dst = dXc + (int64_t)c0 * ld_centered;
    const double* src = hX + (int64_t)c0 * n_samples;
    cudaError_t e;
    if (staged)
      e = staged_block_copy(sp, ring, hX, n_samples, c0, nc, dst, ld_centered, copy_stream);
    else if (ld_centered == n_samples)
      e = cudaMemcpyAsync(dst, src, (size_t)n_samples * nc * 8, cudaMemcpyHostToDevice, copy_stream);
    else
      e = cudaMemcpy2DAsync(dst, (size_t)ld_centered * 8, src, (size_t)n_samples * 8, (size_t)n_samples * 8,
                            (size_t)nc, cudaMemcpyHostToDevice, copy_stream);
    if (e != cudaSuccess) return "H2D copy of a column block failed";
    cudaEventRecord(block_arrived[b], copy_stream);
    cudaStreamWaitEvent(stream, block_arrived[b], 0);
    column_sum_partial_kernel<<<dim3(nc, chunks), 256, 0, stream>>>(dst, ld_centered, n_samples, chunks,
                                                                     d_scratch_part + (int64_t)c0 * chunks);
    column_mean_finish_kernel<<<(nc + 127) / 128, 128, 0, stream>>>(d_scratch_part + (int64_t)c0 * chunks, chunks,
                                                                     nc, n_samples, d_colsum + c0, d_mean + c0);
    center_kernel<<<dim3(gx, nc), 256, 0, stream>>>(dst, ld_centered, n_samples, d_mean + c0, dst, ld_centered);
    gram_dmma_kernel<<<(unsigned)(pl.n_splits * (b + 1)), GRAM_THREADS, GRAM_SMEM, stream>>>(
        tmap, pl.n_tiles, pl.n_pairs, pl.stages_total, pl.n_splits, d_partial, d_error_flag, b);
    if (launches) *launches += 4;
  }
  const int64_t total = (int64_t)P * P;
  gram_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_partial, pl.n_tiles, pl.n_pairs,
                                                                          pl.n_splits, P, d_C, ldc);
  if (launches) *launches += 1;
  return nullptr;
}

}  // namespace bn
