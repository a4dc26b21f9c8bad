// Common macros for libbn_b200.  The *_core.cuh headers are written so that
// the same source compiles for the device (product) and, in tests/ only, for
// the host with a one-lane "warp" -- that is how the sequential chain logic is
// checked against the oracle on a machine without a GPU.  libbn_b200.so never
// runs the host instantiation: every entry point launches CUDA kernels.
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define BN_HD __host__ __device__ __forceinline__
#define BN_D __device__ __forceinline__
#else
#define BN_HD inline
#define BN_D inline
#endif

namespace bn {

// L2-only load for the (read-only, shared by all chains) Gram so that gathers
// do not evict the per-chain state that lives in L1.
template <typename T>
BN_HD T ld_shared_ro(const T* p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(p);
#else
  return *p;
#endif
}

// Contraction-free arithmetic where the reference's last-bit behaviour is
// replicated (prior, Hastings ratio, Wichmann-Hill combine).
BN_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b; return r;
#endif
}
BN_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b; return r;
#endif
}
BN_HD double sub_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  volatile double r = a - b; return r;
#endif
}
BN_HD double div_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  volatile double r = a / b; return r;
#endif
}

BN_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
BN_HD int ffs32(uint32_t v) {  // 1-based index of lowest set bit, 0 if none
#if defined(__CUDA_ARCH__)
  return __ffs((int)v);
#else
  return __builtin_ffs((int)v);
#endif
}

// ---------------------------------------------------------------------------
// Warp abstraction: 32 lanes on the device, 1 lane on the host (tests only).
// ---------------------------------------------------------------------------
struct Warp {
#if defined(__CUDA_ARCH__)
  static constexpr int NL = 32;
  static BN_D int lane() { return (int)(threadIdx.x & 31); }
  static BN_D void sync() { __syncwarp(); }
  static BN_D uint32_t ballot(int pred) { return __ballot_sync(0xffffffffu, pred); }
  static BN_D int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  static BN_D double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  static BN_D int sum(int v) { return (int)__reduce_add_sync(0xffffffffu, (unsigned)v); }
  static BN_D long long sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  static BN_D int min(int v) { return __reduce_min_sync(0xffffffffu, v); }
  static BN_D int incl_scan(int v) {
    int l = lane();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, v, o);
      if (l >= o) v += t;
    }
    return v;
  }
  // fixed-order tree sum of doubles (deterministic)
  static BN_D double sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
#else
  static constexpr int NL = 1;
  static int lane() { return 0; }
  static void sync() {}
  static uint32_t ballot(int pred) { return pred ? 1u : 0u; }
  static int shfl(int v, int) { return v; }
  static double shfl(double v, int) { return v; }
  static int sum(int v) { return v; }
  static long long sum(long long v) { return v; }
  static int min(int v) { return v; }
  static int incl_scan(int v) { return v; }
  static double sum(double v) { return v; }
#endif
};

}  // namespace bn
