// Shared helpers for the ld_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ld_b200.h"

namespace ldb {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define LDB_CUDA(call)                                   \
  do {                                                   \
    int _e = ldb::check_cuda((call), #call);             \
    if (_e) return _e;                                   \
  } while (0)

#define LDB_REQUIRE(cond, code, ...)                     \
  do {                                                   \
    if (!(cond)) {                                       \
      ldb::set_error(__VA_ARGS__);                       \
      return (code);                                     \
    }                                                    \
  } while (0)

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(FULL, v, o);
    v = t < v ? t : v;
  }
  return v;
}

// Pyramid geometry in constant-size POD passed by value to kernels.
struct Pyr {
  int L, N;
  int h[LDB_MAX_LEVELS], w[LDB_MAX_LEVELS], stride[LDB_MAX_LEVELS];
  int off[LDB_MAX_LEVELS + 1];  // prefix sum of A_l; off[L] = A
};

inline int make_pyr(const ldb_pyramid* p, Pyr* out) {
  if (!p || p->num_levels < 1 || p->num_levels > LDB_MAX_LEVELS || p->num_imgs < 1) return LDB_EINVAL;
  out->L = p->num_levels;
  out->N = p->num_imgs;
  int acc = 0;
  for (int l = 0; l < LDB_MAX_LEVELS; ++l) {
    out->off[l] = acc;
    if (l < p->num_levels) {
      if (p->h[l] < 1 || p->w[l] < 1 || p->stride[l] < 1) return LDB_EINVAL;
      out->h[l] = p->h[l];
      out->w[l] = p->w[l];
      out->stride[l] = p->stride[l];
      acc += p->h[l] * p->w[l];
    } else {
      out->h[l] = out->w[l] = out->stride[l] = 0;
    }
  }
  for (int l = p->num_levels; l <= LDB_MAX_LEVELS; ++l) out->off[l] = acc;
  return 0;
}

__device__ __forceinline__ int level_of_anchor(const Pyr& p, int a) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < LDB_MAX_LEVELS; ++i) l += (i < p.L && a >= p.off[i]) ? 1 : 0;
  return l;
}
// row (level-major row space) -> level
__device__ __forceinline__ int level_of_row(const Pyr& p, int r) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < LDB_MAX_LEVELS; ++i) l += (i < p.L && r >= p.off[i] * p.N) ? 1 : 0;
  return l;
}

}  // namespace ldb
