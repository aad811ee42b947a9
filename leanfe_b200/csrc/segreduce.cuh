// segreduce.cuh -- segmented (by-key) reduction building blocks shared by demean.cu and vcov.cu.
#pragma once
#include "common.cuh"

constexpr int PJ_T = 256;
constexpr int PJ_I = 4;  // consecutive entries per thread
constexpr int PJ_MAXNC = 8;

struct ProjOthers {
  const int32_t* code[LFE_MAX_FE];
  const double* alpha[LFE_MAX_FE];
  int n;
};
struct ColPtrs {
  const double* col[PJ_MAXNC];
};
struct MutColPtrs {
  double* col[PJ_MAXNC];
};

__device__ __forceinline__ void load4_i32(const int32_t* __restrict__ p, int64_t base, int64_t n, bool vec, int out[4]) {
  if (vec && base + 3 < n) {
    int4 v = __ldcs(reinterpret_cast<const int4*>(p + base));
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) out[e] = (base + e < n) ? __ldcs(p + base + e) : -1;
  }
}

// Flush thread-local runs and segmented-reduce the open (tail) runs across the warp.
template <int NC>
__device__ __forceinline__ void segmented_flush(const int key[PJ_I], double v[PJ_I][NC], double* __restrict__ acc,
                                                int64_t ld_row, int acc_col0, int64_t ld_col = 1) {
  const int lane = threadIdx.x & 31;
  int cur = key[0];
  double run[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) run[c] = v[0][c];
#pragma unroll
  for (int e = 1; e < PJ_I; ++e) {
    if (key[e] == cur) {
#pragma unroll
      for (int c = 0; c < NC; ++c) run[c] += v[e][c];
    } else {
      if (cur >= 0) {
        double* a = acc + (size_t)cur * ld_row + (size_t)acc_col0 * ld_col;
#pragma unroll
        for (int c = 0; c < NC; ++c) atomicAdd(a + (size_t)c * ld_col, run[c]);
      }
      cur = key[e];
#pragma unroll
      for (int c = 0; c < NC; ++c) run[c] = v[e][c];
    }
  }
  // inclusive segmented scan of the tail runs.  Segments are maximal runs of ADJACENT lanes with equal
  // keys (head flags), so the result is exact for unsorted keys too (equal keys that are not adjacent
  // simply end up in separate atomics).
  const int prv = __shfl_up_sync(0xffffffffu, cur, 1);
  const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prv != cur);
  const int seg_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    bool add = (lane - d >= seg_start);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      double t = shfl_up_d(run[c], d);
      if (add) run[c] += t;
    }
  }
  int nxt = __shfl_down_sync(0xffffffffu, cur, 1);
  if (cur >= 0 && (lane == 31 || nxt != cur)) {
    double* a = acc + (size_t)cur * ld_row + (size_t)acc_col0 * ld_col;
#pragma unroll
    for (int c = 0; c < NC; ++c) atomicAdd(a + (size_t)c * ld_col, run[c]);
  }
}

