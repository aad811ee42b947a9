// tail.cuh -- the regression tail computed straight from (raw columns, FE tables): z~_i = z_i - sum_j alpha_j[code_j(i)]
// is formed in registers, so the demeaned columns are never written to HBM and never read back.
// Shared by gram.cu (X~'WX~, X~'Wy~) and vcov.cu (residuals, scores, meats).
//
// Row-group layout.  One thread per row (the first version of these kernels) needs all C values of a row plus the
// (C+1)(C+2)/2 Gram accumulators in registers -- 176 registers at C = 6, 214 at C = 11, one CTA per SM -- and every
// 128-bit piece of a gathered table row is its own L1TEX wavefront (32 rows = 32 lines per warp request): the
// pass is latency- and LSU-bound far below HBM speed.  Here a lane owns ONE column slot of FOUR-row tiles:
// lane = 4 g + t holds z[row t of the tile][slot g], which is exactly the A (8 x 4) and the B (4 x 8) fragment of
// the fp64 tensor-core instruction mma.m8n8k4, so Z'WZ of a tile is one DMMA with TWO accumulator registers per
// lane; a table row is fetched by the 8 lanes of a row together (one or two sectors per row, 4 rows per request),
// and the kernel runs at full occupancy.  P = C + 1 <= 8 slots fit one slot group, P <= 16 two (three DMMAs per
// tile for the upper triangle).  Per iteration a warp covers 32 consecutive rows: lane (g, t) loads the 8
// consecutive values [base + 8 t, base + 8 t + 8) of its column with four 128-bit loads.
#pragma once
#include "common.cuh"
#include "segreduce.cuh"

constexpr int TL_MAXC = 11;   // columns (y + regressors) the fused tail handles: P = C + 1 <= 16 slots, 11 keeps the
                              // padded table rows (ld 12) within the staged-table budget of the demean kernels
constexpr int TL_T = 256;     // 8 warps per CTA
constexpr int TL_RB = 8;      // rows per lane per iteration (consecutive), 4 t-slots -> 32 rows per warp

struct TailCols {
  const double* col[TL_MAXC];   // col[0] = y, col[1..k] = x
};
// FE tables in ROW order: code[j][i] = group of row i in FE j; alpha[j] = [G_j x ldc] row-major; small tables are
// staged in shared memory (stage_rows > 0), the others are gathered from L2 / HBM
typedef ProjOthers TailTabs;

// D (8 x 8) += A (8 x 4) B (4 x 8) on the fp64 tensor cores; lane = 4 g + t holds A[g][t], B[t][g], D[g][2t], D[g][2t+1]
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// which column a lane's slot g (+ 8 q) of Z = [1 | x_1..x_k | y] is: kind 0 = padding, 1 = intercept, 2 = data
// (raw column `src` of TailCols, table column `tc` -- the tables keep the caller's order y, x_1..x_k)
template <int NS>
struct TailSlots {
  int kind[NS], src[NS];
  __device__ __forceinline__ void init(int g, int C) {
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const int zc = g + 8 * q;
      kind[q] = zc == 0 ? 1 : (zc <= C ? 2 : 0);
      src[q] = zc == C ? 0 : zc;   // the y column closes the row of Z
    }
  }
};

// 8 consecutive values starting at p (p + 8 <= end guaranteed when FULL; 16-byte aligned when VEC)
template <bool VEC, typename T>
__device__ __forceinline__ void load8(const T* __restrict__ p, int64_t nvalid, T out[TL_RB], T fill);

template <>
__device__ __forceinline__ void load8<true, double>(const double* __restrict__ p, int64_t, double out[TL_RB], double) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double2 v = __ldcs(reinterpret_cast<const double2*>(p) + j);
    out[2 * j] = v.x; out[2 * j + 1] = v.y;
  }
}
template <>
__device__ __forceinline__ void load8<false, double>(const double* __restrict__ p, int64_t nvalid, double out[TL_RB],
                                                     double fill) {
#pragma unroll
  for (int i = 0; i < TL_RB; ++i) out[i] = i < nvalid ? __ldcs(p + i) : fill;
}
template <>
__device__ __forceinline__ void load8<true, int>(const int* __restrict__ p, int64_t, int out[TL_RB], int) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int4 v = __ldcs(reinterpret_cast<const int4*>(p) + j);
    out[4 * j] = v.x; out[4 * j + 1] = v.y; out[4 * j + 2] = v.z; out[4 * j + 3] = v.w;
  }
}
template <>
__device__ __forceinline__ void load8<false, int>(const int* __restrict__ p, int64_t nvalid, int out[TL_RB], int fill) {
#pragma unroll
  for (int i = 0; i < TL_RB; ++i) out[i] = i < nvalid ? __ldcs(p + i) : fill;
}

// z[q][i] = Z[rbase + i][slot g + 8 q] of the DEMEANED design for i < 8; rows >= n give 0 in every slot (intercept
// included), so they drop out of every sum.  VEC: vector loads, all 8 rows exist.
template <int NS, bool VEC>
__device__ __forceinline__ void tail_lane_rows(const TailCols& cols, const TailTabs& T, const double* const tabs[LFE_MAX_FE],
                                               int ldc, const TailSlots<NS>& sl, int64_t rbase, int64_t n,
                                               double z[NS][TL_RB]) {
  const int64_t nvalid = VEC ? TL_RB : (n - rbase < TL_RB ? (n - rbase > 0 ? n - rbase : 0) : TL_RB);
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    if (sl.kind[q] == 2) {
      load8<VEC, double>(cols.col[sl.src[q]] + rbase, nvalid, z[q], 0.0);
    } else {
#pragma unroll
      for (int i = 0; i < TL_RB; ++i) z[q][i] = (sl.kind[q] == 1 && i < nvalid) ? 1.0 : 0.0;
    }
  }
  for (int j = 0; j < T.n; ++j) {
    int code[TL_RB];
    load8<VEC, int>(T.code[j] + rbase, nvalid, code, 0);
    const double* tab = tabs[j];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      if (sl.kind[q] == 2) {
        const double* tcol = tab + sl.src[q];
#pragma unroll
        for (int i = 0; i < TL_RB; ++i) z[q][i] -= tcol[(size_t)code[i] * ldc];
      }
    }
  }
  if (!VEC) {
#pragma unroll
    for (int q = 0; q < NS; ++q)
#pragma unroll
      for (int i = 0; i < TL_RB; ++i)
        if (i >= nvalid) z[q][i] = 0.0;
  }
}

// host: alignment test shared by the fused launches (128-bit loads of columns, codes, weights)
static inline bool tail_aligned(const TailCols& cols, int C, const TailTabs& T, const double* w) {
  bool ok = (w == nullptr) || ((((uintptr_t)w) & 15) == 0);
  for (int c = 0; c < C; ++c) ok = ok && ((((uintptr_t)cols.col[c]) & 15) == 0);
  for (int j = 0; j < T.n; ++j) ok = ok && ((((uintptr_t)T.code[j]) & 15) == 0);
  return ok;
}
