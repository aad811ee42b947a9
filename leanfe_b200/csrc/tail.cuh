// tail.cuh -- the regression tail computed straight from (raw columns, FE tables): x~_i = x_i - sum_j alpha_j[code_j(i)]
// is formed in registers, so the demeaned columns are never written to HBM and never read back.
// Shared by gram.cu (X~'WX~, X~'Wy~) and vcov.cu (residuals, scores, meats).
#pragma once
#include "common.cuh"
#include "segreduce.cuh"

constexpr int TL_MAXC = 11;   // columns (y + regressors) the fused tail handles: P = C + 1 <= 12 triangle accumulators
constexpr int TL_T = 256;

struct TailCols {
  const double* col[TL_MAXC];   // col[0] = y, col[1..k] = x
};
// FE tables in ROW order: code[j][i] = group of row i in FE j; alpha[j] = [G_j x ldc] row-major; small tables are
// staged in shared memory (stage_rows > 0), the others are gathered from L2 / HBM
typedef ProjOthers TailTabs;

// rows per thread: all loads of a thread's rows are issued before the first use, so the gathers of R rows overlap
template <int NC>
struct TailCfg {
  static constexpr int R = NC <= 4 ? 4 : (NC <= 8 ? 2 : 1);        // residual / score pass (5-8 columns: 3 CTAs per SM)
  // Gram pass: its (NC+1)(NC+2)/2 accumulators come on top, so 5-6 columns take 2 rows per thread and stay within
  // 128 registers = two CTAs per SM (c2: 0.99 -> 0.76 ms; 4 rows per thread needed 176 registers, one CTA per SM)
  static constexpr int R_GRAM = NC <= 4 ? 4 : (NC <= 8 ? 2 : 1);
};

// x[r][c] = col_c[base + r] - sum_j alpha_j[code_j[base + r]][c]   for r < R; VEC: 16-byte aligned columns / codes and
// all R rows exist.  Rows beyond n are zero-filled (valid[r] = false).
template <int NC, int R, bool VEC>
__device__ __forceinline__ void tail_rows(const TailCols& cols, const TailTabs& T, const double* const tabs[LFE_MAX_FE],
                                          int ldc, bool vec2, int64_t base, int64_t n, double x[R][NC], bool valid[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) valid[r] = VEC || (base + r < n);
  if constexpr (VEC && R == 4) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double2 a = __ldcs(reinterpret_cast<const double2*>(cols.col[c] + base));
      const double2 b = __ldcs(reinterpret_cast<const double2*>(cols.col[c] + base) + 1);
      x[0][c] = a.x; x[1][c] = a.y; x[2][c] = b.x; x[3][c] = b.y;
    }
  } else if constexpr (VEC && R == 2) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double2 a = __ldcs(reinterpret_cast<const double2*>(cols.col[c] + base));
      x[0][c] = a.x; x[1][c] = a.y;
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < NC; ++c) x[r][c] = valid[r] ? __ldcs(cols.col[c] + base + r) : 0.0;
  }
  for (int j = 0; j < T.n; ++j) {
    int k[R];
    if constexpr (VEC && R == 4) {
      const int4 k4 = __ldcs(reinterpret_cast<const int4*>(T.code[j] + base));
      k[0] = k4.x; k[1] = k4.y; k[2] = k4.z; k[3] = k4.w;
    } else if constexpr (VEC && R == 2) {
      const int2 k2 = __ldcs(reinterpret_cast<const int2*>(T.code[j] + base));
      k[0] = k2.x; k[1] = k2.y;
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) k[r] = valid[r] ? __ldcs(T.code[j] + base + r) : 0;
    }
    const double* tab = tabs[j];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double* row = tab + (size_t)k[r] * ldc;
      if (vec2) {
#pragma unroll
        for (int c = 0; c + 1 < NC; c += 2) {
          const double2 t = *reinterpret_cast<const double2*>(row + c);
          x[r][c] -= t.x;
          x[r][c + 1] -= t.y;
        }
        if (NC & 1) x[r][NC - 1] -= row[NC - 1];
      } else {
#pragma unroll
        for (int c = 0; c < NC; ++c) x[r][c] -= row[c];
      }
    }
  }
  if constexpr (!VEC) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (!valid[r]) {
#pragma unroll
        for (int c = 0; c < NC; ++c) x[r][c] = 0.0;
      }
  }
}

// host: table staging plan + alignment test shared by the fused launches
static inline bool tail_aligned(const TailCols& cols, int C, const TailTabs& T, const double* w) {
  bool ok = (w == nullptr) || ((((uintptr_t)w) & 15) == 0);
  for (int c = 0; c < C; ++c) ok = ok && ((((uintptr_t)cols.col[c]) & 15) == 0);
  for (int j = 0; j < T.n; ++j) ok = ok && ((((uintptr_t)T.code[j]) & 15) == 0);
  return ok;
}
