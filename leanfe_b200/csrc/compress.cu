// compress.cu -- YOCO compression: device group-by producing sufficient statistics.
// Reference: python/leanfe/compress.py:325-350 (group_by(x + fe + cluster).agg(len, sum y, sum y^2)) and
// the distinct count of estimate_compression_ratio (compress.py:249-250).
//
// The K code columns are folded into a mixed-radix key (most significant = first column, so ascending key
// order == ascending lexicographic order of the code tuples) and the key is factorised.  Folding is STAGED:
// columns are accumulated while the product of their cardinalities stays below 2^26, then that partial key
// is replaced by its dense sorted rank before more columns are folded in: a direct-address presence table,
// filled and read back by two passes that RECOMPUTE the folded key from the raw columns (k_fold_mark /
// k_fold_lookup), so the 64-bit keys are never written to memory.  uint8 columns (dummies) enter as they are,
// 1 byte per row, without a factorisation pass.  Ranks are order-preserving, so the final group order is still lexicographic, and the
// composite key of, e.g., 23 dummies x 1000 groups x 24 periods (2e11 combinations, 24 000 present) never
// needs a 38-bit radix sort.  Counts / sum w / sum w y / sum w y^2 are by-key reductions.  Keys and counts
// are exact integers, hence bit-reproducible; the floating sums differ from a sequential group-by only by
// summation order.
#include <vector>

#include "common.cuh"
#include "segreduce.cuh"

int64_t factorize_dispatch(lfe_ctx* h, const void* keys, int key_dtype, int64_t n, int order, int32_t* codes,
                           void* uniques, int64_t uniques_cap, cudaStream_t s);

constexpr int CP_MAXK = 32;
constexpr unsigned long long CP_STAGE_SPACE = 1ull << 26;

struct KeyCols {
  // u8 columns first (n_u8 of them), int32 code columns after; mult[j] = product of the radices of the
  // columns that FOLLOW column j in the caller's order, so key = sum_j value_j * mult[j] in any order
  const void* col[CP_MAXK + 1];
  unsigned long long mult[CP_MAXK + 1];
  int n_u8, K;
};

constexpr int FR = 16;          // rows per thread
constexpr int FW = 32 * FR;     // rows per warp tile

// Folded keys of one warp tile of FW consecutive rows.  Lane l owns rows 128 e + 4 l + i (e, i = 0..3; key index
// 4 e + i), so that every load and store of the warp is one fully coalesced request: four 32-bit loads per lane
// from a uint8 column (128 B per warp each), four 128-bit loads from an int32 column (512 B per warp each).
// KT = unsigned (direct-address stages, key space <= 2^26) or unsigned long long.  The loads of four uint8
// columns are issued before they are consumed.
template <typename KT>
__device__ __forceinline__ void acc_u8(KT key[FR], const unsigned w[4], KT m) {
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int i = 0; i < 4; ++i) key[4 * e + i] += (KT)((w[e] >> (8 * i)) & 0xffu) * m;
}
__device__ __forceinline__ void ld_u8(const void* col, int64_t base, int lane, unsigned w[4]) {
  const unsigned* p = reinterpret_cast<const unsigned*>(col) + base / 4 + lane;
#pragma unroll
  for (int e = 0; e < 4; ++e) w[e] = __ldcs(p + 32 * e);
}
template <typename KT>
__device__ __forceinline__ void fold_tile(const KeyCols& kc, int64_t base, int64_t n, bool vec, int lane, KT key[FR]) {
#pragma unroll
  for (int e = 0; e < FR; ++e) key[e] = 0;
  if (vec && base + FW <= n) {
    int j = 0;
    for (; j + 4 <= kc.n_u8; j += 4) {
      unsigned w0[4], w1[4], w2[4], w3[4];
      ld_u8(kc.col[j], base, lane, w0);
      ld_u8(kc.col[j + 1], base, lane, w1);
      ld_u8(kc.col[j + 2], base, lane, w2);
      ld_u8(kc.col[j + 3], base, lane, w3);
      acc_u8<KT>(key, w0, (KT)kc.mult[j]);
      acc_u8<KT>(key, w1, (KT)kc.mult[j + 1]);
      acc_u8<KT>(key, w2, (KT)kc.mult[j + 2]);
      acc_u8<KT>(key, w3, (KT)kc.mult[j + 3]);
    }
    for (; j < kc.n_u8; ++j) {
      unsigned w0[4];
      ld_u8(kc.col[j], base, lane, w0);
      acc_u8<KT>(key, w0, (KT)kc.mult[j]);
    }
    for (; j < kc.K; ++j) {
      const int4* p = reinterpret_cast<const int4*>(kc.col[j]) + base / 4 + lane;
      int4 a[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] = __ldcs(p + 32 * e);
      const KT m = (KT)kc.mult[j];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        key[4 * e] += (KT)(unsigned)a[e].x * m;
        key[4 * e + 1] += (KT)(unsigned)a[e].y * m;
        key[4 * e + 2] += (KT)(unsigned)a[e].z * m;
        key[4 * e + 3] += (KT)(unsigned)a[e].w * m;
      }
    }
  } else {
    for (int e = 0; e < FR; ++e) {
      const int64_t r = base + 128 * (e >> 2) + 4 * lane + (e & 3);
      if (r >= n) continue;
      KT k = 0;
      for (int j = 0; j < kc.K; ++j) {
        unsigned v = j < kc.n_u8 ? (unsigned)reinterpret_cast<const uint8_t*>(kc.col[j])[r]
                                 : (unsigned)reinterpret_cast<const int32_t*>(kc.col[j])[r];
        k += (KT)v * (KT)kc.mult[j];
      }
      key[e] = k;
    }
  }
}
#define FOLD_TILE_LOOP                                                                          \
  const int lane = threadIdx.x & 31;                                                            \
  const int64_t ntile = ceil_div64(n, FW);                                                      \
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;                               \
  for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < ntile; t += wstride)
__device__ __forceinline__ int64_t tile_row(int64_t base, int lane, int e) { return base + 128 * (e >> 2) + 4 * lane + (e & 3); }

// sort-path fallback (key space too large for a presence table): materialise the 64-bit keys
__global__ void __launch_bounds__(256) k_fold_keys(KeyCols kc, int64_t n, bool vec, uint64_t* __restrict__ out) {
  FOLD_TILE_LOOP {
    unsigned long long key[FR];
    fold_tile<unsigned long long>(kc, t * FW, n, vec, lane, key);
#pragma unroll
    for (int e = 0; e < FR; ++e) {
      const int64_t r = tile_row(t * FW, lane, e);
      if (r < n) out[r] = key[e];
    }
  }
}
// direct-address stage, pass 1: presence flags of the folded keys (the keys are never written out)
__global__ void __launch_bounds__(256) k_fold_mark(KeyCols kc, int64_t n, bool vec, uint32_t* __restrict__ flags) {
  FOLD_TILE_LOOP {
    unsigned key[FR];
    fold_tile<unsigned>(kc, t * FW, n, vec, lane, key);
    // test before set: with few distinct keys and many rows, unconditional stores serialise on a few L2 lines
    // (a stale 0 from L1 only costs a redundant store; every writer stores 1)
#pragma unroll
    for (int e = 0; e < FR; ++e)
      if (tile_row(t * FW, lane, e) < n && flags[key[e]] == 0u) flags[key[e]] = 1u;
  }
}
// pass 2: recompute the key, look up its dense sorted rank
__global__ void __launch_bounds__(256) k_fold_lookup(KeyCols kc, int64_t n, bool vec, const uint32_t* __restrict__ rank,
                                                     int32_t* __restrict__ gcode) {
  FOLD_TILE_LOOP {
    unsigned key[FR];
    const int64_t base = t * FW;
    fold_tile<unsigned>(kc, base, n, vec, lane, key);
    if (vec && base + FW <= n) {
      int r[FR];
#pragma unroll
      for (int e = 0; e < FR; ++e) r[e] = (int)__ldg(rank + key[e]);
      int4* o = reinterpret_cast<int4*>(gcode) + base / 4 + lane;
#pragma unroll
      for (int e = 0; e < 4; ++e) o[32 * e] = make_int4(r[4 * e], r[4 * e + 1], r[4 * e + 2], r[4 * e + 3]);
    } else {
      for (int e = 0; e < FR; ++e) {
        const int64_t r = tile_row(base, lane, e);
        if (r < n) gcode[r] = (int)rank[key[e]];
      }
    }
  }
}
__global__ void __launch_bounds__(256) k_emit_stage_uniques(const uint32_t* __restrict__ flags,
                                                            const uint32_t* __restrict__ rank, int64_t range,
                                                            uint64_t* __restrict__ uniq) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < range; i += (int64_t)gridDim.x * blockDim.x)
    if (flags[i]) uniq[rank[i]] = (uint64_t)i;
}

// One decode step on the G final groups: key = uniq[code_in[g]] -> digits.  Digit 0 may be the code of the
// previous stage (written to prev_code_out), the others are original columns (written to their outputs).
struct KeyOut {
  int32_t* code[CP_MAXK + 1];   // nullptr = discard
  unsigned long long radix[CP_MAXK + 1];
  int K;
};
__global__ void __launch_bounds__(256) k_unfold_keys(const uint64_t* __restrict__ uniq,
                                                     const int32_t* __restrict__ code_in, int64_t G, KeyOut ko) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long key = uniq[code_in ? code_in[g] : g];
    for (int j = ko.K - 1; j >= 0; --j) {
      if (ko.code[j]) ko.code[j][g] = (int32_t)(key % ko.radix[j]);
      key /= ko.radix[j];
    }
  }
}

// weighted: stats[0*G + g] += w, stats[1*G + g] += w*y, stats[2*G + g] += w*y*y
// unweighted: stats[0*G + g] += y, stats[1*G + g] += y*y   (n_g is the exact integer count of the histogram)
template <bool WEIGHTED>
__global__ void __launch_bounds__(PJ_T) k_group_stats(const int32_t* __restrict__ key, const double* __restrict__ y,
                                                      const double* __restrict__ w, int64_t n,
                                                      double* __restrict__ stats, int64_t G, bool vec_codes) {
  constexpr int NS = WEIGHTED ? 3 : 2;
  const int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I);
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    const int64_t base = (ch * PJ_T + threadIdx.x) * PJ_I;
    int k[PJ_I];
    load4_i32(key, base, n, vec_codes, k);
    double v[PJ_I][NS];
#pragma unroll
    for (int e = 0; e < PJ_I; ++e) {
      bool valid = base + e < n;
      double yy = valid ? __ldcs(y + base + e) : 0.0;
      if (WEIGHTED) {
        double ww = valid ? __ldcs(w + base + e) : 0.0;
        v[e][0] = ww;
        v[e][1] = ww * yy;
        v[e][2] = ww * yy * yy;
      } else {
        v[e][0] = yy;
        v[e][1] = yy * yy;
      }
    }
    segmented_flush<NS>(k, v, stats, 1, 0, G);
  }
}

// Mid-size group counts (a few 10^4 groups, e.g. the 24 000 cells of BASELINE config 4) with unsorted keys: one RED
// per row and statistic into L2 runs at ~170 G atomics/s.  Instead, one 1024-thread CTA per SM privatises a window
// of [g0, g0 + B) groups x NS statistics in shared memory (up to 227 KB) and the rows are streamed once per window.
template <bool WEIGHTED>
__global__ void __launch_bounds__(1024, 1) k_group_stats_smem(const int32_t* __restrict__ key, const double* __restrict__ y,
                                                              const double* __restrict__ w, int64_t n,
                                                              double* __restrict__ stats, int64_t G, int g0, int B) {
  constexpr int NS = WEIGHTED ? 3 : 2;
  extern __shared__ double bins[];   // [NS][B]
  for (int i = threadIdx.x; i < NS * B; i += blockDim.x) bins[i] = 0.0;
  __syncthreads();
  const int64_t nv = n / 4;
  const int4* kv = reinterpret_cast<const int4*>(key);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nv; q += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = __ldcs(kv + q);
    const double2 ya = __ldcs(reinterpret_cast<const double2*>(y) + 2 * q);
    const double2 yb = __ldcs(reinterpret_cast<const double2*>(y) + 2 * q + 1);
    const int cc[4] = {c.x - g0, c.y - g0, c.z - g0, c.w - g0};
    const double yy[4] = {ya.x, ya.y, yb.x, yb.y};
    double ww[4] = {1.0, 1.0, 1.0, 1.0};
    if (WEIGHTED) {
      const double2 wa = __ldcs(reinterpret_cast<const double2*>(w) + 2 * q);
      const double2 wb = __ldcs(reinterpret_cast<const double2*>(w) + 2 * q + 1);
      ww[0] = wa.x; ww[1] = wa.y; ww[2] = wb.x; ww[3] = wb.y;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if ((unsigned)cc[e] < (unsigned)B) {
        if (WEIGHTED) {
          atomicAdd(&bins[cc[e]], ww[e]);
          atomicAdd(&bins[B + cc[e]], ww[e] * yy[e]);
          atomicAdd(&bins[2 * B + cc[e]], ww[e] * yy[e] * yy[e]);
        } else {
          atomicAdd(&bins[cc[e]], yy[e]);
          atomicAdd(&bins[B + cc[e]], yy[e] * yy[e]);
        }
      }
  }
  if (blockIdx.x == 0)
    for (int64_t i = nv * 4 + threadIdx.x; i < n; i += blockDim.x) {
      const int c = key[i] - g0;
      if ((unsigned)c < (unsigned)B) {
        const double yv = y[i], wv = WEIGHTED ? w[i] : 1.0;
        if (WEIGHTED) {
          atomicAdd(&bins[c], wv);
          atomicAdd(&bins[B + c], wv * yv);
          atomicAdd(&bins[2 * B + c], wv * yv * yv);
        } else {
          atomicAdd(&bins[c], yv);
          atomicAdd(&bins[B + c], yv * yv);
        }
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < NS * B; i += blockDim.x) {
    const int q = i / B, g = i - q * B;
    if (g0 + g < G && bins[i] != 0.0) atomicAdd(&stats[(size_t)q * G + g0 + g], bins[i]);
  }
}

// Re-aggregation of partial group tables (multi-GPU YOCO): stats[q*G + g] += part_q[i], q = n, sum_y, sum_y_sq, cnt
__global__ void __launch_bounds__(PJ_T) k_group_merge(const int32_t* __restrict__ key, const double* __restrict__ pn,
                                                      const double* __restrict__ psy, const double* __restrict__ psyy,
                                                      const int32_t* __restrict__ pcnt, int64_t n,
                                                      double* __restrict__ stats, int64_t G, bool vec_codes) {
  const int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I);
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    const int64_t base = (ch * PJ_T + threadIdx.x) * PJ_I;
    int k[PJ_I];
    load4_i32(key, base, n, vec_codes, k);
    double v[PJ_I][4];
#pragma unroll
    for (int e = 0; e < PJ_I; ++e) {
      bool valid = base + e < n;
      v[e][0] = valid ? pn[base + e] : 0.0;
      v[e][1] = valid ? psy[base + e] : 0.0;
      v[e][2] = valid ? psyy[base + e] : 0.0;
      v[e][3] = valid ? (double)pcnt[base + e] : 0.0;   // row counts < 2^31: exact in fp64
    }
    segmented_flush<4>(k, v, stats, 1, 0, G);
  }
}
__global__ void __launch_bounds__(256) k_f64_to_i32(const double* __restrict__ v, int64_t n, int32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int32_t)llrint(v[i]);
}

void prim_histogram(lfe_ctx* h, const int32_t* codes, int64_t n, int64_t G, uint32_t* counts, cudaStream_t s);
void comm_allreduce(lfe_ctx* h, double* buf, int64_t count, int op_max, cudaStream_t s);                  // api.cu
void comm_allreduce_int(lfe_ctx* h, void* buf, int64_t count, int kind, int op, cudaStream_t s);         // api.cu

// partial tables to re-aggregate instead of raw (y, weights) rows
struct MergeIn {
  const double *n, *sum_y, *sum_y_sq;
  const int32_t* cnt;
};

struct Stage {
  std::vector<int> cols;                 // original column indices folded in this stage (after the previous stage's code)
  bool has_prev = false;                 // digit 0 = code of the previous stage
  unsigned long long prev_radix = 1;
  DevBuf<uint64_t> uniq;                 // sorted distinct folded keys of this stage
  int64_t G = 0;
};

// deterministic mode: stats[q * G + g] = sum over the rows of group g, in row order, of (w | w y | w y^2): one warp per
// group walks its rows of the group-sorted permutation, fixed shuffle tree, plain stores (no floating-point atomics)
template <bool WEIGHTED>
__global__ void __launch_bounds__(256) k_det_group_stats(const uint32_t* __restrict__ off, int64_t G,
                                                         const uint32_t* __restrict__ perm, const double* __restrict__ y,
                                                         const double* __restrict__ w, double* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < G; g += nwarp) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const uint32_t e1 = off[g + 1];
    for (uint32_t e = off[g] + lane; e < e1; e += 32) {
      const size_t row = perm[e];
      const double yv = y[row];
      if (WEIGHTED) {
        const double wv = w[row];
        s0 += wv; s1 += wv * yv; s2 += wv * yv * yv;
      } else {
        s1 += yv; s2 += yv * yv;
      }
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) {
      if (WEIGHTED) { stats[g] = s0; stats[(size_t)G + g] = s1; stats[2 * (size_t)G + g] = s2; }
      else { stats[g] = s1; stats[(size_t)G + g] = s2; }
    }
  }
}

static void compress_core(lfe_ctx* h, const void* const* key_codes, const int* key_dtypes, const int64_t* n_groups, int K,
                          const double* y,
                          const double* weights, const MergeIn* merge, int64_t n, int64_t out_cap,
                          int32_t* const* key_codes_out, double* n_out, int32_t* cnt_out, double* sum_y_out,
                          double* sum_y_sq_out, int64_t* n_compressed_out, cudaStream_t s) {
  LFE_REQUIRE(K >= 0 && K <= CP_MAXK, "lfe_compress: K=%d out of range [0,%d]", K, CP_MAXK);
  LFE_REQUIRE(n >= 0 && n < (int64_t)0x7fffffff, "lfe_compress: n out of range");
  if (n == 0) {
    *n_compressed_out = 0;
    return;
  }
  bool vec = true;   // vector loads need 4-row-aligned columns (always true for whole torch tensors)
  for (int j = 0; j < K; ++j) {
    LFE_REQUIRE(n_groups[j] >= 1, "lfe_compress: n_groups[%d] must be >= 1", j);
    const int dt = key_dtypes ? key_dtypes[j] : LFE_I32;
    LFE_REQUIRE(dt == LFE_I32 || dt == LFE_U8, "lfe_compress: key column %d: dtype must be LFE_I32 or LFE_U8", j);
    LFE_REQUIRE(dt != LFE_U8 || n_groups[j] <= 256, "lfe_compress: uint8 key column %d needs n_groups <= 256", j);
    vec = vec && (((uintptr_t)key_codes[j]) & (dt == LFE_U8 ? 3 : 15)) == 0;   // 32-bit / 128-bit loads
  }
  const int blocks = launch_blocks(h, ceil_div64(n, 256 * FR), 8);
  const bool want_keys = key_codes_out != nullptr;

  // ---- staged folding -------------------------------------------------------------------------------
  std::vector<Stage> stages;
  DevBuf<uint64_t> folded;   // only the sort fallback materialises the folded keys
  DevBuf<int32_t> gcode(n, s), gcode_prev;
  int64_t G = 1;
  int j = 0;
  bool first = true;
  do {
    stages.emplace_back();
    Stage& st = stages.back();
    // digits of this stage in significance order: [previous stage's dense code,] then the new columns
    std::vector<const void*> dcol;
    std::vector<unsigned long long> dradix;
    std::vector<bool> du8;
    unsigned long long space = 1;
    if (!first) {   // carry the previous stage's dense code as the most significant digit
      st.has_prev = true;
      st.prev_radix = (unsigned long long)G;
      std::swap(gcode, gcode_prev);
      if (!gcode.p) gcode.alloc(n, s);
      dcol.push_back(gcode_prev.p);
      dradix.push_back((unsigned long long)G);
      du8.push_back(false);
      space = (unsigned long long)G;
    }
    while (j < K) {
      unsigned long long r = (unsigned long long)n_groups[j];
      bool fits = space <= CP_STAGE_SPACE / r || (dcol.size() == (st.has_prev ? 1u : 0u));   // always take at least one column
      if (!fits) break;
      LFE_REQUIRE(space <= (~0ull) / r, "lfe_compress: composite key space overflows 64 bits");
      dcol.push_back(key_codes[j]);
      dradix.push_back(r);
      du8.push_back(key_dtypes && key_dtypes[j] == LFE_U8);
      space *= r;
      st.cols.push_back(j++);
    }
    KeyCols kc;
    kc.K = 0;
    {
      const int D = (int)dcol.size();
      std::vector<unsigned long long> mult(D, 1ull);
      for (int d = D - 2; d >= 0; --d) mult[d] = mult[d + 1] * dradix[d + 1];
      for (int pass = 0; pass < 2; ++pass)            // uint8 columns first, int32 columns after
        for (int d = 0; d < D; ++d)
          if (du8[d] == (pass == 0)) {
            kc.col[kc.K] = dcol[d];
            kc.mult[kc.K++] = mult[d];
          }
      kc.n_u8 = 0;
      for (int d = 0; d < D; ++d) kc.n_u8 += du8[d] ? 1 : 0;
    }
    // distinct keys of this stage never exceed n; cap the uniques buffer by the key space
    int64_t ucap = (int64_t)((space < (unsigned long long)n) ? space : (unsigned long long)n);
    if (want_keys) st.uniq.alloc((size_t)ucap, s);   // only the decode step below needs the stage's distinct keys
    const unsigned long long direct_limit = (unsigned long long)(16 * n > (1 << 20) ? 16 * n : (1 << 20));
    if (space <= CP_STAGE_SPACE && space <= direct_limit) {
      // direct-address stage: the folded key is recomputed in both passes instead of being written out
      DevBuf<uint32_t> flags((size_t)space, s), rank((size_t)space, s), total(1, s);
      flags.zero();
      k_fold_mark<<<blocks, 256, 0, s>>>(kc, n, vec, flags);
      KERNEL_CHECK();
      prim_exclusive_scan_u32(h, flags, rank, (int64_t)space, total, s);
      k_fold_lookup<<<blocks, 256, 0, s>>>(kc, n, vec && (((uintptr_t)gcode.p) & 15) == 0, rank, gcode);
      KERNEL_CHECK();
      uint32_t ht = 0;
      CUDA_CHECK(cudaMemcpyAsync(&ht, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      G = ht;
      if (want_keys) {
        k_emit_stage_uniques<<<launch_blocks(h, ceil_div64((int64_t)space, 1024), 8), 256, 0, s>>>(flags, rank, (int64_t)space,
                                                                                                    st.uniq);
        KERNEL_CHECK();
      }
      h->n_launches += 3;
    } else {
      if (!folded.p) folded.alloc(n, s);
      k_fold_keys<<<blocks, 256, 0, s>>>(kc, n, vec, folded);
      KERNEL_CHECK();
      h->n_launches += 1;
      G = factorize_dispatch(h, folded, LFE_U64, n, LFE_ORDER_SORTED, gcode, st.uniq.p, ucap, s);
    }
    st.G = G;
    first = false;
  } while (j < K);

  *n_compressed_out = G;
  if (G > out_cap) {
    if (key_codes_out || cnt_out || n_out) LFE_THROW(LFE_ERR_CAPACITY, "lfe_compress: %lld groups exceed out_cap %lld", (long long)G, (long long)out_cap);
    return;
  }
  // ---- decode the final groups back to per-column codes, last stage first --------------------------------
  if (want_keys) {
    DevBuf<int32_t> ca(G, s), cb(G, s);
    const int32_t* code_in = nullptr;   // nullptr: identity (final groups are the last stage's ranks)
    int32_t* code_out = ca;
    for (int si = (int)stages.size() - 1; si >= 0; --si) {
      Stage& st = stages[si];
      KeyOut ko;
      ko.K = 0;
      if (st.has_prev) {
        ko.code[ko.K] = code_out;
        ko.radix[ko.K++] = st.prev_radix;
      }
      for (int c : st.cols) {
        ko.code[ko.K] = key_codes_out[c];
        ko.radix[ko.K++] = (unsigned long long)n_groups[c];
      }
      k_unfold_keys<<<launch_blocks(h, ceil_div64(G, 256), 8), 256, 0, s>>>(st.uniq, code_in, G, ko);
      KERNEL_CHECK();
      h->n_launches += 1;
      code_in = code_out;
      code_out = (code_out == ca.p) ? cb.p : ca.p;
    }
  }
  if (merge) {
    DevBuf<double> stats((size_t)4 * G, s);
    stats.zero();
    int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I);
    k_group_merge<<<launch_blocks(h, nchunk, 8), PJ_T, 0, s>>>(gcode, merge->n, merge->sum_y, merge->sum_y_sq, merge->cnt,
                                                                 n, stats, G, (((uintptr_t)gcode.p) & 15) == 0);
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(n_out, stats.p, G * sizeof(double), cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(sum_y_out, stats.p + G, G * sizeof(double), cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(sum_y_sq_out, stats.p + 2 * G, G * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (cnt_out) {
      k_f64_to_i32<<<launch_blocks(h, ceil_div64(G, 256), 8), 256, 0, s>>>(stats.p + 3 * G, G, cnt_out);
      KERNEL_CHECK();
    }
    h->n_launches += 2;
    return;
  }
  const bool want_stats = y && n_out && sum_y_out && sum_y_sq_out;
  DevBuf<uint32_t> cnt_tmp;
  uint32_t* cnt = (uint32_t*)cnt_out;
  if (!cnt && want_stats && !weights) {
    cnt_tmp.alloc((size_t)G, s);
    cnt = cnt_tmp.p;
  }
  if (cnt) prim_histogram(h, gcode, n, G, cnt, s);
  if (want_stats) {
    const int NS = weights ? 3 : 2;
    DevBuf<double> stats((size_t)NS * G, s);
    stats.zero();
    int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I);
    const bool vc = (((uintptr_t)gcode.p) & 15) == 0;
    // shared-memory windows when a few of them cover all groups and there are many rows per group
    const int win = (int)(((size_t)h->smem_optin - 1024) / (NS * sizeof(double)));
    const int64_t npass = win > 0 ? ceil_div64(G, win) : 1 << 30;
    const bool aligned = vc && (((uintptr_t)y) & 15) == 0 && (!weights || (((uintptr_t)weights) & 15) == 0);
    if (h->opt_deterministic) {
      DevBuf<uint32_t> perm(n, s), off((size_t)G + 1, s);
      off.zero();
      prim_histogram(h, gcode, n, G, off, s);
      prim_exclusive_scan_u32(h, off, off, G + 1, nullptr, s);
      prim_argsort_u32(h, (const uint32_t*)gcode.p, n, bits_for((uint64_t)(G > 0 ? G - 1 : 0)), perm, nullptr, s);
      const int db = launch_blocks(h, ceil_div64(G, 8), 8);
      if (weights) k_det_group_stats<true><<<db, 256, 0, s>>>(off, G, perm, y, weights, stats);
      else k_det_group_stats<false><<<db, 256, 0, s>>>(off, G, perm, y, nullptr, stats);
      KERNEL_CHECK();
      h->n_launches += 1;
    } else if (G > 4096 && npass <= 4 && n >= 64 * G && aligned) {
      const int B = (int)ceil_div64(G, npass);
      const size_t smem = (size_t)NS * B * sizeof(double);
      static bool attr_set = false;
      if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(k_group_stats_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
        CUDA_CHECK(cudaFuncSetAttribute(k_group_stats_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
        attr_set = true;
      }
      for (int64_t ps = 0; ps < npass; ++ps) {
        if (weights) k_group_stats_smem<true><<<h->sm_count, 1024, smem, s>>>(gcode, y, weights, n, stats, G, (int)(ps * B), B);
        else k_group_stats_smem<false><<<h->sm_count, 1024, smem, s>>>(gcode, y, nullptr, n, stats, G, (int)(ps * B), B);
        KERNEL_CHECK();
      }
      h->n_launches += npass;
    } else {
      if (weights) k_group_stats<true><<<launch_blocks(h, nchunk, 8), PJ_T, 0, s>>>(gcode, y, weights, n, stats, G, vc);
      else k_group_stats<false><<<launch_blocks(h, nchunk, 8), PJ_T, 0, s>>>(gcode, y, nullptr, n, stats, G, vc);
      KERNEL_CHECK();
      h->n_launches += 1;
    }
    const double* sy = stats.p + (weights ? G : 0);
    if (weights) CUDA_CHECK(cudaMemcpyAsync(n_out, stats.p, G * sizeof(double), cudaMemcpyDeviceToDevice, s));
    else prim_u32_to_f64(h, cnt, G, n_out, s);   // unweighted: _n is the exact integer group size
    CUDA_CHECK(cudaMemcpyAsync(sum_y_out, sy, G * sizeof(double), cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(sum_y_sq_out, sy + G, G * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }
}

void compress_impl(lfe_ctx* h, const void* const* key_codes, const int* key_dtypes, const int64_t* n_groups, int K,
                   const double* y, const double* weights, int64_t n, int64_t out_cap, int32_t* const* key_codes_out,
                   double* n_out, int32_t* cnt_out, double* sum_y_out, double* sum_y_sq_out, int64_t* n_compressed_out,
                   cudaStream_t s) {
  PhaseTimer t(h, s, 12);
  compress_core(h, key_codes, key_dtypes, n_groups, K, y, weights, nullptr, n, out_cap, key_codes_out, n_out, cnt_out, sum_y_out,
                sum_y_sq_out, n_compressed_out, s);
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU YOCO: every rank compresses its own rows (no data-path collective over the N rows), then the
// partial group tables -- G_r rows each, tiny next to N -- are all-gathered and re-aggregated, so every
// rank ends up with the identical global table.  The gather is a sum all-reduce over a zero-initialised
// buffer in which each rank fills its own slice (codes and counts as uint32, statistics as fp64).
// Without a communicator this merges duplicate keys inside the one table it is given.
// ---------------------------------------------------------------------------------------------
void compress_merge_impl(lfe_ctx* h, const int32_t* const* key_codes, const int64_t* n_groups, int K,
                         const double* n_in, const int32_t* cnt_in, const double* sum_y_in, const double* sum_y_sq_in,
                         int64_t n_rows, int64_t out_cap, int32_t* const* key_codes_out, double* n_out,
                         int32_t* cnt_out, double* sum_y_out, double* sum_y_sq_out, int64_t* n_compressed_out,
                         cudaStream_t s) {
  PhaseTimer t(h, s, 12);
  LFE_REQUIRE(K >= 1 && K <= CP_MAXK, "lfe_compress_merge: K=%d out of range [1,%d]", K, CP_MAXK);
  LFE_REQUIRE(n_rows >= 0 && n_in && cnt_in && sum_y_in && sum_y_sq_in && n_out && sum_y_out && sum_y_sq_out &&
                  key_codes_out, "lfe_compress_merge: NULL argument");
  if (h->world <= 1) {
    MergeIn m{n_in, sum_y_in, sum_y_sq_in, cnt_in};
    compress_core(h, (const void* const*)key_codes, nullptr, n_groups, K, nullptr, nullptr, &m, n_rows, out_cap, key_codes_out, n_out, cnt_out,
                  sum_y_out, sum_y_sq_out, n_compressed_out, s);
    return;
  }
  const int W = h->world;
  DevBuf<double> sizes(W, s);
  sizes.zero();
  const double mine = (double)n_rows;
  CUDA_CHECK(cudaMemcpyAsync(sizes.p + h->rank, &mine, sizeof(double), cudaMemcpyHostToDevice, s));
  comm_allreduce(h, sizes.p, W, 0, s);
  std::vector<double> hs(W);
  CUDA_CHECK(cudaMemcpyAsync(hs.data(), sizes.p, W * sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  int64_t total = 0, off = 0;
  for (int r = 0; r < W; ++r) {
    if (r == h->rank) off = total;
    total += (int64_t)hs[r];
  }
  LFE_REQUIRE(total < (int64_t)0x7fffffff, "lfe_compress_merge: gathered table too large");
  if (total == 0) {
    *n_compressed_out = 0;
    return;
  }
  DevBuf<uint32_t> gi((size_t)(K + 1) * total, s);   // K code columns + raw counts
  DevBuf<double> gd((size_t)3 * total, s);           // n, sum_y, sum_y_sq
  gi.zero();
  gd.zero();
  if (n_rows > 0) {
    for (int j = 0; j < K; ++j)
      CUDA_CHECK(cudaMemcpyAsync(gi.p + (size_t)j * total + off, key_codes[j], n_rows * sizeof(int32_t),
                                 cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(gi.p + (size_t)K * total + off, cnt_in, n_rows * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(gd.p + off, n_in, n_rows * sizeof(double), cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(gd.p + total + off, sum_y_in, n_rows * sizeof(double), cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(gd.p + 2 * total + off, sum_y_sq_in, n_rows * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }
  comm_allreduce_int(h, gi.p, (int64_t)(K + 1) * total, 0, 0, s);
  comm_allreduce(h, gd.p, 3 * total, 0, s);
  std::vector<const void*> kc(K);
  for (int j = 0; j < K; ++j) kc[j] = gi.p + (size_t)j * total;
  MergeIn m{gd.p, gd.p + total, gd.p + 2 * total, (const int32_t*)(gi.p + (size_t)K * total)};
  compress_core(h, kc.data(), nullptr, n_groups, K, nullptr, nullptr, &m, total, out_cap, key_codes_out, n_out, cnt_out,
                sum_y_out, sum_y_sq_out, n_compressed_out, s);
}
