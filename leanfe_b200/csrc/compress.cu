// compress.cu -- YOCO compression: device group-by producing sufficient statistics.
// Reference: python/leanfe/compress.py:325-350 (group_by(x + fe + cluster).agg(len, sum y, sum y^2)) and
// the distinct count of estimate_compression_ratio (compress.py:249-250).
//
// The K code columns are folded into a mixed-radix key (most significant = first column, so ascending key
// order == ascending lexicographic order of the code tuples) and the key is factorised.  Folding is STAGED:
// columns are accumulated while the product of their cardinalities stays below 2^26, then that partial key
// is replaced by its dense sorted rank (direct-address table, three streaming passes) before more columns
// are folded in.  Ranks are order-preserving, so the final group order is still lexicographic, and the
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
  const int32_t* code[CP_MAXK + 1];
  unsigned long long radix[CP_MAXK + 1];
  int K;
};

__global__ void __launch_bounds__(256) k_fold_keys(KeyCols kc, int64_t n, uint64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long key = 0;
    for (int j = 0; j < kc.K; ++j) key = key * kc.radix[j] + (unsigned long long)__ldcs(kc.code[j] + i);
    out[i] = key;
  }
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

// stats[0*G + g] += w, stats[1*G + g] += w*y, stats[2*G + g] += w*y*y
__global__ void __launch_bounds__(PJ_T) k_group_stats(const int32_t* __restrict__ key, const double* __restrict__ y,
                                                      const double* __restrict__ w, int64_t n,
                                                      double* __restrict__ stats, int64_t G, bool vec_codes) {
  const int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I);
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    const int64_t base = (ch * PJ_T + threadIdx.x) * PJ_I;
    int k[PJ_I];
    load4_i32(key, base, n, vec_codes, k);
    double v[PJ_I][3];
#pragma unroll
    for (int e = 0; e < PJ_I; ++e) {
      bool valid = base + e < n;
      double yy = valid ? __ldcs(y + base + e) : 0.0;
      double ww = valid ? (w ? __ldcs(w + base + e) : 1.0) : 0.0;
      v[e][0] = ww;
      v[e][1] = ww * yy;
      v[e][2] = ww * yy * yy;
    }
    segmented_flush<3>(k, v, stats, 1, 0, G);
  }
}

void prim_histogram(lfe_ctx* h, const int32_t* codes, int64_t n, int64_t G, uint32_t* counts, cudaStream_t s);

struct Stage {
  std::vector<int> cols;                 // original column indices folded in this stage (after the previous stage's code)
  bool has_prev = false;                 // digit 0 = code of the previous stage
  unsigned long long prev_radix = 1;
  DevBuf<uint64_t> uniq;                 // sorted distinct folded keys of this stage
  int64_t G = 0;
};

void compress_impl(lfe_ctx* h, const int32_t* const* key_codes, const int64_t* n_groups, int K, const double* y,
                   const double* weights, int64_t n, int64_t out_cap, int32_t* const* key_codes_out, double* n_out,
                   int32_t* cnt_out, double* sum_y_out, double* sum_y_sq_out, int64_t* n_compressed_out,
                   cudaStream_t s) {
  PhaseTimer t(h, s, 12);
  LFE_REQUIRE(K >= 0 && K <= CP_MAXK, "lfe_compress: K=%d out of range [0,%d]", K, CP_MAXK);
  LFE_REQUIRE(n >= 0 && n < (int64_t)0x7fffffff, "lfe_compress: n out of range");
  if (n == 0) {
    *n_compressed_out = 0;
    return;
  }
  for (int j = 0; j < K; ++j) LFE_REQUIRE(n_groups[j] >= 1, "lfe_compress: n_groups[%d] must be >= 1", j);
  const int blocks = launch_blocks(h, ceil_div64(n, 1024), 8);
  const bool want_keys = key_codes_out != nullptr;

  // ---- staged folding -------------------------------------------------------------------------------
  std::vector<Stage> stages;
  DevBuf<uint64_t> folded(n, s);
  DevBuf<int32_t> gcode(n, s), gcode_prev;
  int64_t G = 1;
  int j = 0;
  bool first = true;
  do {
    stages.emplace_back();
    Stage& st = stages.back();
    KeyCols kc;
    kc.K = 0;
    unsigned long long space = 1;
    if (!first) {   // carry the previous stage's dense code as the most significant digit
      st.has_prev = true;
      st.prev_radix = (unsigned long long)G;
      std::swap(gcode, gcode_prev);
      if (!gcode.p) gcode.alloc(n, s);
      kc.code[kc.K] = gcode_prev;
      kc.radix[kc.K++] = (unsigned long long)G;
      space = (unsigned long long)G;
    }
    while (j < K) {
      unsigned long long r = (unsigned long long)n_groups[j];
      bool fits = space <= CP_STAGE_SPACE / r || (kc.K == (st.has_prev ? 1 : 0));   // always take at least one column
      if (!fits) break;
      LFE_REQUIRE(space <= (~0ull) / r, "lfe_compress: composite key space overflows 64 bits");
      kc.code[kc.K] = key_codes[j];
      kc.radix[kc.K++] = r;
      space *= r;
      st.cols.push_back(j++);
    }
    k_fold_keys<<<blocks, 256, 0, s>>>(kc, n, folded);
    KERNEL_CHECK();
    h->n_launches += 1;
    // distinct keys of this stage never exceed n; cap the uniques buffer by the key space
    int64_t ucap = (int64_t)((space < (unsigned long long)n) ? space : (unsigned long long)n);
    if (want_keys) st.uniq.alloc((size_t)ucap, s);   // only the decode step below needs the stage's distinct keys
    G = factorize_dispatch(h, folded, LFE_U64, n, LFE_ORDER_SORTED, gcode, st.uniq.p, ucap, s);
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
  if (cnt_out) prim_histogram(h, gcode, n, G, (uint32_t*)cnt_out, s);
  if (y && n_out && sum_y_out && sum_y_sq_out) {
    DevBuf<double> stats((size_t)3 * G, s);
    stats.zero();
    int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I);
    k_group_stats<<<launch_blocks(h, nchunk, 8), PJ_T, 0, s>>>(gcode, y, weights, n, stats, G,
                                                                 (((uintptr_t)gcode.p) & 15) == 0);
    KERNEL_CHECK();
    h->n_launches += 1;
    CUDA_CHECK(cudaMemcpyAsync(n_out, stats.p, G * sizeof(double), cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(sum_y_out, stats.p + G, G * sizeof(double), cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(sum_y_sq_out, stats.p + 2 * G, G * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }
}
