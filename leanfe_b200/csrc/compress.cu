// compress.cu -- YOCO compression: device group-by producing sufficient statistics.
// Reference: python/leanfe/compress.py:325-350 (group_by(x + fe + cluster).agg(len, sum y, sum y^2)) and
// the distinct count of estimate_compression_ratio (compress.py:249-250).
//
// The K code columns are folded into one mixed-radix uint64 key (most significant = first column, so
// ascending key order == ascending lexicographic order of the code tuples), the key is factorised with
// the same direct-table / radix-sort machinery as the FE ids, and counts / sum w / sum w y / sum w y^2
// are by-key reductions.  Keys and counts are exact integers, hence bit-reproducible; the floating
// sums differ from a sequential group-by only by summation order.
#include "common.cuh"
#include "segreduce.cuh"

int64_t factorize_dispatch(lfe_ctx* h, const void* keys, int key_dtype, int64_t n, int order, int32_t* codes,
                           void* uniques, int64_t uniques_cap, cudaStream_t s);

struct KeyCols {
  const int32_t* code[32];
  unsigned long long radix[32];
  int K;
};
constexpr int CP_MAXK = 32;

__global__ void __launch_bounds__(256) k_fold_keys(KeyCols kc, int64_t n, uint64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long key = 0;
    for (int j = 0; j < kc.K; ++j) key = key * kc.radix[j] + (unsigned long long)__ldcs(kc.code[j] + i);
    out[i] = key;
  }
}

struct KeyOut {
  int32_t* code[32];
  unsigned long long radix[32];
  int K;
};
__global__ void __launch_bounds__(256) k_unfold_keys(const uint64_t* __restrict__ uniq, int64_t G, KeyOut ko) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long key = uniq[g];
    for (int j = ko.K - 1; j >= 0; --j) {
      ko.code[j][g] = (int32_t)(key % ko.radix[j]);
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
  KeyCols kc;
  kc.K = K;
  long double space = 1.0L;
  for (int j = 0; j < K; ++j) {
    LFE_REQUIRE(n_groups[j] >= 1, "lfe_compress: n_groups[%d] must be >= 1", j);
    kc.code[j] = key_codes[j];
    kc.radix[j] = (unsigned long long)n_groups[j];
    space *= (long double)n_groups[j];
  }
  LFE_REQUIRE(space < 1.8e19L, "lfe_compress: composite key space overflows 64 bits; pre-combine key columns");
  DevBuf<uint64_t> folded(n, s), uniq(n, s);
  DevBuf<int32_t> gcode(n, s);
  int blocks = launch_blocks(h, ceil_div64(n, 1024), 8);
  k_fold_keys<<<blocks, 256, 0, s>>>(kc, n, folded);
  KERNEL_CHECK();
  h->n_launches += 1;
  int64_t G = factorize_dispatch(h, folded, LFE_U64, n, LFE_ORDER_SORTED, gcode, uniq, n, s);
  *n_compressed_out = G;
  if (G > out_cap) {
    if (key_codes_out || cnt_out || n_out) LFE_THROW(LFE_ERR_CAPACITY, "lfe_compress: %lld groups exceed out_cap %lld", (long long)G, (long long)out_cap);
    return;
  }
  if (key_codes_out) {
    KeyOut ko;
    ko.K = K;
    for (int j = 0; j < K; ++j) {
      ko.code[j] = key_codes_out[j];
      ko.radix[j] = kc.radix[j];
    }
    k_unfold_keys<<<launch_blocks(h, ceil_div64(G, 256), 8), 256, 0, s>>>(uniq, G, ko);
    KERNEL_CHECK();
    h->n_launches += 1;
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
