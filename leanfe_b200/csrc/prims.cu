// prims.cu -- device-wide primitives written for this library: exclusive scan, stable LSD radix
// argsort, histogram, sortedness test, gathers.  All launches size their grids in multiples of the
// SM count (148 on B200) or one tile per CTA for the tiled sort passes.
#include "common.cuh"

int bits_for(uint64_t v) {
  int b = 0;
  while (v) { ++b; v >>= 1; }
  return b;
}

int launch_blocks(lfe_ctx* h, int64_t work_items, int per_sm) {
  int64_t cap = (int64_t)h->sm_count * per_sm;
  if (work_items < 1) work_items = 1;
  return (int)(work_items < cap ? work_items : cap);
}

// ---------------------------------------------------------------------------------------------
// exclusive scan (uint32)
// ---------------------------------------------------------------------------------------------
constexpr int SC_T = 256;
constexpr int SC_I = 8;
constexpr int SC_TILE = SC_T * SC_I;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t ws[SC_T / 32];
  __shared__ uint32_t tot;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) ws[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t x = lane < SC_T / 32 ? ws[lane] : 0;
    uint32_t xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, xi, d);
      if (lane >= d) xi += o;
    }
    if (lane < SC_T / 32) ws[lane] = xi - x;
    if (lane == 31) tot = xi;
  }
  __syncthreads();
  uint32_t r = ws[w] + inc - v;
  *total = tot;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SC_T) k_scan_reduce(const uint32_t* __restrict__ in, int64_t n,
                                                      uint32_t* __restrict__ tile_sums) {
  int64_t base = (int64_t)blockIdx.x * SC_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SC_I; ++i) {
    int64_t idx = base + (int64_t)i * SC_T + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  __shared__ uint32_t ws[SC_T / 32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i = 0; i < SC_T / 32; ++i) t += ws[i];
    tile_sums[blockIdx.x] = t;
  }
}

// scans one tile; thread t owns SC_I consecutive items
__global__ void __launch_bounds__(SC_T) k_scan_tile(const uint32_t* in, uint32_t* out, int64_t n,
                                                    const uint32_t* __restrict__ tile_offsets,
                                                    uint32_t* total_out) {
  int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_I;
  uint32_t v[SC_I];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SC_I; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0u;
    s += v[i];
  }
  uint32_t tot;
  uint32_t ex = block_exclusive_scan(s, &tot);
  uint32_t off = tile_offsets ? tile_offsets[blockIdx.x] : 0u;
  uint32_t run = off + ex;
#pragma unroll
  for (int i = 0; i < SC_I; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
  if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = off + tot;
}

void prim_exclusive_scan_u32(lfe_ctx* h, const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total_dev,
                             cudaStream_t s) {
  if (n <= 0) {
    if (total_dev) CUDA_CHECK(cudaMemsetAsync(total_dev, 0, sizeof(uint32_t), s));
    return;
  }
  int64_t nt = ceil_div64(n, SC_TILE);
  if (nt == 1) {
    k_scan_tile<<<1, SC_T, 0, s>>>(in, out, n, nullptr, total_dev);
    KERNEL_CHECK();
    h->n_launches += 1;
    return;
  }
  DevBuf<uint32_t> sums(nt, s);
  k_scan_reduce<<<(int)nt, SC_T, 0, s>>>(in, n, sums);
  KERNEL_CHECK();
  prim_exclusive_scan_u32(h, sums, sums, nt, nullptr, s);
  k_scan_tile<<<(int)nt, SC_T, 0, s>>>(in, out, n, sums, total_dev);
  KERNEL_CHECK();
  h->n_launches += 2;
}

__global__ void k_iota(uint32_t* p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = (uint32_t)i;
}
void prim_iota_u32(uint32_t* p, int64_t n, cudaStream_t s) {
  if (n <= 0) return;
  int64_t nb = ceil_div64(n, 1024);
  int blocks = (int)(nb < 148 * 8 ? nb : 148 * 8);
  k_iota<<<blocks, 256, 0, s>>>(p, n);
  KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------
// histogram of dense codes
// ---------------------------------------------------------------------------------------------
constexpr int HG_SMEM_BINS = 8192;

__global__ void __launch_bounds__(256) k_hist_smem(const int32_t* __restrict__ codes, int64_t n, int G,
                                                   uint32_t* __restrict__ counts) {
  extern __shared__ uint32_t sh[];
  for (int i = threadIdx.x; i < G; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  int lane = threadIdx.x & 31;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t nround = ceil_div64(n, stride);
  for (int64_t r = 0; r < nround; ++r) {
    int64_t i = r * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < n;
    uint32_t c = valid ? (uint32_t)__ldcs(codes + i) : (0x80000000u + lane);
    // aggregate in the warp only when neighbouring lanes hold equal codes (sorted / clustered rows); for random
    // codes MATCH.ANY finds nothing and costs more than the atomics it would save
    const uint32_t nb = __shfl_down_sync(0xffffffffu, c, 1);
    if (__any_sync(0xffffffffu, lane < 31 && nb == c)) {
      uint32_t peers = __match_any_sync(0xffffffffu, c);
      if (valid && lane == (__ffs(peers) - 1)) atomicAdd(&sh[c], __popc(peers));
    } else if (valid) {
      atomicAdd(&sh[c], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x)
    if (sh[i]) atomicAdd(&counts[i], sh[i]);
}

__global__ void __launch_bounds__(256) k_hist_global(const int32_t* __restrict__ codes, int64_t n,
                                                     uint32_t* __restrict__ counts) {
  int lane = threadIdx.x & 31;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t nround = ceil_div64(n, stride);
  for (int64_t r = 0; r < nround; ++r) {
    int64_t i = r * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < n;
    uint32_t c = valid ? (uint32_t)__ldcs(codes + i) : (0x80000000u + lane);
    const uint32_t nb = __shfl_down_sync(0xffffffffu, c, 1);
    if (__any_sync(0xffffffffu, lane < 31 && nb == c)) {
      uint32_t peers = __match_any_sync(0xffffffffu, c);
      if (valid && lane == (__ffs(peers) - 1)) atomicAdd(&counts[c], __popc(peers));
    } else if (valid) {
      atomicAdd(&counts[c], 1u);
    }
  }
}

// Mid-size histograms (8k < G <= 48k bins, e.g. the 24 000 YOCO groups of BASELINE config 4): one 1024-thread
// CTA per SM privatises all bins in up to 192 KB of shared memory; 128-bit code loads, native shared atomics.
constexpr int HG_BIG_BINS = 48 * 1024;
__global__ void __launch_bounds__(1024, 1) k_hist_smem_big(const int32_t* __restrict__ codes, int64_t n, int G,
                                                           uint32_t* __restrict__ counts) {
  extern __shared__ uint32_t sh[];
  for (int i = threadIdx.x; i < G; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int64_t nv = n / 4;
  const int4* cv = reinterpret_cast<const int4*>(codes);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nv; q += (int64_t)gridDim.x * blockDim.x) {
    int4 c = __ldcs(cv + q);
    atomicAdd(&sh[c.x], 1u);
    atomicAdd(&sh[c.y], 1u);
    atomicAdd(&sh[c.z], 1u);
    atomicAdd(&sh[c.w], 1u);
  }
  if (blockIdx.x == 0)
    for (int64_t i = nv * 4 + threadIdx.x; i < n; i += blockDim.x) atomicAdd(&sh[codes[i]], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x)
    if (sh[i]) atomicAdd(&counts[i], sh[i]);
}

void prim_histogram(lfe_ctx* h, const int32_t* codes, int64_t n, int64_t G, uint32_t* counts, cudaStream_t s) {
  CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)G * sizeof(uint32_t), s));
  if (n <= 0) return;
  int blocks = launch_blocks(h, ceil_div64(n, 256 * 8), 8);
  if (G > HG_SMEM_BINS && G <= HG_BIG_BINS && n >= 256 * G && (((uintptr_t)codes) & 15) == 0 &&
      (size_t)G * sizeof(uint32_t) <= (size_t)h->smem_optin) {
    static bool attr_set = false;
    if (!attr_set) {
      CUDA_CHECK(cudaFuncSetAttribute(k_hist_smem_big, cudaFuncAttributeMaxDynamicSharedMemorySize, HG_BIG_BINS * 4));
      attr_set = true;
    }
    k_hist_smem_big<<<h->sm_count, 1024, (size_t)G * sizeof(uint32_t), s>>>(codes, n, (int)G, counts);
  } else if (G <= HG_SMEM_BINS) {
    k_hist_smem<<<blocks, 256, (size_t)G * sizeof(uint32_t), s>>>(codes, n, (int)G, counts);
  } else {
    k_hist_global<<<blocks, 256, 0, s>>>(codes, n, counts);
  }
  KERNEL_CHECK();
  h->n_launches += 1;
}

// ---------------------------------------------------------------------------------------------
// sortedness
// ---------------------------------------------------------------------------------------------
__global__ void k_is_sorted(const int32_t* __restrict__ codes, int64_t n, int* __restrict__ flag) {
  int bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += (int64_t)gridDim.x * blockDim.x)
    bad |= (codes[i] < codes[i - 1]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *flag = 1;
}

int prim_is_sorted(lfe_ctx* h, const int32_t* codes, int64_t n, cudaStream_t s) {
  if (n <= 1) return 1;
  DevBuf<int> flag(1, s);
  flag.zero();
  int blocks = launch_blocks(h, ceil_div64(n, 256 * 8), 8);
  k_is_sorted<<<blocks, 256, 0, s>>>(codes, n, flag);
  KERNEL_CHECK();
  h->n_launches += 1;
  int hflag = 0;
  CUDA_CHECK(cudaMemcpyAsync(&hflag, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  return hflag == 0;
}

// ---------------------------------------------------------------------------------------------
// gathers: dst[i] = src[perm[i]]   (4 independent loads in flight per thread)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_gather(const T* __restrict__ src, const uint32_t* __restrict__ perm,
                                                int64_t n, T* __restrict__ dst) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint32_t p0 = __ldcs(perm + i), p1 = __ldcs(perm + i + stride), p2 = __ldcs(perm + i + 2 * stride),
             p3 = __ldcs(perm + i + 3 * stride);
    T a = src[p0], b = src[p1], c = src[p2], d = src[p3];
    __stcs(dst + i, a);
    __stcs(dst + i + stride, b);
    __stcs(dst + i + 2 * stride, c);
    __stcs(dst + i + 3 * stride, d);
  }
  for (; i < n; i += stride) dst[i] = src[perm[i]];
}

void prim_gather_f64(const double* src, const uint32_t* perm, int64_t n, double* dst, cudaStream_t s) {
  if (n <= 0) return;
  int64_t nb = ceil_div64(n, 256 * 4);
  int blocks = (int)(nb < 148 * 16 ? nb : 148 * 16);
  k_gather<double><<<blocks, 256, 0, s>>>(src, perm, n, dst);
  KERNEL_CHECK();
}
void prim_gather_i32(const int32_t* src, const uint32_t* perm, int64_t n, int32_t* dst, cudaStream_t s) {
  if (n <= 0) return;
  int64_t nb = ceil_div64(n, 256 * 4);
  int blocks = (int)(nb < 148 * 16 ? nb : 148 * 16);
  k_gather<int32_t><<<blocks, 256, 0, s>>>(src, perm, n, dst);
  KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------
// small conversions
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cvt_u32_f64(const uint32_t* __restrict__ in, int64_t n, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (double)in[i];
}
__global__ void __launch_bounds__(256) k_cvt_f64_u32(const double* __restrict__ in, int64_t n, uint32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (uint32_t)in[i];
}
void prim_u32_to_f64(lfe_ctx* h, const uint32_t* in, int64_t n, double* out, cudaStream_t s) {
  if (n <= 0) return;
  k_cvt_u32_f64<<<launch_blocks(h, ceil_div64(n, 1024), 8), 256, 0, s>>>(in, n, out);
  KERNEL_CHECK();
  h->n_launches += 1;
}
void prim_f64_to_u32(lfe_ctx* h, const double* in, int64_t n, uint32_t* out, cudaStream_t s) {
  if (n <= 0) return;
  k_cvt_f64_u32<<<launch_blocks(h, ceil_div64(n, 1024), 8), 256, 0, s>>>(in, n, out);
  KERNEL_CHECK();
  h->n_launches += 1;
}

// ---------------------------------------------------------------------------------------------
// factor / interaction expansion: out_j[i] = (code[i] == level_j) * (base ? base[i] : 1)
// Replaces the per-level column expressions of _expand_factors_polars / _expand_interactions_polars
// (python/leanfe/polars_impl.py:27-115): one pass reads the code (and base) of a row once and writes all J dummy
// columns, N (4 + 8 + w J) bytes instead of J passes of compare + multiply.
// ---------------------------------------------------------------------------------------------
constexpr int XP_MAXJ = 32;
struct ExpandSpec {
  int32_t level[XP_MAXJ];
  void* out[XP_MAXJ];
  int J;
};
template <typename OutT>
__global__ void __launch_bounds__(256) k_expand_dummies(const int32_t* __restrict__ code, const double* __restrict__ base,
                                                        int64_t n, ExpandSpec sp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t c = __ldcs(code + i);
    const double b = base ? __ldcs(base + i) : 1.0;
    for (int j = 0; j < sp.J; ++j) __stcs(reinterpret_cast<OutT*>(sp.out[j]) + i, c == sp.level[j] ? (OutT)b : (OutT)0);
  }
}

void expand_dummies_impl(lfe_ctx* h, const int32_t* codes, int64_t n, const int32_t* levels, int J, const double* base,
                         int out_dtype, void* const* out_cols, cudaStream_t s) {
  LFE_REQUIRE(out_dtype == LFE_F64 || (out_dtype == LFE_U8 && base == nullptr),
              "lfe_expand_dummies: outputs are float64, or uint8 for plain (base-less) dummies");
  if (n <= 0) return;
  const int blocks = launch_blocks(h, ceil_div64(n, 256 * 4), 8);
  for (int j0 = 0; j0 < J; j0 += XP_MAXJ) {
    ExpandSpec sp;
    sp.J = J - j0 < XP_MAXJ ? J - j0 : XP_MAXJ;
    for (int j = 0; j < sp.J; ++j) { sp.level[j] = levels[j0 + j]; sp.out[j] = out_cols[j0 + j]; }
    if (out_dtype == LFE_F64) k_expand_dummies<double><<<blocks, 256, 0, s>>>(codes, base, n, sp);
    else k_expand_dummies<uint8_t><<<blocks, 256, 0, s>>>(codes, nullptr, n, sp);
    KERNEL_CHECK();
    h->n_launches += 1;
  }
}
