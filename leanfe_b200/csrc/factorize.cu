// factorize.cu -- group codes, group counts, singleton mask, compaction, row ordering.
//
// lfe_factorize maps raw keys to dense int32 codes that are bit-identical to
// np.unique(keys, return_inverse=True)[1] (order 0) or pandas.factorize(keys)[0] (order 1).
// Keys are first mapped to an order-preserving uint64 image; if their range is small (the common
// case for panel ids) a direct-address presence table + scan gives the ranks in three streaming
// passes; otherwise a stable LSD radix argsort of (key - min) does.
#include <climits>
#include <vector>

#include "common.cuh"

void comm_allreduce_int(lfe_ctx* h, void* buf, int64_t count, int kind, int op, cudaStream_t s);  // api.cu

// ---------------------------------------------------------------------------------------------
// order-preserving uint64 image of a key
// ---------------------------------------------------------------------------------------------
template <int DT>
struct KeyIO;
template <>
struct KeyIO<LFE_I32> {
  typedef int32_t T;
  __device__ static uint64_t enc(T v) { return (uint64_t)((uint32_t)v ^ 0x80000000u); }
  __device__ static T dec(uint64_t u) { return (T)((uint32_t)u ^ 0x80000000u); }
};
template <>
struct KeyIO<LFE_I64> {
  typedef int64_t T;
  __device__ static uint64_t enc(T v) { return (uint64_t)v ^ 0x8000000000000000ull; }
  __device__ static T dec(uint64_t u) { return (T)(u ^ 0x8000000000000000ull); }
};
template <>
struct KeyIO<LFE_U32> {
  typedef uint32_t T;
  __device__ static uint64_t enc(T v) { return v; }
  __device__ static T dec(uint64_t u) { return (T)u; }
};
template <>
struct KeyIO<LFE_U64> {
  typedef uint64_t T;
  __device__ static uint64_t enc(T v) { return v; }
  __device__ static T dec(uint64_t u) { return u; }
};
template <>
struct KeyIO<LFE_U8> {
  typedef uint8_t T;
  __device__ static uint64_t enc(T v) { return v; }
  __device__ static T dec(uint64_t u) { return (T)u; }
};
template <>
struct KeyIO<LFE_F64> {
  typedef double T;
  // -0.0 -> +0.0, every NaN -> the canonical quiet NaN (sorts last, like np.unique)
  __device__ static uint64_t enc(T v) {
    if (v != v) return 0xFFF8000000000000ull;
    v = v + 0.0;
    uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
  }
  __device__ static T dec(uint64_t u) {
    if (u == 0xFFF8000000000000ull) return __longlong_as_double(0x7FF8000000000000ll);
    uint64_t b = (u & 0x8000000000000000ull) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u;
    return __longlong_as_double((long long)b);
  }
};

template <int DT>
__global__ void __launch_bounds__(256) k_minmax(const typename KeyIO<DT>::T* __restrict__ keys, int64_t n,
                                                unsigned long long* __restrict__ mm) {
  unsigned long long lo = ~0ull, hi = 0ull;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long u = KeyIO<DT>::enc(keys[i]);
    lo = u < lo ? u : lo;
    hi = u > hi ? u : hi;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, d), b = __shfl_xor_sync(0xffffffffu, hi, d);
    lo = a < lo ? a : lo;
    hi = b > hi ? b : hi;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&mm[0], lo);
    atomicMax(&mm[1], hi);
  }
}

template <int DT>
__global__ void __launch_bounds__(256) k_mark(const typename KeyIO<DT>::T* __restrict__ keys, int64_t n,
                                              unsigned long long kmin, uint32_t* __restrict__ flags) {
  // test before set: low-cardinality keys would otherwise serialise millions of stores on a few L2 lines
  // (a stale 0 from L1 only costs a redundant store; every writer stores 1)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t* f = flags + (KeyIO<DT>::enc(keys[i]) - kmin);
    if (*f == 0u) *f = 1u;
  }
}

// Direct-address COUNT table instead of presence flags: cnt[key - kmin] += 1.  Presence, ranks and the group
// sizes (which the singleton filter and the demean plan need next) all come from this one pass.
// Large ranges: RED.ADD.U32 into the L2-resident table, warp-aggregated only when neighbouring lanes hold
// equal keys (sorted panels); random ids have no duplicates inside a warp and MATCH.ANY would be pure latency.
template <int DT>
__global__ void __launch_bounds__(256) k_count_global(const typename KeyIO<DT>::T* __restrict__ keys, int64_t n,
                                                      unsigned long long kmin, uint32_t* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t nround = ceil_div64(n, stride);
  for (int64_t r = 0; r < nround; ++r) {
    const int64_t i = r * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const uint32_t c = valid ? (uint32_t)(KeyIO<DT>::enc(__ldcs(keys + i)) - kmin) : (0xffffff00u + lane);
    const uint32_t nb = __shfl_down_sync(0xffffffffu, c, 1);
    if (__any_sync(0xffffffffu, lane < 31 && nb == c)) {
      const uint32_t peers = __match_any_sync(0xffffffffu, c);
      if (valid && lane == (__ffs(peers) - 1)) atomicAdd(&cnt[c], (uint32_t)__popc(peers));
    } else if (valid) {
      atomicAdd(&cnt[c], 1u);
    }
  }
}
// Small ranges (<= 8192 distinct values possible): per-CTA bins in shared memory, native 32-bit shared atomics.
template <int DT>
__global__ void __launch_bounds__(256) k_count_smem(const typename KeyIO<DT>::T* __restrict__ keys, int64_t n,
                                                    unsigned long long kmin, int range, uint32_t* __restrict__ cnt) {
  extern __shared__ uint32_t sh_cnt[];
  for (int i = threadIdx.x; i < range; i += blockDim.x) sh_cnt[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&sh_cnt[(uint32_t)(KeyIO<DT>::enc(__ldcs(keys + i)) - kmin)], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < range; i += blockDim.x)
    if (sh_cnt[i]) atomicAdd(&cnt[i], sh_cnt[i]);
}
__global__ void __launch_bounds__(256) k_flags_from_counts(const uint32_t* __restrict__ cnt, int64_t range,
                                                           uint32_t* __restrict__ flags) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < range; i += (int64_t)gridDim.x * blockDim.x)
    flags[i] = cnt[i] ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k_compact_counts(const uint32_t* __restrict__ cnt,
                                                        const uint32_t* __restrict__ rank, int64_t range,
                                                        uint32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < range; i += (int64_t)gridDim.x * blockDim.x)
    if (cnt[i]) out[rank[i]] = cnt[i];
}

template <int DT>
__global__ void __launch_bounds__(256) k_lookup(const typename KeyIO<DT>::T* __restrict__ keys, int64_t n,
                                                unsigned long long kmin, const uint32_t* __restrict__ rank,
                                                int32_t* __restrict__ codes) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    codes[i] = (int32_t)rank[KeyIO<DT>::enc(keys[i]) - kmin];
}

template <int DT>
__global__ void __launch_bounds__(256) k_emit_uniques_direct(const uint32_t* __restrict__ flags,
                                                             const uint32_t* __restrict__ rank, int64_t range,
                                                             unsigned long long kmin,
                                                             typename KeyIO<DT>::T* __restrict__ uniques) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < range; i += (int64_t)gridDim.x * blockDim.x)
    if (flags[i]) uniques[rank[i]] = KeyIO<DT>::dec(kmin + (unsigned long long)i);
}

template <int DT>
__global__ void __launch_bounds__(256) k_encode_rel(const typename KeyIO<DT>::T* __restrict__ keys, int64_t n,
                                                    unsigned long long kmin, uint64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = KeyIO<DT>::enc(keys[i]) - kmin;
}

__global__ void __launch_bounds__(256) k_boundary_flags(const uint64_t* __restrict__ sk, int64_t n,
                                                        uint32_t* __restrict__ flags) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flags[i] = (i > 0 && sk[i] != sk[i - 1]) ? 1u : 0u;
}

// rank_excl = exclusive scan of boundary flags; the rank of position p is rank_excl[p] + flags[p]
template <int DT>
__global__ void __launch_bounds__(256) k_scatter_sorted_codes(const uint64_t* __restrict__ sk,
                                                              const uint32_t* __restrict__ perm,
                                                              const uint32_t* __restrict__ rank_excl,
                                                              const uint32_t* __restrict__ flags, int64_t n,
                                                              unsigned long long kmin, int32_t* __restrict__ codes,
                                                              typename KeyIO<DT>::T* __restrict__ uniques,
                                                              int64_t uniques_cap) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r = rank_excl[i] + flags[i];
    codes[perm[i]] = (int32_t)r;
    if (uniques && (i == 0 || flags[i]) && (int64_t)r < uniques_cap) uniques[r] = KeyIO<DT>::dec(sk[i] + kmin);
  }
}

// ---- first-appearance relabelling ------------------------------------------------------------
__global__ void __launch_bounds__(256) k_first_index(const int32_t* __restrict__ codes, int64_t n,
                                                     uint32_t* __restrict__ first) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    atomicMin(&first[codes[i]], (uint32_t)i);
}
__global__ void __launch_bounds__(256) k_invert_perm(const uint32_t* __restrict__ perm, int64_t G,
                                                     int32_t* __restrict__ newcode) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < G; r += (int64_t)gridDim.x * blockDim.x)
    newcode[perm[r]] = (int32_t)r;
}
__global__ void __launch_bounds__(256) k_remap(int32_t* __restrict__ codes, int64_t n,
                                               const int32_t* __restrict__ newcode) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    codes[i] = newcode[codes[i]];
}
template <typename T>
__global__ void __launch_bounds__(256) k_permute_small(const T* __restrict__ src, const uint32_t* __restrict__ perm,
                                                       int64_t G, T* __restrict__ dst) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < G; r += (int64_t)gridDim.x * blockDim.x)
    dst[r] = src[perm[r]];
}

__global__ void k_not_u64(unsigned long long* p) { *p = ~*p; }

template <int DT>
static int64_t factorize_impl(lfe_ctx* h, const void* keys_v, int64_t n, int order, int32_t* codes,
                              void* uniques_v, int64_t uniques_cap, cudaStream_t s) {
  typedef typename KeyIO<DT>::T T;
  const T* keys = (const T*)keys_v;
  T* uniques = (T*)uniques_v;
  const bool global = (order & LFE_ORDER_GLOBAL) != 0 && h->world > 1;
  order &= ~LFE_ORDER_GLOBAL;
  if (n == 0 && !global) return 0;
  int blocks = launch_blocks(h, ceil_div64(n, 256 * 4), 8);
  DevBuf<unsigned long long> mm(2, s);
  unsigned long long init[2] = {~0ull, 0ull};
  CUDA_CHECK(cudaMemcpyAsync(mm.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
  if (n > 0) {
    k_minmax<DT><<<blocks, 256, 0, s>>>(keys, n, mm);
    KERNEL_CHECK();
  }
  if (global) {  // one code space for all ranks: global key range (ONE max all-reduce of [~min, max]), merged presence table
    k_not_u64<<<1, 1, 0, s>>>(mm.p);
    comm_allreduce_int(h, mm.p, 2, 1, 2, s);
    k_not_u64<<<1, 1, 0, s>>>(mm.p);
    KERNEL_CHECK();
  }
  unsigned long long hmm[2];
  CUDA_CHECK(cudaMemcpyAsync(hmm, mm.p, sizeof(hmm), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  h->n_launches += 1;
  const unsigned long long kmin = hmm[0];
  const unsigned long long span = hmm[1] - hmm[0];  // range - 1
  int64_t G = 0;
  unsigned long long direct_limit = (unsigned long long)(16 * n > (1 << 20) ? 16 * n : (1 << 20));
  if (direct_limit > (1ull << 28)) direct_limit = 1ull << 28;
  DevBuf<uint32_t> total(1, s);
  if (global) {
    LFE_REQUIRE(order == LFE_ORDER_SORTED, "lfe_factorize: LFE_ORDER_GLOBAL supports sorted-rank codes only");
    LFE_REQUIRE(span < (1ull << 28), "lfe_factorize: LFE_ORDER_GLOBAL needs a key range < 2^28 (got %llu)", span);
    direct_limit = 1ull << 28;
  }
  h->last_counts.release();
  h->last_counts_G = -1;
  if (span < direct_limit) {
    int64_t range = (int64_t)span + 1;
    DevBuf<uint32_t> flags(range, s), rank(range, s), cnt;
    // rank-local sorted-rank codes: count instead of mark, the group sizes are a by-product
    const bool want_counts = !global && order == LFE_ORDER_SORTED && n > 0;
    if (want_counts) {
      cnt.alloc(range, s);
      cnt.zero();
      if (range <= 8192)
        k_count_smem<DT><<<blocks, 256, (size_t)range * sizeof(uint32_t), s>>>(keys, n, kmin, (int)range, cnt);
      else
        k_count_global<DT><<<launch_blocks(h, ceil_div64(n, 256 * 8), 8), 256, 0, s>>>(keys, n, kmin, cnt);
      KERNEL_CHECK();
      k_flags_from_counts<<<launch_blocks(h, ceil_div64(range, 256 * 4), 8), 256, 0, s>>>(cnt, range, flags);
      KERNEL_CHECK();
      h->n_launches += 1;
    } else {
      flags.zero();
      if (n > 0) {
        k_mark<DT><<<blocks, 256, 0, s>>>(keys, n, kmin, flags);
        KERNEL_CHECK();
      }
    }
    if (global) comm_allreduce_int(h, flags.p, range, 0, 2, s);
    prim_exclusive_scan_u32(h, flags, rank, range, total, s);
    if (n > 0) {
      k_lookup<DT><<<blocks, 256, 0, s>>>(keys, n, kmin, rank, codes);
      KERNEL_CHECK();
    }
    uint32_t ht = 0;
    CUDA_CHECK(cudaMemcpyAsync(&ht, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    G = ht;
    h->n_launches += 2;
    if (want_counts && G > 0) {
      h->last_counts.alloc(G, s);
      if (h->last_counts.p) {   // (allocation failure just means "no counts": the caller falls back to lfe_group_counts)
        k_compact_counts<<<launch_blocks(h, ceil_div64(range, 256 * 4), 8), 256, 0, s>>>(cnt, rank, range,
                                                                                        h->last_counts);
        KERNEL_CHECK();
        h->last_counts_G = G;
        h->last_counts_codes = codes;
        h->n_launches += 1;
      } else {
        (void)cudaGetLastError();   // clear the allocation error
      }
    }
    if (uniques && G <= uniques_cap) {
      int b2 = launch_blocks(h, ceil_div64(range, 256 * 4), 8);
      k_emit_uniques_direct<DT><<<b2, 256, 0, s>>>(flags, rank, range, kmin, uniques);
      KERNEL_CHECK();
      h->n_launches += 1;
    }
  } else {
    DevBuf<uint64_t> rel(n, s), sk(n, s);
    DevBuf<uint32_t> perm(n, s), flags(n, s), rank(n, s);
    k_encode_rel<DT><<<blocks, 256, 0, s>>>(keys, n, kmin, rel);
    KERNEL_CHECK();
    prim_argsort_u64(h, rel, n, bits_for(span), perm, sk, s);
    k_boundary_flags<<<blocks, 256, 0, s>>>(sk, n, flags);
    KERNEL_CHECK();
    prim_exclusive_scan_u32(h, flags, rank, n, total, s);
    k_scatter_sorted_codes<DT><<<blocks, 256, 0, s>>>(sk, perm, rank, flags, n, kmin, codes, uniques, uniques_cap);
    KERNEL_CHECK();
    uint32_t ht = 0;
    CUDA_CHECK(cudaMemcpyAsync(&ht, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    G = (int64_t)ht + 1;
    h->n_launches += 3;
  }
  if (order == LFE_ORDER_FIRST && G > 1) {
    DevBuf<uint32_t> first(G, s), gperm(G, s);
    DevBuf<int32_t> newcode(G, s);
    CUDA_CHECK(cudaMemsetAsync(first.p, 0xFF, (size_t)G * sizeof(uint32_t), s));
    k_first_index<<<blocks, 256, 0, s>>>(codes, n, first);
    KERNEL_CHECK();
    prim_argsort_u32(h, first, G, bits_for((uint64_t)n), gperm, nullptr, s);
    int bg = launch_blocks(h, ceil_div64(G, 256), 8);
    k_invert_perm<<<bg, 256, 0, s>>>(gperm, G, newcode);
    KERNEL_CHECK();
    k_remap<<<blocks, 256, 0, s>>>(codes, n, newcode);
    KERNEL_CHECK();
    h->n_launches += 3;
    if (uniques && G <= uniques_cap) {
      DevBuf<T> tmp(G, s);
      CUDA_CHECK(cudaMemcpyAsync(tmp.p, uniques, (size_t)G * sizeof(T), cudaMemcpyDeviceToDevice, s));
      k_permute_small<T><<<bg, 256, 0, s>>>(tmp, gperm, G, uniques);
      KERNEL_CHECK();
      h->n_launches += 1;
    }
  }
  return G;
}

// ---- float keys that are really small integers (dummies, counts, years stored as double) -------------
// flags[0] |= 1 if any value is non-integral / non-finite / outside +-2^52 ; mm = {min, max} as int64
__global__ void __launch_bounds__(256) k_f64_integral_range(const double* __restrict__ v, int64_t n,
                                                            unsigned int* __restrict__ bad, long long* __restrict__ mm) {
  long long lo = LLONG_MAX, hi = LLONG_MIN;
  int b = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double x = __ldcs(v + i);
    if (!(fabs(x) < 4503599627370496.0) || x != rint(x)) { b = 1; continue; }
    long long q = (long long)x;
    lo = q < lo ? q : lo;
    hi = q > hi ? q : hi;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    long long a = __shfl_xor_sync(0xffffffffu, lo, d), c = __shfl_xor_sync(0xffffffffu, hi, d);
    lo = a < lo ? a : lo;
    hi = c > hi ? c : hi;
    b |= __shfl_xor_sync(0xffffffffu, b, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (b) atomicOr(bad, 1u);
    atomicMin(&mm[0], lo);
    atomicMax(&mm[1], hi);
  }
}
__global__ void __launch_bounds__(256) k_f64_to_i64(const double* __restrict__ v, int64_t n, long long* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (long long)__ldcs(v + i);     // -0.0 -> 0: the canonical form
}
__global__ void __launch_bounds__(256) k_i64_to_f64(const long long* __restrict__ v, int64_t n, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (double)v[i];
}

// Integer-valued doubles with a small range take the direct-address path instead of a 64-bit radix sort.
static bool factorize_f64_as_int(lfe_ctx* h, const double* keys, int64_t n, int order, int32_t* codes, double* uniques,
                                 int64_t uniques_cap, cudaStream_t s, int64_t* G_out) {
  // With a shared code space (LFE_ORDER_GLOBAL) the decision is taken on all-reduced quantities, so every
  // rank follows the same branch and issues the same collectives -- also a rank that holds no rows.
  const bool global = (order & LFE_ORDER_GLOBAL) != 0 && h->world > 1;
  if (n <= 0 && !global) return false;
  DevBuf<unsigned int> bad(1, s);
  DevBuf<long long> mm(2, s);
  bad.zero();
  long long init[2] = {LLONG_MAX, LLONG_MIN};
  CUDA_CHECK(cudaMemcpyAsync(mm.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
  int blocks = launch_blocks(h, ceil_div64(n > 0 ? n : 1, 256 * 4), 8);
  if (n > 0) {
    k_f64_integral_range<<<blocks, 256, 0, s>>>(keys, n, bad, mm);
    KERNEL_CHECK();
  }
  if (global) {
    comm_allreduce_int(h, bad.p, 1, 0, 2, s);
    comm_allreduce_int(h, mm.p, 1, 2, 3, s);
    comm_allreduce_int(h, mm.p + 1, 1, 2, 2, s);
  }
  unsigned int hb = 0;
  long long hmm[2];
  CUDA_CHECK(cudaMemcpyAsync(&hb, bad.p, sizeof(hb), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaMemcpyAsync(hmm, mm.p, sizeof(hmm), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  h->n_launches += 1;
  if (hb || hmm[1] < hmm[0] || (unsigned long long)(hmm[1] - hmm[0]) >= (1ull << 24)) {
    LFE_REQUIRE(!global, "lfe_factorize: LFE_ORDER_GLOBAL on float keys needs integer-valued data with a range < 2^24");
    return false;
  }
  DevBuf<long long> ik(n > 0 ? n : 1, s), iu;
  if (n > 0) {
    k_f64_to_i64<<<blocks, 256, 0, s>>>(keys, n, ik);
    KERNEL_CHECK();
  }
  if (uniques) iu.alloc((size_t)(hmm[1] - hmm[0] + 1), s);
  int64_t G = factorize_impl<LFE_I64>(h, ik.p, n, order, codes, uniques ? iu.p : nullptr, hmm[1] - hmm[0] + 1, s);
  if (uniques && G <= uniques_cap) {
    k_i64_to_f64<<<launch_blocks(h, ceil_div64(G, 256), 8), 256, 0, s>>>(iu.p, G, uniques);
    KERNEL_CHECK();
  }
  h->n_launches += 2;
  *G_out = G;
  return true;
}

int64_t factorize_dispatch(lfe_ctx* h, const void* keys, int key_dtype, int64_t n, int order, int32_t* codes,
                           void* uniques, int64_t uniques_cap, cudaStream_t s) {
  PhaseTimer t(h, s, 0);
  if (key_dtype == LFE_F64) {
    int64_t G = 0;
    if (factorize_f64_as_int(h, (const double*)keys, n, order, codes, (double*)uniques, uniques_cap, s, &G)) return G;
  }
  switch (key_dtype) {
    case LFE_I32: return factorize_impl<LFE_I32>(h, keys, n, order, codes, uniques, uniques_cap, s);
    case LFE_I64: return factorize_impl<LFE_I64>(h, keys, n, order, codes, uniques, uniques_cap, s);
    case LFE_U32: return factorize_impl<LFE_U32>(h, keys, n, order, codes, uniques, uniques_cap, s);
    case LFE_U64: return factorize_impl<LFE_U64>(h, keys, n, order, codes, uniques, uniques_cap, s);
    case LFE_F64: return factorize_impl<LFE_F64>(h, keys, n, order, codes, uniques, uniques_cap, s);
    case LFE_U8: return factorize_impl<LFE_U8>(h, keys, n, order, codes, uniques, uniques_cap, s);
    default: LFE_THROW(LFE_ERR_ARG, "lfe_factorize: unknown key dtype %d", key_dtype);
  }
}

// ---------------------------------------------------------------------------------------------
// [min, max] of K integer key columns in one call (one sync): lets the caller use small raw values
// (uint8 dummies) directly as group codes, radix = max + 1, instead of factorising every column.
// ---------------------------------------------------------------------------------------------
struct RangeCols {
  const void* col[32];
  int dtype[32];
};
template <typename T, typename V, int W>   // V = 16-byte vector of W elements of T
__device__ __forceinline__ void range_scan(const T* __restrict__ v, int64_t n, long long& lo, long long& hi) {
  const bool vec = (((uintptr_t)v) & 15) == 0;
  const int64_t nv = vec ? n / W : 0;
  T tlo = (T)0, thi = (T)0;
  bool any = false;
#pragma unroll 4
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nv; q += (int64_t)gridDim.x * blockDim.x) {
    V t = __ldcs(reinterpret_cast<const V*>(v) + q);
    const T* e = reinterpret_cast<const T*>(&t);
    if (!any) { tlo = thi = e[0]; any = true; }
#pragma unroll
    for (int i = 0; i < W; ++i) {
      tlo = e[i] < tlo ? e[i] : tlo;
      thi = e[i] > thi ? e[i] : thi;
    }
  }
  for (int64_t i = nv * W + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T x = v[i];
    if (!any) { tlo = thi = x; any = true; }
    tlo = x < tlo ? x : tlo;
    thi = x > thi ? x : thi;
  }
  if (any) { lo = (long long)tlo; hi = (long long)thi; }
}
// grid = (CTAs per column, K columns); one atomic pair per CTA
__global__ void __launch_bounds__(256) k_ranges(RangeCols rc, int64_t n, long long* __restrict__ mm) {
  const int j = blockIdx.y;
  long long lo = LLONG_MAX, hi = LLONG_MIN;
  switch (rc.dtype[j]) {
    case LFE_U8: range_scan<uint8_t, uint4, 16>((const uint8_t*)rc.col[j], n, lo, hi); break;
    case LFE_I32: range_scan<int32_t, int4, 4>((const int32_t*)rc.col[j], n, lo, hi); break;
    default: range_scan<long long, longlong2, 2>((const long long*)rc.col[j], n, lo, hi); break;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    long long a = __shfl_xor_sync(0xffffffffu, lo, d), c = __shfl_xor_sync(0xffffffffu, hi, d);
    lo = a < lo ? a : lo;
    hi = c > hi ? c : hi;
  }
  __shared__ long long slo[8], shi[8];
  if ((threadIdx.x & 31) == 0) {
    slo[threadIdx.x >> 5] = lo;
    shi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      lo = slo[w] < lo ? slo[w] : lo;
      hi = shi[w] > hi ? shi[w] : hi;
    }
    if (lo <= hi) {
      atomicMin(&mm[2 * j], lo);
      atomicMax(&mm[2 * j + 1], hi);
    }
  }
}

void key_ranges_impl(lfe_ctx* h, const void* const* cols, const int* dtypes, int K, int64_t n, int64_t* min_out,
                     int64_t* max_out, cudaStream_t s) {
  PhaseTimer t(h, s, 0);
  LFE_REQUIRE(K >= 1 && K <= 1024, "lfe_key_ranges: K out of range");
  std::vector<long long> init(2 * K);
  for (int j = 0; j < K; ++j) {
    init[2 * j] = LLONG_MAX;
    init[2 * j + 1] = LLONG_MIN;
  }
  DevBuf<long long> mm(2 * K, s);
  CUDA_CHECK(cudaMemcpyAsync(mm.p, init.data(), init.size() * sizeof(long long), cudaMemcpyHostToDevice, s));
  CUDA_CHECK(cudaStreamSynchronize(s));   // init is a host stack/heap buffer
  for (int j = 0; j < K; ++j)
    LFE_REQUIRE(dtypes[j] == LFE_U8 || dtypes[j] == LFE_I32 || dtypes[j] == LFE_I64,
                "lfe_key_ranges: column %d: dtype must be LFE_U8, LFE_I32 or LFE_I64", j);
  for (int j0 = 0; j0 < K && n > 0; j0 += 32) {   // up to 32 columns per launch
    const int kk = K - j0 < 32 ? K - j0 : 32;
    RangeCols rc;
    for (int j = 0; j < kk; ++j) {
      rc.col[j] = cols[j0 + j];
      rc.dtype[j] = dtypes[j0 + j];
    }
    int bx = (8 * h->sm_count + kk - 1) / kk;
    const int64_t need = ceil_div64(n, 256 * 16);
    if (bx > need) bx = (int)need;
    k_ranges<<<dim3(bx < 1 ? 1 : bx, kk), 256, 0, s>>>(rc, n, mm.p + 2 * j0);
    KERNEL_CHECK();
    h->n_launches += 1;
  }
  if (h->world > 1) {   // the ranges of the whole (sharded) column; interleaved {min, max} pairs -> two strided passes
    std::vector<long long> tmp(2 * K);
    DevBuf<long long> lo(K, s), hi(K, s);
    CUDA_CHECK(cudaMemcpy2DAsync(lo.p, sizeof(long long), mm.p, 2 * sizeof(long long), sizeof(long long), K,
                                 cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpy2DAsync(hi.p, sizeof(long long), mm.p + 1, 2 * sizeof(long long), sizeof(long long), K,
                                 cudaMemcpyDeviceToDevice, s));
    comm_allreduce_int(h, lo.p, K, 2, 3, s);
    comm_allreduce_int(h, hi.p, K, 2, 2, s);
    CUDA_CHECK(cudaMemcpyAsync(min_out, lo.p, K * sizeof(long long), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(max_out, hi.p, K * sizeof(long long), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    return;
  }
  std::vector<long long> host(2 * K);
  CUDA_CHECK(cudaMemcpyAsync(host.data(), mm.p, host.size() * sizeof(long long), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  for (int j = 0; j < K; ++j) {
    min_out[j] = host[2 * j];
    max_out[j] = host[2 * j + 1];
  }
}

// ---------------------------------------------------------------------------------------------
// singleton mask  (polars_impl.py:478-482)
// ---------------------------------------------------------------------------------------------
struct CodePtrs {
  const int32_t* code[LFE_MAX_FE];
  const uint32_t* cnt[LFE_MAX_FE];
};

__global__ void __launch_bounds__(256) k_singleton_mask(CodePtrs P, int F, int64_t n, uint8_t* __restrict__ keep,
                                                        unsigned long long* __restrict__ n_keep) {
  unsigned int local = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    bool k = true;
    for (int j = 0; j < F; ++j) k = k && (P.cnt[j][P.code[j][i]] > 1u);
    keep[i] = k ? 1 : 0;
    local += k ? 1u : 0u;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_keep, (unsigned long long)local);
}

// after the drop: a group survives iff it still has a kept row
__global__ void __launch_bounds__(256) k_mark_kept_groups(const int32_t* __restrict__ code,
                                                          const uint8_t* __restrict__ keep, int64_t n,
                                                          uint32_t* __restrict__ present) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (keep[i]) present[code[i]] = 1u;
}
__global__ void __launch_bounds__(256) k_count_nonzero(const uint32_t* __restrict__ v, int64_t n,
                                                       unsigned long long* __restrict__ out) {
  unsigned int local = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    local += v[i] ? 1u : 0u;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, (unsigned long long)local);
}

__global__ void __launch_bounds__(256) k_count_ones(const uint32_t* __restrict__ v, int64_t n,
                                                    unsigned long long* __restrict__ out) {
  unsigned int local = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    local += v[i] == 1u ? 1u : 0u;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, (unsigned long long)local);
}

// Multi-GPU: group sizes of every FE except `local_fe` are summed over the ranks before the test; n_keep_out[0]
// is the local, n_keep_out[1] the global number of kept rows; fe_dims are global.
void singleton_mask_impl(lfe_ctx* h, const int32_t* const* codes, const int64_t* n_groups, int F, int64_t n,
                         int local_fe, const int32_t* const* counts_in, uint8_t* keep, int64_t* n_keep_out,
                         int64_t* fe_dims_out, cudaStream_t s) {
  PhaseTimer t(h, s, 1);
  LFE_REQUIRE(F >= 1 && F <= LFE_MAX_FE, "lfe_singleton_mask: F=%d out of range", F);
  const bool dist = h->world > 1;
  std::vector<DevBuf<uint32_t>> cnts(F);
  CodePtrs P;
  for (int j = 0; j < F; ++j) {
    cnts[j].alloc((size_t)n_groups[j], s);
    if (counts_in && counts_in[j])   // rank-local group sizes supplied by the caller (lfe_group_counts)
      CUDA_CHECK(cudaMemcpyAsync(cnts[j].p, counts_in[j], (size_t)n_groups[j] * sizeof(uint32_t),
                                 cudaMemcpyDeviceToDevice, s));
    else
      prim_histogram(h, codes[j], n, n_groups[j], cnts[j], s);
    if (dist && j != local_fe) comm_allreduce_int(h, cnts[j].p, n_groups[j], 0, 0, s);
    P.code[j] = codes[j];
    P.cnt[j] = cnts[j];
  }
  DevBuf<unsigned long long> acc(2 + F, s);
  acc.zero();
  int blocks = launch_blocks(h, ceil_div64(n, 256 * 4), 8);
  std::vector<unsigned long long> hacc(2 + F);
  // No group of size one anywhere (the usual panel) => every row is kept: the row pass is skipped.  The group
  // tables are G-sized, the row pass N-sized.  Everything that decision and its results need from the other ranks
  // -- the number of singleton groups (so that all ranks take the same branch), the global row count and the
  // shard key's group count -- travels in ONE all-reduce of three counters, read back with one synchronisation;
  // the non-local FEs' tables are already global, so their counters are the same on every rank.
  int64_t n_tot = n;
  {
    DevBuf<unsigned long long> red(3, s), nz(F, s);
    red.zero(); nz.zero();
    for (int j = 0; j < F; ++j) {
      const int bg = launch_blocks(h, ceil_div64(n_groups[j], 256 * 4), 8);
      k_count_ones<<<bg, 256, 0, s>>>(cnts[j], n_groups[j], red.p);
      KERNEL_CHECK();
      k_count_nonzero<<<bg, 256, 0, s>>>(cnts[j], n_groups[j], nz.p + j);
      KERNEL_CHECK();
    }
    h->n_launches += 2;
    unsigned long long hn = (unsigned long long)n;
    CUDA_CHECK(cudaMemcpyAsync(red.p + 1, &hn, sizeof(hn), cudaMemcpyHostToDevice, s));   // pageable: staged at once
    if (dist && local_fe >= 0 && local_fe < F)
      CUDA_CHECK(cudaMemcpyAsync(red.p + 2, nz.p + local_fe, sizeof(hn), cudaMemcpyDeviceToDevice, s));
    if (dist) comm_allreduce_int(h, red.p, 3, 1, 0, s);
    unsigned long long hred[3];
    std::vector<unsigned long long> hnz(F);
    CUDA_CHECK(cudaMemcpyAsync(hred, red.p, sizeof(hred), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(hnz.data(), nz.p, F * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    n_tot = (int64_t)hred[1];
    n_keep_out[2] = n_tot;
    if (hred[0] == 0) {
      if (n > 0) CUDA_CHECK(cudaMemsetAsync(keep, 1, (size_t)n, s));
      n_keep_out[0] = n;
      n_keep_out[1] = n_tot;
      if (fe_dims_out)
        for (int j = 0; j < F; ++j) fe_dims_out[j] = (int64_t)((dist && j == local_fe) ? hred[2] : hnz[j]);
      return;
    }
  }
  if (n > 0) {
    k_singleton_mask<<<blocks, 256, 0, s>>>(P, F, n, keep, acc);
    KERNEL_CHECK();
    h->n_launches += 1;
  }
  CUDA_CHECK(cudaMemcpyAsync(acc.p + 1, acc.p, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
  if (dist) comm_allreduce_int(h, acc.p + 1, 1, 1, 0, s);
  CUDA_CHECK(cudaMemcpyAsync(hacc.data(), acc.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  if (fe_dims_out) {
    const bool dropped = (int64_t)hacc[1] < n_tot;
    for (int j = 0; j < F; ++j) {
      if (dropped) {  // a group survives iff it still has a kept row: one more pass over the codes
        CUDA_CHECK(cudaMemsetAsync(cnts[j].p, 0, (size_t)n_groups[j] * sizeof(uint32_t), s));
        if (n > 0) {
          k_mark_kept_groups<<<launch_blocks(h, ceil_div64(n, 256 * 4), 8), 256, 0, s>>>(codes[j], keep, n, cnts[j]);
          KERNEL_CHECK();
          h->n_launches += 1;
        }
        if (dist && j != local_fe) comm_allreduce_int(h, cnts[j].p, n_groups[j], 0, 2, s);
      }  // else: nothing dropped, the non-empty groups of the (global) histogram are the answer
      int bg = launch_blocks(h, ceil_div64(n_groups[j], 256 * 4), 8);
      k_count_nonzero<<<bg, 256, 0, s>>>(cnts[j], n_groups[j], acc.p + 2 + j);
      KERNEL_CHECK();
      h->n_launches += 1;
      if (dist && j == local_fe) comm_allreduce_int(h, acc.p + 2 + j, 1, 1, 0, s);
    }
    CUDA_CHECK(cudaMemcpyAsync(hacc.data() + 2, acc.p + 2, F * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
  n_keep_out[0] = (int64_t)hacc[0];
  n_keep_out[1] = (int64_t)hacc[1];
  if (fe_dims_out)
    for (int j = 0; j < F; ++j) fe_dims_out[j] = (int64_t)hacc[2 + j];
}

// ---------------------------------------------------------------------------------------------
// stable compaction
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_keep_to_u32(const uint8_t* __restrict__ keep, int64_t n,
                                                     uint32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = keep[i];
}
template <typename T>
__global__ void __launch_bounds__(256) k_compact(const T* __restrict__ src, const uint8_t* __restrict__ keep,
                                                 const uint32_t* __restrict__ pos, int64_t n, T* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (keep[i]) dst[pos[i]] = src[i];
}

void compact_impl(lfe_ctx* h, const uint8_t* keep, int64_t n, const double* const* f64_in, double* const* f64_out,
                  int C, const int32_t* const* i32_in, int32_t* const* i32_out, int K, cudaStream_t s) {
  PhaseTimer t(h, s, 2);
  if (n <= 0) return;
  DevBuf<uint32_t> pos(n, s);
  int blocks = launch_blocks(h, ceil_div64(n, 256 * 4), 8);
  k_keep_to_u32<<<blocks, 256, 0, s>>>(keep, n, pos);
  KERNEL_CHECK();
  prim_exclusive_scan_u32(h, pos, pos, n, nullptr, s);
  for (int c = 0; c < C; ++c) {
    k_compact<double><<<blocks, 256, 0, s>>>(f64_in[c], keep, pos, n, f64_out[c]);
    KERNEL_CHECK();
  }
  for (int c = 0; c < K; ++c) {
    k_compact<int32_t><<<blocks, 256, 0, s>>>(i32_in[c], keep, pos, n, i32_out[c]);
    KERNEL_CHECK();
  }
  h->n_launches += 1 + C + K;
}

// ---------------------------------------------------------------------------------------------
// row ordering
// ---------------------------------------------------------------------------------------------
void sort_rows_impl(lfe_ctx* h, const int32_t* codes, int64_t n, int64_t n_groups, int32_t* perm_out,
                    int* already_sorted_out, cudaStream_t s) {
  PhaseTimer t(h, s, 3);
  int sorted = prim_is_sorted(h, codes, n, s);
  if (already_sorted_out) *already_sorted_out = sorted;
  if (sorted) {
    prim_iota_u32((uint32_t*)perm_out, n, s);
    h->n_launches += 1;
    return;
  }
  prim_argsort_u32(h, (const uint32_t*)codes, n, bits_for((uint64_t)(n_groups > 0 ? n_groups - 1 : 0)),
                   (uint32_t*)perm_out, nullptr, s);
}

void gather_impl(lfe_ctx* h, const int32_t* perm, int64_t n, const double* const* f64_in, double* const* f64_out,
                 int C, const int32_t* const* i32_in, int32_t* const* i32_out, int K, cudaStream_t s) {
  PhaseTimer t(h, s, 4);
  for (int c = 0; c < C; ++c) prim_gather_f64(f64_in[c], (const uint32_t*)perm, n, f64_out[c], s);
  for (int c = 0; c < K; ++c) prim_gather_i32(i32_in[c], (const uint32_t*)perm, n, i32_out[c], s);
  h->n_launches += C + K;
}
