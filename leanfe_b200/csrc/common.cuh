// common.cuh -- shared host/device helpers for liblfe (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/lfe.h"

#define LFE_SM_COUNT_DEFAULT 148

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
struct lfe_error {
  int code;
  std::string msg;
};
void lfe_set_error(const std::string& m);

#define LFE_THROW(code_, ...)                          \
  do {                                                 \
    char _b[512];                                      \
    snprintf(_b, sizeof(_b), __VA_ARGS__);             \
    throw lfe_error{(code_), std::string(_b)};         \
  } while (0)

#define CUDA_CHECK(x)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (x);                                                                  \
    if (_e != cudaSuccess)                                                                 \
      LFE_THROW(LFE_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(_e)); \
  } while (0)

#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())

#define LFE_REQUIRE(cond, ...)                      \
  do {                                              \
    if (!(cond)) LFE_THROW(LFE_ERR_ARG, __VA_ARGS__); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
struct lfe_ctx {
  int device = 0;
  int sm_count = LFE_SM_COUNT_DEFAULT;
  size_t l2_bytes = 0;
  int smem_optin = 0;
  bool profiling = false;
  double timer_ms[LFE_N_TIMERS] = {0};
  struct ev_pair {
    int slot;
    cudaEvent_t a, b;
  };
  std::vector<ev_pair> pending;
  std::vector<cudaEvent_t> free_events;
  int64_t n_launches = 0;
  int64_t n_sweep_launches = 0;
  int64_t n_collectives = 0;   // NCCL calls issued since the last timer reset
  // NCCL (dlopen'ed)
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  int world_attached = 1;   // size of the attached communicator; world == 1 while collectives are suspended
  // group sizes produced as a by-product of the last lfe_factorize call (direct-address path), see
  // lfe_factorize_counts; owned by the handle, stream-ordered
  struct CountsBuf {
    uint32_t* p = nullptr;
    cudaStream_t s = nullptr;
    void alloc(size_t n, cudaStream_t s_) {
      release();
      s = s_;
      if (n && cudaMallocAsync((void**)&p, n * sizeof(uint32_t), s) != cudaSuccess) p = nullptr;
    }
    void release() {
      if (p) cudaFreeAsync(p, s);
      p = nullptr;
    }
    operator uint32_t*() const { return p; }
  } last_counts;
  int64_t last_counts_G = -1;
  const void* last_counts_codes = nullptr;
  // peer-to-peer receive arena of the row repartition (shard.cu): cudaMalloc'ed, exported with CUDA IPC, written
  // directly by the other ranks' push kernels over NVLink
  struct P2P {
    void* base = nullptr;
    size_t bytes = 0;
    std::vector<void*> peer_base;          // [rank] mapped arena of every peer (own: base)
    std::vector<std::string> peer_handle;  // [rank] the IPC handle bytes currently mapped ("" = none)
    bool disabled = false;                 // IPC unavailable: the grouped ncclSend/ncclRecv path is used
  } p2p;
  // lfe_set_option
  int opt_columns_hint = 0;                 // columns per demean run (0 = unknown): sizes the L2 blocks of the entry lists
  int64_t opt_l2_block_bytes = 48ll << 20;  // a table slice this large is expected to stay L2-resident while it is gathered
  int opt_deterministic = 0;
};

// RAII phase timer: records a CUDA-event pair on the stream when profiling is on.
struct PhaseTimer {
  lfe_ctx* h;
  cudaStream_t s;
  int slot;
  cudaEvent_t a = nullptr, b = nullptr;
  PhaseTimer(lfe_ctx* h_, cudaStream_t s_, int slot_);
  ~PhaseTimer();
};

// ---------------------------------------------------------------------------------------------
// stream-ordered scratch
// ---------------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  DevBuf() {}
  DevBuf(size_t n_, cudaStream_t s_) { alloc(n_, s_); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    release();
    p = o.p; n = o.n; s = o.s; o.p = nullptr;
    return *this;
  }
  void alloc(size_t n_, cudaStream_t s_) {
    release();
    n = n_; s = s_;
    if (n == 0) { p = nullptr; return; }
    CUDA_CHECK(cudaMallocAsync((void**)&p, n * sizeof(T), s));
  }
  void zero() { if (p) CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
  }
  ~DevBuf() { release(); }
  operator T*() const { return p; }
};

__host__ __device__ static inline int ceil_div_i(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
__host__ __device__ static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double shfl_down_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_xor_d(double v, int d) { return __shfl_xor_sync(0xffffffffu, v, d); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += shfl_xor_d(v, d);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmax(v, shfl_xor_d(v, d));
  return v;
}
// max of non-negative doubles through their (order-preserving) bit image
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}
// streaming loads / stores (read-once data: do not pollute L1, evict-first in L2)
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }
#endif

// ---------------------------------------------------------------------------------------------
// primitives (prims.cu)
// ---------------------------------------------------------------------------------------------
// exclusive prefix sum of n uint32 values (in may alias out); optional total written to total_dev.
void prim_exclusive_scan_u32(lfe_ctx* h, const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total_dev,
                             cudaStream_t s);
// stable LSD radix argsort: perm_out[p] = index (in [0,n)) of the p-th smallest key; keys < 2^key_bits.
// keys_sorted_out nullable.
void prim_argsort_u32(lfe_ctx* h, const uint32_t* keys, int64_t n, int key_bits, uint32_t* perm_out,
                      uint32_t* keys_sorted_out, cudaStream_t s);
void prim_sort_pairs_u32(lfe_ctx* h, const uint32_t* keys, const uint32_t* vals_in, int64_t n, int key_bits,
                         uint32_t* keys_out, uint32_t* vals_out, cudaStream_t s);
// keys only, sorted (stably) on bits [bit_lo, bit_lo + key_bits)
void prim_sort_keys_u32(lfe_ctx* h, const uint32_t* keys, int64_t n, int bit_lo, int key_bits, uint32_t* keys_out,
                        cudaStream_t s);
void prim_argsort_u64(lfe_ctx* h, const uint64_t* keys, int64_t n, int key_bits, uint32_t* perm_out,
                      uint64_t* keys_sorted_out, cudaStream_t s);
// counts[g] = #codes == g
void prim_histogram(lfe_ctx* h, const int32_t* codes, int64_t n, int64_t G, uint32_t* counts, cudaStream_t s);
// returns 1 on host if codes non-decreasing (synchronises)
int prim_is_sorted(lfe_ctx* h, const int32_t* codes, int64_t n, cudaStream_t s);
void prim_gather_f64(const double* src, const uint32_t* perm, int64_t n, double* dst, cudaStream_t s);
void prim_gather_i32(const int32_t* src, const uint32_t* perm, int64_t n, int32_t* dst, cudaStream_t s);
void prim_iota_u32(uint32_t* p, int64_t n, cudaStream_t s);
void prim_u32_to_f64(lfe_ctx* h, const uint32_t* in, int64_t n, double* out, cudaStream_t s);
void prim_f64_to_u32(lfe_ctx* h, const double* in, int64_t n, uint32_t* out, cudaStream_t s);
int bits_for(uint64_t max_value);
int launch_blocks(lfe_ctx* h, int64_t work_items, int per_sm);
