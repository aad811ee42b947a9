// sort.cu -- stable LSD radix sort of (key, 32-bit payload) pairs, 8-10 bits per pass.
//
// Used for the per-FE entry lists of the demeaning loop (keys = dense group codes, payload = the
// other FE's code or the row index) and for the general factorisation path (64-bit keys).  Digit
// width is chosen per sort so that the key bits split evenly over the fewest passes (20-bit unit
// codes: 2 x 10; 9-bit period codes: 1 x 9).  Each pass is histogram -> exclusive scan -> scatter;
// ranks inside a tile come from warp-level __match_any_sync in element order, which keeps every pass
// stable and the whole sort deterministic.  (Ranks use BITS warp ballots per element, not MATCH.ANY.)
#include "common.cuh"

constexpr int RS_T = 256;
constexpr int RS_NW = RS_T / 32;
constexpr int RS_I = 16;
constexpr int RS_TILE = RS_T * RS_I;

// Lanes holding the same digit as this lane (the lane itself included), among valid lanes.  BITS ballots:
// fixed cost, unlike MATCH.ANY whose latency grows with the number of distinct values in the warp -- and
// with 9/10-bit digits of random codes nearly all 32 values are distinct.
template <int BITS>
__device__ __forceinline__ uint32_t warp_peers(uint32_t d, bool valid) {
  uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int b = 0; b < BITS; ++b) {
    const bool bit = (d >> b) & 1u;
    const uint32_t m = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? m : ~m;
  }
  return peers;
}

template <typename KeyT, int BITS>
__global__ void __launch_bounds__(RS_T) k_radix_hist(const KeyT* __restrict__ keys, int64_t n, int shift,
                                                     uint32_t* __restrict__ hist, int nblocks) {
  constexpr int BINS = 1 << BITS;
  __shared__ uint32_t cnt[BINS];
  for (int t = threadIdx.x; t < BINS; t += RS_T) cnt[t] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  const int lane = threadIdx.x & 31;
#pragma unroll 4
  for (int i = 0; i < RS_I; ++i) {
    int64_t idx = base + (int64_t)i * RS_T + threadIdx.x;
    bool valid = idx < n;
    uint32_t d = valid ? (uint32_t)((__ldcs(keys + idx) >> shift) & (BINS - 1)) : 0u;
    uint32_t peers = warp_peers<BITS>(d, valid);
    if (valid && lane == (__ffs(peers) - 1)) atomicAdd(&cnt[d], __popc(peers));
  }
  __syncthreads();
  for (int t = threadIdx.x; t < BINS; t += RS_T) hist[(size_t)blockIdx.x * BINS + t] = cnt[t];   // block-major: coalesced
}

template <typename KeyT, int BITS>
__global__ void __launch_bounds__(RS_T, 3) k_radix_scatter(const KeyT* __restrict__ keys_in,
                                                        const uint32_t* __restrict__ vals_in,
                                                        KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                        int64_t n, int shift, const uint32_t* __restrict__ offsets,
                                                        int nblocks) {
  constexpr int BINS = 1 << BITS;
  __shared__ __align__(16) uint32_t wcnt[RS_NW][BINS];
  {
    uint4* z = reinterpret_cast<uint4*>(&wcnt[0][0]);   // 128-bit zero fill
    for (int i = threadIdx.x; i < RS_NW * BINS / 4; i += RS_T) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (32 * RS_I);
  KeyT k[RS_I];
  uint32_t v[RS_I];
  uint32_t rk[RS_I];
#pragma unroll
  for (int r = 0; r < RS_I; ++r) {
    int64_t idx = wbase + r * 32 + lane;
    bool valid = idx < n;
    k[r] = valid ? __ldcs(keys_in + idx) : (KeyT)0;
    v[r] = valid ? (vals_in ? __ldcs(vals_in + idx) : (uint32_t)idx) : 0u;
  }
#pragma unroll
  for (int r = 0; r < RS_I; ++r) {
    int64_t idx = wbase + r * 32 + lane;
    bool valid = idx < n;
    uint32_t d = valid ? (uint32_t)((k[r] >> shift) & (BINS - 1)) : 0u;
    uint32_t peers = warp_peers<BITS>(d, valid);
    // the group leader bumps the warp's digit counter; the returned old value is the rank base of this round.
    // Same-warp atomics on one address retire in program order, so rounds need no barrier between them and
    // rounds that touch different digits overlap freely.
    const int leader = valid ? (__ffs(peers) - 1) : lane;
    uint32_t base = 0u;
    if (valid && lane == leader) base = atomicAdd(&wcnt[warp][d], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    rk[r] = base + __popc(peers & lt);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < BINS; t += RS_T) {
    // absolute output base of (warp, digit): global digit offset of this tile + the earlier warps' counts
    uint32_t run = offsets[(size_t)blockIdx.x * BINS + t];   // block-major copy of the scanned offsets: coalesced
#pragma unroll
    for (int w = 0; w < RS_NW; ++w) {
      uint32_t c = wcnt[w][t];
      wcnt[w][t] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_I; ++r) {
    int64_t idx = wbase + r * 32 + lane;
    if (idx < n) {
      uint32_t d = (uint32_t)((k[r] >> shift) & (BINS - 1));
      uint32_t pos = wcnt[warp][d] + rk[r];
      if (keys_out) keys_out[pos] = k[r];
      vals_out[pos] = v[r];
    }
  }
}

// [rows][cols] -> [cols][rows] through a padded shared-memory tile (both sides coalesced)
__global__ void __launch_bounds__(256) k_transpose_u32(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                       int rows, int cols) {
  __shared__ uint32_t tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int r = ty; r < 32; r += 8)
    if (by + r < rows && bx + tx < cols) tile[r][tx] = in[(size_t)(by + r) * cols + bx + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (bx + r < cols && by + tx < rows) out[(size_t)(bx + r) * rows + by + tx] = tile[tx][r];
}

template <typename KeyT, int BITS>
static void radix_pass(lfe_ctx* h, const KeyT* kin, const uint32_t* vin, KeyT* kout, uint32_t* vout, int64_t n,
                       int shift, uint32_t* hist, int nblocks, cudaStream_t s) {
  constexpr int BINS = 1 << BITS;
  uint32_t* hist_t = hist + (size_t)BINS * nblocks;   // second half of the buffer: digit-major copy for the scan
  k_radix_hist<KeyT, BITS><<<nblocks, RS_T, 0, s>>>(kin, n, shift, hist, nblocks);
  KERNEL_CHECK();
  dim3 g1((BINS + 31) / 32, (nblocks + 31) / 32), g2((nblocks + 31) / 32, (BINS + 31) / 32);
  k_transpose_u32<<<g1, 256, 0, s>>>(hist, hist_t, nblocks, BINS);          // [block][digit] -> [digit][block]
  KERNEL_CHECK();
  prim_exclusive_scan_u32(h, hist_t, hist_t, (int64_t)BINS * nblocks, nullptr, s);
  k_transpose_u32<<<g2, 256, 0, s>>>(hist_t, hist, BINS, nblocks);          // back to [block][digit]
  KERNEL_CHECK();
  k_radix_scatter<KeyT, BITS><<<nblocks, RS_T, 0, s>>>(kin, vin, kout, vout, n, shift, hist, nblocks);
  KERNEL_CHECK();
  h->n_launches += 4;
}

// vals_in == nullptr -> payload = element index (argsort)
template <typename KeyT>
static void sort_pairs_impl(lfe_ctx* h, const KeyT* keys, const uint32_t* vals_in, int64_t n, int key_bits,
                            KeyT* keys_out, uint32_t* vals_out, cudaStream_t s) {
  if (n <= 0) return;
  int npass = (key_bits + 9) / 10;
  if (npass == 0) {
    if (vals_in) CUDA_CHECK(cudaMemcpyAsync(vals_out, vals_in, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    else prim_iota_u32(vals_out, n, s);
    if (keys_out) CUDA_CHECK(cudaMemcpyAsync(keys_out, keys, n * sizeof(KeyT), cudaMemcpyDeviceToDevice, s));
    h->n_launches += 1;
    return;
  }
  int bits = (key_bits + npass - 1) / npass;
  if (bits < 8) bits = 8;
  const int nblocks = (int)ceil_div64(n, RS_TILE);
  DevBuf<uint32_t> hist((size_t)2 * (1 << bits) * nblocks, s);
  DevBuf<KeyT> ka(npass > 1 ? n : 0, s), kb(npass > 2 ? n : 0, s);
  DevBuf<uint32_t> va(npass > 1 ? n : 0, s);
  const KeyT* kin = keys;
  const uint32_t* vin = vals_in;
  for (int p = 0; p < npass; ++p) {
    const bool last = (p == npass - 1);
    uint32_t* vout = (((npass - 1 - p) & 1) == 0) ? vals_out : va.p;   // the last pass lands in vals_out
    KeyT* kout = last ? keys_out : ((p & 1) == 0 ? ka.p : kb.p);
    const int shift = p * bits;
    switch (bits) {
      case 8: radix_pass<KeyT, 8>(h, kin, vin, kout, vout, n, shift, hist, nblocks, s); break;
      case 9: radix_pass<KeyT, 9>(h, kin, vin, kout, vout, n, shift, hist, nblocks, s); break;
      default: radix_pass<KeyT, 10>(h, kin, vin, kout, vout, n, shift, hist, nblocks, s); break;
    }
    kin = kout;
    vin = vout;
  }
}

void prim_sort_pairs_u32(lfe_ctx* h, const uint32_t* keys, const uint32_t* vals_in, int64_t n, int key_bits,
                         uint32_t* keys_out, uint32_t* vals_out, cudaStream_t s) {
  sort_pairs_impl<uint32_t>(h, keys, vals_in, n, key_bits, keys_out, vals_out, s);
}
void prim_argsort_u32(lfe_ctx* h, const uint32_t* keys, int64_t n, int key_bits, uint32_t* perm_out,
                      uint32_t* keys_sorted_out, cudaStream_t s) {
  sort_pairs_impl<uint32_t>(h, keys, nullptr, n, key_bits, keys_sorted_out, perm_out, s);
}
void prim_argsort_u64(lfe_ctx* h, const uint64_t* keys, int64_t n, int key_bits, uint32_t* perm_out,
                      uint64_t* keys_sorted_out, cudaStream_t s) {
  sort_pairs_impl<uint64_t>(h, keys, nullptr, n, key_bits, keys_sorted_out, perm_out, s);
}
