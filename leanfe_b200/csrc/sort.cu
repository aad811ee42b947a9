// sort.cu -- stable LSD radix sort of keys or (key, 32-bit payload) pairs, 8-11 bits per pass.
//
// Used for the per-FE entry lists of the demeaning loop (packed (big code, small code) words sorted on a
// bit range, or dense group codes with a row-index payload) and for the general factorisation path
// (64-bit keys).  Digit width is chosen per sort so that the key bits split evenly over the fewest passes
// (20-bit unit codes: 2 x 10; 22 bits: 2 x 11; 9-bit period codes: 1 x 9).
//
// One pass = per-tile digit histogram -> exclusive scan (digit-major) -> scatter.  The scatter kernel is
// built for B200: a 512-thread CTA ranks an 8192-key tile, reorders it by digit in shared memory and writes
// it out in digit runs, so global stores are sector-sized runs instead of 32 scattered 4-byte stores per warp.
// Ranks come from BITS warp ballots per round (fixed cost, independent of the number of distinct digits) and
// per-warp 16-bit digit counters packed two to a word; the ballot / counter-atomic / shuffle steps of the 16
// rounds are issued as separate straight-line phases (8 rounds at a time) so the shared-memory atomics of different rounds
// overlap instead of forming a 16-deep dependent chain (that chain -- short-scoreboard stalls -- was 52 % of
// the previous kernel's stall samples, profiles/r1_ncu_top_kernels_final.md).  A CTA walks `sub` consecutive
// tiles and keeps the running global digit offsets in registers, so the histogram has one row per
// sub x 8192 keys (6 % of the key bytes at 400M keys instead of 25 %).  Element order inside a tile is
// (warp, round, lane) = ascending index, which keeps every pass stable and the whole sort deterministic.
#include <algorithm>

#include "common.cuh"

constexpr int RS_T = 512;
constexpr int RS_NW = RS_T / 32;
constexpr int RS_I = 16;
constexpr int RS_TILE = RS_T * RS_I;   // 8192 keys per tile
constexpr int RS_WKEYS = 32 * RS_I;    // each warp owns 512 consecutive keys of the tile
constexpr int RS_H = 8;                // ranking rounds whose ballots / atomics / shuffles are issued together

// (MATCH.ANY instead of the ballots was measured again in this kernel: 50M-key passes got 0.13 ms slower.)
// Lanes holding the same digit as this lane (the lane itself included), among valid lanes.  BITS ballots:
// fixed cost, unlike MATCH.ANY whose latency grows with the number of distinct values in the warp -- and
// with 9-11-bit digits of random codes nearly all 32 values are distinct.
// One digit bit: test -> vote -> select -> and, as one PTX block so the predicate is consumed where it is
// produced (left to the compiler, the 160 predicates of a tile were computed up front and parked in general
// registers bit by bit: 7 LOP3 per bit instead of 2).
template <int B>
__device__ __forceinline__ void peer_bits(uint32_t& peers, uint32_t d) {
  if constexpr (B >= 0) {
    peer_bits<B - 1>(peers, d);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 m, x;\n\t"
        "and.b32 x, %1, %2;\n\t"
        "setp.ne.u32 p, x, 0;\n\t"
        "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
        "selp.b32 x, 0, 0xffffffff, p;\n\t"
        "lop3.b32 %0, %0, m, x, 0x60;\n\t"   // peers & (m ^ x): keeps m where the bit is set, ~m where it is clear
        "}"
        : "+r"(peers)
        : "r"(d), "n"(1u << B));
  }
}
template <int BITS>
__device__ __forceinline__ uint32_t warp_peers(uint32_t d, bool valid) {
  uint32_t peers = __ballot_sync(0xffffffffu, valid);
  peer_bits<BITS - 1>(peers, d);
  return peers;
}

// hist[d * nsuper + blockIdx.x] = number of keys with digit d among the CTA's sub x RS_TILE keys
template <typename KeyT, int BITS>
__global__ void __launch_bounds__(RS_T) k_radix_hist(const KeyT* __restrict__ keys, int64_t n, int shift,
                                                     uint32_t* __restrict__ hist, int nsuper, int sub) {
  constexpr int BINS = 1 << BITS;
  __shared__ uint32_t cnt[BINS];
  for (int t = threadIdx.x; t < BINS; t += RS_T) cnt[t] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t base = (int64_t)blockIdx.x * sub * RS_TILE;
  int64_t end = base + (int64_t)sub * RS_TILE;
  if (end > n) end = n;
  for (int64_t i0 = base + (threadIdx.x & ~31); i0 < end; i0 += RS_T) {   // i0 is warp-uniform
    const int64_t idx = i0 + lane;
    const bool valid = idx < end;
    const uint32_t d = valid ? (uint32_t)((__ldcs(keys + idx) >> shift) & (BINS - 1)) : (uint32_t)BINS;
    const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
    if (__all_sync(0xffffffffu, d == d0)) {          // sorted / clustered input: one add per warp
      if (lane == 0) atomicAdd(&cnt[d0], 32u);
    } else if (valid) {
      atomicAdd(&cnt[d], 1u);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < BINS; t += RS_T) hist[(size_t)t * nsuper + blockIdx.x] = cnt[t];
}

template <int BITS>
constexpr size_t rs_smem_bytes(size_t key_size) {
  return ((size_t)RS_NW * ((1 << BITS) / 2) + (1 << BITS) + 32) * sizeof(uint32_t) + (size_t)RS_TILE * key_size;
}

// offsets: the exclusive scan of hist (digit-major), i.e. offsets[d * nsuper + b] = first output slot of
// CTA b's keys with digit d.  vals_in == nullptr with HAS_VAL: the payload is the element index (argsort).
template <typename KeyT, int BITS, bool HAS_VAL>
__global__ void __launch_bounds__(RS_T, (sizeof(KeyT) == 4 ? 2 : 1))
    k_radix_scatter(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                    KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift,
                    const uint32_t* __restrict__ offsets, int nsuper, int sub) {
  constexpr int BINS = 1 << BITS;
  constexpr int WORDS = BINS / 2;                   // two 16-bit counters per word
  constexpr int WPT = (WORDS + RS_T - 1) / RS_T;    // counter words owned by a thread in the digit scan
  extern __shared__ __align__(16) unsigned char rs_smem[];
  uint32_t* wcnt = reinterpret_cast<uint32_t*>(rs_smem);   // [RS_NW][WORDS] per-warp digit counters
  uint32_t* gbase = wcnt + RS_NW * WORDS;                  // [BINS] global slot of tile position 0 of a digit run
  uint32_t* wsum = gbase + BINS;                           // [32] warp totals of the digit scan
  KeyT* buf = reinterpret_cast<KeyT*>(wsum + 32);          // [RS_TILE] the tile reordered by digit
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;

  // running global offsets of the digits this thread owns in the digit scan
  uint32_t gofs[2 * WPT];
#pragma unroll
  for (int j = 0; j < WPT; ++j) {
    const int word = threadIdx.x * WPT + j;
    gofs[2 * j] = gofs[2 * j + 1] = 0u;
    if (word < WORDS) {
      gofs[2 * j] = offsets[(size_t)(2 * word) * nsuper + blockIdx.x];
      gofs[2 * j + 1] = offsets[(size_t)(2 * word + 1) * nsuper + blockIdx.x];
    }
  }

#pragma unroll 1
  for (int st = 0; st < sub; ++st) {
    const int64_t tile_base = ((int64_t)blockIdx.x * sub + st) * RS_TILE;
    if (tile_base >= n) break;
    const int cnt_tile = (int)(n - tile_base < (int64_t)RS_TILE ? n - tile_base : (int64_t)RS_TILE);
    {
      uint4* z = reinterpret_cast<uint4*>(wcnt);   // 128-bit zero fill
      for (int i = threadIdx.x; i < RS_NW * WORDS / 4; i += RS_T) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    const int woff = warp * RS_WKEYS + lane;       // tile position of this lane's round-0 key
    const int rem = cnt_tile - woff;               // round r's key exists iff r * 32 < rem
    const KeyT* kp = keys_in + tile_base + woff;
    KeyT k[RS_I];
#pragma unroll
    for (int r = 0; r < RS_I; ++r) k[r] = r * 32 < rem ? __ldcs(kp + r * 32) : (KeyT)0;
    __syncthreads();

    // ---- rank inside the tile: straight-line phases ------------------------------------------
    uint32_t rk[RS_I];
#pragma unroll
    for (int h0 = 0; h0 < RS_I; h0 += RS_H) {   // RS_H rounds at a time: bounded register footprint
      uint32_t peers[RS_H], old[RS_H];
#pragma unroll
      for (int q = 0; q < RS_H; ++q) {
        const int r = h0 + q;
        const bool valid = r * 32 < rem;
        const uint32_t d = (uint32_t)((k[r] >> shift) & (BINS - 1));
        peers[q] = warp_peers<BITS>(d, valid);
      }
      // the group leader bumps the warp's digit counter; the returned old value is the rank base of the round.
      // Same-warp atomics on one address retire in program order, so rounds need no barrier between them.
#pragma unroll
      for (int q = 0; q < RS_H; ++q) {
        const int r = h0 + q;
        const bool valid = r * 32 < rem;
        const uint32_t d = (uint32_t)((k[r] >> shift) & (BINS - 1));
        old[q] = 0u;
        if (valid && lane == __ffs(peers[q]) - 1)
          old[q] = atomicAdd(&wcnt[warp * WORDS + (d >> 1)], (uint32_t)__popc(peers[q]) << ((d & 1u) * 16));
      }
#pragma unroll
      for (int q = 0; q < RS_H; ++q) {
        const int r = h0 + q;
        const bool valid = r * 32 < rem;
        const uint32_t d = (uint32_t)((k[r] >> shift) & (BINS - 1));
        const int leader = valid ? __ffs(peers[q]) - 1 : lane;
        const uint32_t o = __shfl_sync(0xffffffffu, old[q], leader);
        rk[r] = ((o >> ((d & 1u) * 16)) & 0xffffu) + (uint32_t)__popc(peers[q] & lt);
      }
    }
    __syncthreads();

    // ---- digit scan: counters -> tile positions, global bases ---------------------------------------
    {
      uint32_t tot[WPT];
      uint32_t s = 0;
#pragma unroll
      for (int j = 0; j < WPT; ++j) {
        const int word = threadIdx.x * WPT + j;
        tot[j] = 0u;
        if (word < WORDS) {
          uint32_t run = 0u;   // packed exclusive prefix over the warps (halves stay < 2^16: a tile has 8192 keys)
#pragma unroll
          for (int w = 0; w < RS_NW; ++w) {
            const uint32_t c = wcnt[w * WORDS + word];
            wcnt[w * WORDS + word] = run;
            run += c;
          }
          tot[j] = run;
          s += (run & 0xffffu) + (run >> 16);
        }
      }
      uint32_t inc = s;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
      }
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();
      if (warp == 0) {
        const uint32_t x = lane < RS_NW ? wsum[lane] : 0u;
        uint32_t xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, xi, d);
          if (lane >= d) xi += o;
        }
        if (lane < RS_NW) wsum[lane] = xi - x;
      }
      __syncthreads();
      uint32_t ex = wsum[warp] + inc - s;   // tile position of this thread's first digit run
#pragma unroll
      for (int j = 0; j < WPT; ++j) {
        const int word = threadIdx.x * WPT + j;
        if (word < WORDS) {
          const uint32_t lo = tot[j] & 0xffffu, hi = tot[j] >> 16;
          const uint32_t ds_lo = ex, ds_hi = ex + lo;
          ex += lo + hi;
          const uint32_t add = ds_lo | (ds_hi << 16);
#pragma unroll
          for (int w = 0; w < RS_NW; ++w) wcnt[w * WORDS + word] += add;
          gbase[2 * word] = gofs[2 * j] - ds_lo;
          gbase[2 * word + 1] = gofs[2 * j + 1] - ds_hi;
          gofs[2 * j] += lo;
          gofs[2 * j + 1] += hi;
        }
      }
    }
    __syncthreads();

    // ---- reorder by digit in shared memory, then write out in digit runs ----------------------------
#pragma unroll
    for (int r = 0; r < RS_I; ++r) {
      if (r * 32 < rem) {
        const uint32_t d = (uint32_t)((k[r] >> shift) & (BINS - 1));
        const uint32_t pos = ((wcnt[warp * WORDS + (d >> 1)] >> ((d & 1u) * 16)) & 0xffffu) + rk[r];
        buf[pos] = k[r];
        rk[r] = pos;
      }
    }
    __syncthreads();
    uint32_t gpos[RS_I];
#pragma unroll
    for (int j = 0; j < RS_I; ++j) {
      const int i = threadIdx.x + j * RS_T;
      gpos[j] = 0u;
      if (i < cnt_tile) {
        const KeyT key = buf[i];
        const uint32_t d = (uint32_t)((key >> shift) & (BINS - 1));
        gpos[j] = gbase[d] + (uint32_t)i;
        if (keys_out) keys_out[gpos[j]] = key;
      }
    }
    if (HAS_VAL) {
      __syncthreads();
      uint32_t* vbuf = reinterpret_cast<uint32_t*>(buf);
#pragma unroll
      for (int r = 0; r < RS_I; ++r)
        if (r * 32 < rem)
          vbuf[rk[r]] = vals_in ? __ldcs(vals_in + tile_base + woff + r * 32) : (uint32_t)(tile_base + woff) + r * 32;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < RS_I; ++j) {
        const int i = threadIdx.x + j * RS_T;
        if (i < cnt_tile) vals_out[gpos[j]] = vbuf[i];
      }
    }
    __syncthreads();   // the next tile reuses wcnt / gbase / buf
  }
}

template <typename KeyT, int BITS>
static void radix_pass(lfe_ctx* h, const KeyT* kin, const uint32_t* vin, bool has_val, KeyT* kout, uint32_t* vout,
                       int64_t n, int shift, uint32_t* hist, int nsuper, int sub, cudaStream_t s) {
  constexpr int BINS = 1 << BITS;
  constexpr size_t smem = rs_smem_bytes<BITS>(sizeof(KeyT));
  k_radix_hist<KeyT, BITS><<<nsuper, RS_T, 0, s>>>(kin, n, shift, hist, nsuper, sub);
  KERNEL_CHECK();
  prim_exclusive_scan_u32(h, hist, hist, (int64_t)BINS * nsuper, nullptr, s);
  if (has_val) {
    CUDA_CHECK(cudaFuncSetAttribute(k_radix_scatter<KeyT, BITS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    k_radix_scatter<KeyT, BITS, true><<<nsuper, RS_T, smem, s>>>(kin, vin, kout, vout, n, shift, hist, nsuper, sub);
  } else {
    CUDA_CHECK(cudaFuncSetAttribute(k_radix_scatter<KeyT, BITS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    k_radix_scatter<KeyT, BITS, false><<<nsuper, RS_T, smem, s>>>(kin, nullptr, kout, nullptr, n, shift, hist, nsuper,
                                                                 sub);
  }
  KERNEL_CHECK();
  h->n_launches += 2;
}

// Sort on key bits [bit_lo, bit_lo + key_bits).  has_val: carry a payload (vals_in == nullptr -> the element
// index, i.e. argsort); keys_out nullable when has_val.
template <typename KeyT>
static void sort_impl(lfe_ctx* h, const KeyT* keys, const uint32_t* vals_in, bool has_val, int64_t n, int bit_lo,
                      int key_bits, KeyT* keys_out, uint32_t* vals_out, cudaStream_t s) {
  if (n <= 0) return;
  const int npass = (key_bits + 10) / 11;
  if (npass == 0) {
    if (has_val) {
      if (vals_in) CUDA_CHECK(cudaMemcpyAsync(vals_out, vals_in, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
      else prim_iota_u32(vals_out, n, s);
    }
    if (keys_out) CUDA_CHECK(cudaMemcpyAsync(keys_out, keys, n * sizeof(KeyT), cudaMemcpyDeviceToDevice, s));
    h->n_launches += 1;
    return;
  }
  int bits = (key_bits + npass - 1) / npass;
  if (bits < 8) bits = 8;
  // tiles per CTA: large inputs amortise the histogram rows, small ones keep every SM busy
  const int sub = n >= (int64_t)RS_TILE * 4 * 2 * h->sm_count * 2 ? 4 : 1;
  const int nsuper = (int)ceil_div64(n, (int64_t)RS_TILE * sub);
  DevBuf<uint32_t> hist((size_t)(1 << bits) * nsuper, s);
  DevBuf<KeyT> ka(npass > 1 ? n : 0, s), kb(npass > 2 ? n : 0, s);
  DevBuf<uint32_t> va(has_val && npass > 1 ? n : 0, s);
  const KeyT* kin = keys;
  const uint32_t* vin = vals_in;
  for (int p = 0; p < npass; ++p) {
    const bool last = (p == npass - 1);
    uint32_t* vout = has_val ? ((((npass - 1 - p) & 1) == 0) ? vals_out : va.p) : nullptr;   // last pass -> vals_out
    KeyT* kout = last ? keys_out : ((p & 1) == 0 ? ka.p : kb.p);
    const int shift = bit_lo + p * bits;
    switch (bits) {
      case 8: radix_pass<KeyT, 8>(h, kin, vin, has_val, kout, vout, n, shift, hist, nsuper, sub, s); break;
      case 9: radix_pass<KeyT, 9>(h, kin, vin, has_val, kout, vout, n, shift, hist, nsuper, sub, s); break;
      case 10: radix_pass<KeyT, 10>(h, kin, vin, has_val, kout, vout, n, shift, hist, nsuper, sub, s); break;
      default: radix_pass<KeyT, 11>(h, kin, vin, has_val, kout, vout, n, shift, hist, nsuper, sub, s); break;
    }
    kin = kout;
    vin = vout;
  }
}

void prim_sort_pairs_u32(lfe_ctx* h, const uint32_t* keys, const uint32_t* vals_in, int64_t n, int key_bits,
                         uint32_t* keys_out, uint32_t* vals_out, cudaStream_t s) {
  if (vals_in == nullptr && n > 0) LFE_THROW(LFE_ERR_ARG, "prim_sort_pairs_u32: vals_in is null");
  sort_impl<uint32_t>(h, keys, vals_in, true, n, 0, key_bits, keys_out, vals_out, s);
}
void prim_sort_keys_u32(lfe_ctx* h, const uint32_t* keys, int64_t n, int bit_lo, int key_bits, uint32_t* keys_out,
                        cudaStream_t s) {
  sort_impl<uint32_t>(h, keys, nullptr, false, n, bit_lo, key_bits, keys_out, nullptr, s);
}
void prim_argsort_u32(lfe_ctx* h, const uint32_t* keys, int64_t n, int key_bits, uint32_t* perm_out,
                      uint32_t* keys_sorted_out, cudaStream_t s) {
  sort_impl<uint32_t>(h, keys, nullptr, true, n, 0, key_bits, keys_sorted_out, perm_out, s);
}
void prim_argsort_u64(lfe_ctx* h, const uint64_t* keys, int64_t n, int key_bits, uint32_t* perm_out,
                      uint64_t* keys_sorted_out, cudaStream_t s) {
  sort_impl<uint64_t>(h, keys, nullptr, true, n, 0, key_bits, keys_sorted_out, perm_out, s);
}
