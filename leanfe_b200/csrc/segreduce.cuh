// segreduce.cuh -- segmented (by-key) reduction building blocks shared by demean.cu and vcov.cu.
#pragma once
#include "common.cuh"

constexpr int PJ_T = 256;
constexpr int PJ_I = 4;  // consecutive entries per thread
constexpr int PJ_MAXNC = 8;
constexpr int S0_MAXNC = 12;   // the fused S0 pass takes up to 12 columns at once (c5: C = 11 in ONE pass over the rows)

struct ProjOthers {
  const int32_t* code[LFE_MAX_FE];
  const double* alpha[LFE_MAX_FE];
  int stage_rows[LFE_MAX_FE];  // > 0: copy that many table rows into shared memory at CTA start (small tables)
  int n;
};
constexpr int STAGE_SMEM_BYTES = 48 * 1024;

// Copy the small tables into shared memory; tabs[o] then points at the staged copy (generic address space).
__device__ __forceinline__ void stage_tables(const ProjOthers& oth, int ldc, double* smem, const double* tabs[LFE_MAX_FE]) {
  int off = 0;
  for (int o = 0; o < oth.n; ++o) {
    if (oth.stage_rows[o] > 0) {
      const int cnt = oth.stage_rows[o] * ldc;
      for (int i = threadIdx.x; i < cnt; i += blockDim.x) smem[off + i] = __ldg(oth.alpha[o] + i);
      tabs[o] = smem + off;
      off += (cnt + 1) & ~1;   // keep 16-byte alignment for the double2 gathers
    } else {
      tabs[o] = oth.alpha[o];
    }
  }
  __syncthreads();
}
struct ColPtrs {
  const double* col[S0_MAXNC];
};
struct MutColPtrs {
  double* col[PJ_MAXNC];         // destination
  const double* src[PJ_MAXNC];   // source (== col for in-place)
};

__device__ __forceinline__ void load4_i32(const int32_t* __restrict__ p, int64_t base, int64_t n, bool vec, int out[4]) {
  if (vec && base + 3 < n) {
    int4 v = __ldcs(reinterpret_cast<const int4*>(p + base));
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) out[e] = (base + e < n) ? __ldcs(p + base + e) : -1;
  }
}

// Fold PJ_I consecutive entries into the thread's open run (cur, run): equal keys accumulate, a key change
// flushes the finished run with RED.ADD.F64.  cur < 0 = no open run / padding entry.
template <int NC, int I = PJ_I>
__device__ __forceinline__ void run_fold(int& cur, double* run, const int* key, double (*v)[NC],
                                         double* __restrict__ acc, int64_t ld_row, int acc_col0, int64_t ld_col = 1) {
  bool same = true;
#pragma unroll
  for (int e = 0; e < I; ++e) same = same && (key[e] == cur);
  if (same) {   // inside a long group of a sorted list no run ends here: straight-line adds
#pragma unroll
    for (int e = 0; e < I; ++e)
#pragma unroll
      for (int c = 0; c < NC; ++c) run[c] += v[e][c];
    return;
  }
#pragma unroll
  for (int e = 0; e < I; ++e) {
    if (key[e] == cur) {
#pragma unroll
      for (int c = 0; c < NC; ++c) run[c] += v[e][c];
    } else {
      if (cur >= 0) {
        double* a = acc + (size_t)cur * ld_row + (size_t)acc_col0 * ld_col;
#pragma unroll
        for (int c = 0; c < NC; ++c) atomicAdd(a + (size_t)c * ld_col, run[c]);
      }
      cur = key[e];
#pragma unroll
      for (int c = 0; c < NC; ++c) run[c] = v[e][c];
    }
  }
}

// Segmented-reduce the threads' open runs across the warp and flush one total per (warp, key).
template <int NC>
__device__ __forceinline__ void warp_flush(int cur, double run[NC], double* __restrict__ acc, int64_t ld_row,
                                           int acc_col0, int64_t ld_col = 1) {
  const int lane = threadIdx.x & 31;
  // inclusive segmented scan of the tail runs.  Segments are maximal runs of ADJACENT lanes with equal
  // keys (head flags), so the result is exact for unsorted keys too (equal keys that are not adjacent
  // simply end up in separate atomics).
  const int prv = __shfl_up_sync(0xffffffffu, cur, 1);
  const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prv != cur);
  const int seg_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    bool add = (lane - d >= seg_start);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      double t = shfl_up_d(run[c], d);
      if (add) run[c] += t;
    }
  }
  int nxt = __shfl_down_sync(0xffffffffu, cur, 1);
  if (cur >= 0 && (lane == 31 || nxt != cur)) {
    double* a = acc + (size_t)cur * ld_row + (size_t)acc_col0 * ld_col;
#pragma unroll
    for (int c = 0; c < NC; ++c) atomicAdd(a + (size_t)c * ld_col, run[c]);
  }
}

// one batch of I (default PJ_I) entries per thread: fold + warp flush
template <int NC, int I = PJ_I>
__device__ __forceinline__ void segmented_flush(const int* key, double (*v)[NC], double* __restrict__ acc,
                                                int64_t ld_row, int acc_col0, int64_t ld_col = 1) {
  int cur = -1;
  double run[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) run[c] = 0.0;
  run_fold<NC, I>(cur, run, key, v, acc, ld_row, acc_col0, ld_col);
  warp_flush<NC>(cur, run, acc, ld_row, acc_col0, ld_col);
}


// ---------------------------------------------------------------------------------------------
// One table row added to global memory by the TMA engine: cp.reduce.async.bulk ... .add.f64 (SASS UBLKRED.ADD.F64)
// ---------------------------------------------------------------------------------------------
// NC RED.ADD.F64 per row cost NC lane-operations in the LSU (~1.3 cycles per lane, the floor of the by-key passes);
// the bulk reduce hands the whole 16-byte-multiple row to the L2 as ONE request: 50M rows x 6 columns into a
// 48 MB table take 0.64 ms instead of 1.58 ms, into a 192 MB table 1.70 instead of 2.95 ms
// (scratch/micro_tma_red.cu, same sums bit for bit).  The row is staged in a per-thread shared-memory slot
// (BR_SLOTS deep, so the next row is prepared while the engine reads the previous one).  Requirements: the
// destination row and the slot are 16-byte aligned and NC is padded to an even count (NCB).
constexpr int BR_SLOTS = 2;   // (4 slots: S0 pass 0.6 ms slower at c2 -- the extra shared memory costs occupancy)
template <int NC>
struct BulkRow {
  static constexpr int NCB = (NC + 1) & ~1;                         // doubles staged per row
  static constexpr int SLOT_DOUBLES = BR_SLOTS * NCB;               // per thread
};

template <int NC>
__device__ __forceinline__ void bulk_add_row(double* __restrict__ gdst, const double* v, double* __restrict__ my_slots,
                                             int& slot) {
  constexpr int NCB = BulkRow<NC>::NCB;
  // the slot written BR_SLOTS rows ago must have been read by the engine
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(BR_SLOTS - 1) : "memory");
  double* d = my_slots + slot * NCB;
#pragma unroll
  for (int c = 0; c < NCB; c += 2)
    *reinterpret_cast<double2*>(d + c) = make_double2(v[c], c + 1 < NC ? v[c + 1] : 0.0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(d);
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst), "r"(sa),
               "n"(NCB * 8)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  slot = (slot + 1 == BR_SLOTS) ? 0 : slot + 1;
}
// before the CTA exits (or reuses the slots for something else): all bulk reductions of this thread are complete
__device__ __forceinline__ void bulk_add_drain() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// true when no two neighbouring rows of the warp (thread-internal or across lanes) share a key: every row is its
// own run and can go out as one bulk reduce; otherwise the caller takes the run-merging RED path
template <int I = PJ_I>
__device__ __forceinline__ bool warp_keys_all_distinct_neighbours(const int* key) {
  bool dup = false;
#pragma unroll
  for (int e = 0; e + 1 < I; ++e) dup = dup || (key[e] == key[e + 1] && key[e] >= 0);
  const int prev_last = __shfl_up_sync(0xffffffffu, key[I - 1], 1);
  dup = dup || ((threadIdx.x & 31) > 0 && prev_last == key[0] && key[0] >= 0);
  return !__any_sync(0xffffffffu, dup);
}
