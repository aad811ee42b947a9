// demean.cu -- FWL alternating projections (reference: python/leanfe/polars_impl.py:490-526).
//
// B200 design.  The reference rewrites all C = 1+k columns once per FE per sweep
// (c <- c - mean_g(c)), i.e. N*F*(16C+8) bytes per sweep.  Here the iteration runs in
// FE-coefficient space: the columns stay untouched and each FE j keeps a table alpha_j[G_j x C];
// the current residual of row i is c_i - sum_j alpha_j[code_j(i)].  Projecting on FE j is then
//
//     alpha_j[g] <- ( S0_j[g] - sum_{i in g} w_i * sum_{k != j} alpha_k[code_k(i)] ) / n_j[g]
//
// with S0_j[g] = sum_{i in g} w_i c_i computed once.  In exact arithmetic this is the reference's
// Gauss-Seidel sweep (same FE order, same iterates); per sweep it reads only the int32 code columns
// (in FE-j-sorted order, built once with a radix argsort; the data rows themselves are never permuted) and
// gathers rows of the other FEs' tables (L1/L2-resident) -- no column traffic, no floating-point scatter.
// S0 is one by-key pass over the columns in their original order: shared-memory bins for low-cardinality
// FEs, run-merged RED.ADD.F64 into the L2-resident table otherwise.
// Because every FE's rows are visited in sorted order the group sums are segmented reductions:
// thread-local run compression, a warp-shuffle segmented scan, and one RED.ADD.F64 per (warp, group)
// for groups that straddle warps.  After convergence one streaming pass applies the tables to the
// columns in place.
#include <memory>

#include "common.cuh"
#include "segreduce.cuh"

#include "tail.cuh"

void comm_allreduce(lfe_ctx* h, double* buf, int64_t count, int op_max, cudaStream_t s);  // api.cu
void gram_fused_impl(lfe_ctx* h, const double* const* cols, int C, int64_t n, const double* weights, const TailTabs& T,
                     int ldc, size_t smem, double* gram_out, cudaStream_t s);                                   // gram.cu
void solve_spd_impl(lfe_ctx* h, const double* A, const double* b, int p, double* beta_out, double* Ainv_out,
                    int* used_fallback_out, cudaStream_t s);                                                    // gram.cu
void resid_vcov_fused_impl(lfe_ctx* h, const double* const* cols, int C, int64_t n, const TailTabs& T, int ldc,
                           size_t tab_smem, const double* beta_full, const double* Binv, int vcov_kind, const int32_t* cl,
                           int64_t Gcap, const double* weights, int ssc, int64_t n_obs, int64_t df_resid, double* se_out,
                           double* vcov_out, double* stats_out, int64_t* n_clusters_out, cudaStream_t s);       // vcov.cu

// what lfe_demean_plan_fit wants done with the converged tables instead of applying them to the columns
struct TailRequest {
  int vcov_kind;
  const int32_t* cl_codes;   // one-way cluster codes (row order), nullable
  int64_t n_cl_groups;
  int ssc;
  int64_t n_obs, df_resid;
  double *beta_full_out, *se_out, *vcov_out, *stats_out;   // device
  int64_t* n_clusters_out;                                  // host
};

// A = G[0:p, 0:p], b = G[0:p, p]   (G is (p+1) x (p+1): order [1, x.., y])
__global__ void k_split_gram(const double* __restrict__ G, int p, double* __restrict__ A, double* __restrict__ b) {
  const int P = p + 1;
  for (int i = threadIdx.x; i < p * p; i += blockDim.x) A[i] = G[(i / p) * P + (i % p)];
  for (int i = threadIdx.x; i < p; i += blockDim.x) b[i] = G[i * P + p];
}
// IID standard errors from the Gram alone.  G is (p+1) x (p+1) over [1, x.., y]; beta (p) solves G[:p,:p] beta = G[:p,p].
// stats = {rss, tss, rss, 0}; ok = 0 when rss is too small a difference of its terms to be trusted.
__global__ void __launch_bounds__(256) k_iid_from_gram(const double* __restrict__ G, const double* __restrict__ beta,
                                                       const double* __restrict__ B, int p, double n_obs, double df,
                                                       double* __restrict__ se, double* __restrict__ vcov,
                                                       double* __restrict__ stats, int* __restrict__ ok) {
  __shared__ double rss_s;
  const int P = p + 1, k = p - 1;
  if (threadIdx.x == 0) {
    const double yy = G[p * P + p];
    double bxy = 0.0;
    for (int a = 0; a < p; ++a) bxy += beta[a] * G[a * P + p];
    const double rss = yy - bxy;
    const double sy = G[0 * P + p];
    stats[0] = rss;
    stats[1] = yy - sy * sy / n_obs;
    stats[2] = rss;
    stats[3] = 0.0;
    rss_s = rss;
    *ok = (rss > 1e-6 * yy && rss > 0.0) ? 1 : 0;
  }
  __syncthreads();
  const double sc = rss_s / df;
  for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x) {
    const double v = sc * B[idx];
    if (vcov) vcov[idx] = v;
    if (idx / k == idx % k) se[idx / k] = sqrt(fmax(v, 0.0));
  }
}

// B = Ainv[1:, 1:]
__global__ void k_drop_intercept(const double* __restrict__ Ainv, int p, double* __restrict__ B) {
  const int k = p - 1;
  for (int i = threadIdx.x; i < k * k; i += blockDim.x) B[i] = Ainv[(i / k + 1) * p + (i % k + 1)];
}

// acc[key][acc_col0 + c] += w * sum_k alpha_k[code_k][col0 + c]    over the FE-sorted entry list
constexpr int PJ_NB = 2;  // sub-batches of PJ_I entries per thread: 8 consecutive entries, one warp scan

// One sub-batch: keys + gathered/summed rows of the other FEs' tables.  FULL = all 4 entries exist (no bounds
// or validity branches, so the compiler hoists all 128-bit gathers ahead of the first use).
// PACKED: `key` is the two-FE entry list of packed words (pk.kshift/kmask select this FE's code,
// pk.oshift/omask the other FE's), one 4-byte stream instead of two.
struct PackSpec {
  int kshift, oshift;
  uint32_t kmask, omask;
};

template <int NC, bool WEIGHTED, bool FULL, bool PACKED>
__device__ __forceinline__ void proj_batch(const int32_t* __restrict__ key, const ProjOthers& oth,
                                           const double* const tabs[LFE_MAX_FE], const double* __restrict__ w,
                                           int64_t base, int64_t n, int ldc, int col0, bool vec_codes, bool vec2,
                                           const PackSpec& pk, int k[PJ_I], double v[PJ_I][NC]) {
  int pw[PJ_I];
  load4_i32(key, base, n, vec_codes && FULL, PACKED ? pw : k);
  if (PACKED) {
#pragma unroll
    for (int e = 0; e < PJ_I; ++e)
      k[e] = (FULL || base + e < n) ? (int)(((uint32_t)pw[e] >> pk.kshift) & pk.kmask) : -1;
  }
#pragma unroll
  for (int e = 0; e < PJ_I; ++e)
#pragma unroll
    for (int c = 0; c < NC; ++c) v[e][c] = 0.0;
  for (int o = 0; o < oth.n; ++o) {
    int oc[PJ_I];
    if (PACKED) {
#pragma unroll
      for (int e = 0; e < PJ_I; ++e)
        oc[e] = (FULL || base + e < n) ? (int)(((uint32_t)pw[e] >> pk.oshift) & pk.omask) : -1;
    } else {
      load4_i32(oth.code[o], base, n, vec_codes && FULL, oc);
    }
    const double* tab = tabs[o] + col0;   // global (L2-resident) or the CTA's shared-memory copy
#pragma unroll
    for (int e = 0; e < PJ_I; ++e) {
      if (FULL || oc[e] >= 0) {
        const double* row = tab + (size_t)oc[e] * ldc;
        if (vec2) {
#pragma unroll
          for (int c = 0; c < NC / 2; ++c) {
            double2 t = *(reinterpret_cast<const double2*>(row) + c);
            v[e][2 * c] += t.x;
            v[e][2 * c + 1] += t.y;
          }
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c) v[e][c] += row[c];
        }
      }
    }
  }
  if (WEIGHTED) {
#pragma unroll
    for (int e = 0; e < PJ_I; ++e) {
      double we = (FULL || base + e < n) ? __ldcs(w + base + e) : 0.0;
#pragma unroll
      for (int c = 0; c < NC; ++c) v[e][c] *= we;
    }
  }
}

template <int NC, bool WEIGHTED, bool PACKED>
__global__ void __launch_bounds__(PJ_T, (NC <= 4 ? 4 : 2)) k_proj(const int32_t* __restrict__ key, ProjOthers oth,
                                               const double* __restrict__ w, int64_t n, int ldc, int col0,
                                               double* __restrict__ acc, int ld_acc, int acc_col0, bool vec_codes,
                                               PackSpec pk) {
  constexpr int PER_THREAD = PJ_I * PJ_NB;
  extern __shared__ double sm_tabs[];
  const double* tabs[LFE_MAX_FE];
  stage_tables(oth, ldc, sm_tabs, tabs);
  const int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PER_THREAD);
  const bool vec2 = ((NC & 1) == 0) && ((ldc & 1) == 0) && ((col0 & 1) == 0);
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    const int64_t tbase = (ch * PJ_T + threadIdx.x) * PER_THREAD;
    int cur = -1;
    double run[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) run[c] = 0.0;
    if (tbase + PER_THREAD <= n) {
#pragma unroll 1
      for (int b = 0; b < PJ_NB; ++b) {
        int k[PJ_I];
        double v[PJ_I][NC];
        proj_batch<NC, WEIGHTED, true, PACKED>(key, oth, tabs, w, tbase + b * PJ_I, n, ldc, col0, vec_codes, vec2, pk, k, v);
        run_fold<NC>(cur, run, k, v, acc, ld_acc, acc_col0);
      }
    } else {
      for (int b = 0; b < PJ_NB; ++b) {
        int k[PJ_I];
        double v[PJ_I][NC];
        proj_batch<NC, WEIGHTED, false, PACKED>(key, oth, tabs, w, tbase + b * PJ_I, n, ldc, col0, vec_codes, vec2, pk, k, v);
        run_fold<NC>(cur, run, k, v, acc, ld_acc, acc_col0);
      }
    }
    warp_flush<NC>(cur, run, acc, ld_acc, acc_col0);
  }
}

// Packed two-FE projection whose other table is SMALL (the 500-period table of c2 / c5 while projecting on unit):
// every entry gathers one table row from shared memory at random.  With the plain staged copy (k_proj) the 128-bit
// loads of a quarter-warp land on random bank quads -- row r, piece p sits on quad (3 r + p) mod 8 for 48-byte rows
// -- so a warp-wide LDS.128 takes ~8-11 wavefronts instead of 4 and the kernel is bound by the shared-memory data
// pipe (84 % busy, ncu r1).  Here every 16-byte piece is stored R times, once per bank quad: piece p of row r owns the
// R consecutive 16-byte slots (r NP + p) R .. + R - 1 and lane l reads slot (l + p) mod R.  For R = 8 the eight lanes
// of a quarter-warp therefore always hit eight different bank quads, whatever rows they gather: no conflicts, by
// construction.  (R = 4 / 2 when 8 copies do not fit: conflicts only between lanes l and l + R.)  One 512-thread CTA
// per SM owns the replicated table (500 rows x 3 pieces x 128 B = 192 KB at c2).
constexpr int PR_T = 512;
template <int NC, int R>
__global__ void __launch_bounds__(PR_T, 1) k_proj_rep(const uint32_t* __restrict__ words, const double* __restrict__ tab,
                                                      int rows, int ldc, int64_t n, double* __restrict__ acc, int ld_acc,
                                                      int acc_col0, PackSpec pk) {
  constexpr int NP = (NC + 1) / 2;       // 16-byte pieces per row (an odd NC reads the zero pad column too)
  constexpr int PER_THREAD = PJ_I * PJ_NB;
  extern __shared__ __align__(16) double2 rep[];
  for (int idx = threadIdx.x; idx < rows * NP * R; idx += PR_T) {
    const int p = (idx / R) % NP, row = idx / (R * NP);
    rep[idx] = *reinterpret_cast<const double2*>(tab + (size_t)row * ldc + 2 * p);
  }
  __syncthreads();
  const int q = threadIdx.x & (R - 1);
  const int64_t nchunk = ceil_div64(n, (int64_t)PR_T * PER_THREAD);
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    const int64_t tbase = (ch * PR_T + threadIdx.x) * PER_THREAD;
    int cur = -1;
    double run[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) run[c] = 0.0;
#pragma unroll 1
    for (int b = 0; b < PJ_NB; ++b) {
      const int64_t base = tbase + b * PJ_I;
      int pw[PJ_I], k[PJ_I];
      double v[PJ_I][NC];
      load4_i32(reinterpret_cast<const int32_t*>(words), base, n, true, pw);   // (-1 words beyond n)
#pragma unroll
      for (int e = 0; e < PJ_I; ++e) {
        const bool valid = base + e < n;
        k[e] = valid ? (int)(((uint32_t)pw[e] >> pk.kshift) & pk.kmask) : -1;
        const int oc = valid ? (int)(((uint32_t)pw[e] >> pk.oshift) & pk.omask) : 0;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const double2 t = rep[(oc * NP + p) * R + ((q + p) & (R - 1))];
          v[e][2 * p] = valid ? t.x : 0.0;
          if (2 * p + 1 < NC) v[e][2 * p + 1] = valid ? t.y : 0.0;
        }
      }
      run_fold<NC>(cur, run, k, v, acc, ld_acc, acc_col0);
    }
    warp_flush<NC>(cur, run, acc, ld_acc, acc_col0);
  }
}

// Packed two-FE projection whose other table lives in global memory (L2): lanes cooperate on the gathered rows.
// In k_proj every lane gathers its own 8*NC-byte row with up to NC/2 128-bit loads, so a warp-wide load touches
// 32 different cache lines and the kernel is bound by L1TEX wavefronts (87 % of peak at c2, ncu r1), one per
// line, ~2.5 per entry.  Here a row group of LPR = 4 (NC <= 8), 2 (NC <= 4) or 1 lanes owns an entry: lane p
// loads the 16-byte piece p of the row, so a warp-wide load covers 32 / LPR whole rows with ~1.3 lines per
// entry, and owns columns 2p, 2p+1 of the running group sum.  A row group walks PQ_E * PQ_R consecutive
// entries (words fetched 128-bit per lane and handed round with shuffles), folds equal keys in registers,
// flushes finished runs with RED.ADD.F64, and the open runs of the 32 / LPR groups are merged by a segmented
// shuffle scan before the last flush -- the same run-merging as k_proj, at quarter-warp granularity.
constexpr int PQ_E = 16;   // entries per row group per round (one 128-bit word load per lane when LPR = 4)
constexpr int PQ_R = 2;    // rounds per chunk: a row group owns 32 consecutive entries
// lanes per row: 1 (NC <= 2), 2 (NC <= 4), 4 (NC <= 8), 8 (NC <= 16: c5's 11 columns in ONE pass -- the gathers are
// row-granular, so a second column batch costs as much as the first); entries per round: every lane holds >= 4 words
template <int NC>
struct QuadCfg {
  static constexpr int LPR = NC <= 2 ? 1 : (NC <= 4 ? 2 : (NC <= 8 ? 4 : 8));
  static constexpr int E = LPR <= 4 ? PQ_E : 4 * LPR;
};

// one round of a row group: PQ_E entries folded into the open run (cur, r0, r1).  FULL: all entries exist.
// VEC: table rows are 16-byte aligned (a lane that owns two columns loads them as one double2).
template <int NC, int LPR, bool FULL, bool VEC>
__device__ __forceinline__ void quad_round(const uint32_t* __restrict__ words, const double* __restrict__ tab,
                                           int64_t ebase, int64_t n, int ldc, const PackSpec& pk, int p,
                                           int g, int c0, double* __restrict__ acc, int ld_acc, int acc_col0,
                                           int& cur, double& r0, double& r1) {
  constexpr int E = QuadCfg<NC>::E;
  constexpr int WPL = E / LPR;               // entry words held by a lane per round
  const int ncl = NC - c0 >= 2 ? 2 : (NC - c0 == 1 ? 1 : 0);   // columns this lane owns
  const int rem = FULL ? E : (int)(n - ebase < (int64_t)E ? n - ebase : (int64_t)E);   // may be <= 0
  uint32_t mine[WPL];
#pragma unroll
  for (int q = 0; q < WPL / 4; ++q) {
    const int off = p * WPL + 4 * q;
    if (FULL || off + 3 < rem) {
      const uint4 t = __ldcs(reinterpret_cast<const uint4*>(words + ebase + off));
      mine[4 * q] = t.x; mine[4 * q + 1] = t.y; mine[4 * q + 2] = t.z; mine[4 * q + 3] = t.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) mine[4 * q + e] = off + e < rem ? __ldcs(words + ebase + off + e) : 0u;
    }
  }
#pragma unroll
  for (int h0 = 0; h0 < E; h0 += 8) {
    int k[8];
    double v0[8], v1[8];
    bool same = true;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int sidx = h0 + e;
      const uint32_t w = __shfl_sync(0xffffffffu, mine[sidx % WPL], g * LPR + sidx / WPL);
      const bool valid = FULL || sidx < rem;
      k[e] = valid ? (int)((w >> pk.kshift) & pk.kmask) : -1;
      same = same && (k[e] == cur);
      const uint32_t o = valid ? ((w >> pk.oshift) & pk.omask) : 0u;
      const double* row = tab + (size_t)o * ldc + c0;
      if (VEC) {
        double2 t = make_double2(0.0, 0.0);
        if (valid && ncl == 2) t = *reinterpret_cast<const double2*>(row);
        else if (valid && ncl == 1) t.x = row[0];
        v0[e] = t.x; v1[e] = t.y;
      } else {
        v0[e] = (valid && ncl >= 1) ? row[0] : 0.0;
        v1[e] = (valid && ncl == 2) ? row[1] : 0.0;
      }
    }
    if (same) {   // the usual case inside a long group: no run ends here
#pragma unroll
      for (int e = 0; e < 8; ++e) { r0 += v0[e]; r1 += v1[e]; }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (k[e] == cur) {
          r0 += v0[e]; r1 += v1[e];
        } else {
          if (cur >= 0) {
            double* a = acc + (size_t)cur * ld_acc + acc_col0 + c0;
            if (ncl >= 1) atomicAdd(a, r0);
            if (ncl == 2) atomicAdd(a + 1, r1);
          }
          cur = k[e]; r0 = v0[e]; r1 = v1[e];
        }
      }
    }
  }
}

template <int NC, bool VEC>
__global__ void __launch_bounds__(PJ_T, 3) k_proj_q(const uint32_t* __restrict__ words, const double* __restrict__ tab,
                                                    int64_t n, int ldc, double* __restrict__ acc, int ld_acc,
                                                    int acc_col0, PackSpec pk) {
  constexpr int LPR = QuadCfg<NC>::LPR;
  constexpr int NG = 32 / LPR;
  constexpr int PQ_EE = QuadCfg<NC>::E;
  constexpr int GCH = PQ_EE * PQ_R;          // entries per row group per chunk
  constexpr int WCH = NG * GCH;              // entries per warp per chunk
  const int lane = threadIdx.x & 31;
  const int p = lane % LPR, g = lane / LPR;
  const int c0 = 2 * p;
  const bool has0 = c0 < NC, has1 = c0 + 1 < NC;
  const int64_t nchunk = ceil_div64(n, (int64_t)WCH);
  const int64_t nwarps = (int64_t)gridDim.x * (PJ_T / 32);
  for (int64_t ch = (int64_t)blockIdx.x * (PJ_T / 32) + (threadIdx.x >> 5); ch < nchunk; ch += nwarps) {
    int cur = -1;
    double r0 = 0.0, r1 = 0.0;
    const int64_t gbase = ch * WCH + (int64_t)g * GCH;
    if ((ch + 1) * WCH <= n) {
#pragma unroll 1
      for (int r = 0; r < PQ_R; ++r)
        quad_round<NC, LPR, true, VEC>(words, tab, gbase + (int64_t)r * PQ_EE, n, ldc, pk, p, g, c0, acc, ld_acc,
                                       acc_col0, cur, r0, r1);
    } else {
#pragma unroll 1
      for (int r = 0; r < PQ_R; ++r)
        quad_round<NC, LPR, false, VEC>(words, tab, gbase + (int64_t)r * PQ_EE, n, ldc, pk, p, g, c0, acc, ld_acc,
                                        acc_col0, cur, r0, r1);
    }
    // merge the open runs of adjacent row groups (same p): inclusive segmented scan over the groups
    const int prv = __shfl_up_sync(0xffffffffu, cur, LPR);
    const unsigned heads = __ballot_sync(0xffffffffu, g == 0 || prv != cur);
    const int seg_start = (31 - __clz(heads & (0xffffffffu >> (31 - lane)))) / LPR;   // first group of my segment
#pragma unroll
    for (int d = 1; d < NG; d <<= 1) {
      const double t0 = shfl_up_d(r0, d * LPR), t1 = shfl_up_d(r1, d * LPR);
      if (g - d >= seg_start) { r0 += t0; r1 += t1; }
    }
    const int nxt = __shfl_down_sync(0xffffffffu, cur, LPR);
    if (cur >= 0 && (g == NG - 1 || nxt != cur)) {
      double* a = acc + (size_t)cur * ld_acc + acc_col0 + c0;
      if (has0) atomicAdd(a, r0);
      if (has1) atomicAdd(a + 1, r1);
    }
  }
}

// Projection on a VERY low-cardinality FE j (G_j x NC <= PT_MAX_BINS doubles, e.g. 30 years x 3 columns) taken
// while streaming the entry list of the BIG FE b instead of FE j's own list.  In j's own order every entry
// gathers a row of b's table at random (c3: a 240 MB worker table walked 30 times from DRAM, 2.7 ms per
// projection); in b's order that table is read once, sequentially (neighbouring lanes sit in neighbouring
// groups of b, so a warp-wide row load touches a handful of lines), and the 30 x 3 sums live in THREAD-PRIVATE
// shared-memory bins (layout [bin][thread]: conflict-free, no atomics), reduced per CTA at the end.
constexpr int PT_T = 256;
constexpr int PT_MAX_BINS = 96;   // doubles per thread: 96 x 256 x 8 B = 192 KB of shared memory per CTA

// entries per thread per chunk: all code loads, then all table gathers of a chunk are issued back to back (the
// kernel runs 8 warps per SM, so the memory parallelism has to come from inside the thread)
template <int NC>
struct TinyCfg { static constexpr int E = NC <= 3 ? 16 : (NC <= 6 ? 8 : 4); };

template <int NC, bool WEIGHTED, bool PACKED>
__global__ void __launch_bounds__(PT_T, 1) k_proj_tiny(const int32_t* __restrict__ key, const int32_t* __restrict__ code_j,
                                                       ProjOthers oth, const double* __restrict__ tab_b,
                                                       const double* __restrict__ w, int64_t n, int ldc, int col0,
                                                       int Gj, double* __restrict__ acc, int ld_acc, int acc_col0,
                                                       bool vec_codes, PackSpec pk) {
  constexpr int E = TinyCfg<NC>::E;
  extern __shared__ double priv[];   // [Gj * NC][PT_T]
  const int nbins = Gj * NC;
  for (int i = 0; i < nbins; ++i) priv[i * PT_T + threadIdx.x] = 0.0;   // thread-private: no barrier needed
  // 128-bit gathers: even NC, or an odd NC whose batch ends at the zero pad column of an even-width table
  const bool vec2 = ((ldc & 1) == 0) && ((col0 & 1) == 0) && (((NC & 1) == 0) || col0 + NC + 1 <= ldc);
  const int64_t nchunk = ceil_div64(n, (int64_t)PT_T * E);
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    const int64_t base = (ch * PT_T + threadIdx.x) * E;
    const bool full = base + E <= n;
    int kb[E], kj[E];
#pragma unroll
    for (int q = 0; q < E / 4; ++q) load4_i32(key, base + 4 * q, n, vec_codes && full, kb + 4 * q);
    if (PACKED) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const uint32_t word = (uint32_t)kb[e];
        const bool valid = base + e < n;
        kb[e] = valid ? (int)((word >> pk.kshift) & pk.kmask) : -1;
        kj[e] = valid ? (int)((word >> pk.oshift) & pk.omask) : -1;
      }
    } else {
#pragma unroll
      for (int q = 0; q < E / 4; ++q) load4_i32(code_j, base + 4 * q, n, vec_codes && full, kj + 4 * q);
    }
    double v[E][NC];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const double* row = tab_b + (size_t)(kb[e] >= 0 ? kb[e] : 0) * ldc + col0;
      if (vec2) {
#pragma unroll
        for (int c = 0; c < (NC + 1) / 2; ++c) {
          const double2 t = *(reinterpret_cast<const double2*>(row) + c);
          v[e][2 * c] = t.x;
          if (2 * c + 1 < NC) v[e][2 * c + 1] = t.y;
        }
      } else {
#pragma unroll
        for (int c = 0; c < NC; ++c) v[e][c] = row[c];
      }
    }
    for (int o = 0; o < oth.n; ++o) {
      int oc[E];
#pragma unroll
      for (int q = 0; q < E / 4; ++q) load4_i32(oth.code[o], base + 4 * q, n, vec_codes && full, oc + 4 * q);
      const double* tab = oth.alpha[o] + col0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const double* row = tab + (size_t)(oc[e] >= 0 ? oc[e] : 0) * ldc;
        if (vec2) {
#pragma unroll
          for (int c = 0; c < (NC + 1) / 2; ++c) {
            const double2 t = *(reinterpret_cast<const double2*>(row) + c);
            v[e][2 * c] += t.x;
            if (2 * c + 1 < NC) v[e][2 * c + 1] += t.y;
          }
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c) v[e][c] += row[c];
        }
      }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (kj[e] >= 0) {
        double we = 1.0;
        if (WEIGHTED) we = __ldcs(w + base + e);
        double* b = priv + (size_t)kj[e] * NC * PT_T + threadIdx.x;
#pragma unroll
        for (int c = 0; c < NC; ++c) b[c * PT_T] += WEIGHTED ? v[e][c] * we : v[e][c];
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < nbins; i += PT_T / 32) {
    double sum = 0.0;
#pragma unroll
    for (int t = lane; t < PT_T; t += 32) sum += priv[i * PT_T + t];
    sum = warp_sum(sum);
    if (lane == 0 && sum != 0.0) atomicAdd(acc + (size_t)(i / NC) * ld_acc + acc_col0 + (i % NC), sum);
  }
}

// acc[key][acc_col0 + c] += w * col_c[perm ? perm[e] : e]     (S0 tables; also sum of weights with NC=1)
template <int NC, bool WEIGHTED>
__global__ void __launch_bounds__(PJ_T) k_segsum(const int32_t* __restrict__ key, ColPtrs cols,
                                                 const uint32_t* __restrict__ perm, const double* __restrict__ w,
                                                 int64_t n, double* __restrict__ acc, int ld_acc, int acc_col0,
                                                 bool vec_codes) {
  const int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I);
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    const int64_t base = (ch * PJ_T + threadIdx.x) * PJ_I;
    int k[PJ_I];
    load4_i32(key, base, n, vec_codes, k);
    double v[PJ_I][NC];
#pragma unroll
    for (int e = 0; e < PJ_I; ++e) {
      bool valid = base + e < n;
      int64_t src = valid ? (perm ? (int64_t)__ldcs(perm + base + e) : base + e) : 0;
      double we = 1.0;
      if (WEIGHTED) we = valid ? __ldcs(w + base + e) : 0.0;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double x = valid ? (perm ? __ldg(cols.col[c] + src) : __ldcs(cols.col[c] + src)) : 0.0;
        v[e][c] = WEIGHTED ? x * we : x;
      }
    }
    segmented_flush<NC>(k, v, acc, ld_acc, acc_col0);
  }
}

// ---- deterministic mode (LFE_OPT_DETERMINISTIC): group sums in a FIXED order, no floating-point atomics -----------
// One warp owns one group of the FE-sorted entry list, entries [off[g], off[g+1]): lane l adds entries l, l + 32, ...
// sequentially, then a fixed xor-shuffle tree combines the 32 partials and lane 0 STORES the sum.  Every bit of the
// result is a function of the data and the list order alone, so two runs agree exactly.  (Slower than the
// run-merging kernels for long groups -- one warp walks a whole group -- and meant for tests and audits.)
// acc[g][acc_col0 + c] = sum_e w_e sum_k alpha_k[oth_k[e]][col0 + c]
template <int NC, bool WEIGHTED>
__global__ void __launch_bounds__(256) k_det_proj(const uint32_t* __restrict__ off, int64_t G, ProjOthers oth,
                                                  const double* __restrict__ w, int ldc, int col0,
                                                  double* __restrict__ acc, int ld_acc, int acc_col0) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < G; g += nwarp) {
    double sum[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) sum[c] = 0.0;
    const uint32_t e1 = off[g + 1];
    for (uint32_t e = off[g] + lane; e < e1; e += 32) {
      double v[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) v[c] = 0.0;
      for (int o = 0; o < oth.n; ++o) {
        const double* row = oth.alpha[o] + (size_t)oth.code[o][e] * ldc + col0;
#pragma unroll
        for (int c = 0; c < NC; ++c) v[c] += row[c];
      }
      const double we = WEIGHTED ? w[e] : 1.0;
#pragma unroll
      for (int c = 0; c < NC; ++c) sum[c] += WEIGHTED ? v[c] * we : v[c];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) sum[c] = warp_sum(sum[c]);
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[(size_t)g * ld_acc + acc_col0 + c] = sum[c];
    }
  }
}
// acc[g][acc_col0 + c] = sum_e w[row_e] col_c[row_e],  row_e = perm ? perm[e] : e
template <int NC, bool WEIGHTED>
__global__ void __launch_bounds__(256) k_det_s0(const uint32_t* __restrict__ off, int64_t G,
                                                const uint32_t* __restrict__ perm, ColPtrs cols,
                                                const double* __restrict__ w_rows, double* __restrict__ acc, int ld_acc,
                                                int acc_col0) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < G; g += nwarp) {
    double sum[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) sum[c] = 0.0;
    const uint32_t e1 = off[g + 1];
    for (uint32_t e = off[g] + lane; e < e1; e += 32) {
      const size_t row = perm ? perm[e] : e;
      const double we = WEIGHTED ? w_rows[row] : 1.0;
#pragma unroll
      for (int c = 0; c < NC; ++c) sum[c] += WEIGHTED ? cols.col[c][row] * we : cols.col[c][row];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) sum[c] = warp_sum(sum[c]);
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[(size_t)g * ld_acc + acc_col0 + c] = sum[c];
    }
  }
}

// alpha <- (S0 - acc) / nsum ; acc <- 0
// Low-cardinality FE: per-CTA privatised fp64 bins in shared memory (G*NC*8 bytes), rows in their original
// order with coalesced column loads; the CTA's bins are flushed with one RED.ADD.F64 per non-zero bin.
template <int NC, bool WEIGHTED>
__global__ void __launch_bounds__(256) k_binsum_smem(const int32_t* __restrict__ key, ColPtrs cols,
                                                     const double* __restrict__ w, int64_t n, int G,
                                                     double* __restrict__ acc, int ld_acc, int acc_col0) {
  extern __shared__ double bins[];
  for (int i = threadIdx.x; i < G * NC; i += blockDim.x) bins[i] = 0.0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int g = __ldcs(key + i);
    double we = 1.0;
    if (WEIGHTED) we = __ldcs(w + i);
    double* b = bins + (size_t)g * NC;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      double x = __ldcs(cols.col[c] + i);
      atomicAdd(b + c, WEIGHTED ? x * we : x);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * NC; i += blockDim.x) {
    double v = bins[i];
    if (v != 0.0) atomicAdd(acc + (size_t)(i / NC) * ld_acc + acc_col0 + (i % NC), v);
  }
}

// S0 of ALL fixed effects in one pass over the columns (they are read once instead of once per FE): every
// thread takes PJ_I consecutive rows; an FE whose G x NC bins fit the CTA's shared memory accumulates there
// (smem_off >= 0), the others take the run-merged RED path of k_segsum.
struct S0Spec {
  const int32_t* code[LFE_MAX_FE];
  double* acc[LFE_MAX_FE];
  int G[LFE_MAX_FE];
  int smem_off[LFE_MAX_FE];   // offset (doubles) of the FE's bins in shared memory, -1 = RED into acc
  int F;
  int smem_doubles;
};

template <int NC, bool WEIGHTED>
__global__ void __launch_bounds__(PJ_T) k_s0_fused(S0Spec sp, ColPtrs cols, const double* __restrict__ w, int64_t n,
                                                   int ld_acc, int acc_col0, bool vec, bool bulk) {
  extern __shared__ __align__(16) double bins[];
  for (int i = threadIdx.x; i < sp.smem_doubles; i += blockDim.x) bins[i] = 0.0;
  __syncthreads();
  // bulk: rows of the big tables go out as one TMA reduce each (see bulk_add_row); slots follow the bins
  double* my_slots = bins + ((sp.smem_doubles + 1) & ~1) + (size_t)threadIdx.x * BulkRow<NC>::SLOT_DOUBLES;
  int slot = 0;
  const int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I);
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    const int64_t base = (ch * PJ_T + threadIdx.x) * PJ_I;
    double v[PJ_I][NC];
    if (vec && base + PJ_I <= n) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double2 a = __ldcs(reinterpret_cast<const double2*>(cols.col[c] + base));
        const double2 b = __ldcs(reinterpret_cast<const double2*>(cols.col[c] + base) + 1);
        v[0][c] = a.x; v[1][c] = a.y; v[2][c] = b.x; v[3][c] = b.y;
      }
      if (WEIGHTED) {
        const double2 a = __ldcs(reinterpret_cast<const double2*>(w + base));
        const double2 b = __ldcs(reinterpret_cast<const double2*>(w + base) + 1);
        const double we[PJ_I] = {a.x, a.y, b.x, b.y};
#pragma unroll
        for (int e = 0; e < PJ_I; ++e)
#pragma unroll
          for (int c = 0; c < NC; ++c) v[e][c] *= we[e];
      }
    } else {
#pragma unroll
      for (int e = 0; e < PJ_I; ++e) {
        const bool valid = base + e < n;
        const double we = WEIGHTED ? (valid ? __ldcs(w + base + e) : 0.0) : 1.0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const double x = valid ? __ldcs(cols.col[c] + base + e) : 0.0;
          v[e][c] = WEIGHTED ? x * we : x;
        }
      }
    }
    for (int j = 0; j < sp.F; ++j) {
      int k[PJ_I];
      load4_i32(sp.code[j], base, n, vec, k);
      if (sp.smem_off[j] >= 0) {
        double* b0 = bins + sp.smem_off[j];
#pragma unroll
        for (int e = 0; e < PJ_I; ++e) {
          if (k[e] >= 0) {
            double* b = b0 + (size_t)k[e] * NC;
#pragma unroll
            for (int c = 0; c < NC; ++c) atomicAdd(b + c, v[e][c]);
          }
        }
      } else if (bulk && warp_keys_all_distinct_neighbours(k)) {
#pragma unroll
        for (int e = 0; e < PJ_I; ++e)
          if (k[e] >= 0) bulk_add_row<NC>(sp.acc[j] + (size_t)k[e] * ld_acc + acc_col0, v[e], my_slots, slot);
      } else {
        segmented_flush<NC>(k, v, sp.acc[j], ld_acc, acc_col0);
      }
    }
  }
  if (bulk) bulk_add_drain();
  __syncthreads();
  for (int j = 0; j < sp.F; ++j) {
    if (sp.smem_off[j] < 0) continue;
    const double* b0 = bins + sp.smem_off[j];
    for (int i = threadIdx.x; i < sp.G[j] * NC; i += blockDim.x) {
      const double x = b0[i];
      if (x != 0.0) atomicAdd(sp.acc[j] + (size_t)(i / NC) * ld_acc + acc_col0 + (i % NC), x);
    }
  }
}

// delta_max (nullable): max |alpha_new - alpha_old| = largest weighted group mean of the residual just removed
// alpha_out may alias alpha_in (in place) or be a staging table (speculative projection, see the loop);
// delta_col < 0: the max runs over all columns, else over that column only (the y-only stopping test).
__global__ void __launch_bounds__(256) k_finalize(const double* alpha_in, double* alpha_out, double* __restrict__ acc,
                                                  const double* __restrict__ S0, const double* __restrict__ nsum,
                                                  int64_t G, int C, double* __restrict__ delta_max, int delta_col) {
  const int64_t tot = G * C;
  double m = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
    double nn = nsum[i / C];
    double a = nn > 0.0 ? (S0[i] - acc[i]) / nn : 0.0;
    if (delta_max && nn > 0.0 && (delta_col < 0 || (int)(i % C) == delta_col)) m = fmax(m, fabs(a - alpha_in[i]));
    alpha_out[i] = a;
    acc[i] = 0.0;
  }
  if (delta_max) {
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomic_max_nonneg(delta_max, m);
  }
}

// max_g | (S0u[g] - acc[g]) / cnt[g] - alpha[g][0] |   (polars_impl.py:511-521: y only, unweighted)
__global__ void __launch_bounds__(256) k_check(const double* __restrict__ alpha, double* __restrict__ acc,
                                               const double* __restrict__ S0u, int ld_s0,
                                               const double* __restrict__ cnt, int64_t G, int C,
                                               double* __restrict__ out_max) {
  double m = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x) {
    double nn = cnt[g];
    if (nn > 0.0) {
      double mean = (S0u[g * ld_s0] - acc[g]) / nn - alpha[g * C];
      m = fmax(m, fabs(mean));
    }
    acc[g] = 0.0;
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomic_max_nonneg(out_max, m);
}

__global__ void __launch_bounds__(256) k_u32_to_f64(const uint32_t* __restrict__ in, int64_t n,
                                                    double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (double)in[i];
}

// words[i] = (big[i] << bs) | small[i]
__global__ void __launch_bounds__(256) k_pack_codes(const int32_t* __restrict__ big, const int32_t* __restrict__ small,
                                                    int64_t n, int bs, uint32_t* __restrict__ words) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    __stcs(words + i, ((uint32_t)__ldcs(big + i) << bs) | (uint32_t)__ldcs(small + i));
}

// cols[c][i] -= sum_j alpha_j[code_j[i]][col0 + c]
typedef ProjOthers ApplyTabs;   // code[j] in ROW order here; small tables are staged in shared memory too
template <int NC>
__global__ void __launch_bounds__(256) k_apply(MutColPtrs cols, ApplyTabs T, int64_t n, int ldc, int col0) {
  extern __shared__ double sm_tabs[];
  const double* tabs[LFE_MAX_FE];
  stage_tables(T, ldc, sm_tabs, tabs);
  const bool vec2 = ((NC & 1) == 0) && ((ldc & 1) == 0) && ((col0 & 1) == 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double s[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) s[c] = 0.0;
    for (int j = 0; j < T.n; ++j) {
      const double* row = tabs[j] + (size_t)__ldcs(T.code[j] + i) * ldc + col0;
      if (vec2) {
#pragma unroll
        for (int c = 0; c < NC / 2; ++c) {
          double2 t = *(reinterpret_cast<const double2*>(row) + c);
          s[2 * c] += t.x;
          s[2 * c + 1] += t.y;
        }
      } else {
#pragma unroll
        for (int c = 0; c < NC; ++c) s[c] += row[c];
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      double x = __ldcs(cols.src[c] + i);
      __stcs(cols.col[c] + i, x - s[c]);
    }
  }
}

// Same, four consecutive rows per thread with 128-bit code / column accesses: the per-row chain code load ->
// table gather -> store left k_apply 91 % long-scoreboard-stalled at half occupancy; four independent chains
// per thread give the memory system 4x the requests in flight.  Needs 16-byte aligned columns and codes.
template <int NC>
__global__ void __launch_bounds__(256, 2) k_apply4(MutColPtrs cols, ApplyTabs T, int64_t n, int ldc, int col0) {
  extern __shared__ double sm_tabs[];
  const double* tabs[LFE_MAX_FE];
  stage_tables(T, ldc, sm_tabs, tabs);
  const bool vec2 = ((NC & 1) == 0) && ((ldc & 1) == 0) && ((col0 & 1) == 0);
  const int64_t n4 = n >> 2;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (int64_t)gridDim.x * blockDim.x) {
    // all column loads first: they depend on nothing, and (dst may alias src) none of them may be ordered after
    // a store of this iteration; the gathered table rows are then subtracted in place
    double2 xa[NC], xb[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      xa[c] = __ldcs(reinterpret_cast<const double2*>(cols.src[c]) + 2 * g);
      xb[c] = __ldcs(reinterpret_cast<const double2*>(cols.src[c]) + 2 * g + 1);
    }
    for (int j = 0; j < T.n; ++j) {
      const int4 k4 = __ldcs(reinterpret_cast<const int4*>(T.code[j]) + g);
      const double* tab = tabs[j] + col0;
      const double* r0 = tab + (size_t)k4.x * ldc;
      const double* r1 = tab + (size_t)k4.y * ldc;
      const double* r2 = tab + (size_t)k4.z * ldc;
      const double* r3 = tab + (size_t)k4.w * ldc;
      if (vec2) {
#pragma unroll
        for (int c = 0; c < NC / 2; ++c) {
          const double2 t0 = *(reinterpret_cast<const double2*>(r0) + c);
          const double2 t1 = *(reinterpret_cast<const double2*>(r1) + c);
          const double2 t2 = *(reinterpret_cast<const double2*>(r2) + c);
          const double2 t3 = *(reinterpret_cast<const double2*>(r3) + c);
          xa[2 * c].x -= t0.x; xa[2 * c + 1].x -= t0.y;
          xa[2 * c].y -= t1.x; xa[2 * c + 1].y -= t1.y;
          xb[2 * c].x -= t2.x; xb[2 * c + 1].x -= t2.y;
          xb[2 * c].y -= t3.x; xb[2 * c + 1].y -= t3.y;
        }
      } else {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          xa[c].x -= r0[c];
          xa[c].y -= r1[c];
          xb[c].x -= r2[c];
          xb[c].y -= r3[c];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      __stcs(reinterpret_cast<double2*>(cols.col[c]) + 2 * g, xa[c]);
      __stcs(reinterpret_cast<double2*>(cols.col[c]) + 2 * g + 1, xb[c]);
    }
  }
  // the last n % 4 rows
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    double s[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) s[c] = 0.0;
    for (int j = 0; j < T.n; ++j) {
      const double* row = tabs[j] + (size_t)T.code[j][i] * ldc + col0;
#pragma unroll
      for (int c = 0; c < NC; ++c) s[c] += row[c];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) cols.col[c][i] = cols.src[c][i] - s[c];
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

struct FEState {
  int64_t G = 0;
  bool identity = false;
  DevBuf<uint32_t> perm;
  DevBuf<int32_t> key_buf;
  const int32_t* key = nullptr;
  DevBuf<int32_t> oth_buf[LFE_MAX_FE];
  const int32_t* oth[LFE_MAX_FE] = {nullptr};
  bool packed = false;   // key_buf holds packed (big code << bs | small code) words, see demean_plan_create
  PackSpec pk = {0, 0, 0u, 0u};
  DevBuf<double> w_buf;
  const double* w = nullptr;
  DevBuf<double> alpha, S0, acc, nsum, cnt_buf, S0u_buf, acc_chk;
  DevBuf<uint32_t> off;   // deterministic mode: CSR offsets of the groups in the sorted entry list (G + 1)
  const double* cnt = nullptr;  // unweighted counts
  const double* S0u = nullptr;  // unweighted y sums
  int ld_s0u = 1;
};

// decide which tables are small enough to be staged in shared memory; returns the dynamic smem bytes
static size_t plan_staging(ProjOthers& o, const int64_t* rows, int ldc) {
  size_t used = 0;
  for (int q = 0; q < o.n; ++q) {
    size_t bytes = (((size_t)rows[q] * ldc + 1) & ~(size_t)1) * sizeof(double);
    if (rows[q] > 0 && used + bytes <= (size_t)STAGE_SMEM_BYTES) {
      o.stage_rows[q] = (int)rows[q];
      used += bytes;
    } else {
      o.stage_rows[q] = 0;
    }
  }
  return used;
}

template <bool WEIGHTED>
static void launch_proj(lfe_ctx* h, int nc, const FEState& st, const ProjOthers& oth, int64_t n, int ldc, int col0,
                        double* acc, int ld_acc, int acc_col0, size_t smem, cudaStream_t s) {
  int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I * PJ_NB);
  int blocks = launch_blocks(h, nchunk, nc <= 4 ? 4 : 2);   // persistent: 2-3 CTAs per SM (register-limited), tables staged once
  bool vec = aligned16(st.key);
  if (!st.packed)
    for (int o = 0; o < oth.n; ++o) vec = vec && aligned16(oth.code[o]);
  if (st.packed && oth.n == 1 && oth.stage_rows[0] == 0 && n >= 4096) {
    // the other FE's table is too large for shared memory: row-group cooperative gathers
    const int64_t nch = ceil_div64(n, (int64_t)(PJ_T / 32) * 32 * PQ_E * PQ_R / 4);   // >= CTA chunks for any LPR
    const int qb = launch_blocks(h, nch, 3);
    const bool vec2 = ((ldc & 1) == 0) && ((col0 & 1) == 0) && aligned16(oth.alpha[0]);
#define PQ_CASE(NC_)                                                                                            \
  case NC_:                                                                                                     \
    if (vec2)                                                                                                   \
      k_proj_q<NC_, true><<<qb, PJ_T, 0, s>>>((const uint32_t*)st.key, oth.alpha[0] + col0, n, ldc, acc, ld_acc, \
                                              acc_col0, st.pk);                                                 \
    else                                                                                                        \
      k_proj_q<NC_, false><<<qb, PJ_T, 0, s>>>((const uint32_t*)st.key, oth.alpha[0] + col0, n, ldc, acc, ld_acc, \
                                               acc_col0, st.pk);                                                \
    break;
    switch (nc) {
      PQ_CASE(1) PQ_CASE(2) PQ_CASE(3) PQ_CASE(4) PQ_CASE(5) PQ_CASE(6) PQ_CASE(7) PQ_CASE(8)
      PQ_CASE(9) PQ_CASE(10) PQ_CASE(11) PQ_CASE(12)
      default: LFE_THROW(LFE_ERR_ARG, "launch_proj: nc=%d", nc);
    }
#undef PQ_CASE
    KERNEL_CHECK();
    h->n_launches += 1;
    h->n_sweep_launches += 1;
    return;
  }
  if (st.packed && oth.n == 1 && oth.stage_rows[0] > 0 && n >= 65536 && (ldc & 1) == 0 && (col0 & 1) == 0 &&
      aligned16(oth.alpha[0]) && aligned16(st.key) && col0 + 2 * ((nc + 1) / 2) <= ldc) {
    // small other table: bank-conflict-free replicated copy in shared memory (k_proj_rep)
    static const bool no_rep = getenv("LFE_NO_PROJ_REP") != nullptr;
    const int rows = oth.stage_rows[0], np = (nc + 1) / 2;
    const size_t budget = (size_t)(h->smem_optin > 0 ? h->smem_optin : 0) - 1024;
    int R = 8;
    while (R >= 2 && (size_t)rows * np * R * 16 > budget) R >>= 1;
    if (R >= 2 && !no_rep) {
      const size_t rsmem = (size_t)rows * np * R * 16;
      const int rb = launch_blocks(h, ceil_div64(n, (int64_t)PR_T * PJ_I * PJ_NB), 1);
#define PR_LAUNCH(NC_, R_)                                                                                           \
  do {                                                                                                               \
    CUDA_CHECK(cudaFuncSetAttribute(k_proj_rep<NC_, R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem)); \
    k_proj_rep<NC_, R_><<<rb, PR_T, rsmem, s>>>((const uint32_t*)st.key, oth.alpha[0] + col0, rows, ldc, n, acc,    \
                                                ld_acc, acc_col0, st.pk);                                            \
  } while (0)
#define PR_CASE(NC_)                                                                                  \
  case NC_:                                                                                           \
    if (R == 8) PR_LAUNCH(NC_, 8); else if (R == 4) PR_LAUNCH(NC_, 4); else PR_LAUNCH(NC_, 2);        \
    break;
      switch (nc) {
        PR_CASE(1) PR_CASE(2) PR_CASE(3) PR_CASE(4) PR_CASE(5) PR_CASE(6) PR_CASE(7) PR_CASE(8)
        default: LFE_THROW(LFE_ERR_ARG, "launch_proj: nc=%d", nc);
      }
#undef PR_CASE
#undef PR_LAUNCH
      KERNEL_CHECK();
      h->n_launches += 1;
      h->n_sweep_launches += 1;
      return;
    }
  }
#define PJ_CASE(NC_)                                                                                          \
  case NC_:                                                                                                   \
    if (st.packed)                                                                                            \
      k_proj<NC_, false, true><<<blocks, PJ_T, smem, s>>>(st.key, oth, st.w, n, ldc, col0, acc, ld_acc,      \
                                                          acc_col0, vec, st.pk);                             \
    else                                                                                                      \
      k_proj<NC_, WEIGHTED, false><<<blocks, PJ_T, smem, s>>>(st.key, oth, st.w, n, ldc, col0, acc, ld_acc,  \
                                                              acc_col0, vec, st.pk);                         \
    break;
  switch (nc) {
    PJ_CASE(1) PJ_CASE(2) PJ_CASE(3) PJ_CASE(4) PJ_CASE(5) PJ_CASE(6) PJ_CASE(7) PJ_CASE(8)
    default: LFE_THROW(LFE_ERR_ARG, "launch_proj: nc=%d", nc);
  }
#undef PJ_CASE
  KERNEL_CHECK();
  h->n_launches += 1;
  h->n_sweep_launches += 1;
}

// projection on the tiny FE j over the entry list of the big FE b (see k_proj_tiny)
template <bool WEIGHTED>
static void launch_proj_tiny(lfe_ctx* h, int nc, const std::vector<FEState>& st, int b, int j, int F, int64_t n, int ldc,
                             int col0, double* acc, int ld_acc, int acc_col0, cudaStream_t s) {
  const FEState& B = st[b];
  ProjOthers o;
  o.n = 0;
  bool vec = aligned16(B.key);
  if (!B.packed) {
    vec = vec && aligned16(B.oth[j]);
    for (int k = 0; k < F; ++k) {
      if (k == b || k == j) continue;
      o.code[o.n] = B.oth[k];
      o.alpha[o.n] = st[k].alpha;
      o.stage_rows[o.n] = 0;
      vec = vec && aligned16(B.oth[k]);
      o.n++;
    }
  }
  const int Gj = (int)st[j].G;
  const size_t smem = (size_t)Gj * nc * PT_T * sizeof(double);
  const int blocks = launch_blocks(h, ceil_div64(n, (int64_t)PT_T * PJ_I), 1);
#define PT_CASE(NC_)                                                                                                 \
  case NC_:                                                                                                          \
    if (B.packed) {                                                                                                  \
      CUDA_CHECK(cudaFuncSetAttribute(k_proj_tiny<NC_, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                      PT_MAX_BINS * PT_T * (int)sizeof(double)));                                    \
      k_proj_tiny<NC_, false, true><<<blocks, PT_T, smem, s>>>(B.key, nullptr, o, st[b].alpha, nullptr, n, ldc,     \
                                                               col0, Gj, acc, ld_acc, acc_col0, vec, B.pk);          \
    } else {                                                                                                         \
      CUDA_CHECK(cudaFuncSetAttribute(k_proj_tiny<NC_, WEIGHTED, false>,                                             \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,                                   \
                                      PT_MAX_BINS * PT_T * (int)sizeof(double)));                                    \
      k_proj_tiny<NC_, WEIGHTED, false><<<blocks, PT_T, smem, s>>>(B.key, B.oth[j], o, st[b].alpha, B.w, n, ldc,    \
                                                                   col0, Gj, acc, ld_acc, acc_col0, vec, B.pk);      \
    }                                                                                                                \
    break;
  switch (nc) {
    PT_CASE(1) PT_CASE(2) PT_CASE(3) PT_CASE(4) PT_CASE(5) PT_CASE(6) PT_CASE(7) PT_CASE(8)
    default: LFE_THROW(LFE_ERR_ARG, "launch_proj_tiny: nc=%d", nc);
  }
#undef PT_CASE
  KERNEL_CHECK();
  h->n_launches += 1;
  h->n_sweep_launches += 1;
}

template <bool WEIGHTED>
static void launch_segsum(lfe_ctx* h, int nc, const int32_t* key, const ColPtrs& cols, const uint32_t* perm,
                          const double* w, int64_t n, double* acc, int ld_acc, int acc_col0, cudaStream_t s) {
  int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I);
  int blocks = launch_blocks(h, nchunk, 8);
  bool vec = aligned16(key);
#define SS_CASE(NC_)                                                                                        \
  case NC_:                                                                                                 \
    k_segsum<NC_, WEIGHTED><<<blocks, PJ_T, 0, s>>>(key, cols, perm, w, n, acc, ld_acc, acc_col0, vec);    \
    break;
  switch (nc) {
    SS_CASE(1) SS_CASE(2) SS_CASE(3) SS_CASE(4) SS_CASE(5) SS_CASE(6) SS_CASE(7) SS_CASE(8)
    default: LFE_THROW(LFE_ERR_ARG, "launch_segsum: nc=%d", nc);
  }
#undef SS_CASE
  KERNEL_CHECK();
  h->n_launches += 1;
}

constexpr int BINSUM_MAX_DOUBLES = 8192;  // 64 KB of shared-memory bins -> 3 CTAs per SM

// group sums of columns in ORIGINAL row order: shared-memory bins for small G, run-merged RED otherwise
template <bool WEIGHTED>
static void launch_groupsum(lfe_ctx* h, int nc, const int32_t* key, int64_t G, const ColPtrs& cols, const double* w,
                            int64_t n, double* acc, int ld_acc, int acc_col0, cudaStream_t s) {
  if (n <= 0) return;
  if (G * nc <= BINSUM_MAX_DOUBLES) {
    size_t smem = (size_t)G * nc * sizeof(double);
    int per_sm = smem <= 24 * 1024 ? 8 : (smem <= 48 * 1024 ? 4 : 3);
    int blocks = launch_blocks(h, ceil_div64(n, 256 * 4), per_sm);
#define BS_CASE(NC_)                                                                                              \
  case NC_:                                                                                                       \
    if (smem > 48 * 1024)                                                                                         \
      CUDA_CHECK(cudaFuncSetAttribute(k_binsum_smem<NC_, WEIGHTED>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem));                                                                \
    k_binsum_smem<NC_, WEIGHTED><<<blocks, 256, smem, s>>>(key, cols, w, n, (int)G, acc, ld_acc, acc_col0);      \
    break;
    switch (nc) {
      BS_CASE(1) BS_CASE(2) BS_CASE(3) BS_CASE(4) BS_CASE(5) BS_CASE(6) BS_CASE(7) BS_CASE(8)
      default: LFE_THROW(LFE_ERR_ARG, "launch_groupsum: nc=%d", nc);
    }
#undef BS_CASE
    KERNEL_CHECK();
    h->n_launches += 1;
  } else {
    launch_segsum<WEIGHTED>(h, nc, key, cols, nullptr, w, n, acc, ld_acc, acc_col0, s);
  }
}

// S0 of every FE for one column batch in a single pass over the columns (F >= 2)
template <bool WEIGHTED>
static void launch_s0_fused(lfe_ctx* h, int nc, const int32_t* const* codes, const int64_t* G, double* const* acc,
                            int F, const ColPtrs& cols, const double* w, int64_t n, int ld_acc, int acc_col0,
                            cudaStream_t s) {
  if (n <= 0) return;
  S0Spec sp;
  sp.F = F;
  sp.smem_doubles = 0;
  bool vec = (w == nullptr) || aligned16(w);
  for (int c = 0; c < nc; ++c) vec = vec && aligned16(cols.col[c]);
  for (int j = 0; j < F; ++j) {
    sp.code[j] = codes[j];
    sp.acc[j] = acc[j];
    sp.G[j] = (int)(G[j] < 0x7fffffff ? G[j] : 0x7fffffff);
    vec = vec && aligned16(codes[j]);
    if (G[j] * nc + sp.smem_doubles <= BINSUM_MAX_DOUBLES) {
      sp.smem_off[j] = sp.smem_doubles;
      sp.smem_doubles += (int)(G[j] * nc);
    } else {
      sp.smem_off[j] = -1;
    }
  }
  // TMA bulk reduce for the rows of the big tables: needs 16-byte aligned table rows (even ld / column offset)
  // (for one or two columns the plain REDs are as fast as the bulk request)
  bool bulk = nc >= 3 && (ld_acc % 2 == 0) && (acc_col0 % 2 == 0);
  bool any_red = false;
  for (int j = 0; j < F; ++j)
    if (sp.smem_off[j] < 0) { any_red = true; bulk = bulk && aligned16(acc[j]); }
  bulk = bulk && any_red;
  const size_t slot_bytes = bulk ? (size_t)PJ_T * BR_SLOTS * ((nc + 1) & ~1) * sizeof(double) : 0;
  const size_t smem = (size_t)((sp.smem_doubles + 1) & ~1) * sizeof(double) + slot_bytes;
  const int per_sm = smem <= 24 * 1024 ? 8 : (smem <= 48 * 1024 ? 4 : (smem <= 72 * 1024 ? 3 : 2));
  const int blocks = launch_blocks(h, ceil_div64(n, (int64_t)PJ_T * PJ_I), per_sm);
#define S0_CASE(NC_)                                                                                            \
  case NC_:                                                                                                     \
    if (smem > 48 * 1024)                                                                                       \
      CUDA_CHECK(cudaFuncSetAttribute(k_s0_fused<NC_, WEIGHTED>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                      (int)smem));                                                              \
    k_s0_fused<NC_, WEIGHTED><<<blocks, PJ_T, smem, s>>>(sp, cols, w, n, ld_acc, acc_col0, vec, bulk);         \
    break;
  switch (nc) {
    S0_CASE(1) S0_CASE(2) S0_CASE(3) S0_CASE(4) S0_CASE(5) S0_CASE(6) S0_CASE(7) S0_CASE(8)
    S0_CASE(9) S0_CASE(10) S0_CASE(11) S0_CASE(12)
    default: LFE_THROW(LFE_ERR_ARG, "launch_s0_fused: nc=%d", nc);
  }
#undef S0_CASE
  KERNEL_CHECK();
  h->n_launches += 1;
}

static void launch_apply(lfe_ctx* h, int nc, const MutColPtrs& cols, const ApplyTabs& T, int64_t n, int ldc, int col0,
                         size_t smem, cudaStream_t s) {
  bool vec = true;
  for (int c = 0; c < nc; ++c) vec = vec && aligned16(cols.col[c]) && aligned16(cols.src[c]);
  for (int j = 0; j < T.n; ++j) vec = vec && aligned16(T.code[j]);
  int blocks = vec ? launch_blocks(h, ceil_div64(n, 256 * 4), 2) : launch_blocks(h, ceil_div64(n, 256), smem > 0 ? 4 : 8);
#define AP_CASE(NC_)                                                              \
  case NC_:                                                                       \
    if (vec) k_apply4<NC_><<<blocks, 256, smem, s>>>(cols, T, n, ldc, col0);      \
    else k_apply<NC_><<<blocks, 256, smem, s>>>(cols, T, n, ldc, col0);           \
    break;
  switch (nc) {
    AP_CASE(1) AP_CASE(2) AP_CASE(3) AP_CASE(4) AP_CASE(5) AP_CASE(6) AP_CASE(7) AP_CASE(8)
    default: LFE_THROW(LFE_ERR_ARG, "launch_apply: nc=%d", nc);
  }
#undef AP_CASE
  KERNEL_CHECK();
  h->n_launches += 1;
}

template <bool WEIGHTED>
static void launch_det_proj(lfe_ctx* h, int nc, const FEState& S, const ProjOthers& oth, int ldc, int col0, double* acc,
                            int ld_acc, int acc_col0, cudaStream_t s) {
  const int blocks = launch_blocks(h, ceil_div64(S.G, 8), 8);
#define DP_CASE(NC_) \
  case NC_: k_det_proj<NC_, WEIGHTED><<<blocks, 256, 0, s>>>(S.off, S.G, oth, S.w, ldc, col0, acc, ld_acc, acc_col0); break;
  switch (nc) {
    DP_CASE(1) DP_CASE(2) DP_CASE(3) DP_CASE(4) DP_CASE(5) DP_CASE(6) DP_CASE(7) DP_CASE(8)
    default: LFE_THROW(LFE_ERR_ARG, "launch_det_proj: nc=%d", nc);
  }
#undef DP_CASE
  KERNEL_CHECK();
  h->n_launches += 1;
  h->n_sweep_launches += 1;
}
// w_rows: weights in ROW order (nullable)
template <bool WEIGHTED>
static void launch_det_s0(lfe_ctx* h, int nc, const FEState& S, const ColPtrs& cols, const double* w_rows, double* acc,
                          int ld_acc, int acc_col0, cudaStream_t s) {
  const int blocks = launch_blocks(h, ceil_div64(S.G, 8), 8);
  const uint32_t* perm = S.identity ? nullptr : S.perm.p;
#define DS_CASE(NC_) \
  case NC_: k_det_s0<NC_, WEIGHTED><<<blocks, 256, 0, s>>>(S.off, S.G, perm, cols, w_rows, acc, ld_acc, acc_col0); break;
  switch (nc) {
    DS_CASE(1) DS_CASE(2) DS_CASE(3) DS_CASE(4) DS_CASE(5) DS_CASE(6) DS_CASE(7) DS_CASE(8)
    default: LFE_THROW(LFE_ERR_ARG, "launch_det_s0: nc=%d", nc);
  }
#undef DS_CASE
  KERNEL_CHECK();
  h->n_launches += 1;
}

// The column-independent part of a fit: entry lists and group sizes.  Built from the code columns alone, so
// the host can build it while the data columns are still in flight over PCIe, and reuse it for several
// outcome/regressor sets on the same panel.
struct lfe_plan {
  int F = 0;
  int64_t n = 0;
  bool weighted = false;
  int local_fe = -1;
  int big = 0;   // highest-cardinality FE: its entry list was built first
  int block_bits = 0;   // > 0: the other FEs' lists are ordered in L2 blocks of the big FE's codes (l2_block_bits)
  bool det = false;     // LFE_OPT_DETERMINISTIC at creation: plain (key, other codes, perm) lists + CSR offsets
  const int32_t* codes[LFE_MAX_FE] = {nullptr};
  const double* weights = nullptr;
  std::vector<FEState> st;
};

// L2 blocking of the entry lists.  A projection on a non-big FE gathers rows of the BIG FE's [G x C] table in
// ascending order inside every group; when that table outgrows the L2 (c5: 4M units x 12 columns = 384 MB, c3: 10M
// workers x 3 = 240 MB) the sparse ascending walk re-fetches it from HBM for every group (ncu: the c5 time
// projection moved 77 GB per sweep for 1.6 GB of entry words).  Ordering the list by (block of the big code, own
// code, big code) instead -- one more stable radix pass on the block bits -- makes each table slice of
// opt_l2_block_bytes L2-resident while all entries that gather from it are processed; a group is then split into
// at most 2^bits runs (a few more RED.ADD.F64 per group).  Returns the number of block bits (0 = no blocking).
static int l2_block_bits(const lfe_ctx* h, int64_t G_big, int F) {
  if (h->opt_columns_hint <= 0 || F < 2) return 0;
  const int C = h->opt_columns_hint;
  const int64_t row_bytes = (int64_t)((C >= 5 && (C & 1)) ? C + 1 : C) * 8;
  if (G_big * row_bytes <= h->opt_l2_block_bytes + h->opt_l2_block_bytes / 2) return 0;
  int64_t block_codes = 1;
  while (block_codes * 2 * row_bytes <= h->opt_l2_block_bytes) block_codes *= 2;
  const int bb = bits_for((uint64_t)(G_big - 1)), lb = bits_for((uint64_t)(block_codes - 1));
  const int bits = bb - lb;
  return bits < 0 ? 0 : (bits > 8 ? 8 : bits);
}

// out[i] = code[i] >> shift
__global__ void __launch_bounds__(256) k_block_of(const int32_t* __restrict__ code, int64_t n, int shift,
                                                  uint32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    __stcs(out + i, (uint32_t)__ldcs(code + i) >> shift);
}

lfe_plan* demean_plan_create(lfe_ctx* h, const int32_t* const* codes, const int64_t* n_groups, int F, int64_t n,
                             const double* weights, int local_fe, const int32_t* const* counts_in, cudaStream_t s) {
  LFE_REQUIRE(F >= 1 && F <= LFE_MAX_FE, "lfe_demean: F=%d out of range [1,%d]", F, LFE_MAX_FE);
  LFE_REQUIRE(n >= 0 && n < (int64_t)0x7fffffff, "lfe_demean: n out of range");
  const bool weighted = weights != nullptr;
  const bool dist = h->world > 1;
  std::unique_ptr<lfe_plan> plan(new lfe_plan());
  plan->F = F; plan->n = n; plan->weighted = weighted; plan->local_fe = local_fe; plan->weights = weights;
  for (int j = 0; j < F; ++j) plan->codes[j] = codes[j];
  plan->st.resize(F);
  std::vector<FEState>& st = plan->st;
  {
    PhaseTimer t(h, s, 5);
    // Entry lists.  The highest-cardinality FE ("big") is sorted first; every other FE is then sorted
    // STABLY from the big-sorted sequence, so inside each of its groups the big FE's codes ascend and the
    // gathers into the big FE's table walk memory monotonically.
    int big = 0;
    for (int j = 1; j < F; ++j)
      if (n_groups[j] > n_groups[big]) big = j;
    plan->big = big;
    for (int j = 0; j < F; ++j) {
      st[j].G = n_groups[j];
      LFE_REQUIRE(st[j].G >= 1, "lfe_demean: n_groups[%d] must be >= 1", j);
      st[j].identity = (F == 1) || prim_is_sorted(h, codes[j], n, s) != 0;   // F == 1 never projects
    }
    const bool det = h->opt_deterministic != 0;
    LFE_REQUIRE(!det || !dist, "LFE_OPT_DETERMINISTIC is single-GPU only");
    plan->det = det;
    if (det) {
      // every FE that projects needs its (sorted key, other codes, row permutation) list: F == 1 never projects, but
      // its S0 is summed in list order too
      for (int j = 0; j < F; ++j) st[j].identity = prim_is_sorted(h, codes[j], n, s) != 0;
    }
    if (!det && F == 2 && !weighted && !st[0].identity && !st[1].identity &&
        bits_for((uint64_t)(st[0].G - 1)) + bits_for((uint64_t)(st[1].G - 1)) <= 32) {
      // two FEs whose codes fit one 32-bit word together: every entry is the packed word
      // (big code << bs) | small code.  Sorting the words stably on the big code's bits gives the big FE's
      // entry list, sorting THAT stably on the small code's bits gives the small FE's (big codes ascending
      // inside every small group) -- keys only, no payload, no row permutation, no gathers, and each
      // projection streams 4 bytes per entry.
      const int sm = 1 - big;
      FEState &B = st[big], &M = st[sm];
      const int bs = bits_for((uint64_t)(M.G - 1)), bb = bits_for((uint64_t)(B.G - 1));
      const uint32_t smask = bs >= 32 ? 0xffffffffu : ((1u << bs) - 1u);
      const uint32_t bmask = bb >= 32 ? 0xffffffffu : ((1u << bb) - 1u);
      DevBuf<int32_t> words(n, s);
      k_pack_codes<<<launch_blocks(h, ceil_div64(n, 256 * 4), 8), 256, 0, s>>>(codes[big], codes[sm], n, bs,
                                                                                (uint32_t*)words.p);
      KERNEL_CHECK();
      h->n_launches += 1;
      B.key_buf.alloc(n, s);
      prim_sort_keys_u32(h, (const uint32_t*)words.p, n, bs, bb, (uint32_t*)B.key_buf.p, s);
      M.key_buf.alloc(n, s);
      prim_sort_keys_u32(h, (const uint32_t*)B.key_buf.p, n, 0, bs, (uint32_t*)M.key_buf.p, s);
      const int blk_bits = l2_block_bits(h, B.G, F);
      if (blk_bits > 0) {
        // (small, big) order -> (block of big, small, big): one more stable pass on the top bits of the big code
        DevBuf<int32_t> blocked(n, s);
        prim_sort_keys_u32(h, (const uint32_t*)M.key_buf.p, n, bs + bb - blk_bits, blk_bits, (uint32_t*)blocked.p, s);
        M.key_buf = std::move(blocked);
      }
      plan->block_bits = blk_bits;
      B.key = B.key_buf; B.packed = true; B.pk = PackSpec{bs, 0, bmask, smask};
      M.key = M.key_buf; M.packed = true; M.pk = PackSpec{0, bs, smask, bmask};
      B.oth[sm] = nullptr; M.oth[big] = nullptr;
    } else if (!det && F == 2 && !weighted && !st[0].identity && !st[1].identity) {
      // two FEs, codes too wide to pack: sort (key, other code) pairs directly
      const int sm = 1 - big;
      FEState &B = st[big], &M = st[sm];
      B.key_buf.alloc(n, s); B.oth_buf[sm].alloc(n, s);
      prim_sort_pairs_u32(h, (const uint32_t*)codes[big], (const uint32_t*)codes[sm], n, bits_for((uint64_t)(B.G - 1)),
                          (uint32_t*)B.key_buf.p, (uint32_t*)B.oth_buf[sm].p, s);
      B.key = B.key_buf; B.oth[sm] = B.oth_buf[sm];
      M.key_buf.alloc(n, s); M.oth_buf[big].alloc(n, s);
      prim_sort_pairs_u32(h, (const uint32_t*)B.oth_buf[sm].p, (const uint32_t*)B.key_buf.p, n,
                          bits_for((uint64_t)(M.G - 1)), (uint32_t*)M.key_buf.p, (uint32_t*)M.oth_buf[big].p, s);
      M.key = M.key_buf; M.oth[big] = M.oth_buf[big];
    } else {
      const uint32_t* base_perm = nullptr;   // big-sorted order (nullptr = identity)
      for (int step = 0; step < F; ++step) {
        const int j = step == 0 ? big : (step <= big ? step - 1 : step);
        FEState& S = st[j];
        if (S.identity) {
          S.key = codes[j];
          for (int k = 0; k < F; ++k) S.oth[k] = codes[k];
          S.w = weights;
          continue;
        }
        S.perm.alloc(n, s);
        S.key_buf.alloc(n, s);
        const int kb = bits_for((uint64_t)(S.G - 1));
        if (j == big || base_perm == nullptr) {
          prim_argsort_u32(h, (const uint32_t*)codes[j], n, kb, S.perm, (uint32_t*)S.key_buf.p, s);
          if (j == big) base_perm = S.perm;
        } else {
          DevBuf<int32_t> kin(n, s);
          prim_gather_i32(codes[j], base_perm, n, kin, s);
          prim_sort_pairs_u32(h, (const uint32_t*)kin.p, base_perm, n, kb, (uint32_t*)S.key_buf.p, S.perm, s);
          h->n_launches += 1;
          // (narrow rows -- C <= 4 -- waste little of a 32-byte sector and their gathers are bound by the L1TEX
          // pipe, not by HBM: c3 gained nothing; an FE small enough for k_proj_tiny never uses its own list)
          const bool tiny_j = S.G * (h->opt_columns_hint < PJ_MAXNC ? h->opt_columns_hint : PJ_MAXNC) <= PT_MAX_BINS;
          const int blk_bits = (st[big].identity || h->opt_columns_hint < 5 || tiny_j) ? 0 : l2_block_bits(h, st[big].G, F);
          if (blk_bits > 0) {
            // (own, big) order -> (block of big, own, big): stable pass on the block index of the big code
            const int shift = bits_for((uint64_t)(st[big].G - 1)) - blk_bits;
            DevBuf<uint32_t> bk(n, s), bk_sorted(n, s), perm2(n, s);
            prim_gather_i32(codes[big], S.perm, n, kin, s);
            k_block_of<<<launch_blocks(h, ceil_div64(n, 256 * 4), 8), 256, 0, s>>>(kin, n, shift, bk);
            KERNEL_CHECK();
            prim_sort_pairs_u32(h, bk, S.perm, n, blk_bits, bk_sorted, perm2, s);
            S.perm = std::move(perm2);
            prim_gather_i32(codes[j], S.perm, n, S.key_buf, s);
            h->n_launches += 3;
            plan->block_bits = blk_bits;
          }
        }
        S.key = S.key_buf;
        for (int k = 0; k < F; ++k) {
          if (k == j) continue;
          S.oth_buf[k].alloc(n, s);
          prim_gather_i32(codes[k], S.perm, n, S.oth_buf[k], s);
          S.oth[k] = S.oth_buf[k];
          h->n_launches += 1;
        }
        if (weighted) {
          S.w_buf.alloc(n, s);
          prim_gather_f64(weights, S.perm, n, S.w_buf, s);
          S.w = S.w_buf;
          h->n_launches += 1;
        }
      }
    }
    // group sizes: unweighted counts (and sum of weights per group when weighted)
    for (int j = 0; j < F; ++j) {
      FEState& S = st[j];
      S.nsum.alloc(S.G, s);
      {
        DevBuf<uint32_t> c32;
        const uint32_t* cnt32 = (counts_in && counts_in[j]) ? (const uint32_t*)counts_in[j] : nullptr;
        if (!cnt32) {
          c32.alloc(S.G, s);
          prim_histogram(h, codes[j], n, S.G, c32, s);
          cnt32 = c32;
        }
        DevBuf<double>& dst = weighted ? S.cnt_buf : S.nsum;
        if (weighted) S.cnt_buf.alloc(S.G, s);
        prim_u32_to_f64(h, cnt32, S.G, dst, s);
        if (det) {   // CSR offsets of the groups in the sorted list: exclusive scan of the sizes, off[G] = n
          S.off.alloc(S.G + 1, s);
          S.off.zero();
          CUDA_CHECK(cudaMemcpyAsync(S.off.p, cnt32, (size_t)S.G * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
          prim_exclusive_scan_u32(h, S.off, S.off, S.G + 1, nullptr, s);
        }
      }
      if (weighted) {
        S.nsum.zero();
        ColPtrs cw; cw.col[0] = weights;
        if (det) launch_det_s0<false>(h, 1, S, cw, nullptr, S.nsum, 1, 0, s);
        else launch_groupsum<false>(h, 1, codes[j], S.G, cw, nullptr, n, S.nsum, 1, 0, s);
        S.cnt = S.cnt_buf;
      } else {
        S.cnt = S.nsum;
      }
      if (dist && j != local_fe) {
        PhaseTimer tc(h, s, 13);
        comm_allreduce(h, S.nsum, S.G, 0, s);
        if (weighted) comm_allreduce(h, S.cnt_buf, S.G, 0, s);
      }
    }
  }
  return plan.release();
}

void demean_plan_destroy(lfe_plan* plan) { delete plan; }

// cols_src (nullable): read the input columns from there and write the demeaned columns to `cols`
// (out of place: the caller's data stay untouched without a separate copy); nullptr = in place on `cols`.
static void demean_plan_run_tail(lfe_ctx* h, lfe_plan* plan, double* const* cols, const double* const* cols_src, int C,
                                 const int* fe_order, double tol, int max_iter, int min_iter, int flags,
                                 int* iterations_out, double* max_mean_out, int* converged_out, const TailRequest* tail,
                                 cudaStream_t s) {
  std::vector<const double*> src_v(C);
  for (int c = 0; c < C; ++c) src_v[c] = cols_src ? cols_src[c] : cols[c];
  const double* const* in = src_v.data();
  const bool strict_check = (flags & LFE_DEMEAN_STRICT_CHECK) != 0;
  const int F = plan->F;
  const int64_t n = plan->n;
  const bool weighted = plan->weighted;
  const int local_fe = plan->local_fe;
  const double* weights = plan->weights;
  const int32_t* const* codes = plan->codes;
  std::vector<FEState>& st = plan->st;
  LFE_REQUIRE(C >= 1, "lfe_demean: C must be >= 1");
  // leading dimension of every FE table: an odd C >= 5 is rounded up to even, so table rows are 16-byte aligned
  // (128-bit gathers, TMA bulk reduces); the pad column stays zero throughout.  Narrow tables are left alone:
  // padding 3 -> 4 columns makes c3's 240 MB worker table 320 MB and its random gathers 20 % slower (L2 hit
  // rate), and 1 -> 2 would double every table of the per-column runs.
  static const bool pad_narrow = getenv("LFE_PAD_NARROW") != nullptr;   // experiment: 3 -> 4 columns as well
  // (64-byte rows for C = 5..7 -- no gathered row would straddle a 128-byte line -- were measured too: the 48 MB unit
  // table of c2 becomes 64 MB, L2 hit rates fall and the fit takes 11.6 instead of 10.9 ms)
  const int Cp = ((C >= 5 || (pad_narrow && C >= 3)) && (C & 1)) ? C + 1 : C;
  LFE_REQUIRE(max_iter >= 1, "lfe_demean: max_iter must be >= 1");
  {
    bool seen[LFE_MAX_FE] = {false};
    for (int j = 0; j < F; ++j) {
      LFE_REQUIRE(fe_order[j] >= 0 && fe_order[j] < F && !seen[fe_order[j]], "lfe_demean: fe_order is not a permutation");
      seen[fe_order[j]] = true;
    }
  }
  const bool dist = h->world > 1;
  *iterations_out = 0;
  if (max_mean_out) *max_mean_out = 0.0;
  if (converged_out) *converged_out = 0;
  // release the per-run tables when this call returns, whatever happens
  struct TableGuard {
    std::vector<FEState>& st;
    ~TableGuard() {
      for (auto& S : st) { S.alpha.release(); S.S0.release(); S.acc.release(); S.acc_chk.release(); S.S0u_buf.release(); }
    }
  } guard{st};
  // ---- S0: the only pass over the columns before the final apply -----------------------------------
  {
    PhaseTimer t(h, s, 5);
    for (int j = 0; j < F; ++j) {
      FEState& S = st[j];
      size_t tab = (size_t)S.G * Cp;
      S.alpha.alloc(tab, s); S.S0.alloc(tab, s); S.acc.alloc(tab, s);
      S.alpha.zero(); S.S0.zero(); S.acc.zero();
      S.acc_chk.alloc(S.G, s);
      S.acc_chk.zero();
    }
    // S0 (weighted sums of every column), column batches of <= PJ_MAXNC; with several FEs one fused pass
    // per batch fills every FE's table, so the columns are read once
    // (the fused pass of F >= 2 takes S0_MAXNC = 12 columns at once: one bulk reduce of the whole row per big-table
    // row instead of one per batch -- c5, C = 11: 56 -> 40 ms on one GPU)
    const int s0_batch = (!plan->det && F >= 2) ? S0_MAXNC : PJ_MAXNC;
    for (int c0 = 0; c0 < C; c0 += s0_batch) {
      int nc = C - c0 < s0_batch ? C - c0 : s0_batch;
      ColPtrs cp;
      for (int c = 0; c < nc; ++c) cp.col[c] = in[c0 + c];
      if (plan->det) {
        for (int j = 0; j < F; ++j) {
          if (weighted) launch_det_s0<true>(h, nc, st[j], cp, weights, st[j].S0, Cp, c0, s);
          else launch_det_s0<false>(h, nc, st[j], cp, nullptr, st[j].S0, Cp, c0, s);
        }
      } else if (F >= 2) {
        int64_t Gs[LFE_MAX_FE];
        double* accs[LFE_MAX_FE];
        for (int j = 0; j < F; ++j) { Gs[j] = st[j].G; accs[j] = st[j].S0; }
        if (weighted) launch_s0_fused<true>(h, nc, codes, Gs, accs, F, cp, weights, n, Cp, c0, s);
        else launch_s0_fused<false>(h, nc, codes, Gs, accs, F, cp, nullptr, n, Cp, c0, s);
      } else {
        if (weighted) launch_groupsum<true>(h, nc, codes[0], st[0].G, cp, weights, n, st[0].S0, Cp, c0, s);
        else launch_groupsum<false>(h, nc, codes[0], st[0].G, cp, nullptr, n, st[0].S0, Cp, c0, s);
      }
    }
    for (int j = 0; j < F; ++j) {
      FEState& S = st[j];
      size_t tab = (size_t)S.G * Cp;
      if (weighted) {
        // unweighted y sums for the convergence check (polars_impl.py:513 ignores the weights)
        S.S0u_buf.alloc(S.G, s);
        S.S0u_buf.zero();
        ColPtrs cy; cy.col[0] = in[0];
        if (plan->det) launch_det_s0<false>(h, 1, S, cy, nullptr, S.S0u_buf, 1, 0, s);
        else launch_groupsum<false>(h, 1, codes[j], S.G, cy, nullptr, n, S.S0u_buf, 1, 0, s);
        S.S0u = S.S0u_buf; S.ld_s0u = 1;
      } else {
        S.S0u = S.S0; S.ld_s0u = Cp;
      }
      if (dist && j != local_fe) {
        PhaseTimer tc(h, s, 13);
        comm_allreduce(h, S.S0, (int64_t)tab, 0, s);
        if (weighted) comm_allreduce(h, S.S0u_buf, S.G, 0, s);
      }
    }
  }

  size_t proj_smem = 0;
  auto others_of = [&](int j, bool include_all_for_check) {
    ProjOthers o;
    int64_t rows[LFE_MAX_FE];
    o.n = 0;
    for (int k = 0; k < F; ++k) {
      if (k == j) continue;
      o.code[o.n] = st[j].oth[k];
      o.alpha[o.n] = st[k].alpha;
      rows[o.n] = st[k].G;
      o.n++;
    }
    (void)include_all_for_check;
    proj_smem = plan_staging(o, rows, Cp);
    return o;
  };

  DevBuf<double> dmax(1, s);
  int iterations = 0;
  bool converged = false;
  double max_mean = 0.0;

  // accumulate the projection sums of FE j over all column batches (+ cross-rank sum for non-local FEs)
  // LFE_TRACE_PROJ=1: per-projection device times (CUDA events around every projection, printed when the run ends)
  static const bool trace_proj = getenv("LFE_TRACE_PROJ") != nullptr;
  struct ProjTrace { int j; cudaEvent_t a, b; };
  std::vector<ProjTrace> trace;
  struct TraceDump {
    std::vector<ProjTrace>& tr;
    ~TraceDump() {
      for (auto& e : tr) {
        float ms = 0.f;
        if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess)
          fprintf(stderr, "lfe trace: projection on FE %d: %.4f ms\n", e.j, ms);
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
      }
    }
  } trace_dump{trace};
  auto project = [&](int j) {
    FEState& S = st[j];
    if (F == 1) return;
    ProjTrace tr{j, nullptr, nullptr};
    if (trace_proj) { cudaEventCreate(&tr.a); cudaEventCreate(&tr.b); cudaEventRecord(tr.a, s); }
    struct Rec { ProjTrace& t; std::vector<ProjTrace>& v; cudaStream_t s; bool on;
                 ~Rec() { if (on) { cudaEventRecord(t.b, s); v.push_back(t); } } } rec{tr, trace, s, trace_proj};
    ProjOthers o = others_of(j, false);
    // a very small FE is projected over the big FE's entry list with thread-private bins (k_proj_tiny)
    const int big = plan->big;
    const bool tiny = !plan->det && j != big && n >= 65536 && S.G * (C < PJ_MAXNC ? C : PJ_MAXNC) <= PT_MAX_BINS;
    // the row-group kernel (packed list, other table in L2 / HBM) takes up to 12 columns in one pass
    const bool quad = !plan->det && !tiny && S.packed && o.n == 1 && o.stage_rows[0] == 0 && n >= 4096 && C <= S0_MAXNC;
    const int bw = quad ? S0_MAXNC : PJ_MAXNC;
    for (int c0 = 0; c0 < C; c0 += bw) {
      int nc = C - c0 < bw ? C - c0 : bw;
      if (plan->det) {
        o.stage_rows[0] = 0;   // (no staging: plain gathers)
        if (weighted) launch_det_proj<true>(h, nc, S, o, Cp, c0, S.acc, Cp, c0, s);
        else launch_det_proj<false>(h, nc, S, o, Cp, c0, S.acc, Cp, c0, s);
      } else if (tiny) {
        if (weighted) launch_proj_tiny<true>(h, nc, st, big, j, F, n, Cp, c0, S.acc, Cp, c0, s);
        else launch_proj_tiny<false>(h, nc, st, big, j, F, n, Cp, c0, S.acc, Cp, c0, s);
      } else {
        // an odd batch that ends at the zero pad column takes the pad along: NC even => 128-bit table gathers
        // (the pad column of S0 / acc / alpha is zero and stays zero)
        const int ncl = ((nc & 1) && c0 + nc == C && Cp > C && nc + 1 <= bw && (c0 & 1) == 0) ? nc + 1 : nc;
        if (weighted) launch_proj<true>(h, ncl, S, o, n, Cp, c0, S.acc, Cp, c0, proj_smem, s);
        else launch_proj<false>(h, ncl, S, o, n, Cp, c0, S.acc, Cp, c0, proj_smem, s);
      }
    }
    if (dist && j != local_fe) {
      PhaseTimer tc(h, s, 13);
      comm_allreduce(h, S.acc, S.G * Cp, 0, s);
    }
  };
  // Unweighted fits: the reference's stopping quantity for the FIRST FE of the sweep order,
  // max_g |mean_g(y)| after sweep `it`, IS the y-column update that FE receives at the start of sweep it+1.
  // So the check runs that projection speculatively into a staging table (kept if the loop goes on,
  // dropped if it stops); the LAST FE of the order was projected last, its group means are exactly zero in
  // coefficient space; only the middle FEs (F >= 3) need a y-only pass.
  const bool fast_check = !weighted && !strict_check && F >= 2 && tol > 1e-13;
  DevBuf<double> alpha_next;
  if (fast_check) alpha_next.alloc((size_t)st[fe_order[0]].G * Cp, s);
  bool first_done = false;   // sweep it+1's first projection was already applied by the check of sweep it

  for (int it = 1; it <= max_iter; ++it) {
    if (strict_check) dmax.zero();
    {
      PhaseTimer t(h, s, 6);
      for (int oi = first_done ? 1 : 0; oi < F; ++oi) {
        int j = fe_order[oi];
        FEState& S = st[j];
        project(j);
        k_finalize<<<launch_blocks(h, ceil_div64(S.G * Cp, 256), 8), 256, 0, s>>>(
            S.alpha, S.alpha, S.acc, S.S0, S.nsum, S.G, Cp, strict_check ? dmax.p : nullptr, -1);
        KERNEL_CHECK();
        h->n_launches += 1;
      }
      first_done = false;
    }
    iterations = it;
    if (it >= min_iter && fast_check) {
      PhaseTimer t(h, s, 7);
      // first FE of the order: its speculative projection is the stopping quantity for that FE
      dmax.zero();
      FEState& S1 = st[fe_order[0]];
      project(fe_order[0]);
      k_finalize<<<launch_blocks(h, ceil_div64(S1.G * Cp, 256), 8), 256, 0, s>>>(S1.alpha, alpha_next, S1.acc, S1.S0,
                                                                                  S1.nsum, S1.G, Cp, dmax, 0);
      KERNEL_CHECK();
      h->n_launches += 1;
      // a non-local FE's table comes from all-reduced sums, so its maximum is already the same number on every
      // rank: no second collective per stopping test (only a rank-local first FE needs the max over ranks)
      if (dist && fe_order[0] == local_fe) comm_allreduce(h, dmax, 1, 1, s);
      CUDA_CHECK(cudaMemcpyAsync(&max_mean, dmax.p, sizeof(double), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      if (max_mean < tol && F > 2) {
        // only now can the test pass: the middle FEs need explicit y-only group means.  (While the first FE
        // alone is above tol the maximum over all FEs is too, so these passes are skipped -- a 3-way fit that
        // runs to max_iter never pays for them; max_mean is then the first FE's value, a lower bound >= tol.)
        dmax.zero();
        for (int oi = 1; oi + 1 < F; ++oi) {
          FEState& S = st[fe_order[oi]];
          ProjOthers o = others_of(fe_order[oi], true);
          if (plan->det) launch_det_proj<false>(h, 1, S, o, Cp, 0, S.acc_chk, 1, 0, s);
          else launch_proj<false>(h, 1, S, o, n, Cp, 0, S.acc_chk, 1, 0, proj_smem, s);
          if (dist && fe_order[oi] != local_fe) comm_allreduce(h, S.acc_chk, S.G, 0, s);
          k_check<<<launch_blocks(h, ceil_div64(S.G, 256), 8), 256, 0, s>>>(S.alpha, S.acc_chk, S.S0u, S.ld_s0u, S.cnt,
                                                                             S.G, Cp, dmax);
          KERNEL_CHECK();
          h->n_launches += 1;
        }
        double mid = 0.0;
        bool local_mid = false;
        for (int oi = 1; oi + 1 < F; ++oi) local_mid = local_mid || fe_order[oi] == local_fe;
        if (dist && local_mid) comm_allreduce(h, dmax, 1, 1, s);
        CUDA_CHECK(cudaMemcpyAsync(&mid, dmax.p, sizeof(double), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        if (mid > max_mean) max_mean = mid;
      }
      if (max_mean < tol) {
        converged = true;
        break;
      }
      if (it < max_iter) {   // keep the speculative projection: it is sweep it+1's first step
        std::swap(S1.alpha, alpha_next);
        first_done = true;
      }
    } else if (it >= min_iter && strict_check) {
      // true fixed-point test: the largest (weighted) group mean removed during this sweep, all columns
      PhaseTimer t(h, s, 7);
      if (dist) comm_allreduce(h, dmax, 1, 1, s);
      CUDA_CHECK(cudaMemcpyAsync(&max_mean, dmax.p, sizeof(double), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      if (max_mean < tol) {
        converged = true;
        break;
      }
    } else if (it >= min_iter) {
      PhaseTimer t(h, s, 7);
      dmax.zero();
      for (int j = 0; j < F; ++j) {
        FEState& S = st[j];
        if (F > 1) {
          ProjOthers o = others_of(j, true);
          if (plan->det) launch_det_proj<false>(h, 1, S, o, Cp, 0, S.acc_chk, 1, 0, s);
          else launch_proj<false>(h, 1, S, o, n, Cp, 0, S.acc_chk, 1, 0, proj_smem, s);
          if (dist && j != local_fe) comm_allreduce(h, S.acc_chk, S.G, 0, s);
        }
        k_check<<<launch_blocks(h, ceil_div64(S.G, 256), 8), 256, 0, s>>>(S.alpha, S.acc_chk, S.S0u, S.ld_s0u, S.cnt,
                                                                           S.G, Cp, dmax);
        KERNEL_CHECK();
        h->n_launches += 1;
      }
      if (dist) comm_allreduce(h, dmax, 1, 1, s);
      CUDA_CHECK(cudaMemcpyAsync(&max_mean, dmax.p, sizeof(double), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      if (max_mean < tol) {
        converged = true;
        break;
      }
    }
  }

  *iterations_out = iterations;
  if (max_mean_out) *max_mean_out = max_mean;
  if (converged_out) *converged_out = converged ? 1 : 0;
  if (tail) {
    // ---- regression tail straight from (columns, tables): nothing is written back -------------------
    TailTabs T;
    int64_t rows[LFE_MAX_FE];
    T.n = F;
    for (int j = 0; j < F; ++j) {
      T.code[j] = codes[j];
      T.alpha[j] = st[j].alpha;
      rows[j] = st[j].G;
    }
    const size_t tab_smem = plan_staging(T, rows, Cp);
    const int p = C;   // intercept + k regressors
    DevBuf<double> gram((size_t)(p + 1) * (p + 1), s), A((size_t)p * p, s), b(p, s), Ainv((size_t)p * p, s);
    DevBuf<double> B((size_t)(p > 1 ? (p - 1) * (p - 1) : 1), s);
    gram_fused_impl(h, in, C, n, weights, T, Cp, tab_smem, gram, s);
    k_split_gram<<<1, 256, 0, s>>>(gram, p, A, b);
    KERNEL_CHECK();
    solve_spd_impl(h, A, b, p, tail->beta_full_out, Ainv, nullptr, s);
    k_drop_intercept<<<1, 256, 0, s>>>(Ainv, p, B);
    KERNEL_CHECK();
    h->n_launches += 2;
    if ((tail->vcov_kind & 0xff) == LFE_VCOV_IID && !weighted && p > 1) {
      // IID needs only rss (std_errors.py:183-210): with the normal equations solved, rss = y~'y~ - beta'X~'y~ and
      // tss = y~'y~ - (1'y~)^2 / n are already in the (all-reduced) Gram -- no second pass over the rows.  Guarded:
      // when the fit is so tight that the difference cancels (rss < 1e-6 y~'y~) the explicit residual pass runs.
      PhaseTimer tv(h, s, 11);
      DevBuf<int> ok(1, s);
      DevBuf<double> stats_local(4, s);
      double* stats = tail->stats_out ? tail->stats_out : stats_local.p;
      k_iid_from_gram<<<1, 256, 0, s>>>(gram, tail->beta_full_out, B, p, (double)tail->n_obs, (double)tail->df_resid,
                                        tail->se_out, tail->vcov_out, stats, ok);
      KERNEL_CHECK();
      h->n_launches += 1;
      int hok = 0;
      CUDA_CHECK(cudaMemcpyAsync(&hok, ok.p, sizeof(int), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      if (hok) return;
    }
    resid_vcov_fused_impl(h, in, C, n, T, Cp, tab_smem, tail->beta_full_out, B, tail->vcov_kind, tail->cl_codes,
                          tail->n_cl_groups, weights, tail->ssc, tail->n_obs, tail->df_resid, tail->se_out,
                          tail->vcov_out, tail->stats_out, tail->n_clusters_out, s);
    CUDA_CHECK(cudaStreamSynchronize(s));   // B / Ainv scratch and the caller's host outputs
    return;
  }
  // ---- apply the tables to the columns, in place ------------------------------------------------
  {
    PhaseTimer t(h, s, 8);
    ApplyTabs T;
    int64_t rows[LFE_MAX_FE];
    T.n = F;
    for (int j = 0; j < F; ++j) {
      T.code[j] = codes[j];
      T.alpha[j] = st[j].alpha;
      rows[j] = st[j].G;
    }
    const size_t apply_smem = plan_staging(T, rows, Cp);
    for (int c0 = 0; c0 < C; c0 += PJ_MAXNC) {
      int nc = C - c0 < PJ_MAXNC ? C - c0 : PJ_MAXNC;
      MutColPtrs mp;
      for (int c = 0; c < nc; ++c) { mp.col[c] = cols[c0 + c]; mp.src[c] = in[c0 + c]; }
      if (n > 0) launch_apply(h, nc, mp, T, n, Cp, c0, apply_smem, s);
    }
  }
}

void demean_plan_run(lfe_ctx* h, lfe_plan* plan, double* const* cols, const double* const* cols_src, int C,
                     const int* fe_order, double tol, int max_iter, int min_iter, int flags, int* iterations_out,
                     double* max_mean_out, int* converged_out, cudaStream_t s) {
  demean_plan_run_tail(h, plan, cols, cols_src, C, fe_order, tol, max_iter, min_iter, flags, iterations_out, max_mean_out,
                       converged_out, nullptr, s);
}

// lfe_demean_plan_fit: sweeps + [1 | X~] normal equations + solve + residual statistics + IID / HC1 / one-way
// clustered SEs, with the demeaned columns formed on the fly (never written).  cols are read-only.
void demean_plan_fit(lfe_ctx* h, lfe_plan* plan, const double* const* cols, int C, const int* fe_order, double tol,
                     int max_iter, int min_iter, int flags, int vcov_kind, const int32_t* cl_codes, int64_t n_cl_groups,
                     int ssc, int64_t n_obs, int64_t df_resid, double* beta_full_out, double* se_out, double* vcov_out,
                     double* stats_out, int* iterations_out, double* max_mean_out, int* converged_out,
                     int64_t* n_clusters_out, cudaStream_t s) {
  LFE_REQUIRE(!plan->det || (vcov_kind & 0xff) != LFE_VCOV_CLUSTER,
              "lfe_demean_plan_fit: deterministic clustered SEs go through lfe_demean_plan_run + lfe_resid_vcov");
  LFE_REQUIRE(C >= 1 && C <= TL_MAXC, "lfe_demean_plan_fit: %d columns unsupported (max %d); use lfe_demean_plan_run + "
              "lfe_gram + lfe_resid_vcov", C, TL_MAXC);
  TailRequest tr{vcov_kind, cl_codes, n_cl_groups, ssc, n_obs, df_resid, beta_full_out, se_out, vcov_out, stats_out,
                 n_clusters_out};
  demean_plan_run_tail(h, plan, const_cast<double* const*>(cols), cols, C, fe_order,
                       tol, max_iter, min_iter, flags, iterations_out, max_mean_out, converged_out, &tr, s);
}

void demean_impl(lfe_ctx* h, double* const* cols, int C, const int32_t* const* codes, const int64_t* n_groups,
                 const int* fe_order, int F, int64_t n, const double* weights, double tol, int max_iter, int min_iter,
                 int local_fe, int flags, int* iterations_out, double* max_mean_out, int* converged_out,
                 cudaStream_t s) {
  std::unique_ptr<lfe_plan> plan(demean_plan_create(h, codes, n_groups, F, n, weights, local_fe, nullptr, s));
  demean_plan_run(h, plan.get(), cols, nullptr, C, fe_order, tol, max_iter, min_iter, flags, iterations_out,
                  max_mean_out, converged_out, s);
}
