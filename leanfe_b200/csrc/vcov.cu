// vcov.cu -- residuals, R^2 ingredients and IID / HC1 / clustered sandwich standard errors.
// Reference: python/leanfe/polars_impl.py:229-232, 281-283 and python/leanfe/std_errors.py:183-441.
//
// One streaming pass computes r = y - b0 - X beta, sum r^2, sum w r^2 and the shifted moments of y.
// HC1 reuses the Gram kernels with omega = w r^2.  Cluster scores S_g = sum_{i in g} w x_i r_i are a
// by-key reduction (segmented when the cluster is the row sort key, RED.ADD.F64 otherwise) into a
// row-major G x k table (one or two 32-byte sectors per row) whose Gram is the meat.  The k x k sandwich algebra runs in one CTA.
#include "common.cuh"
#include "segreduce.cuh"
#include "tail.cuh"

void gram_core(lfe_ctx* h, const double* const* cols, int C, int64_t n, const double* w, const double* r,
               int add_intercept, double* gram_out, cudaStream_t s, int64_t stride = 1);
int64_t factorize_dispatch(lfe_ctx* h, const void* keys, int key_dtype, int64_t n, int order, int32_t* codes,
                           void* uniques, int64_t uniques_cap, cudaStream_t s);
void comm_allreduce(lfe_ctx* h, double* buf, int64_t count, int op_max, cudaStream_t s);

constexpr int RV_MAXK = 96;
struct XCols {
  const double* col[RV_MAXK];
};

// partials[block][5] = { sum r^2, sum w r^2, sum (y-c), sum (y-c)^2, 0 }
__global__ void __launch_bounds__(256) k_resid(const double* __restrict__ y, XCols X, int k,
                                               const double* __restrict__ beta_full, const double* __restrict__ w,
                                               int64_t n, double* __restrict__ resid, double* __restrict__ partials,
                                               int use_shift) {
  __shared__ double sb[RV_MAXK + 1];
  for (int j = threadIdx.x; j <= k; j += blockDim.x) sb[j] = beta_full[j];
  __syncthreads();
  const double c = use_shift ? y[0] : 0.0;  // multi-rank: a common shift (0) keeps the moments summable
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double yi = __ldcs(y + i);
    double r = yi - sb[0];
#pragma unroll 4
    for (int j = 0; j < k; ++j) r = fma(-sb[j + 1], __ldcs(X.col[j] + i), r);
    __stcs(resid + i, r);
    double r2 = r * r;
    a0 += r2;
    a1 += w ? __ldcs(w + i) * r2 : r2;
    double d = yi - c;
    a2 += d;
    a3 = fma(d, d, a3);
  }
  __shared__ double sh[8][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
  if (lane == 0) { sh[warp][0] = a0; sh[warp][1] = a1; sh[warp][2] = a2; sh[warp][3] = a3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0;
    for (int q = 0; q < 8; ++q) v += sh[q][threadIdx.x];
    partials[(size_t)blockIdx.x * 4 + threadIdx.x] = v;
  }
}

// stats = { rss, tss, sum w r^2, 0 }
__global__ void k_resid_final(const double* __restrict__ partials, int nblocks, double n_local,
                              double* __restrict__ raw /*4*/) {
  if (threadIdx.x < 4) {
    double v = 0;
    for (int b = 0; b < nblocks; ++b) v += partials[(size_t)b * 4 + threadIdx.x];
    raw[threadIdx.x] = v;
  }
}
__global__ void k_stats_from_raw(const double* __restrict__ raw, double n_total, double* __restrict__ stats) {
  if (threadIdx.x == 0) {
    stats[0] = raw[0];
    stats[1] = raw[3] - raw[2] * raw[2] / n_total;  // sum (y - ybar)^2 from moments shifted by c = y[0]
    stats[2] = raw[1];
    stats[3] = 0.0;
  }
}

// S[key * K + c] += w * x_c * r   (row-major G x K; any key order; runs of equal keys are merged first)
template <int NC>
__global__ void __launch_bounds__(PJ_T) k_scores(const int32_t* __restrict__ key, ColPtrs X,
                                                 const double* __restrict__ r, const double* __restrict__ w,
                                                 int64_t n, double* __restrict__ S, int K, int col0, bool vec_codes) {
  const int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I);
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    const int64_t base = (ch * PJ_T + threadIdx.x) * PJ_I;
    int k[PJ_I];
    load4_i32(key, base, n, vec_codes, k);
    double v[PJ_I][NC];
#pragma unroll
    for (int e = 0; e < PJ_I; ++e) {
      bool valid = base + e < n;
      double rr = valid ? __ldcs(r + base + e) : 0.0;
      if (w && valid) rr *= __ldcs(w + base + e);
#pragma unroll
      for (int c = 0; c < NC; ++c) v[e][c] = valid ? __ldcs(X.col[c] + base + e) * rr : 0.0;
    }
    segmented_flush<NC>(k, v, S, K, col0, 1);
  }
}

__global__ void k_combine_keys(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t Gb, int64_t n,
                               uint64_t* __restrict__ out, int first) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (first ? (uint64_t)a[i] : out[i]) * (uint64_t)Gb + (uint64_t)b[i];
}

__global__ void __launch_bounds__(256) k_count_present(const int32_t* __restrict__ codes, int64_t n,
                                                       uint32_t* __restrict__ present) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    present[codes[i]] = 1u;
}
__global__ void __launch_bounds__(256) k_sum_u32(const uint32_t* __restrict__ v, int64_t n,
                                                 unsigned long long* __restrict__ out) {
  unsigned int local = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    local += v[i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, (unsigned long long)local);
}

// V += sign * B * M * B      (k x k, one CTA, tmp = k*k scratch)
__global__ void __launch_bounds__(256) k_sandwich(const double* __restrict__ B, const double* __restrict__ M, int k,
                                                  double sign, double* __restrict__ tmp, double* __restrict__ V) {
  for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x) {
    int i = idx / k, j = idx % k;
    double s = 0;
    for (int q = 0; q < k; ++q) s = fma(B[i * k + q], M[q * k + j], s);
    tmp[idx] = s;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x) {
    int i = idx / k, j = idx % k;
    double s = 0;
    for (int q = 0; q < k; ++q) s = fma(tmp[i * k + q], B[q * k + j], s);
    V[idx] += sign * s;
  }
}
// vcov = scale * V ; se = sqrt(max(diag, 0))
__global__ void k_scale_se(const double* __restrict__ V, int k, double scale, const double* __restrict__ scale_dev,
                           int scale_idx, double scale_div, double* __restrict__ vcov_out, double* __restrict__ se_out) {
  double sc = scale;
  if (scale_dev) sc *= scale_dev[scale_idx] / scale_div;
  for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x) {
    double v = sc * V[idx];
    if (vcov_out) vcov_out[idx] = v;
    if (idx / k == idx % k) se_out[idx / k] = sqrt(fmax(v, 0.0));
  }
}

// deterministic mode: S[g][col0 + c] = sum over the rows of cluster g, in row order, of w x_c r -- one warp per cluster
// walks its rows of the cluster-sorted permutation, fixed shuffle tree, plain stores (no floating-point atomics)
template <int NC>
__global__ void __launch_bounds__(256) k_det_scores(const uint32_t* __restrict__ off, int64_t G,
                                                    const uint32_t* __restrict__ perm, ColPtrs X,
                                                    const double* __restrict__ r, const double* __restrict__ w,
                                                    double* __restrict__ S, int K, int col0) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < G; g += nwarp) {
    double sum[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) sum[c] = 0.0;
    const uint32_t e1 = off[g + 1];
    for (uint32_t e = off[g] + lane; e < e1; e += 32) {
      const size_t row = perm[e];
      const double rw = w ? r[row] * w[row] : r[row];
#pragma unroll
      for (int c = 0; c < NC; ++c) sum[c] += X.col[c][row] * rw;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) sum[c] = warp_sum(sum[c]);
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < NC; ++c) S[(size_t)g * K + col0 + c] = sum[c];
    }
  }
}

static void scores_meat(lfe_ctx* h, const int32_t* cl, int64_t G, const double* const* xcols, int k, const double* resid,
                        const double* w, int64_t n, double* meat_dev, cudaStream_t s, bool cl_local = false) {
  DevBuf<double> S((size_t)G * k, s);
  S.zero();
  if (h->opt_deterministic && n > 0) {
    LFE_REQUIRE(h->world == 1, "LFE_OPT_DETERMINISTIC is single-GPU only");
    DevBuf<uint32_t> perm(n, s), off(G + 1, s);
    off.zero();
    prim_histogram(h, cl, n, G, off, s);
    prim_exclusive_scan_u32(h, off, off, G + 1, nullptr, s);
    prim_argsort_u32(h, (const uint32_t*)cl, n, bits_for((uint64_t)(G > 0 ? G - 1 : 0)), perm, nullptr, s);
    const int blocks = launch_blocks(h, ceil_div64(G, 8), 8);
    for (int c0 = 0; c0 < k; c0 += PJ_MAXNC) {
      const int nc = k - c0 < PJ_MAXNC ? k - c0 : PJ_MAXNC;
      ColPtrs cp;
      for (int c = 0; c < nc; ++c) cp.col[c] = xcols[c0 + c];
#define DSC_CASE(NC_) \
  case NC_: k_det_scores<NC_><<<blocks, 256, 0, s>>>(off, G, perm, cp, resid, w, S, k, c0); break;
      switch (nc) {
        DSC_CASE(1) DSC_CASE(2) DSC_CASE(3) DSC_CASE(4) DSC_CASE(5) DSC_CASE(6) DSC_CASE(7) DSC_CASE(8)
      }
#undef DSC_CASE
      KERNEL_CHECK();
      h->n_launches += 1;
    }
    std::vector<const double*> scd(k);
    for (int c = 0; c < k; ++c) scd[c] = S.p + c;
    gram_core(h, scd.data(), k, G, nullptr, nullptr, 0, meat_dev, s, k);
    return;
  }
  int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I);
  int blocks = launch_blocks(h, nchunk, 8);
  bool vec = (((uintptr_t)cl) & 15) == 0;
  for (int c0 = 0; c0 < k; c0 += PJ_MAXNC) {
    int nc = k - c0 < PJ_MAXNC ? k - c0 : PJ_MAXNC;
    ColPtrs cp;
    for (int c = 0; c < nc; ++c) cp.col[c] = xcols[c0 + c];
#define SC_CASE(NC_) \
  case NC_: k_scores<NC_><<<blocks, PJ_T, 0, s>>>(cl, cp, resid, w, n, S, k, c0, vec); break;
    switch (nc) {
      SC_CASE(1) SC_CASE(2) SC_CASE(3) SC_CASE(4) SC_CASE(5) SC_CASE(6) SC_CASE(7) SC_CASE(8)
    }
#undef SC_CASE
    KERNEL_CHECK();
    h->n_launches += 1;
  }
  // multi-GPU: clusters shared by ranks -> sum the score table; rank-local clusters -> sum the k x k meats
  if (h->world > 1 && !cl_local) comm_allreduce(h, S, G * k, 0, s);
  std::vector<const double*> sc(k);
  for (int c = 0; c < k; ++c) sc[c] = S.p + c;   // row-major table: column c has stride k
  gram_core(h, sc.data(), k, G, nullptr, nullptr, 0, meat_dev, s, k);
  if (h->world > 1 && cl_local) comm_allreduce(h, meat_dev, (int64_t)k * k, 0, s);
}

void comm_allreduce_int(lfe_ctx* h, void* buf, int64_t count, int kind, int op, cudaStream_t s);
static int64_t count_present(lfe_ctx* h, const int32_t* codes, int64_t n, int64_t G, cudaStream_t s,
                             bool cl_local = false) {
  DevBuf<uint32_t> present(G, s);
  present.zero();
  DevBuf<unsigned long long> tot(1, s);
  tot.zero();
  int blocks = launch_blocks(h, ceil_div64(n, 1024), 8);
  if (n > 0) {
    k_count_present<<<blocks, 256, 0, s>>>(codes, n, present);
    KERNEL_CHECK();
  }
  if (h->world > 1 && !cl_local) comm_allreduce_int(h, present.p, G, 0, 2, s);   // shared code space: OR the flags
  k_sum_u32<<<launch_blocks(h, ceil_div64(G, 1024), 8), 256, 0, s>>>(present, G, tot);
  KERNEL_CHECK();
  if (h->world > 1 && cl_local) comm_allreduce_int(h, tot.p, 1, 1, 0, s);          // disjoint clusters: add the counts
  h->n_launches += 2;
  unsigned long long ht = 0;
  CUDA_CHECK(cudaMemcpyAsync(&ht, tot.p, sizeof(ht), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  return (int64_t)ht;
}


// Clustered sandwich shared by the row-level and the YOCO paths.  r = per-row "score residual"
// (resid, or e0_g = sum_y - n yhat on compressed rows), w optional multiplier.  V must be zeroed.
static void cluster_sandwich(lfe_ctx* h, const int32_t* const* cl_codes, const int64_t* n_cl_groups, int W,
                             const double* const* xc, int k, const double* r, const double* w, int64_t n,
                             const double* Binv, int ssc, int64_t n_obs, int64_t df_resid, double* V, double* tmp,
                             double* meat, double* se_out, double* vcov_out, int64_t* n_clusters_out, cudaStream_t s,
                             bool cl_local = false) {
  if (W == 1) {
    // std_errors.py:289-347 / compress.py:922-953
    int64_t G = count_present(h, cl_codes[0], n, n_cl_groups[0], s, cl_local);
    if (n_clusters_out) n_clusters_out[0] = G;
    LFE_REQUIRE(G > 1, "clustered vcov: a single cluster (division by zero in G/(G-1))");
    scores_meat(h, cl_codes[0], n_cl_groups[0], xc, k, r, w, n, meat, s, cl_local);
    k_sandwich<<<1, 256, 0, s>>>(Binv, meat, k, 1.0, tmp, V);
    KERNEL_CHECK();
    double adj = (double)G / (double)(G - 1);
    if (ssc) adj *= (double)(n_obs - 1) / (double)df_resid;
    k_scale_se<<<1, 256, 0, s>>>(V, k, adj, nullptr, 0, 1.0, vcov_out, se_out);
    KERNEL_CHECK();
    h->n_launches += 2;
    return;
  }
  // multiway, Cameron-Gelbach-Miller inclusion-exclusion      std_errors.py:354-441 / compress.py:956-1042
  LFE_REQUIRE(h->world == 1, "multiway clustering is single-GPU only");
  int64_t Gmin = -1;
  DevBuf<uint64_t> ckey(n, s);
  DevBuf<int32_t> ccode(n, s);
  int kblocks = launch_blocks(h, ceil_div64(n, 1024), 8);
  for (int mask = 1; mask < (1 << W); ++mask) {
    int size = __builtin_popcount(mask);
    double sign = (size & 1) ? 1.0 : -1.0;
    int64_t G, Gcap;
    const int32_t* codes;
    if (size == 1) {
      int a = __builtin_ctz(mask);
      codes = cl_codes[a];
      Gcap = n_cl_groups[a];
      G = count_present(h, codes, n, Gcap, s);
      if (n_clusters_out) n_clusters_out[a] = G;
      Gmin = (Gmin < 0 || G < Gmin) ? G : Gmin;
    } else {
      int first = 1, prev = -1;
      double space = 1.0;
      for (int a = 0; a < W; ++a) {
        if (!(mask & (1 << a))) continue;
        space *= (double)n_cl_groups[a];
        if (prev < 0) { prev = a; continue; }
        k_combine_keys<<<kblocks, 256, 0, s>>>(cl_codes[prev], cl_codes[a], n_cl_groups[a], n, ckey, first);
        KERNEL_CHECK();
        h->n_launches += 1;
        first = 0;
      }
      LFE_REQUIRE(space < 1.8e19, "cluster intersection key space overflows 64 bits");
      G = factorize_dispatch(h, ckey, LFE_U64, n, LFE_ORDER_SORTED, ccode, nullptr, 0, s);
      Gcap = G;
      codes = ccode;
    }
    if (G <= 1) continue;
    scores_meat(h, codes, Gcap, xc, k, r, w, n, meat, s);
    k_sandwich<<<1, 256, 0, s>>>(Binv, meat, k, sign, tmp, V);
    KERNEL_CHECK();
    h->n_launches += 1;
  }
  double adj = 1.0;
  if (Gmin > 2) adj *= (double)Gmin / (double)(Gmin - 1);
  if (ssc) adj *= (double)(n_obs - 1) / (double)df_resid;
  k_scale_se<<<1, 256, 0, s>>>(V, k, adj, nullptr, 0, 1.0, vcov_out, se_out);
  KERNEL_CHECK();
  h->n_launches += 1;
}

// ---- one-way clustering in ONE pass over the columns --------------------------------------------------
// r = y - b0 - x'beta, the rss / tss moments, the cluster presence flags and the scores S_g += w x r are all
// produced while each row is in registers: the regressors are read once instead of twice and the residual
// vector is never materialised.
template <int NC>
__global__ void __launch_bounds__(PJ_T) k_resid_scores(const double* __restrict__ y, ColPtrs X,
                                                       const double* __restrict__ beta_full,
                                                       const double* __restrict__ w, const int32_t* __restrict__ key,
                                                       int64_t n, double* __restrict__ S, int ldS,
                                                       uint32_t* __restrict__ present, double* __restrict__ partials,
                                                       int use_shift, bool vec_codes, bool bulk) {
  // bulk: score rows leave as one TMA reduce each (segreduce.cuh: bulk_add_row) unless neighbouring rows share a
  // cluster; the per-thread staging slots are the dynamic shared memory
  extern __shared__ __align__(16) double rs_slots[];
  double* my_slots = rs_slots + (size_t)threadIdx.x * BulkRow<NC>::SLOT_DOUBLES;
  int slot = 0;
  __shared__ double sb[NC + 1];
  if (threadIdx.x <= NC) sb[threadIdx.x] = beta_full[threadIdx.x];
  __syncthreads();
  const double c0 = use_shift ? y[0] : 0.0;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  const int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I);
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    const int64_t base = (ch * PJ_T + threadIdx.x) * PJ_I;
    int kk[PJ_I];
    load4_i32(key, base, n, vec_codes, kk);
    double v[PJ_I][NC];
#pragma unroll
    for (int e = 0; e < PJ_I; ++e) {
      const bool valid = base + e < n;
      const double yi = valid ? __ldcs(y + base + e) : 0.0;
      double r = yi - sb[0];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        v[e][c] = valid ? __ldcs(X.col[c] + base + e) : 0.0;
        r = fma(-sb[c + 1], v[e][c], r);
      }
      if (!valid) r = 0.0;
      const double we = (w && valid) ? __ldcs(w + base + e) : 1.0;
      const double r2 = r * r;
      a0 += r2;
      a1 += we * r2;
      const double d = valid ? yi - c0 : 0.0;
      a2 += d;
      a3 = fma(d, d, a3);
      const double rw = r * we;
#pragma unroll
      for (int c = 0; c < NC; ++c) v[e][c] *= rw;
      if (valid) present[kk[e]] = 1u;
    }
    if (bulk && warp_keys_all_distinct_neighbours(kk)) {
#pragma unroll
      for (int e = 0; e < PJ_I; ++e)
        if (kk[e] >= 0) bulk_add_row<NC>(S + (size_t)kk[e] * ldS, v[e], my_slots, slot);
    } else {
      segmented_flush<NC>(kk, v, S, ldS, 0, 1);
    }
  }
  if (bulk) bulk_add_drain();
  __shared__ double sh[PJ_T / 32][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
  if (lane == 0) { sh[warp][0] = a0; sh[warp][1] = a1; sh[warp][2] = a2; sh[warp][3] = a3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int q = 0; q < PJ_T / 32; ++q) t += sh[q][threadIdx.x];
    partials[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
  }
}

static void fused_oneway_cluster(lfe_ctx* h, const double* const* cols, int k, int64_t n, const double* beta_full,
                                 const double* Binv, const int32_t* cl, int64_t Gcap, const double* weights, int ssc,
                                 int64_t n_obs, int64_t df_resid, bool cl_local, double* se_out, double* vcov_out,
                                 double* stats_out, int64_t* n_clusters_out, cudaStream_t s) {
  ColPtrs X;
  for (int j = 0; j < k; ++j) X.col[j] = cols[1 + j];
  const int64_t nchunk = ceil_div64(n, (int64_t)PJ_T * PJ_I);
  const int blocks = launch_blocks(h, nchunk, 8);
  // score rows padded to a 16-byte multiple for the TMA bulk reduce (k = 5 -> 6; the pad stays zero); narrow
  // rows keep their width and the RED path
  const int ldS = (k >= 5 && (k & 1)) ? k + 1 : k;
  DevBuf<double> S((size_t)Gcap * ldS, s), partials((size_t)blocks * 4, s), raw(4, s), stats_local(4, s);
  DevBuf<uint32_t> present(Gcap, s);
  DevBuf<unsigned long long> tot(1, s);
  S.zero(); present.zero(); tot.zero();
  const bool vec = (((uintptr_t)cl) & 15) == 0;
  const int shift = h->world == 1 ? 1 : 0;
  const bool bulk = n >= 4096 && k >= 3 && (ldS & 1) == 0;
  const size_t slot_bytes = bulk ? (size_t)PJ_T * BR_SLOTS * ((k + 1) & ~1) * sizeof(double) : 0;
#define RS_CASE(NC_)                                                                                              \
  case NC_:                                                                                                       \
    if (slot_bytes > 40 * 1024)                                                                                   \
      CUDA_CHECK(cudaFuncSetAttribute(k_resid_scores<NC_>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                      (int)slot_bytes));                                                          \
    k_resid_scores<NC_><<<blocks, PJ_T, slot_bytes, s>>>(cols[0], X, beta_full, weights, cl, n, S, ldS, present,  \
                                                         partials, shift, vec, bulk);                             \
    break;
  switch (k) { RS_CASE(1) RS_CASE(2) RS_CASE(3) RS_CASE(4) RS_CASE(5) RS_CASE(6) RS_CASE(7) RS_CASE(8) }
#undef RS_CASE
  KERNEL_CHECK();
  k_resid_final<<<1, 32, 0, s>>>(partials, blocks, (double)n, raw);
  KERNEL_CHECK();
  if (h->world > 1) comm_allreduce(h, raw, 4, 0, s);
  double* stats = stats_out ? stats_out : stats_local.p;
  k_stats_from_raw<<<1, 32, 0, s>>>(raw, (double)n_obs, stats);
  KERNEL_CHECK();
  // number of clusters with at least one row
  if (h->world > 1 && !cl_local) comm_allreduce_int(h, present.p, Gcap, 0, 2, s);
  k_sum_u32<<<launch_blocks(h, ceil_div64(Gcap, 1024), 8), 256, 0, s>>>(present, Gcap, tot);
  KERNEL_CHECK();
  if (h->world > 1 && cl_local) comm_allreduce_int(h, tot.p, 1, 1, 0, s);
  unsigned long long ht = 0;
  CUDA_CHECK(cudaMemcpyAsync(&ht, tot.p, sizeof(ht), cudaMemcpyDeviceToHost, s));
  // meat = S'S (cross-rank: sum the score table for shared clusters, the k x k meats for rank-local ones)
  DevBuf<double> V((size_t)k * k, s), tmp((size_t)k * k, s), meat((size_t)k * k, s);
  V.zero();
  if (h->world > 1 && !cl_local) comm_allreduce(h, S, Gcap * ldS, 0, s);
  std::vector<const double*> sc(k);
  for (int c = 0; c < k; ++c) sc[c] = S.p + c;
  gram_core(h, sc.data(), k, Gcap, nullptr, nullptr, 0, meat, s, ldS);
  if (h->world > 1 && cl_local) comm_allreduce(h, meat, (int64_t)k * k, 0, s);
  k_sandwich<<<1, 256, 0, s>>>(Binv, meat, k, 1.0, tmp, V);
  KERNEL_CHECK();
  CUDA_CHECK(cudaStreamSynchronize(s));
  const int64_t G = (int64_t)ht;
  if (n_clusters_out) n_clusters_out[0] = G;
  LFE_REQUIRE(G > 1, "clustered vcov: a single cluster (division by zero in G/(G-1))");
  double adj = (double)G / (double)(G - 1);
  if (ssc) adj *= (double)(n_obs - 1) / (double)df_resid;
  k_scale_se<<<1, 256, 0, s>>>(V, k, adj, nullptr, 0, 1.0, vcov_out, se_out);
  KERNEL_CHECK();
  h->n_launches += 7;
}

// ---- the same pass fed from (raw columns, FE tables) ---------------------------------------------------
// x~ / y~ are formed in registers (tail.cuh), so lfe_demean_plan_fit never writes the demeaned columns.
// MODE 0: IID (rss / tss moments only); 1: HC1 (meat triangle of x~ with omega = w r^2); 2: one-way cluster scores.
template <int NC, int R, bool VEC, int MODE>
__global__ void __launch_bounds__(TL_T) k_tail_resid(TailCols cols, TailTabs T, const double* __restrict__ beta_full,
                                                     const double* __restrict__ w, const int32_t* __restrict__ key,
                                                     int64_t n, int ldc, double* __restrict__ S, int ldS,
                                                     uint32_t* __restrict__ present, double* __restrict__ partials,
                                                     double* __restrict__ meat_partials, bool bulk) {
  constexpr int K = NC - 1;                       // regressors
  constexpr int KS = K > 0 ? K : 1;
  constexpr int NTM = KS * (KS + 1) / 2;
  extern __shared__ __align__(16) double tl_smem[];
  const double* tabs[LFE_MAX_FE];
  // dynamic shared memory: [bulk-reduce staging slots of all threads][staged tables]
  const int slot_doubles = (MODE == 2 && bulk) ? TL_T * BulkRow<KS>::SLOT_DOUBLES : 0;
  double* my_slots = tl_smem + (size_t)threadIdx.x * BulkRow<KS>::SLOT_DOUBLES;
  stage_tables(T, ldc, tl_smem + slot_doubles, tabs);
  int slot = 0;
  __shared__ double sb[NC];
  if (threadIdx.x < NC) sb[threadIdx.x] = beta_full[threadIdx.x];
  __syncthreads();
  const bool vec2 = (ldc & 1) == 0;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  double macc[MODE == 1 ? NTM : 1];
#pragma unroll
  for (int t = 0; t < (MODE == 1 ? NTM : 1); ++t) macc[t] = 0.0;
  // chunk loop: uniform per CTA, so the warp collectives of the score flush see all 32 lanes
  const int64_t nchunk = ceil_div64(n, (int64_t)TL_T * R);
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    const int64_t base = (ch * TL_T + threadIdx.x) * R;
    double x[R][NC];
    bool valid[R];
    double v[R][KS];
    int kk[R];
    if (VEC && base + R <= n) {
      tail_rows<NC, R, true>(cols, T, tabs, ldc, vec2, base, n, x, valid);
    } else {
      tail_rows<NC, R, false>(cols, T, tabs, ldc, vec2, base, n, x, valid);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double res = x[r][0] - sb[0];
#pragma unroll
      for (int c = 1; c < NC; ++c) res = fma(-sb[c], x[r][c], res);
      if (!valid[r]) res = 0.0;
      const double we = (w && valid[r]) ? __ldcs(w + base + r) : 1.0;
      const double r2 = res * res;
      a0 += r2;
      a1 += we * r2;
      const double d = valid[r] ? x[r][0] : 0.0;
      a2 += d;
      a3 = fma(d, d, a3);
      if (MODE == 1) {
        const double om = we * r2;
        int t = 0;
#pragma unroll
        for (int a = 0; a < K; ++a) {
          const double za = x[r][1 + a] * om;
#pragma unroll
          for (int b = a; b < K; ++b) {
            macc[t] = fma(za, x[r][1 + b], macc[t]);
            ++t;
          }
        }
      }
      if (MODE == 2) {
        const double rw = res * we;
#pragma unroll
        for (int c = 0; c < K; ++c) v[r][c] = x[r][1 + c] * rw;
        kk[r] = valid[r] ? __ldcs(key + base + r) : -1;
        if (valid[r]) present[kk[r]] = 1u;
      }
    }
    if (MODE == 2) {
      if (bulk && warp_keys_all_distinct_neighbours<R>(kk)) {
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (kk[r] >= 0) bulk_add_row<KS>(S + (size_t)kk[r] * ldS, v[r], my_slots, slot);
      } else {
        segmented_flush<KS, R>(kk, v, S, ldS, 0, 1);
      }
    }
  }
  if (MODE == 2 && bulk) bulk_add_drain();
  __shared__ double sh[TL_T / 32][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
  if (lane == 0) { sh[warp][0] = a0; sh[warp][1] = a1; sh[warp][2] = a2; sh[warp][3] = a3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int q = 0; q < TL_T / 32; ++q) t += sh[q][threadIdx.x];
    partials[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
  }
  if (MODE == 1) {
    __shared__ double shm[TL_T / 32][NTM];
#pragma unroll
    for (int t = 0; t < NTM; ++t) {
      const double q = warp_sum(macc[t]);
      if (lane == 0) shm[warp][t] = q;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < NTM; t += TL_T) {
      double q = 0.0;
#pragma unroll
      for (int u = 0; u < TL_T / 32; ++u) q += shm[u][t];
      meat_partials[(size_t)blockIdx.x * NTM + t] = q;
    }
  }
}

__global__ void k_u64_to_f64(const unsigned long long* __restrict__ in, double* __restrict__ out) { *out = (double)*in; }

__global__ void __launch_bounds__(256) k_meat_final(const double* __restrict__ partials, int nblocks, int K,
                                                    double* __restrict__ out) {
  const int NT = K * (K + 1) / 2;
  for (int t = threadIdx.x; t < NT; t += blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * NT + t];
    int a = 0, rem = t;
    while (rem >= K - a) { rem -= K - a; ++a; }
    const int bb = a + rem;
    out[a * K + bb] = s;
    out[bb * K + a] = s;
  }
}

template <int NC, int MODE>
static void launch_tail_resid(lfe_ctx* h, const TailCols& tc, const TailTabs& T, const double* beta_full, const double* w,
                              const int32_t* key, int64_t n, int ldc, size_t tab_smem, bool vec, double* S, int ldS,
                              uint32_t* present, double* partials, double* meat_partials, bool bulk, int blocks,
                              cudaStream_t s) {
  constexpr int R = TailCfg<NC>::R;
  constexpr int KS = NC - 1 > 0 ? NC - 1 : 1;
  const size_t smem = tab_smem + ((MODE == 2 && bulk) ? (size_t)TL_T * BulkRow<KS>::SLOT_DOUBLES * sizeof(double) : 0);
  if (vec && R > 1) {
    if (smem > 32 * 1024)
      CUDA_CHECK(cudaFuncSetAttribute(k_tail_resid<NC, R, true, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tail_resid<NC, R, true, MODE><<<blocks, TL_T, smem, s>>>(tc, T, beta_full, w, key, n, ldc, S, ldS, present, partials,
                                                               meat_partials, bulk);
  } else {
    if (smem > 32 * 1024)
      CUDA_CHECK(cudaFuncSetAttribute(k_tail_resid<NC, 1, false, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tail_resid<NC, 1, false, MODE><<<blocks, TL_T, smem, s>>>(tc, T, beta_full, w, key, n, ldc, S, ldS, present, partials,
                                                                meat_partials, bulk);
  }
  KERNEL_CHECK();
  h->n_launches += 1;
}

// residual statistics + IID / HC1 / one-way clustered SEs of the fit whose demeaned columns are (cols, T): the
// counterpart of resid_vcov_impl without materialised x~.  cols[0] = y (raw), C = k + 1 <= TL_MAXC.
void resid_vcov_fused_impl(lfe_ctx* h, const double* const* cols, int C, int64_t n, const TailTabs& T, int ldc,
                           size_t tab_smem, const double* beta_full, const double* Binv, int vcov_kind, const int32_t* cl,
                           int64_t Gcap, const double* weights, int ssc, int64_t n_obs, int64_t df_resid, double* se_out,
                           double* vcov_out, double* stats_out, int64_t* n_clusters_out, cudaStream_t s) {
  PhaseTimer t(h, s, 11);
  const bool cl_local = (vcov_kind & LFE_VCOV_CLUSTER_LOCAL) != 0;
  vcov_kind &= ~LFE_VCOV_CLUSTER_LOCAL;
  const int k = C - 1;
  LFE_REQUIRE(C >= 1 && C <= TL_MAXC, "resid_vcov_fused: C=%d out of range", C);
  TailCols tc;
  for (int c = 0; c < C; ++c) tc.col[c] = cols[c];
  const bool vec = tail_aligned(tc, C, T, weights) && (vcov_kind != LFE_VCOV_CLUSTER || ((((uintptr_t)cl) & 15) == 0));
  // 5-8 columns: 2 rows per thread, 78 registers -> three CTAs per SM (c2: 1.97 -> 1.74 ms; four thrash the L2 between
  // the unit table and the score table: 2.36 ms)
  const int blocks = launch_blocks(h, ceil_div64(n, (int64_t)TL_T * 2), C <= 4 ? 2 : (C <= 8 ? 3 : 1));
  const int ldS = (k >= 5 && (k & 1)) ? k + 1 : k;
  const int ntm = k > 0 ? k * (k + 1) / 2 : 1;
  DevBuf<double> partials((size_t)blocks * 4, s), raw(4, s), stats_local(4, s);
  DevBuf<double> S, meat_partials;
  DevBuf<uint32_t> present;
  DevBuf<unsigned long long> tot(1, s);
  tot.zero();
  const bool bulk = vcov_kind == LFE_VCOV_CLUSTER && n >= 4096 && k >= 3 && (ldS & 1) == 0;
  if (vcov_kind == LFE_VCOV_CLUSTER) {
    S.alloc((size_t)Gcap * ldS, s); S.zero();
    present.alloc(Gcap, s); present.zero();
  }
  if (vcov_kind == LFE_VCOV_HC1) meat_partials.alloc((size_t)blocks * ntm, s);
  if (n > 0) {
#define TR_CASE(NC_)                                                                                                   \
  case NC_:                                                                                                            \
    if (vcov_kind == LFE_VCOV_IID || k == 0)                                                                           \
      launch_tail_resid<NC_, 0>(h, tc, T, beta_full, weights, cl, n, ldc, tab_smem, vec, S, ldS, present, partials,    \
                                meat_partials, false, blocks, s);                                                      \
    else if (vcov_kind == LFE_VCOV_HC1)                                                                                \
      launch_tail_resid<NC_, 1>(h, tc, T, beta_full, weights, cl, n, ldc, tab_smem, vec, S, ldS, present, partials,    \
                                meat_partials, false, blocks, s);                                                      \
    else                                                                                                               \
      launch_tail_resid<NC_, 2>(h, tc, T, beta_full, weights, cl, n, ldc, tab_smem, vec, S, ldS, present, partials,    \
                                meat_partials, bulk, blocks, s);                                                       \
    break;
    switch (C) {
      TR_CASE(1) TR_CASE(2) TR_CASE(3) TR_CASE(4) TR_CASE(5) TR_CASE(6) TR_CASE(7) TR_CASE(8) TR_CASE(9) TR_CASE(10)
      TR_CASE(11)
    }
#undef TR_CASE
    k_resid_final<<<1, 32, 0, s>>>(partials, blocks, (double)n, raw);
    KERNEL_CHECK();
  } else {
    raw.zero();
  }
  // Everything the ranks must combine is summed in ONE all-reduce of pack = [raw moments (4) | #clusters (1) | meat
  // (k x k)]: the meat is additive over ranks for HC1 and for rank-local clusters (cluster == shard key); clusters
  // shared by ranks need their score table summed first (a second, G x k collective).
  const bool dist = h->world > 1;
  DevBuf<double> pack((size_t)5 + (size_t)k * k, s);
  pack.zero();
  CUDA_CHECK(cudaMemcpyAsync(pack.p, raw.p, 4 * sizeof(double), cudaMemcpyDeviceToDevice, s));
  double* meat = pack.p + 5;
  double* stats = stats_out ? stats_out : stats_local.p;
  DevBuf<double> V((size_t)(k > 0 ? k * k : 1), s), tmp((size_t)(k > 0 ? k * k : 1), s);
  V.zero();
  if (k > 0 && vcov_kind == LFE_VCOV_HC1 && n > 0) {
    k_meat_final<<<1, 256, 0, s>>>(meat_partials, blocks, k, meat);
    KERNEL_CHECK();
    h->n_launches += 1;
  }
  if (k > 0 && vcov_kind == LFE_VCOV_CLUSTER) {
    if (dist && !cl_local) {
      comm_allreduce_int(h, present.p, Gcap, 0, 2, s);
      comm_allreduce(h, S, Gcap * ldS, 0, s);
    }
    k_sum_u32<<<launch_blocks(h, ceil_div64(Gcap, 1024), 8), 256, 0, s>>>(present, Gcap, tot);
    KERNEL_CHECK();
    k_u64_to_f64<<<1, 1, 0, s>>>(tot, pack.p + 4);
    KERNEL_CHECK();
    std::vector<const double*> sc(k);
    for (int c = 0; c < k; ++c) sc[c] = S.p + c;
    gram_core(h, sc.data(), k, Gcap, nullptr, nullptr, 0, meat, s, ldS);
    h->n_launches += 2;
    if (dist && !cl_local) {
      // the table is already global: every rank holds the full meat and count; keep one copy in the sum below
      if (h->rank != 0) CUDA_CHECK(cudaMemsetAsync(pack.p + 4, 0, (size_t)(1 + k * k) * sizeof(double), s));
    }
  }
  if (dist) comm_allreduce(h, pack, 5 + (int64_t)k * k, 0, s);
  k_stats_from_raw<<<1, 32, 0, s>>>(pack, (double)n_obs, stats);
  KERNEL_CHECK();
  h->n_launches += 2;
  if (k == 0) return;
  if (vcov_kind == LFE_VCOV_IID) {
    CUDA_CHECK(cudaMemcpyAsync(V.p, Binv, (size_t)k * k * sizeof(double), cudaMemcpyDeviceToDevice, s));
    k_scale_se<<<1, 256, 0, s>>>(V, k, 1.0, stats, 2, (double)df_resid, vcov_out, se_out);
    KERNEL_CHECK();
    h->n_launches += 1;
    return;
  }
  k_sandwich<<<1, 256, 0, s>>>(Binv, meat, k, 1.0, tmp, V);
  KERNEL_CHECK();
  if (vcov_kind == LFE_VCOV_HC1) {
    k_scale_se<<<1, 256, 0, s>>>(V, k, (double)n_obs / (double)df_resid, nullptr, 0, 1.0, vcov_out, se_out);
    KERNEL_CHECK();
    h->n_launches += 2;
    return;
  }
  double hG = 0.0;
  CUDA_CHECK(cudaMemcpyAsync(&hG, pack.p + 4, sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  const int64_t G = (int64_t)(hG + 0.5);
  if (n_clusters_out) n_clusters_out[0] = G;
  LFE_REQUIRE(G > 1, "clustered vcov: a single cluster (division by zero in G/(G-1))");
  double adj = (double)G / (double)(G - 1);
  if (ssc) adj *= (double)(n_obs - 1) / (double)df_resid;
  k_scale_se<<<1, 256, 0, s>>>(V, k, adj, nullptr, 0, 1.0, vcov_out, se_out);
  KERNEL_CHECK();
  h->n_launches += 2;
}

void resid_vcov_impl(lfe_ctx* h, const double* const* cols, int k, int64_t n, const double* beta_full,
                     const double* Binv, int vcov_kind, const int32_t* const* cl_codes, const int64_t* n_cl_groups, int W,
                     const double* weights, int ssc, int64_t n_obs, int64_t df_resid, double* se_out, double* vcov_out,
                     double* stats_out, double* resid_out, int64_t* n_clusters_out, cudaStream_t s) {
  PhaseTimer t(h, s, 11);
  const bool cl_local = (vcov_kind & LFE_VCOV_CLUSTER_LOCAL) != 0;
  vcov_kind &= ~LFE_VCOV_CLUSTER_LOCAL;
  LFE_REQUIRE(k >= 0 && k <= RV_MAXK, "lfe_resid_vcov: k=%d out of range", k);
  LFE_REQUIRE(vcov_kind >= 0 && vcov_kind <= 2, "lfe_resid_vcov: unknown vcov kind %d", vcov_kind);
  LFE_REQUIRE(n >= 1 || h->world > 1, "lfe_resid_vcov: no rows");
  if (vcov_kind == LFE_VCOV_CLUSTER) LFE_REQUIRE(W >= 1 && W <= LFE_MAX_CLUSTER_WAYS, "lfe_resid_vcov: cluster_cols required (1..%d ways)", LFE_MAX_CLUSTER_WAYS);
  if (vcov_kind == LFE_VCOV_CLUSTER && W == 1 && k >= 1 && k <= PJ_MAXNC && !resid_out && !h->opt_deterministic) {
    fused_oneway_cluster(h, cols, k, n, beta_full, Binv, cl_codes[0], n_cl_groups[0], weights, ssc, n_obs, df_resid,
                         cl_local, se_out, vcov_out, stats_out, n_clusters_out, s);
    return;
  }
  DevBuf<double> resid_buf;
  double* resid = resid_out;
  if (!resid) {
    resid_buf.alloc(n, s);
    resid = resid_buf;
  }
  XCols X;
  for (int j = 0; j < k; ++j) X.col[j] = cols[1 + j];
  int blocks = launch_blocks(h, ceil_div64(n, 512), 4);
  DevBuf<double> partials((size_t)blocks * 4, s), raw(4, s), stats_local(4, s);
  k_resid<<<blocks, 256, 0, s>>>(cols[0], X, k, beta_full, weights, n, resid, partials, h->world == 1 ? 1 : 0);
  KERNEL_CHECK();
  k_resid_final<<<1, 32, 0, s>>>(partials, blocks, (double)n, raw);
  KERNEL_CHECK();
  if (h->world > 1) comm_allreduce(h, raw, 4, 0, s);
  double* stats = stats_out ? stats_out : stats_local.p;
  k_stats_from_raw<<<1, 32, 0, s>>>(raw, (double)n_obs, stats);
  KERNEL_CHECK();
  h->n_launches += 3;
  if (k == 0) return;

  DevBuf<double> V((size_t)k * k, s), tmp((size_t)k * k, s), meat((size_t)k * k, s);
  V.zero();
  std::vector<const double*> xc(k);
  for (int j = 0; j < k; ++j) xc[j] = cols[1 + j];

  if (vcov_kind == LFE_VCOV_IID) {
    // se = sqrt(max(sum(w r^2)/df * diag B, 0))      std_errors.py:183-210
    CUDA_CHECK(cudaMemcpyAsync(V.p, Binv, (size_t)k * k * sizeof(double), cudaMemcpyDeviceToDevice, s));
    k_scale_se<<<1, 256, 0, s>>>(V, k, 1.0, stats, 2, (double)df_resid, vcov_out, se_out);
    KERNEL_CHECK();
    h->n_launches += 1;
    return;
  }
  if (vcov_kind == LFE_VCOV_HC1) {
    // meat_ij = sum w x_i x_j r^2 ; V = B meat B * n/df      std_errors.py:217-282
    gram_core(h, xc.data(), k, n, weights, resid, 0, meat, s);
    if (h->world > 1) comm_allreduce(h, meat, (int64_t)k * k, 0, s);
    k_sandwich<<<1, 256, 0, s>>>(Binv, meat, k, 1.0, tmp, V);
    KERNEL_CHECK();
    k_scale_se<<<1, 256, 0, s>>>(V, k, (double)n_obs / (double)df_resid, nullptr, 0, 1.0, vcov_out, se_out);
    KERNEL_CHECK();
    h->n_launches += 2;
    return;
  }
  cluster_sandwich(h, cl_codes, n_cl_groups, W, xc.data(), k, resid, weights, n, Binv, ssc, n_obs, df_resid, V, tmp,
                   meat, se_out, vcov_out, n_clusters_out, s, cl_local);
}

// ---------------------------------------------------------------------------------------------
// YOCO standard errors on the compressed table   (compress.py:754-810, 854-1042, 1118-1137)
// ---------------------------------------------------------------------------------------------
// The reference solves the full dummy-variable WLS and sandwiches with rows of (M'WM)^-1.  By
// Frisch-Waugh-Lovell the x-block of that inverse is B = (X~' W X~)^-1 and (M'WM)^-1[x,:] M' = B X~',
// with X~ the W-weighted FE-demeaned regressors, so every estimator below is identical to the
// reference's without ever materialising the FE dummies:
//   e_g   = y~_g - x~_g'beta           (== mean_y_g - yhat_g)
//   rss_g = sum_y_sq - 2 yhat sum_y + n yhat^2 ,  e0_g = sum_y - n yhat
//   iid : rss/df * B ; hc1 : n/df * B (X~' diag(rss_g) X~) B ; cluster : scores = sum_g x~_g e0_g
__global__ void __launch_bounds__(256) k_yoco_rows(const double* __restrict__ yt, XCols X, int k,
                                                   const double* __restrict__ beta, const double* __restrict__ nn,
                                                   const double* __restrict__ sum_y, const double* __restrict__ sum_y_sq,
                                                   int64_t G, double* __restrict__ rss_g, double* __restrict__ e0_g,
                                                   double* __restrict__ rss_total) {
  double local = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x) {
    double e = yt[g];
    for (int j = 0; j < k; ++j) e = fma(-beta[j], X.col[j][g], e);
    double n = nn[g], sy = sum_y[g];
    double yhat = sy / n - e;
    double rg = sum_y_sq[g] - 2.0 * yhat * sy + n * (yhat * yhat);
    rss_g[g] = rg;
    e0_g[g] = sy - n * yhat;
    local += rg;
  }
  local = warp_sum(local);
  if (gridDim.x == 1) {   // deterministic mode launches ONE block: the warps' partials are added in warp order
    __shared__ double sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += sh[q];
      *rss_total += t;
    }
    return;
  }
  if ((threadIdx.x & 31) == 0) atomicAdd(rss_total, local);
}

void yoco_vcov_impl(lfe_ctx* h, const double* const* cols, int k, int64_t G, const double* beta, const double* Binv,
                    const double* nn, const double* sum_y, const double* sum_y_sq, int vcov_kind,
                    const int32_t* const* cl_codes, const int64_t* n_cl_groups, int W, int ssc, int64_t n_obs,
                    int64_t df_resid, double* se_out, double* vcov_out, double* rss_out, int64_t* n_clusters_out,
                    cudaStream_t s) {
  PhaseTimer t(h, s, 11);
  LFE_REQUIRE(k >= 1 && k <= RV_MAXK, "lfe_yoco_vcov: k=%d out of range", k);
  LFE_REQUIRE(G >= 1, "lfe_yoco_vcov: empty table");
  LFE_REQUIRE(vcov_kind >= 0 && vcov_kind <= 2, "lfe_yoco_vcov: unknown vcov kind %d", vcov_kind);
  if (vcov_kind == LFE_VCOV_CLUSTER) LFE_REQUIRE(W >= 1 && W <= LFE_MAX_CLUSTER_WAYS, "lfe_yoco_vcov: cluster codes required");
  XCols X;
  std::vector<const double*> xc(k);
  for (int j = 0; j < k; ++j) X.col[j] = xc[j] = cols[1 + j];
  DevBuf<double> rss_g(G, s), e0_g(G, s), rss_loc(1, s);
  double* rss = rss_out ? rss_out : rss_loc.p;
  CUDA_CHECK(cudaMemsetAsync(rss, 0, sizeof(double), s));
  k_yoco_rows<<<h->opt_deterministic ? 1 : launch_blocks(h, ceil_div64(G, 256), 4), 256, 0, s>>>(cols[0], X, k, beta, nn, sum_y, sum_y_sq, G, rss_g,
                                                                      e0_g, rss);
  KERNEL_CHECK();
  h->n_launches += 1;
  DevBuf<double> V((size_t)k * k, s), tmp((size_t)k * k, s), meat((size_t)k * k, s);
  V.zero();
  if (vcov_kind == LFE_VCOV_IID) {
    CUDA_CHECK(cudaMemcpyAsync(V.p, Binv, (size_t)k * k * sizeof(double), cudaMemcpyDeviceToDevice, s));
    k_scale_se<<<1, 256, 0, s>>>(V, k, 1.0, rss, 0, (double)df_resid, vcov_out, se_out);
    KERNEL_CHECK();
    h->n_launches += 1;
    return;
  }
  if (vcov_kind == LFE_VCOV_HC1) {
    gram_core(h, xc.data(), k, G, rss_g, nullptr, 0, meat, s);
    k_sandwich<<<1, 256, 0, s>>>(Binv, meat, k, 1.0, tmp, V);
    KERNEL_CHECK();
    k_scale_se<<<1, 256, 0, s>>>(V, k, (double)n_obs / (double)df_resid, nullptr, 0, 1.0, vcov_out, se_out);
    KERNEL_CHECK();
    h->n_launches += 2;
    return;
  }
  cluster_sandwich(h, cl_codes, n_cl_groups, W, xc.data(), k, e0_g, nullptr, G, Binv, ssc, n_obs, df_resid, V, tmp, meat,
                   se_out, vcov_out, n_clusters_out, s);
}
