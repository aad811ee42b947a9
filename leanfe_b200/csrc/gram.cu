// gram.cu -- tall-skinny fp64 Gram reductions and the small dense solve.
//
// Z'diag(omega)Z for Z = [1 | cols] (N x P, P <= ~90) is HBM-bound: (P+1)P/2 FMAs against 8(P-1)
// bytes per row, so it runs on the fp64 FMA pipe with register-resident upper-triangle accumulators
// (P <= 12) or a shared-memory row tile (larger P); no tensor cores.  omega_i = w_i * r_i^2 with both
// factors optional, which makes the same kernels serve X'X / X'y (polars_impl.py:201-209), the HC1
// meat sum x x' r^2 (std_errors.py:239-273) and S'S of the cluster scores (std_errors.py:336).
// Block partials are reduced in a fixed order by a second kernel, so results are run-to-run
// deterministic.
#include "common.cuh"
#include "tail.cuh"

void comm_allreduce(lfe_ctx* h, double* buf, int64_t count, int op_max, cudaStream_t s);  // api.cu

constexpr int GR_T = 256;
constexpr int GR_MAXP_TRI = 12;
constexpr int GR_MAXCOLS = 96;

struct GramCols {
  const double* col[GR_MAXP_TRI];
};

template <int P, bool INTERCEPT>
__global__ void __launch_bounds__(GR_T) k_gram_tri(GramCols cols, const double* __restrict__ w,
                                                   const double* __restrict__ r, int64_t n, int64_t ld,
                                                   double* __restrict__ partials) {
  constexpr int T = P * (P + 1) / 2;
  constexpr int NCOL = INTERCEPT ? P - 1 : P;
  double acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double z[P];
    if (INTERCEPT) z[0] = 1.0;
#pragma unroll
    for (int c = 0; c < NCOL; ++c) z[c + (INTERCEPT ? 1 : 0)] = __ldcs(cols.col[c] + i * ld);
    double om = 1.0;
    if (w) om = __ldcs(w + i);
    if (r) {
      double rr = __ldcs(r + i);
      om *= rr * rr;
    }
    int t = 0;
#pragma unroll
    for (int a = 0; a < P; ++a) {
      double za = (w || r) ? z[a] * om : z[a];
#pragma unroll
      for (int b = a; b < P; ++b) {
        acc[t] = fma(za, z[b], acc[t]);
        ++t;
      }
    }
  }
  __shared__ double sh[GR_T / 32][T];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    double v = warp_sum(acc[t]);
    if (lane == 0) sh[warp][t] = v;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += GR_T) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < GR_T / 32; ++q) v += sh[q][t];
    partials[(size_t)blockIdx.x * T + t] = v;
  }
}

// generic: row tile in shared memory, pairs (a,b) distributed over threads
constexpr int GG_ROWS = 64;
constexpr int GG_PAIRS = 20;  // pairs per thread -> T <= 5120 (P <= 100)

__global__ void __launch_bounds__(GR_T) k_gram_generic(const double* const* __restrict__ cols, int ncol,
                                                       int intercept, const double* __restrict__ w,
                                                       const double* __restrict__ r, int64_t n, int64_t stride,
                                                       double* __restrict__ partials) {
  extern __shared__ double tile[];  // [GG_ROWS][P+1] ; last column = omega
  const int P = ncol + intercept;
  const int T = P * (P + 1) / 2;
  const int ld = P + 1;
  double acc[GG_PAIRS];
  int pa[GG_PAIRS], pb[GG_PAIRS];
#pragma unroll
  for (int q = 0; q < GG_PAIRS; ++q) {
    acc[q] = 0.0;
    int t = threadIdx.x + q * GR_T;
    // invert t -> (a,b), a <= b, row-major upper triangle
    int a = 0, rem = t;
    if (t < T) {
      while (rem >= P - a) { rem -= P - a; ++a; }
      pa[q] = a; pb[q] = a + rem;
    } else {
      pa[q] = -1; pb[q] = 0;
    }
  }
  const int64_t ntile = ceil_div64(n, GG_ROWS);
  for (int64_t tl = blockIdx.x; tl < ntile; tl += gridDim.x) {
    const int64_t r0 = tl * GG_ROWS;
    __syncthreads();
    for (int idx = threadIdx.x; idx < GG_ROWS * ld; idx += GR_T) {
      int c = idx / GG_ROWS, rr = idx % GG_ROWS;  // column-major fill -> coalesced global reads
      int64_t i = r0 + rr;
      double v = 0.0;
      if (i < n) {
        if (c == P) {
          v = w ? w[i] : 1.0;
          if (r) v *= r[i] * r[i];
        } else if (intercept && c == 0) {
          v = 1.0;
        } else {
          v = cols[c - intercept][i * stride];
        }
      }
      tile[rr * ld + c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < GG_PAIRS; ++q) {
      if (pa[q] >= 0) {
        double s = 0.0;
        for (int rr = 0; rr < GG_ROWS; ++rr) {
          const double* row = tile + rr * ld;
          s = fma(row[pa[q]] * row[P], row[pb[q]], s);
        }
        acc[q] += s;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < GG_PAIRS; ++q) {
    int t = threadIdx.x + q * GR_T;
    if (t < T) partials[(size_t)blockIdx.x * T + t] = acc[q];
  }
}

// out[a][b] = out[b][a] = sum_blocks partials[block][t(a,b)]
__global__ void __launch_bounds__(256) k_gram_final(const double* __restrict__ partials, int nblocks, int P,
                                                    double* __restrict__ out) {
  const int T = P * (P + 1) / 2;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * T + t];
    int a = 0, rem = t;
    while (rem >= P - a) { rem -= P - a; ++a; }
    int bb = a + rem;
    out[a * P + bb] = s;
    out[bb * P + a] = s;
  }
}

// Z' diag(w * r^2) Z ; cols are host pointers to device columns
void gram_core(lfe_ctx* h, const double* const* cols, int C, int64_t n, const double* w, const double* r,
               int add_intercept, double* gram_out, cudaStream_t s, int64_t stride) {
  const int P = C + (add_intercept ? 1 : 0);
  LFE_REQUIRE(P >= 1 && C <= GR_MAXCOLS, "lfe_gram: %d columns unsupported (max %d)", C, GR_MAXCOLS);
  const int T = P * (P + 1) / 2;
  if (n <= 0) {
    CUDA_CHECK(cudaMemsetAsync(gram_out, 0, (size_t)P * P * sizeof(double), s));
    return;
  }
  if (P <= GR_MAXP_TRI) {
    int per_sm = P <= 8 ? 2 : 1;
    int blocks = launch_blocks(h, ceil_div64(n, GR_T * 2), per_sm);
    DevBuf<double> partials((size_t)blocks * T, s);
    GramCols gc;
    for (int c = 0; c < C; ++c) gc.col[c] = cols[c];
#define TRI_CASE(P_)                                                                                   \
  case P_:                                                                                             \
    if (add_intercept) k_gram_tri<P_, true><<<blocks, GR_T, 0, s>>>(gc, w, r, n, stride, partials);           \
    else k_gram_tri<P_, false><<<blocks, GR_T, 0, s>>>(gc, w, r, n, stride, partials);                        \
    break;
    switch (P) {
      TRI_CASE(1) TRI_CASE(2) TRI_CASE(3) TRI_CASE(4) TRI_CASE(5) TRI_CASE(6) TRI_CASE(7) TRI_CASE(8) TRI_CASE(9)
      TRI_CASE(10) TRI_CASE(11) TRI_CASE(12)
    }
#undef TRI_CASE
    KERNEL_CHECK();
    k_gram_final<<<1, 256, 0, s>>>(partials, blocks, P, gram_out);
    KERNEL_CHECK();
    h->n_launches += 2;
  } else {
    LFE_REQUIRE(T <= GG_PAIRS * GR_T, "lfe_gram: P=%d too large", P);
    int blocks = launch_blocks(h, ceil_div64(n, GG_ROWS), 1);
    DevBuf<double> partials((size_t)blocks * T, s);
    DevBuf<const double*> dcols(C, s);
    CUDA_CHECK(cudaMemcpyAsync(dcols.p, cols, C * sizeof(double*), cudaMemcpyHostToDevice, s));
    size_t smem = (size_t)GG_ROWS * (P + 1) * sizeof(double);
    CUDA_CHECK(cudaFuncSetAttribute(k_gram_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gram_generic<<<blocks, GR_T, smem, s>>>(dcols, C, add_intercept ? 1 : 0, w, r, n, stride, partials);
    KERNEL_CHECK();
    k_gram_final<<<1, 256, 0, s>>>(partials, blocks, P, gram_out);
    KERNEL_CHECK();
    h->n_launches += 2;
    CUDA_CHECK(cudaStreamSynchronize(s));  // dcols staged from a host array owned by the caller
  }
}

void gram_impl(lfe_ctx* h, const double* const* cols, int C, int64_t n, const double* weights, int add_intercept,
               double* gram_out, cudaStream_t s) {
  PhaseTimer t(h, s, 9);
  gram_core(h, cols, C, n, weights, nullptr, add_intercept, gram_out, s, 1);
  if (h->world > 1) {
    int P = C + (add_intercept ? 1 : 0);
    comm_allreduce(h, gram_out, (int64_t)P * P, 0, s);
  }
}

// ---------------------------------------------------------------------------------------------
// Gram of the DEMEANED columns straight from (raw columns, FE tables): Z = [1 | x~_1..x~_k | y~] with
// x~_i = x_i - sum_j alpha_j[code_j(i)] formed in registers (tail.cuh).  Replaces k_apply4 (write N x C) +
// k_gram_tri (read N x C): N*(8C + 4F) bytes instead of N*(24C + 4F).
// ---------------------------------------------------------------------------------------------
template <int NC, int R, bool VEC, bool WEIGHTED>
__global__ void __launch_bounds__(TL_T, (NC <= 6 ? 2 : 1)) k_tail_gram(TailCols cols, TailTabs T, const double* __restrict__ w, int64_t n,
                                                    int ldc, double* __restrict__ partials) {
  constexpr int P = NC + 1;
  constexpr int NT = P * (P + 1) / 2;
  extern __shared__ double sm_tabs[];
  const double* tabs[LFE_MAX_FE];
  stage_tables(T, ldc, sm_tabs, tabs);
  const bool vec2 = (ldc & 1) == 0;
  double acc[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) acc[t] = 0.0;
  const int64_t ngroups = VEC ? n / R : ceil_div64(n, R);
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += (int64_t)gridDim.x * blockDim.x) {
    double x[R][NC];
    bool valid[R];
    tail_rows<NC, R, VEC>(cols, T, tabs, ldc, vec2, g * R, n, x, valid);
    double om[R];
#pragma unroll
    for (int r = 0; r < R; ++r) om[r] = WEIGHTED ? (valid[r] ? __ldcs(w + g * R + r) : 0.0) : (valid[r] ? 1.0 : 0.0);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double z[P];
      z[0] = 1.0;
#pragma unroll
      for (int c = 1; c < NC; ++c) z[c] = x[r][c];
      z[P - 1] = x[r][0];
      int t = 0;
#pragma unroll
      for (int a = 0; a < P; ++a) {
        const double za = z[a] * om[r];
#pragma unroll
        for (int b = a; b < P; ++b) {
          acc[t] = fma(za, z[b], acc[t]);
          ++t;
        }
      }
    }
  }
  if (VEC && blockIdx.x == 0 && threadIdx.x < (int)(n % R)) {   // the last n % R rows
    double x[1][NC];
    bool valid[1];
    const int64_t i = (n / R) * R + threadIdx.x;
    tail_rows<NC, 1, false>(cols, T, tabs, ldc, vec2, i, n, x, valid);
    const double om = WEIGHTED ? w[i] : 1.0;
    double z[P];
    z[0] = 1.0;
#pragma unroll
    for (int c = 1; c < NC; ++c) z[c] = x[0][c];
    z[P - 1] = x[0][0];
    int t = 0;
#pragma unroll
    for (int a = 0; a < P; ++a) {
#pragma unroll
      for (int b = a; b < P; ++b) {
        acc[t] = fma(z[a] * om, z[b], acc[t]);
        ++t;
      }
    }
  }
  __shared__ double sh[TL_T / 32][NT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const double v = warp_sum(acc[t]);
    if (lane == 0) sh[warp][t] = v;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < NT; t += TL_T) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < TL_T / 32; ++q) v += sh[q][t];
    partials[(size_t)blockIdx.x * NT + t] = v;
  }
}

template <int NC, bool WEIGHTED>
static void launch_tail_gram(lfe_ctx* h, const TailCols& cols, const TailTabs& T, const double* w, int64_t n, int ldc,
                             size_t smem, bool vec, double* gram_out, cudaStream_t s) {
  constexpr int R = TailCfg<NC>::R_GRAM;
  constexpr int P = NC + 1, NT = P * (P + 1) / 2;
  const int blocks = launch_blocks(h, ceil_div64(n, (int64_t)TL_T * R), NC <= 8 ? 2 : 1);
  DevBuf<double> partials((size_t)blocks * NT, s);
  // staged tables (<= 48 KB dynamic) + the static reduction scratch can exceed the 48 KB default carve-out
  if (vec && R > 1) {
    if (smem > 32 * 1024)
      CUDA_CHECK(cudaFuncSetAttribute(k_tail_gram<NC, R, true, WEIGHTED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tail_gram<NC, R, true, WEIGHTED><<<blocks, TL_T, smem, s>>>(cols, T, w, n, ldc, partials);
  } else {
    if (smem > 32 * 1024)
      CUDA_CHECK(cudaFuncSetAttribute(k_tail_gram<NC, 1, false, WEIGHTED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tail_gram<NC, 1, false, WEIGHTED><<<blocks, TL_T, smem, s>>>(cols, T, w, n, ldc, partials);
  }
  KERNEL_CHECK();
  k_gram_final<<<1, 256, 0, s>>>(partials, blocks, P, gram_out);
  KERNEL_CHECK();
  h->n_launches += 2;
}

// gram_out (P x P, P = C + 1, order [1, x_1..x_k, y]) = Z' diag(w) Z of the demeaned columns; cols[0] = y (raw)
void gram_fused_impl(lfe_ctx* h, const double* const* cols, int C, int64_t n, const double* weights, const TailTabs& T,
                     int ldc, size_t smem, double* gram_out, cudaStream_t s) {
  PhaseTimer t(h, s, 9);
  LFE_REQUIRE(C >= 1 && C <= TL_MAXC, "gram_fused: C=%d out of range", C);
  const int P = C + 1;
  if (n <= 0) {
    CUDA_CHECK(cudaMemsetAsync(gram_out, 0, (size_t)P * P * sizeof(double), s));
  } else {
    TailCols tc;
    for (int c = 0; c < C; ++c) tc.col[c] = cols[c];
    const bool vec = tail_aligned(tc, C, T, weights);
#define TG_CASE(NC_)                                                                                     \
  case NC_:                                                                                              \
    if (weights) launch_tail_gram<NC_, true>(h, tc, T, weights, n, ldc, smem, vec, gram_out, s);         \
    else launch_tail_gram<NC_, false>(h, tc, T, nullptr, n, ldc, smem, vec, gram_out, s);                \
    break;
    switch (C) {
      TG_CASE(1) TG_CASE(2) TG_CASE(3) TG_CASE(4) TG_CASE(5) TG_CASE(6) TG_CASE(7) TG_CASE(8) TG_CASE(9) TG_CASE(10)
      TG_CASE(11)
    }
#undef TG_CASE
  }
  if (h->world > 1) comm_allreduce(h, gram_out, (int64_t)P * P, 0, s);
}

// ---------------------------------------------------------------------------------------------
// small SPD solve: Cholesky, beta, inverse; LU fallback  (polars_impl.py:212-220)
// ---------------------------------------------------------------------------------------------
// One CTA.  L (p x p, row-major, lower) lives in global scratch (L2-resident, p <= 512).
__global__ void __launch_bounds__(256) k_chol_solve(const double* __restrict__ A, const double* __restrict__ b, int p,
                                                    double* __restrict__ L, double* __restrict__ beta,
                                                    double* __restrict__ Ainv, int* __restrict__ status) {
  __shared__ int fail;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) fail = 0;
  for (int i = tid; i < p * p; i += nt) L[i] = A[i];
  __syncthreads();
  for (int j = 0; j < p; ++j) {
    double d = L[j * p + j];
    if (!(d > 0.0)) {
      if (tid == 0) fail = 1;
    }
    __syncthreads();
    if (fail) break;
    double sd = sqrt(d);
    __syncthreads();
    for (int i = j + tid; i < p; i += nt) L[i * p + j] = (i == j) ? sd : L[i * p + j] / sd;
    __syncthreads();
    // trailing update, lower triangle only
    const int m = p - j - 1;
    for (int idx = tid; idx < m * m; idx += nt) {
      int ii = j + 1 + idx / m, kk = j + 1 + idx % m;
      if (kk <= ii) L[ii * p + kk] -= L[ii * p + j] * L[kk * p + j];
    }
    __syncthreads();
  }
  if (fail) {
    if (tid == 0) *status = 1;
    return;
  }
  // right-hand sides: column q < p of the identity -> Ainv[:, q]; q == p -> b -> beta
  for (int q = tid; q <= p; q += nt) {
    double* x = (q < p) ? (Ainv + q) : beta;  // Ainv column q has stride p
    const int xs = (q < p) ? p : 1;
    // forward: L z = rhs
    for (int i = 0; i < p; ++i) {
      double v = (q < p) ? (i == q ? 1.0 : 0.0) : b[i];
      for (int k = 0; k < i; ++k) v -= L[i * p + k] * x[k * xs];
      x[i * xs] = v / L[i * p + i];
    }
    // backward: L' x = z
    for (int i = p - 1; i >= 0; --i) {
      double v = x[i * xs];
      for (int k = i + 1; k < p; ++k) v -= L[k * p + i] * x[k * xs];
      x[i * xs] = v / L[i * p + i];
    }
  }
  if (tid == 0) *status = 0;
}

// Gauss-Jordan with partial pivoting on [A | I | b]  (np.linalg.solve / inv fallback)
__global__ void __launch_bounds__(256) k_lu_solve(const double* __restrict__ A, const double* __restrict__ b, int p,
                                                  double* __restrict__ M, double* __restrict__ beta,
                                                  double* __restrict__ Ainv, int* __restrict__ status) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int W = 2 * p + 1;
  __shared__ int piv;
  __shared__ int singular;
  if (tid == 0) singular = 0;
  for (int idx = tid; idx < p * W; idx += nt) {
    int i = idx / W, c = idx % W;
    M[idx] = c < p ? A[i * p + c] : (c < 2 * p ? (c - p == i ? 1.0 : 0.0) : b[i]);
  }
  __syncthreads();
  for (int j = 0; j < p; ++j) {
    if (tid == 0) {
      int best = j;
      double bv = fabs(M[j * W + j]);
      for (int i = j + 1; i < p; ++i) {
        double v = fabs(M[i * W + j]);
        if (v > bv) { bv = v; best = i; }
      }
      piv = best;
      if (bv == 0.0) singular = 1;
    }
    __syncthreads();
    if (singular) break;
    if (piv != j)
      for (int c = tid; c < W; c += nt) {
        double t = M[j * W + c];
        M[j * W + c] = M[piv * W + c];
        M[piv * W + c] = t;
      }
    __syncthreads();
    double d = M[j * W + j];
    __syncthreads();
    for (int c = tid; c < W; c += nt) M[j * W + c] /= d;
    __syncthreads();
    for (int idx = tid; idx < p * W; idx += nt) {
      int i = idx / W, c = idx % W;
      if (i != j && c != j) M[idx] -= M[i * W + j] * M[j * W + c];
    }
    __syncthreads();
    for (int i = tid; i < p; i += nt)
      if (i != j) M[i * W + j] = 0.0;
    __syncthreads();
  }
  if (singular) {
    if (tid == 0) *status = 3;
    return;
  }
  for (int idx = tid; idx < p * p; idx += nt) Ainv[idx] = M[(idx / p) * W + p + idx % p];
  for (int i = tid; i < p; i += nt) beta[i] = M[i * W + 2 * p];
  if (tid == 0) *status = 2;
}

void solve_spd_impl(lfe_ctx* h, const double* A, const double* b, int p, double* beta_out, double* Ainv_out,
                    int* used_fallback_out, cudaStream_t s) {
  PhaseTimer t(h, s, 10);
  LFE_REQUIRE(p >= 1 && p <= 512, "lfe_solve_spd: p=%d out of range [1,512]", p);
  DevBuf<double> L((size_t)p * (2 * p + 1), s);
  DevBuf<int> status(1, s);
  status.zero();
  k_chol_solve<<<1, 256, 0, s>>>(A, b, p, L, beta_out, Ainv_out, status);
  KERNEL_CHECK();
  h->n_launches += 1;
  int hs = 0;
  CUDA_CHECK(cudaMemcpyAsync(&hs, status.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  if (hs == 1) {
    k_lu_solve<<<1, 256, 0, s>>>(A, b, p, L, beta_out, Ainv_out, status);
    KERNEL_CHECK();
    h->n_launches += 1;
    CUDA_CHECK(cudaMemcpyAsync(&hs, status.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (hs == 3) LFE_THROW(LFE_ERR_ARG, "lfe_solve_spd: singular matrix (np.linalg.LinAlgError in the reference)");
  }
  if (used_fallback_out) *used_fallback_out = (hs == 2) ? 1 : 0;
}
