/*
 * ref_port.c -- multi-threaded C (OpenMP) restatement of the reference's alt_proj hot path.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).  It exists so that bench.py can time
 * the reference ALGORITHM on all host cores (`cpu_baseline`, `--impl reference`): Polars 1.36.1, which
 * executes that algorithm in the reference, is not installable here, and a single-threaded NumPy
 * restatement would flatter the GPU.  It is checked against oracle/ref_numpy.py in the CPU tests.
 *
 * Follows, step for step (paths relative to /root/reference):
 *   group codes by sorted rank            python/leanfe/compress.py:842 (np.unique inverse)
 *   singleton filter, one pass            python/leanfe/polars_impl.py:478-482
 *   FE order = ascending pre-drop card.   python/leanfe/polars_impl.py:485
 *   in-place Gauss-Seidel sweeps          python/leanfe/polars_impl.py:490-508
 *   y-only unweighted check from it >= 3  python/leanfe/polars_impl.py:511-525
 *   [1|X] normal equations, Cholesky      python/leanfe/polars_impl.py:166-226
 *   IID / HC1 / one-way cluster SEs       python/leanfe/std_errors.py:183-347
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXF 8
#define MAXC 64

static int64_t factorize_dense(const int64_t* keys, int64_t n, int32_t* codes) {
  /* sorted-rank codes through a direct-address table (keys of the benchmark panels are small ints) */
  int64_t lo = keys[0], hi = keys[0];
#pragma omp parallel for reduction(min : lo) reduction(max : hi)
  for (int64_t i = 0; i < n; ++i) {
    if (keys[i] < lo) lo = keys[i];
    if (keys[i] > hi) hi = keys[i];
  }
  int64_t range = hi - lo + 1;
  if (range > ((int64_t)1 << 30)) return -1;
  int32_t* rank = (int32_t*)calloc((size_t)range, sizeof(int32_t));
#pragma omp parallel for
  for (int64_t i = 0; i < n; ++i) rank[keys[i] - lo] = 1;
  int64_t G = 0;
  for (int64_t r = 0; r < range; ++r) {
    int32_t f = rank[r];
    rank[r] = (int32_t)G;
    G += f;
  }
#pragma omp parallel for
  for (int64_t i = 0; i < n; ++i) codes[i] = rank[keys[i] - lo];
  free(rank);
  return G;
}

/* c -= mean_g(c) for every column, one FE.  Bins are private per thread when they fit, else atomics. */
static void project(double** cols, int C, const int32_t* code, int64_t G, const double* cnt, int64_t n) {
  const int T = omp_get_max_threads();
  const int priv = ((double)G * C * 8.0 * T) <= 2.0e9;
  double* sums = (double*)calloc((size_t)G * C, sizeof(double));
  if (priv) {
    double* all = (double*)calloc((size_t)G * C * T, sizeof(double));
#pragma omp parallel
    {
      double* mine = all + (size_t)omp_get_thread_num() * G * C;
#pragma omp for schedule(static)
      for (int64_t i = 0; i < n; ++i) {
        double* b = mine + (size_t)code[i] * C;
        for (int c = 0; c < C; ++c) b[c] += cols[c][i];
      }
#pragma omp for schedule(static)
      for (int64_t q = 0; q < G * C; ++q) {
        double s = 0;
        for (int t = 0; t < T; ++t) s += all[(size_t)t * G * C + q];
        sums[q] = s;
      }
    }
    free(all);
  } else {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      double* b = sums + (size_t)code[i] * C;
      for (int c = 0; c < C; ++c) {
#pragma omp atomic
        b[c] += cols[c][i];
      }
    }
  }
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < G * C; ++q) sums[q] = cnt[q / C] > 0 ? sums[q] / cnt[q / C] : 0.0;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const double* b = sums + (size_t)code[i] * C;
    for (int c = 0; c < C; ++c) cols[c][i] -= b[c];
  }
  free(sums);
}

static double max_abs_group_mean(const double* y, const int32_t* code, int64_t G, const double* cnt, int64_t n) {
  const int T = omp_get_max_threads();
  double* all = (double*)calloc((size_t)G * T, sizeof(double));
  double m = 0.0;
#pragma omp parallel
  {
    double* mine = all + (size_t)omp_get_thread_num() * G;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n; ++i) mine[code[i]] += y[i];
#pragma omp for schedule(static) reduction(max : m)
    for (int64_t g = 0; g < G; ++g) {
      double s = 0;
      for (int t = 0; t < T; ++t) s += all[(size_t)t * G + g];
      if (cnt[g] > 0) {
        double a = fabs(s / cnt[g]);
        if (a > m) m = a;
      }
    }
  }
  free(all);
  return m;
}

static int cholesky_solve(const double* A, const double* b, int p, double* beta, double* Ainv) {
  double* L = (double*)calloc((size_t)p * p, sizeof(double));
  for (int j = 0; j < p; ++j) {
    double d = A[j * p + j];
    for (int k = 0; k < j; ++k) d -= L[j * p + k] * L[j * p + k];
    if (!(d > 0)) { free(L); return 1; }
    L[j * p + j] = sqrt(d);
    for (int i = j + 1; i < p; ++i) {
      double s = A[i * p + j];
      for (int k = 0; k < j; ++k) s -= L[i * p + k] * L[j * p + k];
      L[i * p + j] = s / L[j * p + j];
    }
  }
  for (int q = 0; q <= p; ++q) {
    double x[MAXC + 1];
    for (int i = 0; i < p; ++i) {
      double v = q < p ? (i == q ? 1.0 : 0.0) : b[i];
      for (int k = 0; k < i; ++k) v -= L[i * p + k] * x[k];
      x[i] = v / L[i * p + i];
    }
    for (int i = p - 1; i >= 0; --i) {
      double v = x[i];
      for (int k = i + 1; k < p; ++k) v -= L[k * p + i] * x[k];
      x[i] = v / L[i * p + i];
    }
    for (int i = 0; i < p; ++i) {
      if (q < p) Ainv[i * p + q] = x[i]; else beta[i] = x[i];
    }
  }
  free(L);
  return 0;
}

/*
 * cols_in[0] = y, cols_in[1..k] = x.  fe_keys: F raw int64 id columns.  cluster_keys nullable (vcov 2).
 * vcov: 0 iid, 1 hc1, 2 one-way cluster.  Outputs: beta[k], se[k], info[6] = {iterations, n_obs, df_resid,
 * n_clusters, rss, max_mean}.  Returns 0 on success.
 */
int lfe_port_fit(const double* const* cols_in, int k, const int64_t* const* fe_keys, int F, const int64_t* cluster_keys,
                 int64_t n0, double tol, int max_iter, int vcov, int ssc, double* beta, double* se, double* info) {
  const int C = k + 1;
  if (F < 1 || F > MAXF || C > MAXC) return -1;
  int32_t* code0[MAXF];
  int64_t G[MAXF];
  for (int j = 0; j < F; ++j) {
    code0[j] = (int32_t*)malloc((size_t)n0 * sizeof(int32_t));
    G[j] = factorize_dense(fe_keys[j], n0, code0[j]);
    if (G[j] < 0) return -2;
  }
  int32_t* cl0 = NULL;
  int64_t Gcl = 0;
  if (vcov == 2) {
    cl0 = (int32_t*)malloc((size_t)n0 * sizeof(int32_t));
    Gcl = factorize_dense(cluster_keys, n0, cl0);
    if (Gcl < 0) return -2;
  }
  /* singleton filter */
  uint8_t* keep = (uint8_t*)malloc((size_t)n0);
  memset(keep, 1, (size_t)n0);
  for (int j = 0; j < F; ++j) {
    int32_t* cnt = (int32_t*)calloc((size_t)G[j], sizeof(int32_t));
#pragma omp parallel for
    for (int64_t i = 0; i < n0; ++i) {
#pragma omp atomic
      cnt[code0[j][i]]++;
    }
#pragma omp parallel for
    for (int64_t i = 0; i < n0; ++i)
      if (cnt[code0[j][i]] <= 1) keep[i] = 0;
    free(cnt);
  }
  int64_t n = 0;
  int64_t* pos = (int64_t*)malloc((size_t)n0 * sizeof(int64_t));
  {
    /* exclusive prefix sum of keep[] in two parallel passes over per-thread blocks */
    const int T = omp_get_max_threads();
    int64_t* part = (int64_t*)calloc((size_t)T + 1, sizeof(int64_t));
    const int64_t blk = (n0 + T - 1) / T;
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < T; ++t) {
      int64_t a = t * blk, b = a + blk < n0 ? a + blk : n0, c = 0;
      for (int64_t i = a; i < b; ++i) c += keep[i];
      part[t + 1] = c;
    }
    for (int t = 0; t < T; ++t) part[t + 1] += part[t];
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < T; ++t) {
      int64_t a = t * blk, b = a + blk < n0 ? a + blk : n0, c = part[t];
      for (int64_t i = a; i < b; ++i) { pos[i] = c; c += keep[i]; }
    }
    n = part[T];
    free(part);
  }
  double* cols[MAXC];
  int32_t* code[MAXF];
  for (int c = 0; c < C; ++c) {
    cols[c] = (double*)malloc((size_t)n * sizeof(double));
#pragma omp parallel for
    for (int64_t i = 0; i < n0; ++i)
      if (keep[i]) cols[c][pos[i]] = cols_in[c][i];
  }
  double* cnt[MAXF];
  int64_t fe_dims = 0;
  for (int j = 0; j < F; ++j) {
    code[j] = (int32_t*)malloc((size_t)n * sizeof(int32_t));
#pragma omp parallel for
    for (int64_t i = 0; i < n0; ++i)
      if (keep[i]) code[j][pos[i]] = code0[j][i];
    free(code0[j]);
    cnt[j] = (double*)calloc((size_t)G[j], sizeof(double));
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) {
#pragma omp atomic
      cnt[j][code[j][i]] += 1.0;
    }
    for (int64_t g = 0; g < G[j]; ++g) fe_dims += cnt[j][g] > 0;
  }
  int32_t* cl = NULL;
  if (vcov == 2) {
    cl = (int32_t*)malloc((size_t)n * sizeof(int32_t));
#pragma omp parallel for
    for (int64_t i = 0; i < n0; ++i)
      if (keep[i]) cl[pos[i]] = cl0[i];
    free(cl0);
  }
  free(pos);
  free(keep);
  /* FE order: stable ascending by pre-drop cardinality */
  int order[MAXF];
  for (int j = 0; j < F; ++j) order[j] = j;
  for (int a = 1; a < F; ++a)
    for (int b = a; b > 0 && G[order[b]] < G[order[b - 1]]; --b) { int t = order[b]; order[b] = order[b - 1]; order[b - 1] = t; }
  int iterations = 0;
  double max_mean = 0.0;
  for (int it = 1; it <= max_iter; ++it) {
    for (int o = 0; o < F; ++o) project(cols, C, code[order[o]], G[order[o]], cnt[order[o]], n);
    iterations = it;
    if (it >= 3) {
      max_mean = 0.0;
      for (int j = 0; j < F; ++j) {
        double m = max_abs_group_mean(cols[0], code[j], G[j], cnt[j], n);
        if (m > max_mean) max_mean = m;
      }
      if (max_mean < tol) break;
    }
  }
  /* normal equations on [1 | X] */
  const int p = k + 1;
  double XtX[(MAXC + 1) * (MAXC + 1)], Xty[MAXC + 1], bf[MAXC + 1], Ainv[(MAXC + 1) * (MAXC + 1)];
  memset(XtX, 0, sizeof(XtX));
  memset(Xty, 0, sizeof(Xty));
#pragma omp parallel
  {
    double lx[(MAXC + 1) * (MAXC + 1)], ly[MAXC + 1];
    memset(lx, 0, sizeof(lx));
    memset(ly, 0, sizeof(ly));
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      double z[MAXC + 1];
      z[0] = 1.0;
      for (int c = 1; c < p; ++c) z[c] = cols[c][i];
      for (int a = 0; a < p; ++a) {
        for (int b = a; b < p; ++b) lx[a * p + b] += z[a] * z[b];
        ly[a] += z[a] * cols[0][i];
      }
    }
#pragma omp critical
    {
      for (int a = 0; a < p; ++a) {
        for (int b = a; b < p; ++b) XtX[a * p + b] += lx[a * p + b];
        Xty[a] += ly[a];
      }
    }
  }
  for (int a = 0; a < p; ++a)
    for (int b = 0; b < a; ++b) XtX[a * p + b] = XtX[b * p + a];
  if (cholesky_solve(XtX, Xty, p, bf, Ainv)) return -3;
  const int64_t df = n - p - (fe_dims - F);
  double* resid = (double*)malloc((size_t)n * sizeof(double));
  double rss = 0.0;
#pragma omp parallel for reduction(+ : rss) schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double r = cols[0][i] - bf[0];
    for (int c = 1; c < p; ++c) r -= bf[c] * cols[c][i];
    resid[i] = r;
    rss += r * r;
  }
  double B[MAXC * MAXC], meat[MAXC * MAXC], tmp[MAXC * MAXC];
  for (int a = 0; a < k; ++a)
    for (int b = 0; b < k; ++b) B[a * k + b] = Ainv[(a + 1) * p + (b + 1)];
  int64_t n_clusters = 0;
  if (vcov == 0) {
    for (int a = 0; a < k; ++a) se[a] = sqrt(fmax(rss / (double)df * B[a * k + a], 0.0));
  } else {
    memset(meat, 0, sizeof(meat));
    double adj;
    if (vcov == 1) {
#pragma omp parallel
      {
        double lm[MAXC * MAXC];
        memset(lm, 0, sizeof(lm));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
          double r2 = resid[i] * resid[i];
          for (int a = 0; a < k; ++a)
            for (int b = a; b < k; ++b) lm[a * k + b] += cols[a + 1][i] * cols[b + 1][i] * r2;
        }
#pragma omp critical
        for (int q = 0; q < k * k; ++q) meat[q] += lm[q];
      }
      for (int a = 0; a < k; ++a)
        for (int b = 0; b < a; ++b) meat[a * k + b] = meat[b * k + a];
      adj = (double)n / (double)df;
    } else {
      double* S = (double*)calloc((size_t)Gcl * k, sizeof(double));
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) {
        double* s = S + (size_t)cl[i] * k;
        for (int a = 0; a < k; ++a) {
#pragma omp atomic
          s[a] += cols[a + 1][i] * resid[i];
        }
      }
      uint8_t* seen = (uint8_t*)calloc((size_t)Gcl, 1);
#pragma omp parallel for
      for (int64_t i = 0; i < n; ++i) seen[cl[i]] = 1;
      for (int64_t g = 0; g < Gcl; ++g) {
        n_clusters += seen[g];
        const double* s = S + (size_t)g * k;
        for (int a = 0; a < k; ++a)
          for (int b = 0; b < k; ++b) meat[a * k + b] += s[a] * s[b];
      }
      free(seen);
      free(S);
      adj = (double)n_clusters / (double)(n_clusters - 1);
      if (ssc) adj *= (double)(n - 1) / (double)df;
    }
    for (int a = 0; a < k; ++a)
      for (int b = 0; b < k; ++b) {
        double s = 0;
        for (int q = 0; q < k; ++q) s += B[a * k + q] * meat[q * k + b];
        tmp[a * k + b] = s;
      }
    for (int a = 0; a < k; ++a) {
      double s = 0;
      for (int q = 0; q < k; ++q) s += tmp[a * k + q] * B[q * k + a];
      se[a] = sqrt(fmax(adj * s, 0.0));
    }
  }
  for (int a = 0; a < k; ++a) beta[a] = bf[a + 1];
  info[0] = iterations; info[1] = (double)n; info[2] = (double)df; info[3] = (double)n_clusters; info[4] = rss;
  info[5] = max_mean;
  free(resid);
  if (cl) free(cl);
  for (int c = 0; c < C; ++c) free(cols[c]);
  for (int j = 0; j < F; ++j) { free(code[j]); free(cnt[j]); }
  return 0;
}

int lfe_port_threads(void) { return omp_get_max_threads(); }

/* bench.py sets the thread count explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers */
void lfe_port_set_threads(int t) { if (t > 0) omp_set_num_threads(t); }
