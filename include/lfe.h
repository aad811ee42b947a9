/*
 * lfe.h -- C ABI of liblfe.so: B200 (sm_100a) kernels for the leanfe estimation hot path.
 *
 * This is the drop-in boundary.  The reference (jorgenhost/leanfe) has no FFI of its own: its
 * hot path is Polars expressions + NumPy calls inside python/leanfe/polars_impl.py,
 * python/leanfe/std_errors.py and python/leanfe/compress.py.  Each entry point below replaces the
 * reference code cited beside it (paths relative to the reference repo root) and is what a
 * maintainer would bind with ctypes from those call sites.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (LFE_ERR_*); lfe_last_error() gives the
 *     thread-local message of the last failing call.
 *   - array arguments are DEVICE pointers owned by the caller unless marked [host]; the library
 *     never frees or keeps caller memory.  Scratch is stream-ordered (cudaMallocAsync on `stream`)
 *     and released before the call returns.
 *   - columns are struct-of-arrays: one contiguous double array per column, one int32 array per
 *     group-code column.  Row counts must be < 2^31.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls are
 *     asynchronous except where a [host] scalar is returned (those synchronise `stream`).
 *   - a "pointer array" argument (e.g. double* const* cols) is a HOST array of device pointers.
 */
#ifndef LFE_H
#define LFE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lfe_ctx* lfe_handle;

enum {
  LFE_OK = 0,
  LFE_ERR_ARG = -1,       /* bad argument */
  LFE_ERR_CUDA = -2,      /* CUDA runtime error */
  LFE_ERR_CAPACITY = -3,  /* caller buffer too small (e.g. uniques_cap) */
  LFE_ERR_UNSUPPORTED = -4,
  LFE_ERR_NCCL = -5
};

/* key dtypes for lfe_factorize / lfe_compress */
enum { LFE_I32 = 0, LFE_I64 = 1, LFE_U32 = 2, LFE_U64 = 3, LFE_F64 = 4, LFE_U8 = 5 };
/* code order: 0 = rank among ascending sorted uniques (== np.unique(..., return_inverse=True)),
 *             1 = first-appearance rank (== pandas.factorize)                                    */
enum { LFE_ORDER_SORTED = 0, LFE_ORDER_FIRST = 1 };
/* OR-ed into `order`: with a communicator attached, every rank gets codes from ONE shared code space
 * (keys need not be present on every rank); sorted-rank order only, key range < 2^28.             */
#define LFE_ORDER_GLOBAL 0x100
/* vcov kinds (python/leanfe/std_errors.py:83-176) */
enum { LFE_VCOV_IID = 0, LFE_VCOV_HC1 = 1, LFE_VCOV_CLUSTER = 2 };
/* OR-ed into vcov_kind (multi-GPU, one-way): the cluster's groups never span ranks (e.g. cluster == shard
 * key), so only the k x k meat is all-reduced instead of the G x k score table                      */
#define LFE_VCOV_CLUSTER_LOCAL 0x100

#define LFE_MAX_FE 8
#define LFE_MAX_CLUSTER_WAYS 4
#define LFE_N_TIMERS 16

int lfe_version(void);
const char* lfe_last_error(void);

/* one handle per device; not thread-safe per handle */
int lfe_create(int device, lfe_handle* out);
int lfe_destroy(lfe_handle h);

/* Tuning knobs of a handle (all optional).
 * LFE_OPT_COLUMNS_HINT: number of columns C the next demean plans will run with (y + regressors); with it,
 *   lfe_demean_plan_create can order the entry lists in L2-sized blocks of the highest-cardinality FE when that FE's
 *   [G x C] table does not fit the L2 (0 = unknown: no blocking).
 * LFE_OPT_L2_BLOCK_BYTES: bytes of one such table slice (default 48 MB; the B200 L2 is 126 MB over two partitions).
 * LFE_OPT_DETERMINISTIC: 1 = run-to-run bit-identical results (fixed-order reductions instead of floating-point
 *   atomics) on the paths that support it, see lfe_demean.                                                  */
enum { LFE_OPT_COLUMNS_HINT = 1, LFE_OPT_L2_BLOCK_BYTES = 2, LFE_OPT_DETERMINISTIC = 3 };
int lfe_set_option(lfe_handle h, int option, int64_t value);

/* Per-phase CUDA-event timers (ms, accumulated since the last reset); enabling costs two
 * cudaEventRecord per phase.  lfe_get_timers synchronises the recorded events.
 * slots: 0 factorize, 1 counts/singletons, 2 compact, 3 sort, 4 gather, 5 demean-setup (CSR + S0),
 *        6 demean sweeps, 7 convergence checks, 8 demean apply, 9 gram, 10 solve, 11 resid/vcov,
 *        12 compress, 13 collectives, 14 repartition.  lfe_get_counters: kernel launches since reset ([0]), sweep
 *        projection launches ([1]), NCCL calls ([2]).                                                             */
int lfe_set_profiling(lfe_handle h, int enabled);
int lfe_reset_timers(lfe_handle h);
int lfe_get_timers(lfe_handle h, double* ms_out /*[host] LFE_N_TIMERS*/);
int lfe_get_counters(lfe_handle h, int64_t* out /*[host] 4*/);

/* ---- (1) group codes -------------------------------------------------------------------------
 * Replaces: `cast(pl.Categorical).to_physical()` polars_impl.py:118-139 (order 1), the implicit
 * hashing of every `.over(fe)` / `group_by` (polars_impl.py:434,479,503; std_errors.py:329,408),
 * `np.unique(..., return_inverse=True)` compress.py:585-590,842 (order 0), and `n_unique`
 * polars_impl.py:371-373.  Float keys are canonicalised (-0.0 -> +0.0, all NaN equal).
 * uniques_out (nullable) receives the distinct keys in code order, dtype = key dtype; if
 * n_groups > uniques_cap the call fails with LFE_ERR_CAPACITY (n_groups_out is still set).       */
int lfe_factorize(lfe_handle h, const void* keys, int key_dtype, int64_t n, int order,
                  int32_t* codes_out, int64_t* n_groups_out /*[host]*/,
                  void* uniques_out, int64_t uniques_cap, void* stream);

/* Group sizes as a by-product of the LAST lfe_factorize call on this handle: when that call took the
 * direct-address path (rank-local sorted-rank codes of keys with a small range -- panel ids) it counted
 * occurrences instead of marking presence, so `len().over(fe)` (polars_impl.py:478-482) costs no extra pass.
 * `codes` must be the codes_out pointer of that call and n_groups its result.  available_out [host] = 1 and
 * counts_out[g] = #rows with code g when the counts exist, else 0 (use lfe_group_counts).  The handle's copy
 * is released by this call.                                                                          */
int lfe_factorize_counts(lfe_handle h, const int32_t* codes, int64_t n_groups, int32_t* counts_out,
                         int* available_out /*[host]*/, void* stream);

/* counts_out[g] = #rows with codes == g  (int32, length n_groups) */
int lfe_group_counts(lfe_handle h, const int32_t* codes, int64_t n, int64_t n_groups,
                     int32_t* counts_out, void* stream);

/* ---- singleton filter -------------------------------------------------------------------------
 * Replaces: `lf.filter(all_horizontal(len().over(fe) > 1))` polars_impl.py:478-482 (434-435).
 * keep_out[i] = 1 iff every FE's group of row i has more than one row; ONE pass on the input.
 * fe_dims_out[j] (nullable) = number of distinct codes of FE j among kept rows (:531-534).
 * n_keep_out[0] = kept rows on this rank, [1] = kept rows over all ranks, [2] = input rows over all
 * ranks (so `[1] < [2]` is the same "something was dropped" decision on every rank).  With a
 * communicator attached, group sizes of every FE except `local_fe` (-1: none) are summed over the
 * ranks before the test and fe_dims are global.                                                  */
int lfe_singleton_mask(lfe_handle h, const int32_t* const* codes, const int64_t* n_groups, int F,
                       int64_t n, int local_fe, const int32_t* const* counts /*nullable: rank-local group
                       sizes from lfe_group_counts, saves one pass per FE*/, uint8_t* keep_out,
                       int64_t* n_keep_out /*[host] 3*/, int64_t* fe_dims_out /*[host] F*/, void* stream);

/* stable compaction of C double columns and K int32 columns by keep[]; outputs have n_keep rows */
int lfe_compact(lfe_handle h, const uint8_t* keep, int64_t n,
                const double* const* f64_in, double* const* f64_out, int C,
                const int32_t* const* i32_in, int32_t* const* i32_out, int K, void* stream);

/* ---- row ordering -----------------------------------------------------------------------------
 * Stable argsort of rows by a dense code column (the highest-cardinality FE: its groups become
 * contiguous segments, and it is the multi-GPU shard key).  perm_out[p] = source row of sorted
 * position p.  already_sorted_out [host] = 1 when the input was already non-decreasing (perm is
 * then the identity and the caller may skip the gathers).                                        */
int lfe_sort_rows(lfe_handle h, const int32_t* codes, int64_t n, int64_t n_groups,
                  int32_t* perm_out, int* already_sorted_out /*[host]*/, void* stream);
int lfe_gather(lfe_handle h, const int32_t* perm, int64_t n,
               const double* const* f64_in, double* const* f64_out, int C,
               const int32_t* const* i32_in, int32_t* const* i32_out, int K, void* stream);

/* ---- (2) FWL alternating projections ----------------------------------------------------------
 * Replaces: the demeaning loop polars_impl.py:490-526 (and the single-FE within transform
 * :438-457 with max_iter = 1, min_iter > 1).  cols[0] is y.  In place: on return every column
 * holds c - sum_j alpha_j[code_j].  fe_order = FE indices in sweep order (ascending pre-drop
 * cardinality, :485).  From sweep `min_iter` (reference: 3) on, after each full sweep
 * max_fe max_g |mean_g(y)| (UNWEIGHTED, :511-521) is compared with tol.  weights nullable
 * (weighted mean = sum(w c)/sum(w), :492-500).  n_groups[j] bounds codes[j]; empty groups are
 * allowed.  With a communicator attached (lfe_comm_init) rows are assumed sharded so that FE
 * `local_fe` (>= 0) has all rows of each of its groups on one rank; all other FE tables are
 * all-reduced once per projection.  max_mean_out = the stopping quantity of the last test; when the fit
 * stops unconverged with F >= 3 it is the first FE's group-mean maximum only (already >= tol, so the other
 * FEs' y-only passes were skipped), i.e. a lower bound of the reference's quantity.               */
#define LFE_DEMEAN_STRICT_CHECK 1 /* flags bit: stop on max |weighted group mean| over ALL columns < tol
                                    (true fixed point; used for the YOCO solve) instead of the reference test */
int lfe_demean(lfe_handle h, double* const* cols, int C, const int32_t* const* codes,
               const int64_t* n_groups, const int* fe_order, int F, int64_t n,
               const double* weights, double tol, int max_iter, int min_iter, int local_fe, int flags,
               int* iterations_out /*[host]*/, double* max_mean_out /*[host]*/,
               int* converged_out /*[host]*/, void* stream);

/* The same operator in two steps.  The plan holds everything that depends on the code columns only (the
 * per-FE sorted entry lists and group sizes), so it can be built while the data columns are still being
 * uploaded, and re-run for several column sets on one panel.  codes / weights must stay valid and
 * unchanged until lfe_demean_plan_destroy.  lfe_demean == create + run + destroy.                    */
typedef struct lfe_plan* lfe_plan_t;
int lfe_demean_plan_create(lfe_handle h, const int32_t* const* codes, const int64_t* n_groups, int F,
                           int64_t n, const double* weights, int local_fe,
                           const int32_t* const* counts /*nullable, as in lfe_singleton_mask*/,
                           lfe_plan_t* plan_out, void* stream);
int lfe_demean_plan_run(lfe_handle h, lfe_plan_t plan, double* const* cols,
                        const double* const* cols_src /*nullable: out of place, read from here, write to cols*/,
                        int C, const int* fe_order,
                        double tol, int max_iter, int min_iter, int flags, int* iterations_out /*[host]*/,
                        double* max_mean_out /*[host]*/, int* converged_out /*[host]*/, void* stream);
/* The whole fit after the plan, in one call and WITHOUT materialising the demeaned columns: sweeps as in
 * lfe_demean_plan_run, then -- x~_i = x_i - sum_j alpha_j[code_j(i)] formed in registers from the raw columns and
 * the converged FE tables -- the [1 | X~] normal equations and their Cholesky solve (polars_impl.py:166-226), the
 * residual statistics (:229-283) and the IID / HC1 / one-way clustered sandwich (std_errors.py:183-347).  Replaces
 * lfe_demean_plan_run + lfe_gram + lfe_solve_spd + lfe_resid_vcov for C = 1 + k <= 11 columns; cols (cols[0] = y)
 * are read-only.  vcov_kind as in lfe_resid_vcov (cl_codes: ONE code column in row order, or NULL).
 * Outputs (device): beta_full_out (k + 1, intercept first), se_out (k), vcov_out (k x k, nullable), stats_out
 * (4: rss, tss, sum w r^2, 0).  Synchronises `stream`.                                                   */
int lfe_demean_plan_fit(lfe_handle h, lfe_plan_t plan, const double* const* cols, int C, const int* fe_order,
                        double tol, int max_iter, int min_iter, int flags, int vcov_kind, const int32_t* cl_codes,
                        int64_t n_cl_groups, int ssc, int64_t n_obs, int64_t df_resid, double* beta_full_out,
                        double* se_out, double* vcov_out, double* stats_out, int* iterations_out /*[host]*/,
                        double* max_mean_out /*[host]*/, int* converged_out /*[host]*/,
                        int64_t* n_clusters_out /*[host], nullable*/, void* stream);
int lfe_demean_plan_destroy(lfe_handle h, lfe_plan_t plan);

/* ---- (4) Gram + solve --------------------------------------------------------------------------
 * Replaces: X = [1 | X~], X'X, X'y  polars_impl.py:166-209 (weighted: sqrt(w) scaling :201-206).
 * Z = [1 | cols[0] .. cols[C-1]] when add_intercept else [cols]; gram_out (device, P x P row-major,
 * P = C + add_intercept) = Z' diag(w) Z.  All-reduced over the communicator if attached.         */
int lfe_gram(lfe_handle h, const double* const* cols, int C, int64_t n, const double* weights,
             int add_intercept, double* gram_out, void* stream);

/* Replaces: np.linalg.cholesky + two triangular solves for beta and (X'X)^-1, with the
 * LinAlgError -> np.linalg.solve / inv fallback, polars_impl.py:212-220.  A (p x p), b (p),
 * beta_out (p), Ainv_out (p x p) are device pointers; p <= 512.                                  */
int lfe_solve_spd(lfe_handle h, const double* A, const double* b, int p, double* beta_out,
                  double* Ainv_out, int* used_fallback_out /*[host], nullable*/, void* stream);

/* ---- (5) residuals + sandwich vcov ------------------------------------------------------------
 * Replaces: resid / R^2 polars_impl.py:229, 281-283 and compute_standard_errors_polars
 * std_errors.py:183-441.  cols[0] = y~, cols[1..k] = x~ (demeaned).  beta_full (device, k+1:
 * intercept first), Binv (device k x k = (X'X)^-1[1:,1:]).  cl_codes: W cluster code columns
 * (W = 0 unless vcov_kind == CLUSTER; W > 1 = Cameron-Gelbach-Miller multiway, :354-441).
 * Outputs (device): se_out (k), vcov_out (k x k, nullable), stats_out (4: rss, tss, sum_w_r2, 0).
 * n_clusters_out [host] (W).  resid_out nullable (device, n).                                    */
int lfe_resid_vcov(lfe_handle h, const double* const* cols, int k, int64_t n,
                   const double* beta_full, const double* Binv, int vcov_kind,
                   const int32_t* const* cl_codes, const int64_t* n_cl_groups, int W,
                   const double* weights, int ssc, int64_t n_obs, int64_t df_resid,
                   double* se_out, double* vcov_out, double* stats_out, double* resid_out,
                   int64_t* n_clusters_out /*[host]*/, void* stream);

/* ---- (3) YOCO compression ---------------------------------------------------------------------
 * Replaces: lf.group_by(x_cols + fe_cols + cluster_cols).agg(len, sum y, sum y^2)
 * compress.py:325-350 and the distinct count of estimate_compression_ratio compress.py:249-250.
 * key_codes: K code columns with cardinalities n_groups[K]: int32 codes from lfe_factorize (order 0), or --
 * key_dtypes[j] == LFE_U8 -- raw uint8 values used as codes (dummies; n_groups[j] = max + 1 from
 * lfe_key_ranges, no factorisation pass).  key_dtypes NULL = all LFE_I32.
 * Groups are emitted in ascending lexicographic order of the code tuples.  Outputs have capacity
 * out_cap rows: key_codes_out[K] (int32), n_out (sum of weights, or count), cnt_out (int32 raw
 * count), sum_y_out, sum_y_sq_out.  y may be NULL (count-distinct only; stats outputs NULL).     */
int lfe_compress(lfe_handle h, const void* const* key_codes, const int* key_dtypes /*nullable*/,
                 const int64_t* n_groups, int K, const double* y, const double* weights, int64_t n, int64_t out_cap,
                 int32_t* const* key_codes_out, double* n_out, int32_t* cnt_out,
                 double* sum_y_out, double* sum_y_sq_out, int64_t* n_compressed_out /*[host]*/,
                 void* stream);

/* [min, max] of K integer key columns (LFE_U8 / LFE_I32 / LFE_I64) in one call; with a communicator attached,
 * of the whole sharded columns.  An empty column reports min = INT64_MAX, max = INT64_MIN.  Serves the
 * key-cardinality bookkeeping of compress.py:232-253, 334-336 for columns that need no dictionary.   */
int lfe_key_ranges(lfe_handle h, const void* const* cols, const int* dtypes, int K, int64_t n,
                   int64_t* min_out /*[host] K*/, int64_t* max_out /*[host] K*/, void* stream);

/* Multi-GPU YOCO: rows are sharded arbitrarily, every rank runs lfe_compress on its own rows
 * (key codes from ONE shared code space: lfe_factorize with LFE_ORDER_GLOBAL), then this call all-gathers the
 * partial tables (n_rows rows on this rank) and re-aggregates them: every rank receives the identical global
 * table, groups in ascending lexicographic order, with n / cnt / sum_y / sum_y_sq summed over ranks.  Replaces
 * nothing in the reference (single process); it is what makes compress.py:325-350 shard.  Without a
 * communicator it merges duplicate keys inside the one table given.                                */
int lfe_compress_merge(lfe_handle h, const int32_t* const* key_codes, const int64_t* n_groups, int K,
                       const double* n_in, const int32_t* cnt_in, const double* sum_y_in,
                       const double* sum_y_sq_in, int64_t n_rows, int64_t out_cap,
                       int32_t* const* key_codes_out, double* n_out, int32_t* cnt_out /*nullable*/,
                       double* sum_y_out, double* sum_y_sq_out, int64_t* n_compressed_out /*[host]*/,
                       void* stream);

/* ---- YOCO standard errors --------------------------------------------------------------------
 * Replaces: compute_rss_grouped compress.py:754-810, e0_g compress.py:1124-1137 and compute_se_compress
 * compress.py:854-1042 -- without the FE dummy matrix of compress.py:503-619: cols[0] = W-weighted
 * FE-demeaned group means y~, cols[1..k] = demeaned regressors x~ (compressed rows, G of them);
 * beta (device, k) and Binv (device, k x k) = (X~'WX~)^-1; nn / sum_y / sum_y_sq = the sufficient
 * statistics of lfe_compress.  cl_codes index compressed rows.  rss_out (device scalar, nullable). */
int lfe_yoco_vcov(lfe_handle h, const double* const* cols, int k, int64_t n_compressed,
                  const double* beta, const double* Binv, const double* nn, const double* sum_y,
                  const double* sum_y_sq, int vcov_kind, const int32_t* const* cl_codes,
                  const int64_t* n_cl_groups, int W, int ssc, int64_t n_obs, int64_t df_resid,
                  double* se_out, double* vcov_out, double* rss_out,
                  int64_t* n_clusters_out /*[host]*/, void* stream);

/* ---- multi-GPU ----------------------------------------------------------------------------------
 * One process per GPU.  Rank 0 calls lfe_comm_unique_id, the 128 bytes are broadcast by the host
 * (torch.distributed), every rank calls lfe_comm_init.  NCCL is dlopen'ed at run time.          */
int lfe_comm_unique_id(void* id_out /*[host] 128 bytes*/);
int lfe_comm_init(lfe_handle h, const void* id /*[host] 128 bytes*/, int rank, int world);
int lfe_comm_destroy(lfe_handle h);
/* active = 0: the calls that follow run rank-locally (no collectives) although a communicator is attached --
 * used for work that is replicated on every rank, e.g. the solve on the merged YOCO table; 1 = back on. */
int lfe_comm_set_active(lfe_handle h, int active);
/* Row repartition: rows arrive sharded ARBITRARILY; afterwards all rows of every group of the shard key
 * (the highest-cardinality FE) sit on one rank, which is what lfe_singleton_mask / lfe_demean_plan_create /
 * lfe_resid_vcov expect of `local_fe`.  Replaces nothing in the reference (single process): it is what lets the
 * filter and FE order of polars_impl.py:478-485 be evaluated on the GLOBAL frame when the frame is spread over GPUs.
 * lfe_shard_plan: `codes` = this rank's shard-key codes from ONE shared code space (lfe_factorize with
 * LFE_ORDER_GLOBAL), n_groups its size.  The per-rank group sizes are all-reduced and the code axis is cut into
 * `world` contiguous ranges balanced by row count: rank r owns codes [cuts_out[r], cuts_out[r+1]).  order_out
 * (device, n) = this rank's row indices grouped by destination rank, original order kept inside each destination;
 * send_counts_out[r] / recv_counts_out[r] = rows this rank sends to / receives from rank r.  Synchronises.
 * lfe_shard_exchange: moves K columns (dtypes LFE_F64 / LFE_I32 / LFE_I64 / LFE_U8; cols_in n rows, cols_out
 * sum(recv_counts) rows, caller-allocated) with one grouped ncclSend/ncclRecv round per column.  i32_sub (nullable,
 * [host] K): value subtracted from an LFE_I32 column on receipt (cuts[rank] re-bases the shard key's codes to the
 * rank-local range [0, cuts[rank+1] - cuts[rank])).  Without a communicator: order = identity, exchange = copy. */
int lfe_shard_plan(lfe_handle h, const int32_t* codes, int64_t n, int64_t n_groups, int64_t* cuts_out /*[host] world+1*/,
                   int64_t* send_counts_out /*[host] world*/, int64_t* recv_counts_out /*[host] world*/,
                   int32_t* order_out, void* stream);
int lfe_shard_exchange(lfe_handle h, const int32_t* order, const int64_t* send_counts /*[host]*/,
                       const int64_t* recv_counts /*[host]*/, const void* const* cols_in, const int* dtypes,
                       const int64_t* i32_sub /*[host], nullable*/, int K, int64_t n, void* const* cols_out, void* stream);
/* ---- factor / interaction expansion -----------------------------------------------------------------------
 * Replaces: _expand_factors_polars and _expand_interactions_polars, polars_impl.py:27-115 (one `when(col == cat)`
 * expression per level).  out_cols[j][i] = (codes[i] == levels[j]) * (base ? base[i] : 1): all J dummy columns of
 * `i(var)` (base NULL) or `var:i(factor)` (base = var) in ONE pass over the codes.  levels [host] J = the codes of
 * the non-reference categories; out_dtype LFE_F64, or LFE_U8 for base-less dummies (YOCO groups on raw bytes). */
int lfe_expand_dummies(lfe_handle h, const int32_t* codes, int64_t n, const int32_t* levels /*[host]*/, int J,
                       const double* base /*nullable*/, int out_dtype, void* const* out_cols, void* stream);

/* ---- host ingest: pageable memory -> device -------------------------------------------------
 * The reference's API receives DataFrames / NumPy columns (polars_impl.py:342-347): pageable host memory, which a
 * plain cudaMemcpyAsync stages synchronously on one driver thread.  The uploader owns a ring of `n_slots` pinned
 * chunks of `chunk_bytes`; `n_threads` host threads memcpy the pieces of a chunk into its slot and each chunk leaves
 * with one cudaMemcpyAsync on `copy_stream`.  enqueue returns at once (host_src must stay valid and unchanged until
 * the job's copy completed: lfe_uploader_stream_wait + stream completion, or lfe_uploader_drain); columns are
 * transferred in enqueue order.  lfe_uploader_stream_wait blocks the calling HOST thread until the job's last chunk
 * has been issued and then makes `compute_stream` wait for it on the device.  lfe_uploader_drain waits for every
 * job and forgets the tickets.                                                                          */
typedef struct lfe_uploader* lfe_uploader_t;
int lfe_uploader_create(lfe_handle h, int64_t chunk_bytes, int n_slots, int n_threads, lfe_uploader_t* out);
int lfe_uploader_destroy(lfe_uploader_t u);
int lfe_uploader_enqueue(lfe_uploader_t u, const void* host_src /*[host]*/, int64_t nbytes, void* dev_dst,
                         void* copy_stream, int64_t* ticket_out /*[host]*/);
int lfe_uploader_stream_wait(lfe_uploader_t u, int64_t ticket, void* compute_stream);
int lfe_uploader_drain(lfe_uploader_t u);

/* sum all-reduce of a device double buffer over the attached communicator (no-op if none) */
int lfe_allreduce_f64(lfe_handle h, double* buf, int64_t count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LFE_H */
