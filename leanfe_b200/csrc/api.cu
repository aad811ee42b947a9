// api.cu -- the extern "C" boundary declared in include/lfe.h: handle, errors, timers, NCCL plumbing.
#include <dlfcn.h>

#include "common.cuh"

// ---- implementations living in the other translation units -------------------------------------
int64_t factorize_dispatch(lfe_ctx* h, const void* keys, int key_dtype, int64_t n, int order, int32_t* codes,
                           void* uniques, int64_t uniques_cap, cudaStream_t s);
void singleton_mask_impl(lfe_ctx* h, const int32_t* const* codes, const int64_t* n_groups, int F, int64_t n,
                         int local_fe, const int32_t* const* counts_in, uint8_t* keep, int64_t* n_keep_out,
                         int64_t* fe_dims_out, cudaStream_t s);
void compact_impl(lfe_ctx* h, const uint8_t* keep, int64_t n, const double* const* f64_in, double* const* f64_out,
                  int C, const int32_t* const* i32_in, int32_t* const* i32_out, int K, cudaStream_t s);
void sort_rows_impl(lfe_ctx* h, const int32_t* codes, int64_t n, int64_t n_groups, int32_t* perm_out,
                    int* already_sorted_out, cudaStream_t s);
void gather_impl(lfe_ctx* h, const int32_t* perm, int64_t n, const double* const* f64_in, double* const* f64_out,
                 int C, const int32_t* const* i32_in, int32_t* const* i32_out, int K, cudaStream_t s);
void demean_impl(lfe_ctx* h, double* const* cols, int C, const int32_t* const* codes, const int64_t* n_groups,
                 const int* fe_order, int F, int64_t n, const double* weights, double tol, int max_iter, int min_iter,
                 int local_fe, int flags, int* iterations_out, double* max_mean_out, int* converged_out,
                 cudaStream_t s);
lfe_plan* demean_plan_create(lfe_ctx* h, const int32_t* const* codes, const int64_t* n_groups, int F, int64_t n,
                             const double* weights, int local_fe, const int32_t* const* counts_in, cudaStream_t s);
void demean_plan_run(lfe_ctx* h, lfe_plan* plan, double* const* cols, const double* const* cols_src, int C,
                     const int* fe_order, double tol, int max_iter, int min_iter, int flags, int* iterations_out,
                     double* max_mean_out, int* converged_out, cudaStream_t s);
void demean_plan_destroy(lfe_plan* plan);
void demean_plan_fit(lfe_ctx* h, lfe_plan* plan, const double* const* cols, int C, const int* fe_order, double tol,
                     int max_iter, int min_iter, int flags, int vcov_kind, const int32_t* cl_codes, int64_t n_cl_groups,
                     int ssc, int64_t n_obs, int64_t df_resid, double* beta_full_out, double* se_out, double* vcov_out,
                     double* stats_out, int* iterations_out, double* max_mean_out, int* converged_out,
                     int64_t* n_clusters_out, cudaStream_t s);
void gram_impl(lfe_ctx* h, const double* const* cols, int C, int64_t n, const double* weights, int add_intercept,
               double* gram_out, cudaStream_t s);
void solve_spd_impl(lfe_ctx* h, const double* A, const double* b, int p, double* beta_out, double* Ainv_out,
                    int* used_fallback_out, cudaStream_t s);
void resid_vcov_impl(lfe_ctx* h, const double* const* cols, int k, int64_t n, const double* beta_full,
                     const double* Binv, int vcov_kind, const int32_t* const* cl_codes, const int64_t* n_cl_groups, int W,
                     const double* weights, int ssc, int64_t n_obs, int64_t df_resid, double* se_out, double* vcov_out,
                     double* stats_out, double* resid_out, int64_t* n_clusters_out, cudaStream_t s);
void yoco_vcov_impl(lfe_ctx* h, const double* const* cols, int k, int64_t G, const double* beta, const double* Binv,
                    const double* nn, const double* sum_y, const double* sum_y_sq, int vcov_kind,
                    const int32_t* const* cl_codes, const int64_t* n_cl_groups, int W, int ssc, int64_t n_obs,
                    int64_t df_resid, double* se_out, double* vcov_out, double* rss_out, int64_t* n_clusters_out,
                    cudaStream_t s);
void compress_impl(lfe_ctx* h, const void* const* key_codes, const int* key_dtypes, const int64_t* n_groups, int K,
                   const double* y, const double* weights, int64_t n, int64_t out_cap, int32_t* const* key_codes_out,
                   double* n_out, int32_t* cnt_out, double* sum_y_out, double* sum_y_sq_out, int64_t* n_compressed_out,
                   cudaStream_t s);
void key_ranges_impl(lfe_ctx* h, const void* const* cols, const int* dtypes, int K, int64_t n, int64_t* min_out,
                     int64_t* max_out, cudaStream_t s);
void compress_merge_impl(lfe_ctx* h, const int32_t* const* key_codes, const int64_t* n_groups, int K,
                         const double* n_in, const int32_t* cnt_in, const double* sum_y_in, const double* sum_y_sq_in,
                         int64_t n_rows, int64_t out_cap, int32_t* const* key_codes_out, double* n_out,
                         int32_t* cnt_out, double* sum_y_out, double* sum_y_sq_out, int64_t* n_compressed_out,
                         cudaStream_t s);

// ---- errors ---------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
void lfe_set_error(const std::string& m) { g_last_error = m; }

#define LFE_API_BEGIN try {
#define LFE_API_END                                   \
  }                                                   \
  catch (const lfe_error& e) {                        \
    g_last_error = e.msg;                             \
    return e.code;                                    \
  }                                                   \
  catch (const std::exception& e) {                   \
    g_last_error = std::string("std::exception: ") + e.what(); \
    return LFE_ERR_ARG;                               \
  }                                                   \
  return LFE_OK;

static inline cudaStream_t S(void* s) { return (cudaStream_t)s; }

// ---- phase timers ---------------------------------------------------------------------------------
PhaseTimer::PhaseTimer(lfe_ctx* h_, cudaStream_t s_, int slot_) : h(h_), s(s_), slot(slot_) {
  if (!h->profiling) return;
  auto get = [&]() {
    cudaEvent_t e;
    if (!h->free_events.empty()) {
      e = h->free_events.back();
      h->free_events.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  };
  a = get();
  b = get();
  cudaEventRecord(a, s);
}
PhaseTimer::~PhaseTimer() {
  if (!a) return;
  cudaEventRecord(b, s);
  h->pending.push_back({slot, a, b});
}

static void drain_timers(lfe_ctx* h) {
  for (auto& p : h->pending) {
    if (cudaEventSynchronize(p.b) == cudaSuccess) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) h->timer_ms[p.slot] += ms;
    }
    h->free_events.push_back(p.a);
    h->free_events.push_back(p.b);
  }
  h->pending.clear();
}

// ---- NCCL via dlopen ---------------------------------------------------------------------------------
typedef struct { char internal[128]; } nccl_uid_t;
typedef int (*nccl_get_uid_fn)(nccl_uid_t*);
typedef int (*nccl_init_rank_fn)(void**, int, nccl_uid_t, int);
typedef int (*nccl_destroy_fn)(void*);
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);
typedef int (*nccl_sendrecv_fn)(void*, size_t, int, int, void*, cudaStream_t);   // ncclSend (const buf) / ncclRecv
typedef int (*nccl_group_fn)(void);
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
static struct {
  void* lib = nullptr;
  nccl_get_uid_fn get_uid = nullptr;
  nccl_init_rank_fn init_rank = nullptr;
  nccl_destroy_fn destroy = nullptr;
  nccl_allreduce_fn allreduce = nullptr;
  nccl_errstr_fn errstr = nullptr;
  nccl_sendrecv_fn send = nullptr, recv = nullptr;
  nccl_group_fn group_start = nullptr, group_end = nullptr;
  nccl_allgather_fn allgather = nullptr;
} g_nccl;
enum { NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2 };

static void load_nccl() {
  if (g_nccl.lib) return;
  const char* env = getenv("LFE_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm) continue;
    g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) LFE_THROW(LFE_ERR_NCCL, "cannot dlopen libnccl (%s); set LFE_NCCL_LIB", dlerror());
  g_nccl.get_uid = (nccl_get_uid_fn)dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.init_rank = (nccl_init_rank_fn)dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.destroy = (nccl_destroy_fn)dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.allreduce = (nccl_allreduce_fn)dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.errstr = (nccl_errstr_fn)dlsym(g_nccl.lib, "ncclGetErrorString");
  g_nccl.send = (nccl_sendrecv_fn)dlsym(g_nccl.lib, "ncclSend");
  g_nccl.recv = (nccl_sendrecv_fn)dlsym(g_nccl.lib, "ncclRecv");
  g_nccl.group_start = (nccl_group_fn)dlsym(g_nccl.lib, "ncclGroupStart");
  g_nccl.group_end = (nccl_group_fn)dlsym(g_nccl.lib, "ncclGroupEnd");
  g_nccl.allgather = (nccl_allgather_fn)dlsym(g_nccl.lib, "ncclAllGather");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.allreduce || !g_nccl.destroy)
    LFE_THROW(LFE_ERR_NCCL, "libnccl is missing required symbols");
}
#define NCCL_CHECK(x)                                                                                   \
  do {                                                                                                  \
    int _r = (x);                                                                                       \
    if (_r != 0) LFE_THROW(LFE_ERR_NCCL, "%s -> %s", #x, g_nccl.errstr ? g_nccl.errstr(_r) : "nccl error"); \
  } while (0)

void comm_allreduce(lfe_ctx* h, double* buf, int64_t count, int op_max, cudaStream_t s) {
  if (h->world <= 1 || count <= 0) return;
  h->n_collectives += 1;
  NCCL_CHECK(g_nccl.allreduce(buf, buf, (size_t)count, NCCL_FLOAT64, op_max ? NCCL_MAX : NCCL_SUM, h->nccl_comm, s));
}
// kind: 0 = uint32, 1 = uint64, 2 = int64 ; op: 0 sum, 2 max, 3 min (NCCL numbering)
void comm_allreduce_int(lfe_ctx* h, void* buf, int64_t count, int kind, int op, cudaStream_t s) {
  if (h->world <= 1 || count <= 0) return;
  const int nccl_type = kind == 0 ? 3 : (kind == 1 ? 5 : 4);
  h->n_collectives += 1;
  NCCL_CHECK(g_nccl.allreduce(buf, buf, (size_t)count, nccl_type, op, h->nccl_comm, s));
}

// all-to-all of K buffers with per-peer byte counts ([k * world + r]) in ONE grouped round of ncclSend / ncclRecv
// (self included): NCCL fuses the K x world transfers into one launch
void comm_alltoallv_multi(lfe_ctx* h, int K, const void* const* send, const int64_t* send_off, const int64_t* send_cnt,
                          void* const* recv, const int64_t* recv_off, const int64_t* recv_cnt, cudaStream_t s) {
  if (h->world <= 1 || K <= 0) return;
  if (!g_nccl.send || !g_nccl.recv || !g_nccl.group_start || !g_nccl.group_end)
    LFE_THROW(LFE_ERR_NCCL, "libnccl has no ncclSend/ncclRecv (need NCCL >= 2.7)");
  h->n_collectives += 1;
  enum { NCCL_UINT8 = 1 };
  NCCL_CHECK(g_nccl.group_start());
  for (int k = 0; k < K; ++k) {
    for (int r = 0; r < h->world; ++r) {
      const size_t i = (size_t)k * h->world + r;
      if (send_cnt[i] > 0)
        NCCL_CHECK(g_nccl.send((void*)((const char*)send[k] + send_off[i]), (size_t)send_cnt[i], NCCL_UINT8, r, h->nccl_comm, s));
      if (recv_cnt[i] > 0)
        NCCL_CHECK(g_nccl.recv((char*)recv[k] + recv_off[i], (size_t)recv_cnt[i], NCCL_UINT8, r, h->nccl_comm, s));
    }
  }
  NCCL_CHECK(g_nccl.group_end());
}

// all-gather of `bytes` bytes per rank (device buffers); false when the symbol is missing
bool comm_allgather_bytes(lfe_ctx* h, const void* send, void* recv, size_t bytes, cudaStream_t s) {
  if (h->world <= 1 || !g_nccl.allgather) return false;
  h->n_collectives += 1;
  NCCL_CHECK(g_nccl.allgather(send, recv, bytes, 1 /*ncclUint8*/, h->nccl_comm, s));
  return true;
}

void shard_plan_impl(lfe_ctx* h, const int32_t* codes, int64_t n, int64_t G, int64_t* cuts_out, int64_t* send_counts_out,
                     int64_t* recv_counts_out, int32_t* order_out, cudaStream_t s);
void shard_exchange_impl(lfe_ctx* h, const int32_t* order, const int64_t* send_counts, const int64_t* recv_counts,
                         const void* const* cols_in, const int* dtypes, const int64_t* i32_sub, int K, int64_t n,
                         void* const* cols_out, cudaStream_t s);

void expand_dummies_impl(lfe_ctx* h, const int32_t* codes, int64_t n, const int32_t* levels, int J, const double* base,
                         int out_dtype, void* const* out_cols, cudaStream_t s);
struct lfe_uploader;
lfe_uploader* uploader_create(lfe_ctx* h, int64_t chunk_bytes, int n_slots, int n_threads);
void uploader_destroy(lfe_uploader* u);
void uploader_drain(lfe_uploader* u);
int64_t uploader_enqueue(lfe_uploader* u, const void* host_src, int64_t nbytes, void* dev_dst, cudaStream_t copy_stream);
void uploader_stream_wait(lfe_uploader* u, int64_t ticket, cudaStream_t compute_stream);

// ---- exported functions -------------------------------------------------------------------------------
extern "C" {

int lfe_version(void) { return 100; }
const char* lfe_last_error(void) { return g_last_error.c_str(); }

int lfe_create(int device, lfe_handle* out) {
  LFE_API_BEGIN
  LFE_REQUIRE(out != nullptr, "lfe_create: out is NULL");
  CUDA_CHECK(cudaSetDevice(device));
  lfe_ctx* h = new lfe_ctx();
  h->device = device;
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  h->l2_bytes = prop.l2CacheSize;
  h->smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (prop.major < 10)
    fprintf(stderr, "liblfe: warning: built for sm_100a (B200), running on sm_%d%d\n", prop.major, prop.minor);
  // keep freed scratch cached in the stream-ordered pool instead of returning it to the OS
  cudaMemPool_t pool;
  CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thresh = UINT64_MAX;
  CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  *out = h;
  LFE_API_END
}

int lfe_destroy(lfe_handle h) {
  LFE_API_BEGIN
  if (!h) return LFE_OK;
  drain_timers(h);
  for (auto e : h->free_events) cudaEventDestroy(e);
  if (h->nccl_comm && g_nccl.destroy) g_nccl.destroy(h->nccl_comm);
  for (size_t r = 0; r < h->p2p.peer_base.size(); ++r)
    if ((int)r != h->rank && h->p2p.peer_base[r]) cudaIpcCloseMemHandle(h->p2p.peer_base[r]);
  if (h->p2p.base) cudaFree(h->p2p.base);
  h->last_counts.release();
  delete h;
  LFE_API_END
}

int lfe_set_option(lfe_handle h, int option, int64_t value) {
  LFE_API_BEGIN
  LFE_REQUIRE(h != nullptr, "lfe_set_option: handle is NULL");
  switch (option) {
    case LFE_OPT_COLUMNS_HINT: h->opt_columns_hint = (int)(value < 0 ? 0 : value); break;
    case LFE_OPT_L2_BLOCK_BYTES:
      LFE_REQUIRE(value >= (1 << 20), "lfe_set_option: L2 block must be >= 1 MB");
      h->opt_l2_block_bytes = value;
      break;
    case LFE_OPT_DETERMINISTIC: h->opt_deterministic = value != 0; break;
    default: LFE_THROW(LFE_ERR_ARG, "lfe_set_option: unknown option %d", option);
  }
  LFE_API_END
}
int lfe_set_profiling(lfe_handle h, int enabled) {
  LFE_API_BEGIN
  h->profiling = enabled != 0;
  LFE_API_END
}
int lfe_reset_timers(lfe_handle h) {
  LFE_API_BEGIN
  drain_timers(h);
  for (int i = 0; i < LFE_N_TIMERS; ++i) h->timer_ms[i] = 0.0;
  h->n_launches = 0;
  h->n_sweep_launches = 0;
  h->n_collectives = 0;
  LFE_API_END
}
int lfe_get_timers(lfe_handle h, double* ms_out) {
  LFE_API_BEGIN
  drain_timers(h);
  for (int i = 0; i < LFE_N_TIMERS; ++i) ms_out[i] = h->timer_ms[i];
  LFE_API_END
}
int lfe_get_counters(lfe_handle h, int64_t* out) {
  LFE_API_BEGIN
  out[0] = h->n_launches;
  out[1] = h->n_sweep_launches;
  out[2] = h->n_collectives;
  out[3] = 0;
  LFE_API_END
}

int lfe_factorize(lfe_handle h, const void* keys, int key_dtype, int64_t n, int order, int32_t* codes_out,
                  int64_t* n_groups_out, void* uniques_out, int64_t uniques_cap, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && n_groups_out, "lfe_factorize: NULL argument");
  LFE_REQUIRE(n >= 0 && n < (int64_t)0x7fffffff, "lfe_factorize: n out of range");
  LFE_REQUIRE((order & ~LFE_ORDER_GLOBAL) == LFE_ORDER_SORTED || (order & ~LFE_ORDER_GLOBAL) == LFE_ORDER_FIRST,
              "lfe_factorize: bad order %d", order);
  CUDA_CHECK(cudaSetDevice(h->device));
  int64_t G = factorize_dispatch(h, keys, key_dtype, n, order, codes_out, uniques_out, uniques_cap, S(stream));
  *n_groups_out = G;
  if (uniques_out && G > uniques_cap)
    LFE_THROW(LFE_ERR_CAPACITY, "lfe_factorize: %lld groups exceed uniques_cap %lld", (long long)G, (long long)uniques_cap);
  LFE_API_END
}

int lfe_factorize_counts(lfe_handle h, const int32_t* codes, int64_t n_groups, int32_t* counts_out,
                         int* available_out, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && available_out, "lfe_factorize_counts: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  *available_out = 0;
  if (h->last_counts.p && h->last_counts_G == n_groups && h->last_counts_codes == (const void*)codes && counts_out) {
    CUDA_CHECK(cudaMemcpyAsync(counts_out, h->last_counts.p, (size_t)n_groups * sizeof(int32_t),
                               cudaMemcpyDeviceToDevice, S(stream)));
    *available_out = 1;
  }
  h->last_counts.release();
  h->last_counts_G = -1;
  h->last_counts_codes = nullptr;
  LFE_API_END
}

int lfe_group_counts(lfe_handle h, const int32_t* codes, int64_t n, int64_t n_groups, int32_t* counts_out,
                     void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && n_groups >= 0, "lfe_group_counts: bad argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  PhaseTimer t(h, S(stream), 1);
  prim_histogram(h, codes, n, n_groups, (uint32_t*)counts_out, S(stream));
  LFE_API_END
}

int lfe_singleton_mask(lfe_handle h, const int32_t* const* codes, const int64_t* n_groups, int F, int64_t n,
                       int local_fe, const int32_t* const* counts, uint8_t* keep_out, int64_t* n_keep_out,
                       int64_t* fe_dims_out, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && codes && n_groups && keep_out && n_keep_out, "lfe_singleton_mask: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  singleton_mask_impl(h, codes, n_groups, F, n, local_fe, counts, keep_out, n_keep_out, fe_dims_out, S(stream));
  LFE_API_END
}

int lfe_compact(lfe_handle h, const uint8_t* keep, int64_t n, const double* const* f64_in, double* const* f64_out,
                int C, const int32_t* const* i32_in, int32_t* const* i32_out, int K, void* stream) {
  LFE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  compact_impl(h, keep, n, f64_in, f64_out, C, i32_in, i32_out, K, S(stream));
  LFE_API_END
}

int lfe_sort_rows(lfe_handle h, const int32_t* codes, int64_t n, int64_t n_groups, int32_t* perm_out,
                  int* already_sorted_out, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(n >= 0 && n < (int64_t)0x7fffffff, "lfe_sort_rows: n out of range");
  CUDA_CHECK(cudaSetDevice(h->device));
  sort_rows_impl(h, codes, n, n_groups, perm_out, already_sorted_out, S(stream));
  LFE_API_END
}

int lfe_gather(lfe_handle h, const int32_t* perm, int64_t n, const double* const* f64_in, double* const* f64_out,
               int C, const int32_t* const* i32_in, int32_t* const* i32_out, int K, void* stream) {
  LFE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  gather_impl(h, perm, n, f64_in, f64_out, C, i32_in, i32_out, K, S(stream));
  LFE_API_END
}

int lfe_demean(lfe_handle h, double* const* cols, int C, const int32_t* const* codes, const int64_t* n_groups,
               const int* fe_order, int F, int64_t n, const double* weights, double tol, int max_iter, int min_iter,
               int local_fe, int flags, int* iterations_out, double* max_mean_out, int* converged_out, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && cols && codes && n_groups && fe_order && iterations_out, "lfe_demean: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  demean_impl(h, cols, C, codes, n_groups, fe_order, F, n, weights, tol, max_iter, min_iter, local_fe, flags,
              iterations_out, max_mean_out, converged_out, S(stream));
  LFE_API_END
}

int lfe_demean_plan_create(lfe_handle h, const int32_t* const* codes, const int64_t* n_groups, int F, int64_t n,
                           const double* weights, int local_fe, const int32_t* const* counts, lfe_plan_t* plan_out,
                           void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && codes && n_groups && plan_out, "lfe_demean_plan_create: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  *plan_out = demean_plan_create(h, codes, n_groups, F, n, weights, local_fe, counts, S(stream));
  LFE_API_END
}
int lfe_demean_plan_run(lfe_handle h, lfe_plan_t plan, double* const* cols, const double* const* cols_src, int C,
                        const int* fe_order, double tol, int max_iter, int min_iter, int flags, int* iterations_out,
                        double* max_mean_out, int* converged_out, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && plan && cols && fe_order && iterations_out, "lfe_demean_plan_run: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  demean_plan_run(h, plan, cols, cols_src, C, fe_order, tol, max_iter, min_iter, flags, iterations_out, max_mean_out,
                  converged_out, S(stream));
  LFE_API_END
}
int lfe_demean_plan_fit(lfe_handle h, lfe_plan_t plan, const double* const* cols, int C, const int* fe_order, double tol,
                        int max_iter, int min_iter, int flags, int vcov_kind, const int32_t* cl_codes, int64_t n_cl_groups,
                        int ssc, int64_t n_obs, int64_t df_resid, double* beta_full_out, double* se_out, double* vcov_out,
                        double* stats_out, int* iterations_out, double* max_mean_out, int* converged_out,
                        int64_t* n_clusters_out, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && plan && cols && fe_order && iterations_out && beta_full_out && (C == 1 || se_out),
              "lfe_demean_plan_fit: NULL argument");
  const int kind = vcov_kind & ~LFE_VCOV_CLUSTER_LOCAL;
  LFE_REQUIRE(kind >= 0 && kind <= 2, "lfe_demean_plan_fit: unknown vcov kind %d", kind);
  LFE_REQUIRE(kind != LFE_VCOV_CLUSTER || (cl_codes && n_cl_groups >= 1), "lfe_demean_plan_fit: cluster codes required");
  CUDA_CHECK(cudaSetDevice(h->device));
  demean_plan_fit(h, plan, cols, C, fe_order, tol, max_iter, min_iter, flags, vcov_kind, cl_codes, n_cl_groups, ssc, n_obs,
                  df_resid, beta_full_out, se_out, vcov_out, stats_out, iterations_out, max_mean_out, converged_out,
                  n_clusters_out, S(stream));
  LFE_API_END
}
int lfe_demean_plan_destroy(lfe_handle h, lfe_plan_t plan) {
  LFE_API_BEGIN
  if (h) CUDA_CHECK(cudaSetDevice(h->device));
  demean_plan_destroy(plan);
  LFE_API_END
}

int lfe_gram(lfe_handle h, const double* const* cols, int C, int64_t n, const double* weights, int add_intercept,
             double* gram_out, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && gram_out && (cols || C == 0), "lfe_gram: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  gram_impl(h, cols, C, n, weights, add_intercept, gram_out, S(stream));
  LFE_API_END
}

int lfe_solve_spd(lfe_handle h, const double* A, const double* b, int p, double* beta_out, double* Ainv_out,
                  int* used_fallback_out, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && A && b && beta_out && Ainv_out, "lfe_solve_spd: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  solve_spd_impl(h, A, b, p, beta_out, Ainv_out, used_fallback_out, S(stream));
  LFE_API_END
}

int lfe_resid_vcov(lfe_handle h, const double* const* cols, int k, int64_t n, const double* beta_full,
                   const double* Binv, int vcov_kind, const int32_t* const* cl_codes, const int64_t* n_cl_groups, int W,
                   const double* weights, int ssc, int64_t n_obs, int64_t df_resid, double* se_out, double* vcov_out,
                   double* stats_out, double* resid_out, int64_t* n_clusters_out, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && cols && beta_full && (k == 0 || (Binv && se_out)), "lfe_resid_vcov: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  resid_vcov_impl(h, cols, k, n, beta_full, Binv, vcov_kind, cl_codes, n_cl_groups, W, weights, ssc, n_obs, df_resid,
                  se_out, vcov_out, stats_out, resid_out, n_clusters_out, S(stream));
  LFE_API_END
}

int lfe_compress(lfe_handle h, const void* const* key_codes, const int* key_dtypes, const int64_t* n_groups, int K,
                 const double* y, const double* weights, int64_t n, int64_t out_cap, int32_t* const* key_codes_out,
                 double* n_out, int32_t* cnt_out, double* sum_y_out, double* sum_y_sq_out, int64_t* n_compressed_out,
                 void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && n_compressed_out && (K == 0 || (key_codes && n_groups)), "lfe_compress: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  compress_impl(h, key_codes, key_dtypes, n_groups, K, y, weights, n, out_cap, key_codes_out, n_out, cnt_out,
                sum_y_out, sum_y_sq_out, n_compressed_out, S(stream));
  LFE_API_END
}

int lfe_key_ranges(lfe_handle h, const void* const* cols, const int* dtypes, int K, int64_t n, int64_t* min_out,
                   int64_t* max_out, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && cols && dtypes && min_out && max_out && n >= 0, "lfe_key_ranges: bad argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  key_ranges_impl(h, cols, dtypes, K, n, min_out, max_out, S(stream));
  LFE_API_END
}

int lfe_compress_merge(lfe_handle h, const int32_t* const* key_codes, const int64_t* n_groups, int K,
                       const double* n_in, const int32_t* cnt_in, const double* sum_y_in, const double* sum_y_sq_in,
                       int64_t n_rows, int64_t out_cap, int32_t* const* key_codes_out, double* n_out, int32_t* cnt_out,
                       double* sum_y_out, double* sum_y_sq_out, int64_t* n_compressed_out, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && n_compressed_out && key_codes && n_groups, "lfe_compress_merge: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  compress_merge_impl(h, key_codes, n_groups, K, n_in, cnt_in, sum_y_in, sum_y_sq_in, n_rows, out_cap, key_codes_out,
                      n_out, cnt_out, sum_y_out, sum_y_sq_out, n_compressed_out, S(stream));
  LFE_API_END
}

int lfe_yoco_vcov(lfe_handle h, const double* const* cols, int k, int64_t n_compressed, const double* beta,
                  const double* Binv, const double* nn, const double* sum_y, const double* sum_y_sq, int vcov_kind,
                  const int32_t* const* cl_codes, const int64_t* n_cl_groups, int W, int ssc, int64_t n_obs,
                  int64_t df_resid, double* se_out, double* vcov_out, double* rss_out, int64_t* n_clusters_out,
                  void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && cols && beta && Binv && nn && sum_y && sum_y_sq && se_out, "lfe_yoco_vcov: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  yoco_vcov_impl(h, cols, k, n_compressed, beta, Binv, nn, sum_y, sum_y_sq, vcov_kind, cl_codes, n_cl_groups, W, ssc,
                 n_obs, df_resid, se_out, vcov_out, rss_out, n_clusters_out, S(stream));
  LFE_API_END
}

int lfe_comm_unique_id(void* id_out) {
  LFE_API_BEGIN
  load_nccl();
  nccl_uid_t id;
  NCCL_CHECK(g_nccl.get_uid(&id));
  memcpy(id_out, &id, sizeof(id));
  LFE_API_END
}
int lfe_comm_init(lfe_handle h, const void* id, int rank, int world) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && id && world >= 1 && rank >= 0 && rank < world, "lfe_comm_init: bad argument");
  load_nccl();
  CUDA_CHECK(cudaSetDevice(h->device));
  nccl_uid_t uid;
  memcpy(&uid, id, sizeof(uid));
  NCCL_CHECK(g_nccl.init_rank(&h->nccl_comm, world, uid, rank));
  h->rank = rank;
  h->world = world;
  h->world_attached = world;
  LFE_API_END
}
int lfe_comm_set_active(lfe_handle h, int active) {
  LFE_API_BEGIN
  LFE_REQUIRE(h != nullptr, "lfe_comm_set_active: handle is NULL");
  h->world = (active && h->nccl_comm) ? h->world_attached : 1;
  LFE_API_END
}
int lfe_comm_destroy(lfe_handle h) {
  LFE_API_BEGIN
  if (h->nccl_comm) NCCL_CHECK(g_nccl.destroy(h->nccl_comm));
  h->nccl_comm = nullptr;
  h->world = 1;
  h->world_attached = 1;
  h->rank = 0;
  LFE_API_END
}
int lfe_shard_plan(lfe_handle h, const int32_t* codes, int64_t n, int64_t n_groups, int64_t* cuts_out,
                   int64_t* send_counts_out, int64_t* recv_counts_out, int32_t* order_out, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && cuts_out && send_counts_out && recv_counts_out && (order_out || n == 0) && (codes || n == 0),
              "lfe_shard_plan: NULL argument");
  LFE_REQUIRE(n >= 0 && n < (int64_t)0x7fffffff, "lfe_shard_plan: n out of range");
  CUDA_CHECK(cudaSetDevice(h->device));
  shard_plan_impl(h, codes, n, n_groups, cuts_out, send_counts_out, recv_counts_out, order_out, S(stream));
  LFE_API_END
}
int lfe_shard_exchange(lfe_handle h, const int32_t* order, const int64_t* send_counts, const int64_t* recv_counts,
                       const void* const* cols_in, const int* dtypes, const int64_t* i32_sub, int K, int64_t n,
                       void* const* cols_out, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && send_counts && recv_counts && (K == 0 || (cols_in && dtypes && cols_out)),
              "lfe_shard_exchange: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  shard_exchange_impl(h, order, send_counts, recv_counts, cols_in, dtypes, i32_sub, K, n, cols_out, S(stream));
  LFE_API_END
}
int lfe_expand_dummies(lfe_handle h, const int32_t* codes, int64_t n, const int32_t* levels, int J, const double* base,
                       int out_dtype, void* const* out_cols, void* stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && (J == 0 || (levels && out_cols)) && (codes || n == 0) && J >= 0, "lfe_expand_dummies: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  expand_dummies_impl(h, codes, n, levels, J, base, out_dtype, out_cols, S(stream));
  LFE_API_END
}
int lfe_uploader_create(lfe_handle h, int64_t chunk_bytes, int n_slots, int n_threads, lfe_uploader_t* out) {
  LFE_API_BEGIN
  LFE_REQUIRE(h && out, "lfe_uploader_create: NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  *out = uploader_create(h, chunk_bytes, n_slots, n_threads);
  LFE_API_END
}
int lfe_uploader_destroy(lfe_uploader_t u) {
  LFE_API_BEGIN
  uploader_destroy(u);
  LFE_API_END
}
int lfe_uploader_enqueue(lfe_uploader_t u, const void* host_src, int64_t nbytes, void* dev_dst, void* copy_stream,
                         int64_t* ticket_out) {
  LFE_API_BEGIN
  LFE_REQUIRE(u && ticket_out, "lfe_uploader_enqueue: NULL argument");
  *ticket_out = uploader_enqueue(u, host_src, nbytes, dev_dst, S(copy_stream));
  LFE_API_END
}
int lfe_uploader_stream_wait(lfe_uploader_t u, int64_t ticket, void* compute_stream) {
  LFE_API_BEGIN
  LFE_REQUIRE(u != nullptr, "lfe_uploader_stream_wait: NULL argument");
  uploader_stream_wait(u, ticket, S(compute_stream));
  LFE_API_END
}
int lfe_uploader_drain(lfe_uploader_t u) {
  LFE_API_BEGIN
  LFE_REQUIRE(u != nullptr, "lfe_uploader_drain: NULL argument");
  uploader_drain(u);
  LFE_API_END
}
int lfe_allreduce_f64(lfe_handle h, double* buf, int64_t count, void* stream) {
  LFE_API_BEGIN
  PhaseTimer t(h, S(stream), 13);
  comm_allreduce(h, buf, count, 0, S(stream));
  LFE_API_END
}

}  // extern "C"
