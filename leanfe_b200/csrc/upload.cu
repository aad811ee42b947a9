// upload.cu -- pageable host memory -> device through a chunked pinned ring filled by host threads.
//
// The reference's API receives DataFrames / NumPy columns (python/leanfe/polars_impl.py:342-347), i.e. PAGEABLE
// memory.  A cudaMemcpyAsync from pageable memory is staged by the driver, synchronously and on one thread
// (~10 GB/s measured: 308 ms for the 3.2 GB of BASELINE config 2), so the host->device leg of such a fit ran 5x
// slower than from pinned buffers.  Here the staging is ours: a ring of `n_slots` pinned chunks; a dispatcher
// thread walks the queued columns chunk by chunk, `n_threads` workers memcpy the pieces of a chunk into its slot
// in parallel, and the chunk leaves with ONE cudaMemcpyAsync on the caller's copy stream; a slot is reused when the
// event recorded after its copy has completed.  The caller's thread never blocks on staging: it enqueues all
// columns up front (ids first) and, right before a column's first use, makes its compute stream wait for the
// event recorded behind the column's last chunk (lfe_uploader_stream_wait).
#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>

#include "common.cuh"

struct UploadJob {
  int64_t ticket;
  const char* src;
  char* dst;
  int64_t nbytes;
  cudaStream_t stream;
  cudaEvent_t done = nullptr;   // recorded on `stream` behind the last chunk
  bool issued = false;          // every chunk has been handed to the stream (done is recorded)
  int error = 0;
};

struct lfe_uploader {
  int device = 0;
  int64_t chunk = 0;
  int n_slots = 0, n_threads = 0;
  std::vector<char*> slots;
  std::vector<cudaEvent_t> slot_ev;
  std::vector<bool> slot_used;
  // job queue
  std::mutex mu;
  std::condition_variable cv_jobs, cv_issued;
  std::deque<UploadJob*> queue;
  std::vector<UploadJob*> all;      // ticket -> job (kept until drain)
  int64_t next_ticket = 0;
  bool stop = false;
  std::thread dispatcher;
  // worker pool: one "round" = the pieces of one chunk
  std::mutex wmu;
  std::condition_variable cv_work, cv_done;
  std::vector<std::thread> workers;
  const char* w_src = nullptr;
  char* w_dst = nullptr;
  int64_t w_len = 0;
  uint64_t w_round = 0;
  int w_pending = 0;
  bool w_stop = false;
};

static void worker_main(lfe_uploader* u, int idx) {
  uint64_t seen = 0;
  for (;;) {
    const char* src;
    char* dst;
    int64_t len;
    {
      std::unique_lock<std::mutex> lk(u->wmu);
      u->cv_work.wait(lk, [&] { return u->w_stop || u->w_round != seen; });
      if (u->w_stop) return;
      seen = u->w_round;
      src = u->w_src; dst = u->w_dst; len = u->w_len;
    }
    // piece idx of the chunk, 4 KB aligned boundaries
    const int T = u->n_threads;
    int64_t per = ((len / T) + 4095) & ~(int64_t)4095;
    int64_t a = per * idx, b = a + per < len ? a + per : len;
    if (a < b) memcpy(dst + a, src + a, (size_t)(b - a));
    {
      std::lock_guard<std::mutex> lk(u->wmu);
      if (--u->w_pending == 0) u->cv_done.notify_one();
    }
  }
}

static void stage_chunk(lfe_uploader* u, const char* src, char* dst, int64_t len) {
  if (u->n_threads <= 1 || len < (1 << 20)) {
    memcpy(dst, src, (size_t)len);
    return;
  }
  std::unique_lock<std::mutex> lk(u->wmu);
  u->w_src = src; u->w_dst = dst; u->w_len = len;
  u->w_pending = u->n_threads;
  u->w_round++;
  u->cv_work.notify_all();
  u->cv_done.wait(lk, [&] { return u->w_pending == 0; });
}

static void dispatcher_main(lfe_uploader* u) {
  cudaSetDevice(u->device);
  int slot = 0;
  for (;;) {
    UploadJob* job;
    {
      std::unique_lock<std::mutex> lk(u->mu);
      u->cv_jobs.wait(lk, [&] { return u->stop || !u->queue.empty(); });
      if (u->queue.empty()) return;   // stop requested and nothing left
      job = u->queue.front();
      u->queue.pop_front();
    }
    int err = 0;
    for (int64_t off = 0; off < job->nbytes && !err; off += u->chunk) {
      const int64_t len = job->nbytes - off < u->chunk ? job->nbytes - off : u->chunk;
      if (u->slot_used[slot] && cudaEventSynchronize(u->slot_ev[slot]) != cudaSuccess) err = 1;
      stage_chunk(u, job->src + off, u->slots[slot], len);
      if (!err && cudaMemcpyAsync(job->dst + off, u->slots[slot], (size_t)len, cudaMemcpyHostToDevice, job->stream) != cudaSuccess) err = 1;
      if (!err && cudaEventRecord(u->slot_ev[slot], job->stream) != cudaSuccess) err = 1;
      u->slot_used[slot] = true;
      slot = (slot + 1) % u->n_slots;
    }
    if (!err && cudaEventRecord(job->done, job->stream) != cudaSuccess) err = 1;
    {
      std::lock_guard<std::mutex> lk(u->mu);
      job->error = err;
      job->issued = true;
    }
    u->cv_issued.notify_all();
  }
}

lfe_uploader* uploader_create(lfe_ctx* h, int64_t chunk_bytes, int n_slots, int n_threads) {
  LFE_REQUIRE(chunk_bytes >= (1 << 16) && n_slots >= 2 && n_slots <= 64 && n_threads >= 1 && n_threads <= 64,
              "lfe_uploader_create: chunk_bytes >= 64 KB, 2..64 slots, 1..64 threads");
  std::unique_ptr<lfe_uploader> u(new lfe_uploader());
  u->device = h->device;
  u->chunk = (chunk_bytes + 4095) & ~(int64_t)4095;
  u->n_slots = n_slots;
  u->n_threads = n_threads;
  u->slot_used.assign(n_slots, false);
  for (int i = 0; i < n_slots; ++i) {
    char* p = nullptr;
    CUDA_CHECK(cudaHostAlloc((void**)&p, (size_t)u->chunk, cudaHostAllocDefault));
    u->slots.push_back(p);
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    u->slot_ev.push_back(e);
  }
  lfe_uploader* raw = u.release();
  if (n_threads > 1)
    for (int i = 0; i < n_threads; ++i) raw->workers.emplace_back(worker_main, raw, i);
  raw->dispatcher = std::thread(dispatcher_main, raw);
  return raw;
}

static void release_jobs(lfe_uploader* u) {
  for (UploadJob* j : u->all) {
    if (j->done) cudaEventDestroy(j->done);
    delete j;
  }
  u->all.clear();
}

// wait until every queued job has been issued and its copies have completed; forget the tickets
void uploader_drain(lfe_uploader* u) {
  std::unique_lock<std::mutex> lk(u->mu);
  u->cv_issued.wait(lk, [&] {
    for (UploadJob* j : u->all)
      if (!j->issued) return false;
    return true;
  });
  for (UploadJob* j : u->all)
    if (!j->error) cudaEventSynchronize(j->done);
  release_jobs(u);
}

void uploader_destroy(lfe_uploader* u) {
  if (!u) return;
  uploader_drain(u);
  {
    std::lock_guard<std::mutex> lk(u->mu);
    u->stop = true;
  }
  u->cv_jobs.notify_all();
  if (u->dispatcher.joinable()) u->dispatcher.join();
  {
    std::lock_guard<std::mutex> lk(u->wmu);
    u->w_stop = true;
  }
  u->cv_work.notify_all();
  for (auto& t : u->workers) t.join();
  for (char* p : u->slots) cudaFreeHost(p);
  for (cudaEvent_t e : u->slot_ev) cudaEventDestroy(e);
  delete u;
}

int64_t uploader_enqueue(lfe_uploader* u, const void* host_src, int64_t nbytes, void* dev_dst, cudaStream_t copy_stream) {
  LFE_REQUIRE(nbytes >= 0 && (nbytes == 0 || (host_src && dev_dst)), "lfe_uploader_enqueue: bad argument");
  std::unique_ptr<UploadJob> j(new UploadJob());
  j->src = (const char*)host_src;
  j->dst = (char*)dev_dst;
  j->nbytes = nbytes;
  j->stream = copy_stream;
  CUDA_CHECK(cudaEventCreateWithFlags(&j->done, cudaEventDisableTiming));
  int64_t t;
  {
    std::lock_guard<std::mutex> lk(u->mu);
    t = j->ticket = u->next_ticket++;
    u->all.push_back(j.get());
    u->queue.push_back(j.release());
  }
  u->cv_jobs.notify_one();
  return t;
}

// host-blocks until the job's last chunk has been issued, then orders `compute_stream` after its completion event
void uploader_stream_wait(lfe_uploader* u, int64_t ticket, cudaStream_t compute_stream) {
  UploadJob* job = nullptr;
  {
    std::unique_lock<std::mutex> lk(u->mu);
    for (UploadJob* j : u->all)
      if (j->ticket == ticket) job = j;
    LFE_REQUIRE(job != nullptr, "lfe_uploader_stream_wait: unknown ticket %lld", (long long)ticket);
    u->cv_issued.wait(lk, [&] { return job->issued; });
  }
  if (job->error) LFE_THROW(LFE_ERR_CUDA, "lfe_uploader: a host->device chunk copy failed");
  CUDA_CHECK(cudaStreamWaitEvent(compute_stream, job->done, 0));
}
