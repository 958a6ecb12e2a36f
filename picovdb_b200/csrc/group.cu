// pvdb_group_t: a row-sharded store over several GPUs of ONE process (SURVEY.md 8(b) level 2,
// `devices=[...]`).  A library user with a single Python process gets the same row partition,
// kernels and peer-memory exchange as the one-process-per-GPU mode (sharded.py): the group owns one
// pvdb_store_t and one exchange end per device, connected through peer access, and one worker thread
// per device so that the per-shard host work (staging copies, the guard's stream synchronisation)
// runs concurrently -- the kernels of all shards must be in flight together, they wait for each
// other's lists (exchange.cuh).
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "exchange.cuh"
#include "store.cuh"

using namespace pvdb;

struct pvdb_group {
  int world = 0;
  int dim = 0;
  int64_t per = 0;        // rows per shard (multiple of 32)
  int64_t capacity = 0;   // total rows the partition was laid out for
  int64_t slot_keys = 0;
  std::vector<pvdb_store*> stores;
  std::vector<pvdb_exchange*> exs;

  // one worker per shard
  struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = 0;
    std::string err;
  };
  std::vector<Worker*> workers;
  std::mutex call_mu;  // one group call at a time (the exchange sequence must be the same on all shards)

  void start_workers() {
    for (int i = 0; i < world; ++i) {
      Worker* w = new Worker();
      workers.push_back(w);
      w->th = std::thread([w]() {
        std::unique_lock<std::mutex> lk(w->mu);
        for (;;) {
          w->cv.wait(lk, [w]() { return w->has_job || w->quit; });
          if (w->quit) return;
          std::function<int()> job = std::move(w->job);
          w->has_job = false;
          lk.unlock();
          const int rc = job();
          std::string err = rc != PVDB_OK ? g_last_error : std::string();
          lk.lock();
          w->rc = rc;
          w->err = std::move(err);
          w->done = true;
          w->cv.notify_all();
        }
      });
    }
  }
  // Run fn(i) for every shard on its worker; returns the first failure (message copied to this thread).
  int run_all(const std::function<int(int)>& fn) {
    for (int i = 0; i < world; ++i) {
      Worker* w = workers[i];
      std::lock_guard<std::mutex> lk(w->mu);
      w->job = [fn, i]() { return fn(i); };
      w->has_job = true;
      w->done = false;
      w->cv.notify_all();
    }
    int rc = PVDB_OK;
    for (int i = 0; i < world; ++i) {
      Worker* w = workers[i];
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [w]() { return w->done; });
      if (w->rc != PVDB_OK && rc == PVDB_OK) {
        rc = w->rc;
        g_last_error = w->err;
      }
    }
    return rc;
  }
  void stop_workers() {
    for (Worker* w : workers) {
      {
        std::lock_guard<std::mutex> lk(w->mu);
        w->quit = true;
        w->cv.notify_all();
      }
      if (w->th.joinable()) w->th.join();
      delete w;
    }
    workers.clear();
  }
};

extern "C" int pvdb_group_destroy(pvdb_group_t* g) {
  if (!g) return PVDB_OK;
  g->stop_workers();
  for (pvdb_exchange* ex : g->exs) pvdb_exchange_destroy(ex);
  for (pvdb_store* s : g->stores) pvdb_store_destroy(s);
  delete g;
  return PVDB_OK;
}

extern "C" int pvdb_group_create(pvdb_group_t** out, const int* devices, int ndev, int dim, int64_t capacity_rows,
                                 int flags, int64_t slot_keys) {
  if (!out) return fail(PVDB_ERR_INVALID, "out is null");
  *out = nullptr;
  if (!devices || ndev < 1 || ndev > kMaxWorld) return fail(PVDB_ERR_INVALID, "group: 1..%d devices", kMaxWorld);
  if (capacity_rows < 1) return fail(PVDB_ERR_INVALID, "group: the row partition needs the total capacity");
  if (slot_keys < 1) slot_keys = 4096 * 16;
  pvdb_group* g = new pvdb_group();
  g->world = ndev;
  g->dim = dim;
  g->capacity = capacity_rows;
  g->slot_keys = slot_keys;
  // the same partition as sharded.py::shard_range: ceil(capacity / world) rounded up to 32 rows
  g->per = ((capacity_rows + ndev - 1) / ndev + 31) / 32 * 32;
  int rc = PVDB_OK;
  for (int i = 0; i < ndev && rc == PVDB_OK; ++i) {
    pvdb_store* s = nullptr;
    const int64_t row0 = std::min<int64_t>(capacity_rows, static_cast<int64_t>(i) * g->per);
    const int64_t n_local = std::min<int64_t>(capacity_rows, row0 + g->per) - row0;
    rc = pvdb_store_create(&s, devices[i], dim, std::max<int64_t>(n_local, 1), flags);
    if (rc != PVDB_OK) break;
    g->stores.push_back(s);
    rc = pvdb_store_set_row_base(s, row0);
    if (rc != PVDB_OK) break;
    pvdb_exchange* ex = nullptr;
    rc = pvdb_exchange_create(&ex, devices[i], ndev, i, slot_keys);
    if (rc != PVDB_OK) break;
    g->exs.push_back(ex);
  }
  if (rc == PVDB_OK && ndev > 1) rc = pvdb_exchange_connect_local(g->exs.data(), ndev);
  if (rc != PVDB_OK) {
    const std::string keep = g_last_error;
    pvdb_group_destroy(g);
    g_last_error = keep;
    return rc;
  }
  g->start_workers();
  *out = g;
  return PVDB_OK;
}

extern "C" int pvdb_group_size(pvdb_group_t* g, int* out_world, int64_t* out_rows_per_shard) {
  if (!g) return fail(PVDB_ERR_INVALID, "null group handle");
  if (out_world) *out_world = g->world;
  if (out_rows_per_shard) *out_rows_per_shard = g->per;
  return PVDB_OK;
}

extern "C" pvdb_store_t* pvdb_group_store(pvdb_group_t* g, int shard) {
  if (!g || shard < 0 || shard >= g->world) return nullptr;
  return g->stores[shard];
}

extern "C" int pvdb_group_search(pvdb_group_t* g, const float* queries, int64_t nq, int k,
                                 const uint32_t* prefilter_bits, int flags, float* out_scores, int64_t* out_rows) {
  if (!g) return fail(PVDB_ERR_INVALID, "null group handle");
  if (nq < 0 || k < 1 || (nq > 0 && (!queries || !out_scores || !out_rows)))
    return fail(PVDB_ERR_INVALID, "group_search: bad arguments (nq=%lld, k=%d)", (long long)nq, k);
  if (nq == 0) return PVDB_OK;
  if (nq * k > g->slot_keys)
    return fail(PVDB_ERR_INVALID, "group_search: %lld x %d results exceed the exchange slot (%lld keys)", (long long)nq,
                k, (long long)g->slot_keys);
  std::lock_guard<std::mutex> call(g->call_mu);
  // Every shard answers from its rows and exchanges lists with the others inside its kernels; all
  // shards end with the same merged result, shard 0 hands it to the caller.  prefilter_bits is the
  // GLOBAL bitmap: shard i reads the words of its own rows (shards start on word boundaries).
  std::vector<float> sink_s;
  std::vector<int64_t> sink_r;
  if (g->world > 1) {
    sink_s.resize(static_cast<size_t>(g->world - 1) * nq * k);
    sink_r.resize(static_cast<size_t>(g->world - 1) * nq * k);
  }
  return g->run_all([&](int i) -> int {
    const uint32_t* pf = prefilter_bits ? prefilter_bits + (static_cast<int64_t>(i) * g->per) / 32 : nullptr;
    float* os = i == 0 ? out_scores : sink_s.data() + static_cast<size_t>(i - 1) * nq * k;
    int64_t* orow = i == 0 ? out_rows : sink_r.data() + static_cast<size_t>(i - 1) * nq * k;
    return pvdb_search_exchange(g->stores[i], g->exs[i], queries, nq, k, pf, flags, os, orow);
  });
}
