// C ABI of include/enterp.h: handle management, one-time upload of the resolved tables, kernel
// dispatch, and the host-buffer pipelines. No CPU evaluation path exists in this file or anywhere
// in the library: every eval entry point ends in a kernel launch or an error status.
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <map>
#include <mutex>
#include <tuple>
#include <string>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <sched.h>

#include "../../include/enterp.h"
#include "launch.hpp"
#include "resolve.hpp"

namespace enterp {

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
  static std::mutex mu;
  static std::vector<int> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  if ((int)cache.size() <= dev) cache.resize(dev + 1, 0);
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

int cached_occupancy(const void* kernel, int block, size_t smem, size_t opt_in_smem, cudaError_t* err) {
  struct Key {
    const void* k;
    int dev, block;
    size_t smem;
    bool operator<(const Key& o) const { return std::tie(k, dev, block, smem) < std::tie(o.k, o.dev, o.block, o.smem); }
  };
  static std::mutex mu;
  static std::map<Key, int> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  const Key key{kernel, dev, block, smem};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
  }
  if (opt_in_smem > 0) {
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)opt_in_smem);
    if (e != cudaSuccess) {
      if (err) *err = e;
      return 1;
    }
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;  // not cached: a transient failure must not stick
    return per_sm;
  }
  std::lock_guard<std::mutex> lk(mu);
  cache[key] = per_sm;
  return per_sm;
}

// Debug/ablation switch (env ENTERP_DISABLE_ROWS=1): route B-splines through the generic kernel.
static const bool g_disable_rows = [] {
  const char* v = std::getenv("ENTERP_DISABLE_ROWS");
  return v && v[0] == '1';
}();

// Ablation switch (env ENTERP_DISABLE_LEAN=1): Linear-equidistant / Bezier take() through the generic kernels.
static const bool g_disable_lean = [] {
  const char* v = std::getenv("ENTERP_DISABLE_LEAN");
  return v && v[0] == '1';
}();

// Ablation switch (env ENTERP_NO_RECORDS=1): gather knots and elements separately in the Linear kernel.
static const bool g_no_records = [] {
  const char* v = std::getenv("ENTERP_NO_RECORDS");
  return v && v[0] == '1';
}();

// Ablation switch (env ENTERP_NO_BUCKETS=1): sorted searches descend the pivot tree only.
static const bool g_no_buckets = [] {
  const char* v = std::getenv("ENTERP_NO_BUCKETS");
  return v && v[0] == '1';
}();

// Ablation switch (env ENTERP_NO_SLOTS=1): Linear batches over sorted knots without the grid-indexed slot table.
static const bool g_no_slots = [] {
  const char* v = std::getenv("ENTERP_NO_SLOTS");
  return v && v[0] == '1';
}();

// Ablation switch (env ENTERP_BALLOT_SEARCH=1): sorted searches through the warp-cooperative ballot/shuffle 32-ary
// search of the north star (kernels.cuh count_le_ballot) instead of bucket index / records / slots / pivot tree.
static const bool g_ballot_search = [] {
  const char* v = std::getenv("ENTERP_BALLOT_SEARCH");
  return v && v[0] == '1';
}();

// Ablation switch (env ENTERP_NO_SMEM_PIVOTS=1): leave every pivot level in global memory / L1.
static const bool g_no_smem_pivots = [] {
  const char* v = std::getenv("ENTERP_NO_SMEM_PIVOTS");
  return v && v[0] == '1';
}();
constexpr size_t SEARCH_SMEM_BUDGET = 96u * 1024u;

static thread_local std::string t_cuda_error;
static enterp_status cuda_fail(cudaError_t e, const char* where) {
  t_cuda_error = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  cudaGetLastError();  // clear sticky-less errors
  return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice ? ENTERP_NO_DEVICE
                                                                                                    : ENTERP_CUDA_ERROR;
}
#define CK(call, where)                                  \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return cuda_fail(e__, where); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Scratch of the host-buffer pipelines (double buffered). Pipes are pooled per DEVICE and shared by every curve,
// so that a freshly built curve does not pay for device/pinned allocations (a build() + collect() step is then
// bound by PCIe alone). A call borrows an idle pipe (or creates one): concurrent host-buffer callers on one
// device no longer serialise on a single mutex; idle pipes are freed by enterp_host_scratch_release().
// Every buffer carries its OWN size, reset to 0 whenever the buffer is freed, so a failed (re)allocation can
// never leave a null pointer behind a non-zero recorded size.
// Plain pipeline depth: chunks the copy engine has queued ahead. Two suffice on a quiet host; four ride out the
// scheduling hiccups of a shared one (a host thread that is late by a few milliseconds no longer starves the DMA).
constexpr int PLAIN_SLOTS = 4;

struct HostBuf {
  void* p = nullptr;
  size_t bytes = 0;
};
struct HostPipe {
  HostBuf dev_in[PLAIN_SLOTS], dev_out[PLAIN_SLOTS], stage_in[PLAIN_SLOTS], stage_out[PLAIN_SLOTS];
  HostBuf rle_t[2], rle_v[2];  // run-length pipeline: stepper values and their evaluations on the device
  HostBuf rle_stage[2];        // ... and the pinned landing buffers of the distinct values
  cudaStream_t rstream[2] = {nullptr, nullptr};
  cudaEvent_t rdone[2] = {nullptr, nullptr};
  // what this pipe has learnt about its host (kept across calls): the plain pipeline's rate, and how often in a row the
  // hybrid pipeline failed to beat it (hosts whose memory system is busy elsewhere)
  double plain_gbs = 0;
  int hybrid_bad = 0;
  cudaStream_t stream[PLAIN_SLOTS] = {};
  cudaEvent_t done[PLAIN_SLOTS] = {};
  bool busy = false;
};

struct Replica {
  int device = -1;
  void* block = nullptr;                // ONE device allocation holding every table below (and the invalid counter)
  void* store = nullptr;
  void* elems = nullptr;
  void* rows = nullptr;
  void* records = nullptr;
  void* slots = nullptr;                // Linear, sorted knots: grid-indexed copies of the interval records (resolve.hpp)
  void* knot_rows = nullptr;
  void* kslots = nullptr;               // B-splines, sorted knots: grid-indexed copies of the knot rows (resolve.hpp)
  void* pyramid = nullptr;              // all pivot levels, top first, one allocation
  void* buckets = nullptr;              // bucket index of the sorted search (resolve.hpp)
  uint32_t lvl_off[MAX_LEVELS] = {};    // element offsets of the levels inside it
  uint32_t lvl_last[MAX_LEVELS] = {};   // last node index of each level
  int n_lvl = 0;
  uint32_t top_bytes = 0;               // size of levels [0, n_lvl-1) when they are worth staging in smem
  int n_top = 0;
  unsigned long long* invalid = nullptr;
};

struct PipePool {
  std::mutex mu;
  std::vector<std::vector<HostPipe*>> by_device;
};
static PipePool& pipe_pool() {
  static PipePool* pool = new PipePool();  // lives until process exit (no static-destruction order games with CUDA)
  return *pool;
}
static HostPipe* acquire_pipe(int device) {
  PipePool& pool = pipe_pool();
  std::lock_guard<std::mutex> lk(pool.mu);
  if ((int)pool.by_device.size() <= device) pool.by_device.resize(device + 1);
  for (HostPipe* hp : pool.by_device[device])
    if (!hp->busy) {
      hp->busy = true;
      return hp;
    }
  HostPipe* hp = new HostPipe();
  hp->busy = true;
  pool.by_device[device].push_back(hp);
  return hp;
}
static void release_pipe(HostPipe* hp) {
  std::lock_guard<std::mutex> lk(pipe_pool().mu);
  hp->busy = false;
}
static void free_buf(HostBuf& b, bool host) {
  if (b.p) {
    if (host) cudaFreeHost(b.p); else cudaFree(b.p);
  }
  b.p = nullptr;
  b.bytes = 0;
}
// Frees the idle pipes of `device` (all devices when device < 0); returns how many were freed.
static int release_idle_pipes(int device) {
  PipePool& pool = pipe_pool();
  std::lock_guard<std::mutex> lk(pool.mu);
  int freed = 0;
  for (int d = 0; d < (int)pool.by_device.size(); ++d) {
    if (device >= 0 && d != device) continue;
    std::vector<HostPipe*>& v = pool.by_device[d];
    DeviceGuard g(d);
    for (size_t i = 0; i < v.size();) {
      HostPipe* hp = v[i];
      if (hp->busy) {
        ++i;
        continue;
      }
      for (int b = 0; b < PLAIN_SLOTS; ++b) {
        free_buf(hp->dev_in[b], false);
        free_buf(hp->dev_out[b], false);
        free_buf(hp->stage_in[b], true);
        free_buf(hp->stage_out[b], true);
        if (hp->stream[b]) cudaStreamDestroy(hp->stream[b]);
        if (hp->done[b]) cudaEventDestroy(hp->done[b]);
      }
      for (int b = 0; b < 2; ++b) {
        free_buf(hp->rle_stage[b], true);
        free_buf(hp->rle_t[b], false);
        free_buf(hp->rle_v[b], false);
        if (hp->rstream[b]) cudaStreamDestroy(hp->rstream[b]);
        if (hp->rdone[b]) cudaEventDestroy(hp->rdone[b]);
      }
      delete hp;
      v.erase(v.begin() + i);
      ++freed;
    }
  }
  return freed;
}

// ---- run-length host transfer ------------------------------------------------------------------------------------
// An f32 take beyond index 2^24 repeats every parameter in runs of 2^ls samples (rows.cuh RUNS). For host-buffer calls
// only the DISTINCT values cross PCIe; the host replicates them into the caller's buffer (plain copies: no curve
// arithmetic runs on the CPU). The pool below does that replication with a few threads and streaming stores, so the
// end-to-end rate is bounded by host memory bandwidth instead of the PCIe D2H rate.
static std::atomic<unsigned long long> g_h2d_bytes{0}, g_d2h_bytes{0};

// Ablation switch (env ENTERP_NO_RLE=1): host-buffer takes always copy every sample over PCIe.
static const bool g_no_rle = [] {
  const char* v = std::getenv("ENTERP_NO_RLE");
  return v && v[0] == '1';
}();

// Ablation switch (env ENTERP_RLE_ONLY=1): eligible chunks all take the run-length path (no plain chunks beside them).
static const bool g_rle_only = [] {
  const char* v = std::getenv("ENTERP_RLE_ONLY");
  return v && v[0] == '1';
}();

class FillPool {
 public:
  static FillPool& get() {
    static FillPool* p = new FillPool();  // lives until process exit
    return *p;
  }
  // fn(lo, hi) over [0, total) in blocks of `block`; the caller works too. One run at a time. `between` (may be
  // empty) is called by the CALLING thread after every block it fills and while it waits for the workers: the host
  // pipeline uses it to keep the copy engine fed during a fill.
  void run(uint64_t total, uint64_t block, const std::function<void(uint64_t, uint64_t)>& fn,
           const std::function<void()>& between = std::function<void()>()) {
    std::lock_guard<std::mutex> serial(run_mu_);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      total_ = total;
      block_ = block;
      next_.store(0, std::memory_order_relaxed);
      pending_ = (int)workers_.size();
      ++gen_;
    }
    cv_.notify_all();
    for (;;) {
      const uint64_t lo = next_.fetch_add(block_, std::memory_order_relaxed);
      if (lo >= total_) break;
      fn(lo, lo + block_ < total_ ? lo + block_ : total_);
      if (between) between();
    }
    std::unique_lock<std::mutex> lk(mu_);
    while (!done_cv_.wait_for(lk, std::chrono::microseconds(50), [&] { return pending_ == 0; })) {
      if (between) {
        lk.unlock();
        between();
        lk.lock();
      }
    }
    fn_ = nullptr;
  }

 private:
  FillPool() {
    int n = 0;
    if (const char* e = std::getenv("ENTERP_HOST_THREADS")) n = std::atoi(e);
    if (n <= 0) {
      cpu_set_t set;
      CPU_ZERO(&set);
      n = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
      if (n > 16) n = 16;
    }
    if (n < 1) n = 1;
    for (int i = 1; i < n; ++i) workers_.emplace_back([this] { loop(); });
    for (std::thread& t : workers_) t.detach();
  }
  void work() {
    for (;;) {
      const uint64_t lo = next_.fetch_add(block_, std::memory_order_relaxed);
      if (lo >= total_) break;
      const uint64_t hi = lo + block_ < total_ ? lo + block_ : total_;
      (*fn_)(lo, hi);
    }
  }
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
      }
      work();
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_cv_.notify_all();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(uint64_t, uint64_t)>* fn_ = nullptr;
  std::atomic<uint64_t> next_{0};
  uint64_t total_ = 0, block_ = 1;
  int pending_ = 0;
  unsigned long long gen_ = 0;
};

// Index of the distinct value sample i reads: RN-even((i - vb) / 2^ls) — what from_usize(i) rounds to, in units of
// 2^ls above vb (the same integer arithmetic as rows.cuh RUNS; tests/test_stepper_runs.py restates it).
static inline uint64_t run_value_index(uint64_t i, uint64_t vb, unsigned ls) {
  const uint64_t dd = i - vb;
  const uint64_t hm1 = ((1ull << ls) >> 1) - 1ull;
  const uint64_t par = ((dd >> ls) + (vb >> ls)) & 1ull;
  return (dd + hm1 + par) >> ls;
}

// out[(i - a) * BYTES ..] = vals[m(i) * BYTES ..] for i in [a, b): walks the runs (value m ends at
// vb + m * 2^ls + 2^(ls-1), plus the tie index when the mantissa (vb >> ls) + m is even).
template <int BYTES>
static void expand_runs(unsigned char* out, const unsigned char* vals, uint64_t a, uint64_t b, uint64_t vb, unsigned ls) {
  const uint64_t half = (1ull << ls) >> 1, vq = vb >> ls;
  uint64_t i = a, m = run_value_index(a, vb, ls);
  while (i < b) {
    uint64_t end = vb + (m << ls) + half + (((m + vq) & 1ull) ? 0ull : 1ull);
    if (end > b) end = b;
    const unsigned char* v = vals + m * BYTES;
    unsigned char* o = out + (i - a) * BYTES;
    const uint64_t n = end - i;
#if defined(__SSE2__)
    if (BYTES == 16 && (reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
      const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i*>(v));
      __m128i* p = reinterpret_cast<__m128i*>(o);
      for (uint64_t k = 0; k < n; ++k) _mm_stream_si128(p + k, x);  // streaming: the buffer is written once, never read here
    } else
#endif
    {
      for (uint64_t k = 0; k < n; ++k) std::memcpy(o + k * BYTES, v, BYTES);
    }
    i = end;
    ++m;
  }
#if defined(__SSE2__)
  _mm_sfence();
#endif
}

static void expand_runs_any(size_t bytes, unsigned char* out, const unsigned char* vals, uint64_t a, uint64_t b, uint64_t vb,
                            unsigned ls) {
  switch (bytes) {
    case 4: expand_runs<4>(out, vals, a, b, vb, ls); break;
    case 8: expand_runs<8>(out, vals, a, b, vb, ls); break;
    case 12: expand_runs<12>(out, vals, a, b, vb, ls); break;
    default: expand_runs<16>(out, vals, a, b, vb, ls); break;
  }
}

}  // namespace enterp

using namespace enterp;

// Tables + device replicas, shared by a curve and the handles derived from it (clamp / slice).
struct CurveCore {
  Resolved r;
  std::mutex mu;
  std::vector<Replica*> replicas;  // index = device ordinal
  ~CurveCore() {
    for (Replica* rp : replicas) {
      if (!rp) continue;
      DeviceGuard g(rp->device);
      cudaFree(rp->block);
      delete rp;
    }
  }
};

// One input adaptor (base/adaptors.rs): kind 0 = Clamp(a = lo, b = hi), 1 = affine t * a + b.
struct PreOp {
  int kind;
  double a, b;  // exact R values widened to double
};

struct enterp_curve {
  std::shared_ptr<CurveCore> core;
  std::vector<PreOp> pre;  // outermost adaptor first; a Bezier .domain(a,b) transform is the innermost
};

namespace {

size_t rsize(const Resolved& r) { return r.scalar == ENTERP_F32 ? 4 : 8; }

enterp_status upload_locked(enterp_curve* c, int device, Replica** out) {
  if (device < 0) return ENTERP_NO_DEVICE;
  int n = 0;
  CK(cudaGetDeviceCount(&n), "cudaGetDeviceCount");
  if (device >= n) return ENTERP_NO_DEVICE;
  if ((int)c->core->replicas.size() <= device) c->core->replicas.resize(device + 1, nullptr);
  if (c->core->replicas[device]) {
    *out = c->core->replicas[device];
    return ENTERP_OK;
  }
  DeviceGuard g(device);
  if (!g.ok) return ENTERP_NO_DEVICE;
  Replica* rp = new Replica();
  rp->device = device;
  // One allocation for all tables (256-byte aligned pieces) + the invalid counter: a build() + first eval costs one
  // cudaMalloc, and destroying the curve one cudaFree (each of those synchronises the device; the end-to-end step of
  // bench.py builds a fresh curve every time).
  const Resolved& r = c->core->r;
  struct Piece {
    const void* src;
    size_t bytes;
    void** dst;
    size_t off;
  };
  std::vector<Piece> pieces;
  auto want = [&](const void* src, size_t bytes, void** dst) {
    *dst = nullptr;
    if (bytes) pieces.push_back(Piece{src, bytes, dst, 0});
  };
  want(r.store.data(), r.store.size(), &rp->store);
  want(r.elems.data(), r.elems.size(), &rp->elems);
  if (r.rows_mode) want(r.rows.data(), r.rows.size(), &rp->rows);
  if (r.rec_stride && !g_no_records) want(r.records.data(), r.records.size(), &rp->records);
  if (r.krow_stride && !g_no_records) want(r.knot_rows.data(), r.knot_rows.size(), &rp->knot_rows);
  if (r.n_slots && !g_no_records && !g_no_buckets && !g_no_slots) want(r.slots.data(), r.slots.size(), &rp->slots);
  if (r.n_kslots && !g_no_records && !g_no_buckets && !g_no_slots) want(r.kslots.data(), r.kslots.size(), &rp->kslots);
  rp->n_lvl = (int)r.levels.level.size();
  std::vector<unsigned char> all;  // the pivot levels, top first, contiguous
  if (rp->n_lvl > 0) {
    const size_t rs = rsize(r);
    for (int l = 0; l < rp->n_lvl; ++l) {
      rp->lvl_off[l] = (uint32_t)(all.size() / rs);
      rp->lvl_last[l] = (uint32_t)(r.levels.level[l].size() / 32) - 1;  // nodes are 32-byte sectors
      all.insert(all.end(), r.levels.level[l].begin(), r.levels.level[l].end());
    }
    want(all.data(), all.size(), &rp->pyramid);
    // Stage as many top levels as fit the per-CTA budget in shared memory (two fat CTAs per SM). Small
    // trees stay in L1 on their own; only trees whose upper levels exceed ~8 KB are worth staging.
    int n_top = 0;
    for (int l = 0; l + 1 < rp->n_lvl; ++l)
      if ((size_t)rp->lvl_off[l + 1] * rs <= SEARCH_SMEM_BUDGET) n_top = l + 1;
    const size_t top = n_top ? (size_t)rp->lvl_off[n_top] * rs : 0;
    if (top >= 8u * 1024u && !g_no_smem_pivots) {
      rp->top_bytes = (uint32_t)top;
      rp->n_top = n_top;
    }
  }
  if (!r.buckets.empty() && !g_no_buckets) {
    want(r.buckets.data(), r.buckets.size() * sizeof(uint32_t), &rp->buckets);
    // with the bucket index the tree is only a fallback for crowded buckets: do not spend shared memory on it
    rp->top_bytes = 0;
    rp->n_top = 0;
  }
  size_t total = 256;  // [0, 8): the invalid counter
  for (Piece& pc : pieces) {
    pc.off = total;
    total += (pc.bytes + 255) / 256 * 256;
  }
  cudaError_t e = cudaMalloc(&rp->block, total);
  if (e == cudaSuccess) {
    unsigned char* base = static_cast<unsigned char*>(rp->block);
    if (total <= (1u << 20)) {  // small curves: one staged copy instead of one per table
      std::vector<unsigned char> stage(total, 0);
      for (const Piece& pc : pieces) std::memcpy(stage.data() + pc.off, pc.src, pc.bytes);
      e = cudaMemcpy(base, stage.data(), total, cudaMemcpyHostToDevice);
    } else {
      e = cudaMemset(base, 0, 256);
      for (const Piece& pc : pieces)
        if (e == cudaSuccess) e = cudaMemcpy(base + pc.off, pc.src, pc.bytes, cudaMemcpyHostToDevice);
    }
    for (const Piece& pc : pieces) *pc.dst = base + pc.off;
    rp->invalid = reinterpret_cast<unsigned long long*>(base);
  }
  if (e != cudaSuccess) {
    cudaFree(rp->block);
    delete rp;
    return cuda_fail(e, "curve upload");
  }
  c->core->replicas[device] = rp;
  *out = rp;
  return ENTERP_OK;
}

// `stream`: the stream the caller is about to enqueue on. The first use of a curve on a device uploads its tables
// with blocking cudaMalloc/cudaMemcpy calls, which are illegal while that stream is being captured into a CUDA
// graph (they would invalidate the capture): report it instead — enterp_curve_upload() belongs before the capture.
enterp_status get_replica(enterp_curve* c, int device, Replica** out, cudaStream_t stream = nullptr) {
  std::lock_guard<std::mutex> lk(c->core->mu);
  const bool have = device >= 0 && device < (int)c->core->replicas.size() && c->core->replicas[device] != nullptr;
  if (!have && stream != nullptr) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) return ENTERP_NOT_UPLOADED;
    cudaGetLastError();
  }
  return upload_locked(c, device, out);
}

template <class R>
CurveDev<R> make_dev(const enterp_curve& cv, const Replica& rp) {
  const Resolved& r = cv.core->r;
  CurveDev<R> d{};
  const R* store = static_cast<const R*>(rp.store);
  d.vk = store ? store + r.vk_off : nullptr;
  d.ik = store ? store + r.ik_off : nullptr;
  d.elems = static_cast<const R*>(rp.elems);
  d.knot_rows = static_cast<const R*>(rp.knot_rows);
  d.krow_lo = (uint32_t)r.degree;
  d.records = static_cast<const R*>(rp.records);
  d.rec_stride = r.rec_stride;
  d.slots = static_cast<const R*>(rp.slots);
  d.n_slots = rp.slots ? r.n_slots : 0;
  d.slot_inv = (R)r.slot_inv;
  d.kslots = static_cast<const R*>(rp.kslots);
  d.n_kslots = rp.kslots ? r.n_kslots : 0;
  d.kslot_inv = (R)r.kslot_inv;
  if (r.n_store > 0) d.k_last = reinterpret_cast<const R*>(r.store.data())[r.ik_off + r.inner_len - 1];
  d.n_lvl = rp.n_lvl;
  for (int l = 0; l < MAX_LEVELS; ++l) {
    d.lvl_off[l] = rp.lvl_off[l];
    d.lvl_last[l] = rp.lvl_last[l];
    d.lvl[l] = l < rp.n_lvl ? static_cast<const R*>(rp.pyramid) + rp.lvl_off[l] : nullptr;
  }
  d.n_top = rp.n_top;
  d.top_bytes = rp.top_bytes;
  d.bkt = static_cast<const uint32_t*>(rp.buckets);
  d.n_bkt = rp.buckets ? r.n_buckets : 0;
  d.bkt_first = (R)r.bkt_first;
  d.bkt_inv = (R)r.bkt_inv;
  d.n_elem = (uint32_t)r.n_elem;
  d.inner_len = (uint32_t)r.inner_len;
  d.virt_len = (uint32_t)r.virt_len;
  d.degree = (uint32_t)r.degree;
  d.imin = (uint32_t)r.imin;
  d.imax = (uint32_t)r.imax;
  d.map0 = (uint32_t)r.map0;
  d.mapL = (uint32_t)r.mapL;
  d.map_add = (int32_t)r.map_add;
  d.eq_step = (R)r.eq_step;
  d.eq_offset = (R)r.eq_offset;
  d.eq_first = (R)r.eq_first;
  d.eq_lenm2 = r.inner_len >= 2 ? (R)(r.inner_len - 2) : R(0);  // from_usize(len - 2), list.rs:546
  d.eq_const = r.knot_base == ENTERP_KNOTS_CONST_EQUIDISTANT ? 1 : 0;
  d.eq_mode = r.eq_mode;
  d.eq_rstep = (R)r.eq_rstep;
  if (g_ballot_search && r.knot_base == ENTERP_KNOTS_SORTED) {
    d.ballot = 1;  // every structure in front of the plain knot array is switched off
    d.records = nullptr;
    d.slots = nullptr;
    d.kslots = nullptr;
    d.bkt = nullptr;
    d.top_bytes = 0;
    d.n_top = 0;
  }
  d.n_pre = (int)cv.pre.size();
  for (int i = 0; i < d.n_pre && i < MAX_PRE; ++i) {
    d.pre_kind[i] = cv.pre[i].kind;
    d.pre_a[i] = (R)cv.pre[i].a;
    d.pre_b[i] = (R)cv.pre[i].b;
  }
  d.easing = r.easing;
  d.easing_n = r.easing_n;
  d.ease_min = (R)r.ease_min;
  d.ease_max = (R)r.ease_max;
  d.invalid = rp.invalid;
  return d;
}

// Stepper::new(n_total, start, end) -> Equidistant::new (list.rs:393-399): step in R arithmetic.
template <class R>
SampleSrc<R> take_src(const Resolved& r, uint64_t n_total, uint64_t first, uint64_t count) {
  SampleSrc<R> s{};
  const R start = (R)r.domain[0], end = (R)r.domain[1];
  const R span = end - start;
  const R denom = (R)(n_total - 1);
  s.t = nullptr;
  s.step = span / denom;
  s.start = start;
  s.first = first;
  s.count = count;
  return s;
}
template <class R>
SampleSrc<R> batch_src(const void* t, uint64_t n) {
  SampleSrc<R> s{};
  s.t = static_cast<const R*>(t);
  s.step = R(0);
  s.start = R(0);
  s.first = 0;
  s.count = n;
  return s;
}

template <class R>
enterp_status dispatch_eval(const enterp_curve& cv, const Replica& rp, const SampleSrc<R>& s, void* out, cudaStream_t st);

// Ablation switch (env ENTERP_NO_GENERIC_RUNS=1): f32 takes beyond 2^26 through the generic kernels sample by sample.
static const bool g_no_generic_runs = [] {
  const char* v = std::getenv("ENTERP_NO_GENERIC_RUNS");
  return v && v[0] == '1';
}();
constexpr uint64_t GENERIC_RUNS_FROM = 1ull << 24;   // where the runs begin (2 samples each up to 2^25)
constexpr uint64_t GENERIC_RUNS_VALUES = 1ull << 22;  // distinct values per pass: <= 80 MB of scratch, L2-resident values

// Stepper runs for every take the span-row kernels do not serve (Linear / Bezier outside the lean kernels, B-splines
// of degree > ROWS_MAX_DEGREE or with runtime degree, ConstEquidistant knots, ...): beyond index 2^24 an f32 stepper
// repeats its parameter in runs (rows.cuh RUNS), so per binade and per pass of <= 2^22 values the
// distinct stepper values are generated (stepper_strided_kernel), evaluated ONCE by whatever batch kernel the curve
// has, and expanded into the output at the store roof (runs_expand_kernel). Scratch comes from the stream-ordered
// allocator (cudaMallocAsync on the caller's stream). Only for curves without input adaptors whose in-domain
// parameters cannot make the reference panic (domain_cannot_panic below): the invalid counter then needs no per-sample
// accounting.
// The device's default stream-ordered pool returns freed memory to the OS at the next synchronisation unless told to
// keep some: let it keep the scratch of one pass, so that repeated takes do not re-allocate it every call.
static void keep_scratch_in_pool(int device) {
  static std::mutex mu;
  static std::vector<char> done;
  std::lock_guard<std::mutex> lk(mu);
  if ((int)done.size() <= device) done.resize(device + 1, 0);
  if (done[device]) return;
  done[device] = 1;
  cudaMemPool_t pool = nullptr;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
    uint64_t keep = 0;
    cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    const uint64_t want = 128ull << 20;
    if (keep < want) {
      keep = want;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  cudaGetLastError();
}

// Can a parameter inside the curve's (finite) domain make the reference panic? Only equidistant knots can
// (`to_usize().unwrap()`, list.rs:456,486,539-540; a B-spline index below the lower cut, bspline/mod.rs:131), and only
// degenerate ones: a zero, negative or non-finite step (the builders accept domain(a, a) and domain(a, b) with a > b), or
// a step so small against the magnitude of the knots that they collapse onto each other. The run paths evaluate VALUES,
// not samples, so they leave such curves to the per-sample kernels and their counter. Sufficient condition (f32, the
// only scalar with runs): step >= 2^-20 * max(|offset|, |domain|) keeps the rounding of (x - offset) / step below 1/8
// for every parameter within a few ulps of the domain, so the quotient stays inside [min - 1/8, len + 1/8].
static bool domain_cannot_panic(const Resolved& r) {
  if (!std::isfinite(r.domain[0]) || !std::isfinite(r.domain[1])) return false;
  if (r.kind == ENTERP_BEZIER || r.knot_base != ENTERP_KNOTS_EQUIDISTANT) return true;
  const double step = r.eq_step;
  if (!(step > 0) || !std::isfinite(step) || !std::isfinite(r.eq_offset)) return false;
  const double m = std::fmax(std::fabs(r.eq_offset), std::fmax(std::fabs(r.domain[0]), std::fabs(r.domain[1])));
  return step >= 1.17549435e-38 && step >= m * (1.0 / 1048576.0);
}

enterp_status generic_runs_take(const enterp_curve& cv, const Replica& rp, const SampleSrc<float>& s, void* out, cudaStream_t st) {
  const Resolved& r = cv.core->r;
  keep_scratch_in_pool(rp.device);
  const size_t stride = sizeof(float) * (size_t)r.dim;
  const uint64_t A = s.first, B = s.first + s.count;
  uint64_t pos = A;
  if (pos < GENERIC_RUNS_FROM) {  // the head keeps the per-sample kernels
    SampleSrc<float> head = s;
    head.count = GENERIC_RUNS_FROM - A;
    const enterp_status es = dispatch_eval<float>(cv, rp, head, out, st);
    if (es) return es;
    pos = GENERIC_RUNS_FROM;
  }
  while (pos < B) {
    const int hb = 63 - __builtin_clzll(pos);
    const unsigned ls = (unsigned)(hb - 23);
    const uint64_t binade_end = hb >= 63 ? B : (1ull << (hb + 1));
    uint64_t end = B < binade_end ? B : binade_end;
    if (end - pos > (GENERIC_RUNS_VALUES << ls)) end = pos + (GENERIC_RUNS_VALUES << ls);
    const uint64_t vb = (pos >> ls) << ls;
    const uint64_t nvals = ((end - 1 - vb) >> ls) + 2;
    void* scratch = nullptr;
    const size_t t_bytes = (nvals * sizeof(float) + 255) / 256 * 256;
    CK(cudaMallocAsync(&scratch, t_bytes + nvals * stride, st), "cudaMallocAsync");
    float* t = static_cast<float*>(scratch);
    void* vals = static_cast<char*>(scratch) + t_bytes;
    SampleSrc<float> ts = s;  // same step / start, indices in units of 2^ls
    ts.first = vb >> ls;
    ts.count = nvals;
    cudaError_t e = launch_stepper_strided<float>(ts, ls, t, st);
    enterp_status es = e == cudaSuccess ? dispatch_eval<float>(cv, rp, batch_src<float>(t, nvals), vals, st) : cuda_fail(e, "stepper launch");
    if (es == ENTERP_OK) {
      e = launch_runs_expand((int)r.dim, vals, static_cast<char*>(out) + (pos - A) * stride, pos, end - pos, vb, ls, st);
      if (e != cudaSuccess) es = cuda_fail(e, "runs expand launch");
    }
    cudaFreeAsync(scratch, st);
    if (es) return es;
    pos = end;
  }
  return ENTERP_OK;
}

template <class R>
enterp_status dispatch_eval(const enterp_curve& cv, const Replica& rp, const SampleSrc<R>& s, void* out, cudaStream_t st) {
  const Resolved& r = cv.core->r;
  // The lean Linear / Bezier take kernels index with 32 bits and stop at 2^32 - 2^24 (inst_lean.cuh): a take that
  // crosses that index (take(2^32) itself) is cut there, so that all but its last 2^24 samples keep the lean kernel
  constexpr uint64_t LEAN_SPLIT = (1ull << 32) - (1ull << 24);
  if (s.t == nullptr && r.kind != ENTERP_BSPLINE && !g_disable_lean && s.first < LEAN_SPLIT && s.count > LEAN_SPLIT - s.first) {
    SampleSrc<R> a = s, b = s;
    a.count = LEAN_SPLIT - s.first;
    b.first = LEAN_SPLIT;
    b.count = s.count - a.count;
    const enterp_status sa = dispatch_eval<R>(cv, rp, a, out, st);
    if (sa) return sa;
    return dispatch_eval<R>(cv, rp, b, static_cast<R*>(out) + a.count * (uint64_t)r.dim, st);
  }
  const CurveDev<R> d = make_dev<R>(cv, rp);
  const bool vec_ok = (reinterpret_cast<uintptr_t>(out) % 16u) == 0;
  const bool proj = r.weighted != 0;
  const bool equi = r.knot_base != ENTERP_KNOTS_SORTED;
  // a take the generic kernels would evaluate sample by sample although its f32 stepper repeats itself (see above)
  const bool generic_runs = sizeof(R) == 4 && s.t == nullptr && !g_no_generic_runs && cv.pre.empty() && vec_ok &&
                            domain_cannot_panic(r) && s.first + s.count >= GENERIC_RUNS_FROM + (1ull << 16) &&
                            s.first + s.count < (1ull << 62);
  auto take_runs = [&]() -> enterp_status {
    if constexpr (sizeof(R) == 4) return generic_runs_take(cv, rp, s, out, st);
    return ENTERP_UNSUPPORTED;
  };
  cudaError_t e;
  switch (r.kind) {
    case ENTERP_BSPLINE: {
      if (r.degree > (uint64_t)MAX_DYN_DEGREE) return ENTERP_UNSUPPORTED;
      const int deg = r.degree <= 7 ? (int)r.degree : 0;
      // Span-row fast path (rows.cuh). Preconditions checked here: 16-byte aligned output, no reachable
      // map_from special case (adaptors.rs:31-39 needs inner == 0 or inner == inner_len), and for
      // equidistant knots a span division the kernel can do exactly without the IEEE divide.
      const bool map_simple = (r.imin > 0 && r.imax < r.inner_len) || r.adaptor == AD_NONE;
      const bool eq_ok = !equi || (r.rows_mode == 1 ? r.eq_mode == 1 : r.eq_mode >= 1);
      if (r.rows_mode && rp.rows && deg > 0 && vec_ok && map_simple && eq_ok && !g_disable_rows && !g_ballot_search) {
        RowsDev<R> rw{};
        rw.table = static_cast<const R*>(rp.rows);
        rw.n_rows = r.rows_n;
        rw.lo = r.rows_lo;
        rw.bytes = (uint32_t)r.rows.size();
        rw.one = R(1);
        rw.eq_rstep = (R)r.eq_rstep;
        rw.ri_off = (uint32_t)((int64_t)r.map_add - (int64_t)r.rows_lo);
        rw.pow2_ops = r.rows_pow2_ops;
        e = launch_bspline_rows<R>(r.dh, proj, deg, equi, r.rows_mode == 1, d, s, static_cast<R*>(out), rw, st);
        if (e == cudaErrorNotSupported) {  // table too big for shared memory and not a plain take(): generic kernel
          if (generic_runs) return take_runs();
          e = launch_bspline<R>(r.dh, proj, deg, equi, d, s, static_cast<R*>(out), vec_ok, st);
        }
      } else {
        if (generic_runs) return take_runs();
        e = launch_bspline<R>(r.dh, proj, deg, equi, d, s, static_cast<R*>(out), vec_ok, st);
      }
      break;
    }
    case ENTERP_LINEAR:
      if (!g_disable_lean && !g_ballot_search && lean_linear<R>(r.dh, proj, equi, d, s, static_cast<R*>(out), vec_ok, st, &e)) break;
      if (generic_runs) return take_runs();
      e = launch_linear<R>(r.dh, proj, equi, d, s, static_cast<R*>(out), vec_ok, st);
      break;
    default: {
      if (r.n_elem > (uint64_t)MAX_DYN_BEZIER) return ENTERP_UNSUPPORTED;
      const int n = r.n_elem <= 16 ? (int)r.n_elem : 0;
      if (!g_disable_lean && lean_bezier<R>(r.dh, proj, n, reinterpret_cast<const R*>(r.elems.data()), d, s,
                                                 static_cast<R*>(out), vec_ok, st, &e))
        break;
      if (generic_runs) return take_runs();
      e = launch_bezier<R>(r.dh, proj, n, d, s, static_cast<R*>(out), vec_ok, st);
      break;
    }
  }
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  return ENTERP_OK;
}

enterp_status eval_common(enterp_curve* c, int device, bool take, const void* t, uint64_t n_total, uint64_t first,
                          uint64_t count, void* out, void* stream) {
  if (!c) return ENTERP_INVALID_ARGUMENT;
  if (count == 0) return ENTERP_OK;
  if (!out || (!take && !t)) return ENTERP_INVALID_ARGUMENT;
  if (take && (n_total == 0 || first > n_total || count > n_total - first)) return ENTERP_INVALID_ARGUMENT;
  Replica* rp = nullptr;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  enterp_status st = get_replica(c, device, &rp, cs);
  if (st) return st;
  DeviceGuard g(device);
  if (!g.ok) return ENTERP_NO_DEVICE;
  if (c->core->r.scalar == ENTERP_F32) {
    auto s = take ? take_src<float>(c->core->r, n_total, first, count) : batch_src<float>(t, count);
    return dispatch_eval<float>(*c, *rp, s, out, cs);
  }
  auto s = take ? take_src<double>(c->core->r, n_total, first, count) : batch_src<double>(t, count);
  return dispatch_eval<double>(*c, *rp, s, out, cs);
}

enterp_status span_common(enterp_curve* c, int device, bool take, const void* t, uint64_t n_total, uint64_t first,
                          uint64_t count, uint32_t* idx, void* factor, void* stream) {
  if (!c) return ENTERP_INVALID_ARGUMENT;
  if (count == 0) return ENTERP_OK;
  if (!idx || (!take && !t)) return ENTERP_INVALID_ARGUMENT;
  if (take && (n_total == 0 || first > n_total || count > n_total - first)) return ENTERP_INVALID_ARGUMENT;
  Replica* rp = nullptr;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  enterp_status st = get_replica(c, device, &rp, cs);
  if (st) return st;
  DeviceGuard g(device);
  if (!g.ok) return ENTERP_NO_DEVICE;
  const bool equi = c->core->r.knot_base != ENTERP_KNOTS_SORTED;
  cudaError_t e;
  if (c->core->r.scalar == ENTERP_F32) {
    auto s = take ? take_src<float>(c->core->r, n_total, first, count) : batch_src<float>(t, count);
    e = launch_span<float>(c->core->r.kind, equi, make_dev<float>(*c, *rp), s, idx, static_cast<float*>(factor), cs);
  } else {
    auto s = take ? take_src<double>(c->core->r, n_total, first, count) : batch_src<double>(t, count);
    e = launch_span<double>(c->core->r.kind, equi, make_dev<double>(*c, *rp), s, idx, static_cast<double*>(factor), cs);
  }
  if (e != cudaSuccess) return cuda_fail(e, "span kernel launch");
  return ENTERP_OK;
}

// ---- host-buffer pipelines ---------------------------------------------------------------------
constexpr uint64_t HOST_CHUNK = 1ull << 22;  // samples per pipeline stage

bool is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

cudaError_t ensure(HostBuf& b, size_t need, bool host) {
  if (b.bytes >= need && b.p) return cudaSuccess;
  free_buf(b, host);  // resets the recorded size together with the pointer
  const cudaError_t e = host ? cudaMallocHost(&b.p, need) : cudaMalloc(&b.p, need);
  if (e == cudaSuccess) {
    b.bytes = need;
  } else {
    b.p = nullptr;
    b.bytes = 0;
  }
  return e;
}

// One device's share of a host-buffer call: samples [off0, off0 + count) of the caller's arrays (take mode: global
// stepper indices first + off0 ...), pushed through a borrowed pipe chunk by chunk. Two double-buffered pipelines work
// on the share's chunks: the PLAIN one (kernel over every sample, D2H of every sample: bound by the PCIe copy engine,
// no CPU time when the caller's buffer is pinned) and, for ranges where an f32 stepper repeats its parameter, the
// RUN-LENGTH one (kernel and D2H over the distinct values only, the host replicates them: bound by host memory
// bandwidth). Chunks are claimed dynamically: whenever a plain slot is free the copy engine gets the next chunk,
// otherwise the CPU path takes it, so the call runs at (PCIe rate + host fill rate) and never below the PCIe rate,
// whatever the host's cores are doing.
struct HostLane {
  int device = 0;
  Replica* rp = nullptr;
  HostPipe* hp = nullptr;
  uint64_t off0 = 0, count = 0, chunk = 0, next = 0, k = 0, kr = 0;
  enterp_status err = ENTERP_OK;  // first failure of a chunk issued from inside a host fill (j.feed)
  // Adaptive choice between the hybrid and the plain pipeline: claims are paced by completions (the slots are few), so
  // the bytes claimed per window of HOST_WINDOW chunks over the wall-clock time of the window is the current rate.
  std::chrono::steady_clock::time_point win_t0;
  uint64_t win_bytes = 0;
  int win_n = 0;
  bool plain_mode = false;  // eligible chunks take the plain pipeline for now
  int plain_windows = 0;    // ... for this many more windows
  struct Pending {
    bool live = false;
    uint64_t off = 0, n = 0;
    uint64_t vb = 0;      // run-length chunks: representable index at or below the chunk's first sample
    unsigned ls = 0;      // log2 of the run length (fl() spacing of the chunk's binade)
  } pend[PLAIN_SLOTS], rpend[2];
};

struct HostJob {
  enterp_curve* c;
  bool take;
  const void* t_host;
  uint64_t n_total, first;
  void* out_host;
  size_t rs, out_stride;
  bool pin_in, pin_out;
  bool rle;     // f32 take of a curve without input adaptors over a finite domain: chunks beyond 2^26 may ship runs only
  bool hybrid;  // ... and the copy engine takes plain chunks beside them
  std::function<void()> feed;  // hands every idle plain slot (of every lane) its next chunk; called during host fills
};

constexpr unsigned RLE_MIN_SHIFT = 2;  // runs of at least 4 samples (indices from 2^26 on)

cudaError_t lane_prepare(const HostJob& j, HostLane& l) {
  HostPipe& hp = *l.hp;
  cudaError_t e;
  for (int b = 0; b < PLAIN_SLOTS; ++b) {
    if (!hp.stream[b] && (e = cudaStreamCreateWithFlags(&hp.stream[b], cudaStreamNonBlocking)) != cudaSuccess) return e;
    if (!hp.done[b] && (e = cudaEventCreateWithFlags(&hp.done[b], cudaEventDisableTiming)) != cudaSuccess) return e;
    if (!j.take && (e = ensure(hp.dev_in[b], l.chunk * j.rs, false)) != cudaSuccess) return e;
    if ((e = ensure(hp.dev_out[b], l.chunk * j.out_stride, false)) != cudaSuccess) return e;
    if (!j.take && !j.pin_in && (e = ensure(hp.stage_in[b], l.chunk * j.rs, true)) != cudaSuccess) return e;
    if (!j.pin_out && (e = ensure(hp.stage_out[b], l.chunk * j.out_stride, true)) != cudaSuccess) return e;
  }
  if (j.rle) {
    const size_t max_vals = (size_t)(l.chunk >> RLE_MIN_SHIFT) + 2;
    for (int b = 0; b < 2; ++b) {
      if (!hp.rstream[b] && (e = cudaStreamCreateWithFlags(&hp.rstream[b], cudaStreamNonBlocking)) != cudaSuccess) return e;
      if (!hp.rdone[b] && (e = cudaEventCreateWithFlags(&hp.rdone[b], cudaEventDisableTiming)) != cudaSuccess) return e;
      if ((e = ensure(hp.rle_t[b], max_vals * j.rs, false)) != cudaSuccess) return e;
      if ((e = ensure(hp.rle_v[b], max_vals * j.out_stride, false)) != cudaSuccess) return e;
      if ((e = ensure(hp.rle_stage[b], max_vals * j.out_stride, true)) != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

// Completes plain slot b (when `wait`, blocks until its copy has landed). *done = the slot is free afterwards.
cudaError_t plain_finish(const HostJob& j, HostLane& l, int b, bool wait, bool* done) {
  *done = true;
  if (!l.pend[b].live) return cudaSuccess;
  cudaError_t e = wait ? cudaEventSynchronize(l.hp->done[b]) : cudaEventQuery(l.hp->done[b]);
  if (e == cudaErrorNotReady) {
    *done = false;
    return cudaSuccess;
  }
  if (e != cudaSuccess) return e;
  if (!j.pin_out)
    std::memcpy(static_cast<char*>(j.out_host) + (l.off0 + l.pend[b].off) * j.out_stride, l.hp->stage_out[b].p,
                l.pend[b].n * j.out_stride);
  l.pend[b].live = false;
  return cudaSuccess;
}

// Completes run-length slot b: waits for its (small) copy and replicates the distinct values into the caller's
// buffer (pinned or pageable alike), a few threads wide.
cudaError_t rle_finish(const HostJob& j, HostLane& l, int b) {
  if (!l.rpend[b].live) return cudaSuccess;
  const cudaError_t e = cudaEventSynchronize(l.hp->rdone[b]);
  if (e != cudaSuccess) return e;
  const HostLane::Pending pd = l.rpend[b];
  const uint64_t a = j.first + l.off0 + pd.off;
  unsigned char* dst = static_cast<unsigned char*>(j.out_host) + (l.off0 + pd.off) * j.out_stride;
  const unsigned char* vals = static_cast<const unsigned char*>(l.hp->rle_stage[b].p);
  const size_t stride = j.out_stride;
  l.rpend[b].live = false;
  FillPool::get().run(
      pd.n, 1u << 16,
      [&](uint64_t lo, uint64_t hi) { expand_runs_any(stride, dst + lo * stride, vals, a + lo, a + hi, pd.vb, pd.ls); }, j.feed);
  return cudaSuccess;
}

// All of chunk [a, a + n) inside one binade of the stepper index, runs of at least 2^RLE_MIN_SHIFT samples?
bool rle_eligible(const HostJob& j, uint64_t a, uint64_t n, unsigned* ls) {
  if (!j.rle || n < 4096) return false;
  const int hb = 63 - __builtin_clzll(a | 1ull);
  if (hb < 24 + (int)RLE_MIN_SHIFT || hb != 63 - __builtin_clzll((a + n - 1) | 1ull)) return false;
  *ls = (unsigned)(hb - 23);
  return true;
}

// Plain chunk [off, off + n) of the lane on slot b (free): H2D of its parameters (batches), kernel, D2H.
enterp_status plain_issue(const HostJob& j, HostLane& l, int b, uint64_t off, uint64_t n) {
  HostPipe& hp = *l.hp;
  const Resolved& r = j.c->core->r;
  const void* t_dev = nullptr;
  if (!j.take) {
    const char* src = static_cast<const char*>(j.t_host) + (l.off0 + off) * j.rs;
    if (!j.pin_in) {
      std::memcpy(hp.stage_in[b].p, src, n * j.rs);
      src = static_cast<const char*>(hp.stage_in[b].p);
    }
    CK(cudaMemcpyAsync(hp.dev_in[b].p, src, n * j.rs, cudaMemcpyHostToDevice, hp.stream[b]), "H2D");
    g_h2d_bytes.fetch_add(n * j.rs, std::memory_order_relaxed);
    t_dev = hp.dev_in[b].p;
  }
  enterp_status es;
  if (r.scalar == ENTERP_F32) {
    auto s = j.take ? take_src<float>(r, j.n_total, j.first + l.off0 + off, n) : batch_src<float>(t_dev, n);
    es = dispatch_eval<float>(*j.c, *l.rp, s, hp.dev_out[b].p, hp.stream[b]);
  } else {
    auto s = j.take ? take_src<double>(r, j.n_total, j.first + l.off0 + off, n) : batch_src<double>(t_dev, n);
    es = dispatch_eval<double>(*j.c, *l.rp, s, hp.dev_out[b].p, hp.stream[b]);
  }
  if (es) return es;
  void* dst = j.pin_out ? static_cast<void*>(static_cast<char*>(j.out_host) + (l.off0 + off) * j.out_stride) : hp.stage_out[b].p;
  CK(cudaMemcpyAsync(dst, hp.dev_out[b].p, n * j.out_stride, cudaMemcpyDeviceToHost, hp.stream[b]), "D2H");
  g_d2h_bytes.fetch_add(n * j.out_stride, std::memory_order_relaxed);
  CK(cudaEventRecord(hp.done[b], hp.stream[b]), "cudaEventRecord");
  l.pend[b].live = true;
  l.pend[b].off = off;
  l.pend[b].n = n;
  return ENTERP_OK;
}

// Run-length chunk on slot b (free): the stepper values at the representable indices, their evaluation as a batch,
// and the D2H of those values only.
enterp_status rle_issue(const HostJob& j, HostLane& l, int b, uint64_t off, uint64_t n, unsigned ls) {
  HostPipe& hp = *l.hp;
  const Resolved& r = j.c->core->r;
  const uint64_t a = j.first + l.off0 + off;
  const uint64_t vb = (a >> ls) << ls;
  const uint64_t nvals = ((a + n - 1 - vb) >> ls) + 2;
  SampleSrc<float> ts = take_src<float>(r, j.n_total, vb >> ls, nvals);  // first/count in units of 2^ls
  const cudaError_t e = launch_stepper_strided<float>(ts, ls, static_cast<float*>(hp.rle_t[b].p), hp.rstream[b]);
  if (e != cudaSuccess) return cuda_fail(e, "stepper launch");
  const enterp_status es = dispatch_eval<float>(*j.c, *l.rp, batch_src<float>(hp.rle_t[b].p, nvals), hp.rle_v[b].p, hp.rstream[b]);
  if (es) return es;
  CK(cudaMemcpyAsync(hp.rle_stage[b].p, hp.rle_v[b].p, nvals * j.out_stride, cudaMemcpyDeviceToHost, hp.rstream[b]), "D2H");
  g_d2h_bytes.fetch_add(nvals * j.out_stride, std::memory_order_relaxed);
  CK(cudaEventRecord(hp.rdone[b], hp.rstream[b]), "cudaEventRecord");
  l.rpend[b].live = true;
  l.rpend[b].off = off;
  l.rpend[b].n = n;
  l.rpend[b].vb = vb;
  l.rpend[b].ls = ls;
  return ENTERP_OK;
}

constexpr int HOST_WINDOW = 8;
// Debug switch (env ENTERP_HOST_TRACE=1): one stderr line per window of the host pipeline.
static const bool g_host_trace = [] {
  const char* v = std::getenv("ENTERP_HOST_TRACE");
  return v && v[0] == '1';
}();

// Called for every claimed chunk of a run-length capable call. Hosts differ (and change) a lot in what their cores can
// add to the copy engine's rate — from 3x down to less than nothing when the memory system is busy elsewhere — so the
// lane measures: a plain window gives the reference rate (remembered per pipe), a hybrid window that does not beat it
// sends the next 4 (then 8, 16, 32) windows down the plain pipeline before the hybrid one is tried again.
void lane_account(const HostJob& j, HostLane& l, uint64_t n) {
  if (!j.hybrid) return;
  const auto now = std::chrono::steady_clock::now();
  if (l.win_n == 0) l.win_t0 = now;
  if (l.win_n < HOST_WINDOW) {
    l.win_bytes += n * j.out_stride;
    ++l.win_n;
    return;
  }
  const double secs = std::chrono::duration<double>(now - l.win_t0).count();
  const double gbs = secs > 0 ? (double)l.win_bytes / secs / 1e9 : 0;
  HostPipe& hp = *l.hp;
  if (g_host_trace)
    std::fprintf(stderr, "[enterp host] dev %d window %s %.1f GB/s (plain reference %.1f, hybrid_bad %d)\n", l.device,
                 l.plain_mode ? "plain " : "hybrid", gbs, hp.plain_gbs, hp.hybrid_bad);
  if (l.plain_mode) {
    hp.plain_gbs = gbs;
    if (--l.plain_windows <= 0) l.plain_mode = false;
  } else if (hp.plain_gbs <= 0) {
    l.plain_mode = true;  // no reference yet on this pipe: one plain window
    l.plain_windows = 1;
  } else if (gbs < 1.02 * hp.plain_gbs) {
    l.plain_mode = true;
    l.plain_windows = 4 << hp.hybrid_bad;  // 4, 8, 16, 32 windows: the longer the fills have not paid, the rarer the probes
    if (hp.hybrid_bad < 3) ++hp.hybrid_bad;
  } else {
    hp.hybrid_bad = 0;
  }
  l.win_t0 = now;
  l.win_bytes = n * j.out_stride;
  l.win_n = 1;
}

// Hands the lane's next chunk to the copy engine if one of its plain slots is free (non-blocking).
enterp_status feed_lane(const HostJob& j, HostLane& l) {
  if (!l.hp || l.next >= l.count) return ENTERP_OK;
  DeviceGuard g(l.device);
  if (!g.ok) return ENTERP_NO_DEVICE;
  for (int b = 0; b < PLAIN_SLOTS; ++b) {
    bool done = true;
    CK(plain_finish(j, l, b, false, &done), "pipeline poll");
    if (done) {
      const uint64_t off = l.next;
      const uint64_t n = l.count - off < l.chunk ? l.count - off : l.chunk;
      const enterp_status es = plain_issue(j, l, b, off, n);
      if (es) return es;
      l.next = off + n;
      lane_account(j, l, n);
      return ENTERP_OK;
    }
  }
  return ENTERP_OK;
}

// One unit of work for the lane: claims its next chunk for the copy engine or for the CPU path (see HostLane).
enterp_status lane_step(const HostJob& j, HostLane& l) {
  DeviceGuard g(l.device);
  if (!g.ok) return ENTERP_NO_DEVICE;
  const uint64_t off = l.next;
  const uint64_t n = l.count - off < l.chunk ? l.count - off : l.chunk;
  unsigned ls = 0;
  const bool eligible = !l.plain_mode && rle_eligible(j, j.first + l.off0 + off, n, &ls);
  bool done = true;
  if (!eligible) {  // the classic double-buffered pipeline: wait for the slot used two chunks ago
    if (l.plain_mode) {  // leaving the hybrid pipeline: its pending fills first, so that they do not trail the call
      for (int b = 0; b < 2; ++b) CK(rle_finish(j, l, (int)((l.kr + b) & 1)), "run-length wait");
      if (l.err) return l.err;
      if (l.next != off) return ENTERP_OK;  // the fills' feed claimed chunks meanwhile: look again
    }
    const int b = (int)(l.k % PLAIN_SLOTS);
    CK(plain_finish(j, l, b, true, &done), "pipeline wait");
    const enterp_status es = plain_issue(j, l, b, off, n);
    if (es) return es;
    ++l.k;
    l.next = off + n;
    lane_account(j, l, n);
    return ENTERP_OK;
  }
  if (j.hybrid) {  // a free plain slot means an idle copy engine: feed it first
    const uint64_t before = l.next;
    const enterp_status es = feed_lane(j, l);
    if (es) return es;
    if (l.next != before) return ENTERP_OK;
  }
  const int rb = (int)(l.kr & 1);
  // replicate the chunk issued two run-length steps ago (CPU work; j.feed keeps the copy engine busy meanwhile and
  // may claim chunks, so the next chunk is looked at again afterwards)
  CK(rle_finish(j, l, rb), "run-length wait");
  if (l.err) return l.err;
  if (l.next >= l.count) return ENTERP_OK;
  const uint64_t off2 = l.next;
  const uint64_t n2 = l.count - off2 < l.chunk ? l.count - off2 : l.chunk;
  if (!rle_eligible(j, j.first + l.off0 + off2, n2, &ls)) return ENTERP_OK;  // the next lane_step takes it as a plain chunk
  const enterp_status es = rle_issue(j, l, rb, off2, n2, ls);
  if (es) return es;
  ++l.kr;
  l.next = off2 + n2;
  lane_account(j, l, n2);
  return ENTERP_OK;
}

// `Signal::sample(..).collect()` / `Curve::take(n).collect()` with host buffers over one or several GPUs of the box:
// the sample range is cut into contiguous per-device shares ([g*N/G, (g+1)*N/G), global stepper
// index), each share runs through its own double-buffered pipe, and ONE host thread issues the chunks round-robin
// over the devices. Blocks until out_host is complete.
enterp_status host_common(enterp_curve* c, const int* devices, int n_devices, bool take, const void* t_host,
                          uint64_t n_total, uint64_t first, uint64_t count, void* out_host) {
  if (!c || !devices || n_devices < 1) return ENTERP_INVALID_ARGUMENT;
  if (count == 0) return ENTERP_OK;
  if (!out_host || (!take && !t_host)) return ENTERP_INVALID_ARGUMENT;
  if (take && (n_total == 0 || first > n_total || count > n_total - first)) return ENTERP_INVALID_ARGUMENT;
  for (int a = 0; a < n_devices; ++a)
    for (int b = a + 1; b < n_devices; ++b)
      if (devices[a] == devices[b]) return ENTERP_INVALID_ARGUMENT;
  HostJob j{};
  j.c = c;
  j.take = take;
  j.t_host = t_host;
  j.n_total = n_total;
  j.first = first;
  j.out_host = out_host;
  j.rs = rsize(c->core->r);
  j.out_stride = j.rs * c->core->r.dim;
  j.pin_in = take || is_pinned(t_host);
  j.pin_out = is_pinned(out_host);
  {
    // run-length transfer: f32 takes that reach past 2^26, no input adaptors, finite domain (then every parameter is
    // finite and inside the domain, no sample can make the reference panic, and the per-value batch evaluation
    // leaves the invalid counter where the per-sample take would)
    const Resolved& rr = c->core->r;
    j.rle = take && !g_no_rle && rr.scalar == ENTERP_F32 && c->pre.empty() && domain_cannot_panic(rr) && n_total > (1ull << (24 + RLE_MIN_SHIFT)) && first + count > (1ull << (24 + RLE_MIN_SHIFT));
    j.hybrid = j.rle && !g_rle_only;
  }
  std::vector<HostLane> lanes((size_t)n_devices);
  enterp_status st = ENTERP_OK;
  for (int g = 0; g < n_devices && st == ENTERP_OK; ++g) {
    HostLane& l = lanes[(size_t)g];
    l.device = devices[g];
    const uint64_t lo = (uint64_t)(((unsigned __int128)count * (unsigned)g) / (unsigned)n_devices);
    const uint64_t hi = (uint64_t)(((unsigned __int128)count * (unsigned)(g + 1)) / (unsigned)n_devices);
    l.off0 = lo;
    l.count = hi - lo;
    l.chunk = l.count < HOST_CHUNK ? l.count : HOST_CHUNK;
    if (l.count == 0) continue;
    st = get_replica(c, l.device, &l.rp);
    if (st) break;
    DeviceGuard dg(l.device);
    if (!dg.ok) {
      st = ENTERP_NO_DEVICE;
      break;
    }
    l.hp = acquire_pipe(l.device);
    const cudaError_t e = lane_prepare(j, l);
    if (e != cudaSuccess) st = cuda_fail(e, "host pipeline scratch");
    if (j.hybrid && l.hp->hybrid_bad >= 2) {  // this host did not profit from the fills lately: start plain, probe later
      l.plain_mode = true;
      l.plain_windows = 4;
    }
  }
  if (j.hybrid)
    j.feed = [&j, &lanes]() {
      for (HostLane& fl : lanes)
        if (fl.err == ENTERP_OK) fl.err = feed_lane(j, fl);
    };
  // round-robin issue: every device gets its next chunk before anyone gets a second one
  bool more = st == ENTERP_OK;
  while (more) {
    more = false;
    for (HostLane& l : lanes) {
      if (!l.hp || l.next >= l.count) continue;
      st = lane_step(j, l);
      if (st == ENTERP_OK) st = l.err;
      if (st) break;
      more = more || l.next < l.count;
    }
    if (st) break;
  }
  // drain (also on errors: nothing may still be writing into the caller's buffers when we return)
  for (HostLane& l : lanes) {
    if (!l.hp) continue;
    DeviceGuard dg(l.device);
    if (st == ENTERP_OK) {
      for (int b = 0; b < 2 && st == ENTERP_OK; ++b) {  // CPU work first: the plain copies keep flowing meanwhile
        const cudaError_t e = rle_finish(j, l, (int)((l.kr + b) & 1));
        if (e != cudaSuccess) st = cuda_fail(e, "run-length drain");
        if (st == ENTERP_OK) st = l.err;
      }
      for (int b = 0; b < PLAIN_SLOTS && st == ENTERP_OK; ++b) {
        bool done = true;
        const cudaError_t e = plain_finish(j, l, b, true, &done);
        if (e != cudaSuccess) st = cuda_fail(e, "pipeline drain");
      }
    }
    if (st != ENTERP_OK) {
      for (int b = 0; b < PLAIN_SLOTS; ++b)
        if (l.hp->stream[b]) cudaStreamSynchronize(l.hp->stream[b]);
      for (int b = 0; b < 2; ++b)
        if (l.hp->rstream[b]) cudaStreamSynchronize(l.hp->rstream[b]);
    }
    release_pipe(l.hp);
  }
  return st;
}

}  // namespace

extern "C" {

ENTERP_API enterp_status enterp_curve_validate(const enterp_curve_desc* desc, int stage, enterp_error_info* err) {
  if (!desc) return ENTERP_INVALID_ARGUMENT;
  return validate(*desc, stage, err);
}

ENTERP_API enterp_status enterp_curve_create(const enterp_curve_desc* desc, enterp_curve** out, enterp_error_info* err) {
  if (out) *out = nullptr;
  if (!desc || !out) return ENTERP_INVALID_ARGUMENT;
  enterp_curve* c = new enterp_curve();
  c->core = std::make_shared<CurveCore>();
  enterp_status st = resolve(*desc, c->core->r, err);
  if (st) {
    delete c;
    return st;
  }
  if (c->core->r.has_transform) c->pre.push_back(PreOp{1, c->core->r.in_mul, c->core->r.in_add});
  *out = c;
  return ENTERP_OK;
}

ENTERP_API void enterp_curve_destroy(enterp_curve* c) { delete c; }  // the core goes with its last handle

// Curve::clamp() -> Clamp (base/adaptors.rs:14-44): clamp(input, domain) before the wrapped eval.
ENTERP_API enterp_status enterp_curve_clamp(const enterp_curve* base, enterp_curve** out) {
  if (out) *out = nullptr;
  if (!base || !out) return ENTERP_INVALID_ARGUMENT;
  if (base->pre.size() >= (size_t)MAX_PRE) return ENTERP_UNSUPPORTED;
  enterp_curve* c = new enterp_curve(*base);
  c->pre.insert(c->pre.begin(), PreOp{0, base->core->r.domain[0], base->core->r.domain[1]});
  *out = c;
  return ENTERP_OK;
}

// Curve::slice(bounds) -> Slice::new (base/adaptors.rs:61-81): TransformInput::new(signal,
// bound_start - signal_start, (bound_end - bound_start) / (signal_end - signal_start)), all in R arithmetic.
ENTERP_API enterp_status enterp_curve_slice(const enterp_curve* base, int has_start, double start, int has_end,
                                            double end, enterp_curve** out) {
  if (out) *out = nullptr;
  if (!base || !out) return ENTERP_INVALID_ARGUMENT;
  if (base->pre.size() >= (size_t)MAX_PRE) return ENTERP_UNSUPPORTED;
  const Resolved& r = base->core->r;
  double mul, add;
  if (r.scalar == ENTERP_F32) {
    const float s0 = (float)r.domain[0], s1 = (float)r.domain[1];
    const float b0 = has_start ? (float)start : s0, b1 = has_end ? (float)end : s1;
    const float num = b1 - b0, den = s1 - s0;
    mul = num / den;
    add = b0 - s0;
  } else {
    const double s0 = r.domain[0], s1 = r.domain[1];
    const double b0 = has_start ? start : s0, b1 = has_end ? end : s1;
    const double num = b1 - b0, den = s1 - s0;
    mul = num / den;
    add = b0 - s0;
  }
  enterp_curve* c = new enterp_curve(*base);
  c->pre.insert(c->pre.begin(), PreOp{1, mul, add});
  *out = c;
  return ENTERP_OK;
}

ENTERP_API enterp_status enterp_curve_get_resolved(const enterp_curve* c, enterp_curve_resolved* out) {
  if (!c || !out) return ENTERP_INVALID_ARGUMENT;
  const Resolved& r = c->core->r;
  const bool equi = r.kind != ENTERP_BEZIER && r.knot_base != ENTERP_KNOTS_SORTED;
  out->eq_len = equi ? r.inner_len : 0;
  out->eq_step = equi ? r.eq_step : 0.0;
  out->eq_offset = equi ? r.eq_offset : 0.0;
  out->has_transform = r.has_transform;
  out->easing = r.easing;
  out->in_addition = r.in_add;
  out->in_multiplication = r.in_mul;
  out->space_len = r.space_len;
  out->ease_min = r.ease_min;
  out->ease_max = r.ease_max;
  return ENTERP_OK;
}

ENTERP_API enterp_status enterp_curve_get_info(const enterp_curve* c, enterp_curve_info* info) {
  if (!c || !info) return ENTERP_INVALID_ARGUMENT;
  const Resolved& r = c->core->r;
  info->kind = r.kind;
  info->scalar = r.scalar;
  info->dim = r.dim;
  info->weighted = r.weighted;
  info->knot_base = r.knot_base;
  info->adaptor = r.adaptor;
  info->n_elements = r.n_elem;
  info->n_knots = r.virt_len;
  info->degree = r.degree;
  info->border = r.border_n;
  info->domain[0] = r.domain[0];
  info->domain[1] = r.domain[1];
  return ENTERP_OK;
}

ENTERP_API enterp_status enterp_curve_domain(const enterp_curve* c, double out[2]) {
  if (!c || !out) return ENTERP_INVALID_ARGUMENT;
  out[0] = c->core->r.domain[0];
  out[1] = c->core->r.domain[1];
  return ENTERP_OK;
}

ENTERP_API enterp_status enterp_curve_upload(enterp_curve* c, int device) {
  if (!c) return ENTERP_INVALID_ARGUMENT;
  Replica* rp = nullptr;
  return get_replica(c, device, &rp);
}

ENTERP_API enterp_status enterp_eval_batch(enterp_curve* c, int device, const void* t_dev, uint64_t n, void* out_dev,
                                           void* stream) {
  return eval_common(c, device, false, t_dev, 0, 0, n, out_dev, stream);
}

ENTERP_API enterp_status enterp_eval_take(enterp_curve* c, int device, uint64_t n_total, uint64_t first, uint64_t count,
                                          void* out_dev, void* stream) {
  return eval_common(c, device, true, nullptr, n_total, first, count, out_dev, stream);
}

ENTERP_API enterp_status enterp_span_batch(enterp_curve* c, int device, const void* t_dev, uint64_t n, uint32_t* idx_dev,
                                           void* factor_dev, void* stream) {
  return span_common(c, device, false, t_dev, 0, 0, n, idx_dev, factor_dev, stream);
}

ENTERP_API enterp_status enterp_span_take(enterp_curve* c, int device, uint64_t n_total, uint64_t first, uint64_t count,
                                          uint32_t* idx_dev, void* factor_dev, void* stream) {
  return span_common(c, device, true, nullptr, n_total, first, count, idx_dev, factor_dev, stream);
}

ENTERP_API enterp_status enterp_stepper(int scalar, int device, double start, double end, uint64_t n_total,
                                        uint64_t first, uint64_t count, void* t_out_dev, void* stream) {
  if (count == 0) return ENTERP_OK;
  if (!t_out_dev || n_total == 0 || first > n_total || count > n_total - first) return ENTERP_INVALID_ARGUMENT;
  if (scalar != ENTERP_F32 && scalar != ENTERP_F64) return ENTERP_INVALID_ARGUMENT;
  DeviceGuard g(device);
  if (!g.ok) return ENTERP_NO_DEVICE;
  Resolved r;
  r.domain[0] = start;
  r.domain[1] = end;
  cudaError_t e;
  if (scalar == ENTERP_F32)
    e = launch_stepper<float>(take_src<float>(r, n_total, first, count), static_cast<float*>(t_out_dev),
                              static_cast<cudaStream_t>(stream));
  else
    e = launch_stepper<double>(take_src<double>(r, n_total, first, count), static_cast<double*>(t_out_dev),
                               static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "stepper launch");
  return ENTERP_OK;
}

ENTERP_API enterp_status enterp_bezier_many_take(int scalar, int dim, uint64_t n_curves, uint32_t n_ctrl,
                                                 const void* ctrl_dev, uint32_t samples_per_curve, int device,
                                                 void* out_dev, void* stream) {
  if (n_curves == 0 || samples_per_curve == 0) return ENTERP_OK;
  if (!ctrl_dev || !out_dev || dim < 1 || dim > 4) return ENTERP_INVALID_ARGUMENT;
  if (scalar != ENTERP_F32 && scalar != ENTERP_F64) return ENTERP_INVALID_ARGUMENT;
  if (n_ctrl < 1) return ENTERP_EMPTY;  // bezier/builder.rs:141
  if (n_ctrl > (uint32_t)MAX_DYN_BEZIER) return ENTERP_UNSUPPORTED;
  DeviceGuard g(device);
  if (!g.ok) return ENTERP_NO_DEVICE;
  const int n = n_ctrl <= 16 ? (int)n_ctrl : 0;
  cudaError_t e;
  if (scalar == ENTERP_F32) {
    const float step = (1.0f - 0.0f) / (float)(uint64_t)(samples_per_curve - 1);
    e = launch_bezier_many<float>(dim, n, static_cast<const float*>(ctrl_dev), n_curves, n_ctrl, samples_per_curve, step,
                                  static_cast<float*>(out_dev), static_cast<cudaStream_t>(stream));
  } else {
    const double step = (1.0 - 0.0) / (double)(uint64_t)(samples_per_curve - 1);
    e = launch_bezier_many<double>(dim, n, static_cast<const double*>(ctrl_dev), n_curves, n_ctrl, samples_per_curve,
                                   step, static_cast<double*>(out_dev), static_cast<cudaStream_t>(stream));
  }
  if (e != cudaSuccess) return cuda_fail(e, "bezier_many launch");
  return ENTERP_OK;
}

static enterp_status bezier_deriv_common(enterp_curve* c, int device, const void* t_dev, uint64_t n, uint32_t k,
                                         int tangent, void* out_dev, void* stream) {
  if (!c) return ENTERP_INVALID_ARGUMENT;
  const Resolved& r = c->core->r;
  if (r.kind != ENTERP_BEZIER || r.weighted || !c->pre.empty()) return ENTERP_UNSUPPORTED;
  if (n == 0) return ENTERP_OK;
  if (!t_dev || !out_dev) return ENTERP_INVALID_ARGUMENT;
  if (r.n_elem > (uint64_t)MAX_DERIV_N || k > (uint32_t)MAX_DERIV_N) return ENTERP_UNSUPPORTED;
  if (k == 0) return ENTERP_OK;  // `[T; 0]`: nothing to write
  // gen_with_deriatives::<K> slices `grad[..=min(K, len-1)]`: out of bounds (a panic upstream) unless K >= len
  if (!tangent && r.n_elem - 1 >= (uint64_t)k) return ENTERP_INVALID_ARGUMENT;
  Replica* rp = nullptr;
  enterp_status st = get_replica(c, device, &rp, static_cast<cudaStream_t>(stream));
  if (st) return st;
  DeviceGuard g(device);
  if (!g.ok) return ENTERP_NO_DEVICE;
  cudaError_t e;
  if (r.scalar == ENTERP_F32)
    e = launch_bezier_deriv<float>(r.dim, make_dev<float>(*c, *rp), batch_src<float>(t_dev, n), (int)k, tangent,
                                   static_cast<float*>(out_dev), static_cast<cudaStream_t>(stream));
  else
    e = launch_bezier_deriv<double>(r.dim, make_dev<double>(*c, *rp), batch_src<double>(t_dev, n), (int)k, tangent,
                                    static_cast<double*>(out_dev), static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "bezier derivative launch");
  return ENTERP_OK;
}

ENTERP_API enterp_status enterp_bezier_tangent_batch(enterp_curve* c, int device, const void* t_dev, uint64_t n,
                                                     void* out_dev, void* stream) {
  return bezier_deriv_common(c, device, t_dev, n, 2, 1, out_dev, stream);
}

ENTERP_API enterp_status enterp_bezier_derivatives_batch(enterp_curve* c, int device, const void* t_dev, uint64_t n,
                                                         uint32_t k, void* out_dev, void* stream) {
  return bezier_deriv_common(c, device, t_dev, n, k, 0, out_dev, stream);
}

ENTERP_API enterp_status enterp_sample_host(enterp_curve* c, int device, const void* t_host, uint64_t n, void* out_host) {
  return host_common(c, &device, 1, false, t_host, 0, 0, n, out_host);
}

ENTERP_API enterp_status enterp_take_host(enterp_curve* c, int device, uint64_t n_total, uint64_t first, uint64_t count,
                                          void* out_host) {
  return host_common(c, &device, 1, true, nullptr, n_total, first, count, out_host);
}

ENTERP_API enterp_status enterp_sample_host_multi(enterp_curve* c, const int* devices, int n_devices, const void* t_host,
                                                  uint64_t n, void* out_host) {
  return host_common(c, devices, n_devices, false, t_host, 0, 0, n, out_host);
}

ENTERP_API enterp_status enterp_take_host_multi(enterp_curve* c, const int* devices, int n_devices, uint64_t n_total,
                                                uint64_t first, uint64_t count, void* out_host) {
  return host_common(c, devices, n_devices, true, nullptr, n_total, first, count, out_host);
}

ENTERP_API int enterp_host_scratch_release(int device) { return release_idle_pipes(device); }

ENTERP_API void enterp_host_transfer_bytes(unsigned long long* h2d, unsigned long long* d2h) {
  if (h2d) *h2d = g_h2d_bytes.load();
  if (d2h) *d2h = g_d2h_bytes.load();
}

ENTERP_API enterp_status enterp_host_alloc(void** ptr, uint64_t bytes) {
  if (!ptr) return ENTERP_INVALID_ARGUMENT;
  *ptr = nullptr;
  CK(cudaMallocHost(ptr, bytes), "cudaMallocHost");
  return ENTERP_OK;
}

ENTERP_API void enterp_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

ENTERP_API enterp_status enterp_invalid_count(enterp_curve* c, int device, unsigned long long* n) {
  if (!c || !n) return ENTERP_INVALID_ARGUMENT;
  Replica* rp = nullptr;
  enterp_status st = get_replica(c, device, &rp);
  if (st) return st;
  DeviceGuard g(device);
  if (!g.ok) return ENTERP_NO_DEVICE;
  CK(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  CK(cudaMemcpy(n, rp->invalid, sizeof(unsigned long long), cudaMemcpyDeviceToHost), "cudaMemcpy");
  return ENTERP_OK;
}

ENTERP_API enterp_status enterp_invalid_count_stream(enterp_curve* c, int device, void* stream, unsigned long long* n) {
  if (!c || !n) return ENTERP_INVALID_ARGUMENT;
  Replica* rp = nullptr;
  enterp_status st = get_replica(c, device, &rp, static_cast<cudaStream_t>(stream));
  if (st) return st;
  DeviceGuard g(device);
  if (!g.ok) return ENTERP_NO_DEVICE;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  CK(cudaMemcpyAsync(n, rp->invalid, sizeof(unsigned long long), cudaMemcpyDeviceToHost, cs), "cudaMemcpyAsync");
  CK(cudaStreamSynchronize(cs), "cudaStreamSynchronize");
  return ENTERP_OK;
}

ENTERP_API enterp_status enterp_invalid_reset(enterp_curve* c, int device, void* stream) {
  if (!c) return ENTERP_INVALID_ARGUMENT;
  Replica* rp = nullptr;
  enterp_status st = get_replica(c, device, &rp, static_cast<cudaStream_t>(stream));
  if (st) return st;
  DeviceGuard g(device);
  if (!g.ok) return ENTERP_NO_DEVICE;
  CK(cudaMemsetAsync(rp->invalid, 0, sizeof(unsigned long long), static_cast<cudaStream_t>(stream)), "cudaMemsetAsync");
  return ENTERP_OK;
}

ENTERP_API const char* enterp_status_str(enterp_status s) {
  switch (s) {
    case ENTERP_OK: return "ok";
    case ENTERP_TOO_FEW_ELEMENTS: return "TooFewElements";
    case ENTERP_TOO_FEW_KNOTS: return "TooFewKnots";
    case ENTERP_KNOT_ELEMENT_INEQUALITY: return "KnotElementInequality";
    case ENTERP_NOT_SORTED: return "NotSorted";
    case ENTERP_INCONGRUOUS_ELEMENTS_KNOTS: return "IncongruousElementsKnots";
    case ENTERP_INCONGRUOUS_ELEMENTS_DEGREE: return "IncongruousElementsDegree";
    case ENTERP_INVALID_DEGREE: return "InvalidDegree";
    case ENTERP_TOO_SMALL_WORKSPACE: return "TooSmallWorkspace";
    case ENTERP_EMPTY: return "Empty";
    case ENTERP_INVALID_ARGUMENT: return "invalid argument";
    case ENTERP_UNSUPPORTED: return "unsupported";
    case ENTERP_NO_DEVICE: return "no CUDA device";
    case ENTERP_CUDA_ERROR: return "CUDA error";
    case ENTERP_NOT_UPLOADED: return "curve not uploaded to this device (call enterp_curve_upload before capturing the stream)";
  }
  return "unknown status";
}

ENTERP_API const char* enterp_last_cuda_error(void) { return t_cuda_error.c_str(); }

ENTERP_API int enterp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

ENTERP_API unsigned long long enterp_launch_count(void) { return g_launches.load(); }

ENTERP_API const char* enterp_version(void) { return "enterpolation_b200 0.1.0 (sm_100a; reference enterpolation 0.3.0)"; }

}  // extern "C"
