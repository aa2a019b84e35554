// host_transport.cpp -- host half of the compact result transports of mg_step_host (MG_TRANSPORT_PACKED / _DELTA).
//
// The step itself always runs on the device (collect_kernels.cu).  What crosses PCIe is not the expanded observation
// (3 bytes per cell) but either the packed grid plane (1 byte per cell) or a per-env record of the <= 3A cells the step
// wrote; this file turns those back into the (W, H, 3) uint8 `Grid.encode()` arrays the caller asked for (grid.py:223-252:
// OBJECT_IDX, COLOR_IDX, STATE per cell), in the caller's host buffer, on a small pool of host threads.  It is a decoder of
// the wire format, not an env implementation: no rule of the game lives here.
#include "host_transport.h"

#include <immintrin.h>
#include <sched.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace mg {

// ------------------------------------------------------------------------------------------------ thread pool
// Persistent workers that pull CHUNKS of tasks: a task is a body over [0, n) units cut into chunks of `grain` units, claimed one at
// a time by whichever thread is free - workers, the submitting thread while it waits (run / help_until), or both.  Nothing waits for a
// particular thread: a worker that the OS has descheduled (eight ranks sharing a host, a busy box) delays at most the chunk it holds,
// where a static split with a spin barrier stalls the whole step for a scheduler quantum.  Several tasks can be in flight (the decodes
// of different env batches); `then` runs on the thread that finished a task's last chunk and may submit a follow-up task.
// Idle workers spin briefly for the next task (steps arrive back to back in an RL loop), then sleep on a condition variable.
static const int g_probe_gap = [] { const char* v = std::getenv("MG_PROBE_GAP"); return v ? std::atoi(v) : 64; }();

class HostPool {
 public:
  static HostPool& get() { static HostPool p; return p; }

  void ensure(int threads) {   // `threads` counts the submitting thread: threads - 1 workers
    std::lock_guard<std::mutex> lk(mu_);
    if (threads > 64) threads = 64;
    while ((int)workers_.size() < threads - 1) workers_.emplace_back([this] { loop(); });
  }
  int size() { std::lock_guard<std::mutex> lk(mu_); return (int)workers_.size() + 1; }

  // the task must stay alive until its `then` has been called (or, without one, until `left` reads 0)
  void submit(HostTask* t) {
    t->chunks = (t->n + t->grain - 1) / t->grain;
    t->next = 0;
    t->active.store(0, std::memory_order_relaxed);
    t->left.store(t->chunks, std::memory_order_release);
    if (t->chunks == 0) { if (t->then) t->then(); return; }
    bool wake;
    {
      std::lock_guard<std::mutex> lk(mu_);
      tasks_.push_back(t);
      avail_.fetch_add(1, std::memory_order_release);
      wake = sleepers_ > 0;
    }
    if (wake) cv_.notify_all();
  }

  // runs one chunk of some task; false = nothing to claim right now
  bool help_once() {
    if (avail_.load(std::memory_order_acquire) == 0) return false;
    HostTask* t = nullptr;
    size_t c = 0;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (size_t i = 0; i < tasks_.size(); ++i) {
        HostTask* q = tasks_[i];
        if (q->active.load(std::memory_order_relaxed) >= q->cap) continue;
        t = q; c = q->next++;
        q->active.fetch_add(1, std::memory_order_relaxed);
        if (q->next == q->chunks) { tasks_.erase(tasks_.begin() + (long)i); avail_.fetch_sub(1, std::memory_order_release); }
        break;
      }
    }
    if (!t) return false;
    const size_t lo = c * t->grain, hi = lo + t->grain < t->n ? lo + t->grain : t->n;
    t->body(lo, hi);
    const std::function<void()> then = t->then;   // a copy: once `left` reads 0 the owner may reuse the task, also from inside `then`
    t->active.fetch_sub(1, std::memory_order_relaxed);
    if (t->left.fetch_sub(1, std::memory_order_acq_rel) == 1 && then) then();   // the decrement is the last access to *t
    return true;
  }

  // the calling thread works on whatever is queued until `done` is set (by a task's `then`)
  void help_until(const std::atomic<int>& done) {
    int idle = 0;
    while (!done.load(std::memory_order_acquire)) {
      if (help_once()) { idle = 0; continue; }
      if (try_poll()) continue;
      if (++idle < 2000) _mm_pause(); else std::this_thread::yield();
    }
  }

  // Steps whose device half is still running are watched by whichever thread has nothing to decode - one at a time, so the others can
  // go to sleep: `probe` returns a step that has landed in host memory (or null), `start` turns it into tasks.  No thread is
  // dedicated to waiting and none sleeps in the driver (a blocking event wait costs ~150 us here).
  void set_poller(void* (*probe)(), void (*start)(void*)) { probe_ = probe; start_ = start; }
  void poll_add(int d) { polls_.fetch_add(d, std::memory_order_acq_rel); }
  bool try_poll() {
    if (polls_.load(std::memory_order_acquire) <= 0 || !probe_) return false;
    if (!poll_mu_.try_lock()) return false;
    void* ready = probe_();
    poll_mu_.unlock();
    if (ready) start_(ready);
    else for (int i = 0; i < g_probe_gap; ++i) _mm_pause();   // a few us between probes: the driver's locks are also the enqueueing thread's
    return true;
  }

  // synchronous: body over [0, n) on at most `threads` threads (the caller is one of them)
  void run(int threads, size_t n, size_t grain, const std::function<void(size_t, size_t)>& body) {
    if (n == 0) return;
    if (threads <= 1 || n <= grain) { body(0, n); return; }
    std::atomic<int> done{0};
    HostTask t;
    t.body = body; t.n = n; t.grain = grain; t.cap = threads;
    t.then = [&done] { done.store(1, std::memory_order_release); };
    submit(&t);
    help_until(done);
  }

 private:
  HostPool() = default;
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) if (t.joinable()) t.join();
  }
  void loop() {
    for (;;) {
      if (help_once()) continue;
      int spins = 0;
      while (avail_.load(std::memory_order_acquire) == 0 && !stop_ && spins < 4000) {
        if (try_poll()) { spins = 0; continue; }   // the watcher stays awake while steps are in flight
        _mm_pause(); ++spins;
      }
      if (avail_.load(std::memory_order_acquire) != 0) {
        // queued chunks this thread may not take (the task's thread cap is reached): do not spin on the lock
        if (!help_once()) std::this_thread::yield();
        continue;
      }
      std::unique_lock<std::mutex> lk(mu_);
      if (stop_) return;
      ++sleepers_;
      cv_.wait(lk, [&] { return stop_ || avail_.load(std::memory_order_acquire) != 0; });
      --sleepers_;
      if (stop_) return;
    }
  }

  std::mutex mu_, poll_mu_;
  std::atomic<int> polls_{0};
  void* (*probe_)() = nullptr;
  void (*start_)(void*) = nullptr;
  std::condition_variable cv_;
  std::vector<std::thread> workers_;
  std::vector<HostTask*> tasks_;      // tasks with unclaimed chunks, oldest first
  std::atomic<int> avail_{0};         // == tasks_.size(), readable without the lock
  int sleepers_ = 0;
  std::atomic<bool> stop_{false};
};

void host_submit(HostTask* t) { HostPool::get().submit(t); }
void host_help_until(const std::atomic<int>& done) { HostPool::get().help_until(done); }
void host_set_poller(void* (*probe)(), void (*start)(void*)) { HostPool::get().set_poller(probe, start); }
void host_poll_add(int d) { HostPool::get().poll_add(d); }

int host_default_threads() {
  cpu_set_t set;
  CPU_ZERO(&set);
  int n = 0;
  if (sched_getaffinity(0, sizeof set, &set) == 0) n = CPU_COUNT(&set);
  if (n < 1) n = (int)std::thread::hardware_concurrency();
  if (n < 1) n = 1;
  return n > 32 ? 32 : n;
}

void host_pool_ensure(int threads) { HostPool::get().ensure(threads); }

// ------------------------------------------------------------------------------------------------ expansion
// packed cell = type | colour << 2 | state << 6  ->  (type, colour, state); a ball's bit 6 is internal (mg_device.cuh: expand4)
static inline void put3(uint8_t* o, uint8_t c) {
  o[0] = c & 3; o[1] = (c >> 2) & 15; o[2] = (c & 3) == 2 ? 0 : (uint8_t)(c >> 6);
}

static void expand_scalar(const uint8_t* in, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; ++i) put3(out + 3 * i, in[i]);
}

__attribute__((target("avx2"))) static void expand_avx2(const uint8_t* in, uint8_t* out, size_t n) {
  const __m256i m3 = _mm256_set1_epi8(3), m15 = _mm256_set1_epi8(15), two = _mm256_set1_epi8(2);
  // per 128-bit lane: 16 cells -> 48 bytes (t0 c0 s0 t1 c1 s1 ...), as three 16-byte pieces gathered with pshufb
  const __m256i st0 = _mm256_setr_epi8(0, -1, -1, 1, -1, -1, 2, -1, -1, 3, -1, -1, 4, -1, -1, 5, 0, -1, -1, 1, -1, -1, 2, -1, -1, 3, -1, -1, 4, -1, -1, 5);
  const __m256i sc0 = _mm256_setr_epi8(-1, 0, -1, -1, 1, -1, -1, 2, -1, -1, 3, -1, -1, 4, -1, -1, -1, 0, -1, -1, 1, -1, -1, 2, -1, -1, 3, -1, -1, 4, -1, -1);
  const __m256i ss0 = _mm256_setr_epi8(-1, -1, 0, -1, -1, 1, -1, -1, 2, -1, -1, 3, -1, -1, 4, -1, -1, -1, 0, -1, -1, 1, -1, -1, 2, -1, -1, 3, -1, -1, 4, -1);
  const __m256i st1 = _mm256_setr_epi8(-1, -1, 6, -1, -1, 7, -1, -1, 8, -1, -1, 9, -1, -1, 10, -1, -1, -1, 6, -1, -1, 7, -1, -1, 8, -1, -1, 9, -1, -1, 10, -1);
  const __m256i sc1 = _mm256_setr_epi8(5, -1, -1, 6, -1, -1, 7, -1, -1, 8, -1, -1, 9, -1, -1, 10, 5, -1, -1, 6, -1, -1, 7, -1, -1, 8, -1, -1, 9, -1, -1, 10);
  const __m256i ss1 = _mm256_setr_epi8(-1, 5, -1, -1, 6, -1, -1, 7, -1, -1, 8, -1, -1, 9, -1, -1, -1, 5, -1, -1, 6, -1, -1, 7, -1, -1, 8, -1, -1, 9, -1, -1);
  const __m256i st2 = _mm256_setr_epi8(-1, 11, -1, -1, 12, -1, -1, 13, -1, -1, 14, -1, -1, 15, -1, -1, -1, 11, -1, -1, 12, -1, -1, 13, -1, -1, 14, -1, -1, 15, -1, -1);
  const __m256i sc2 = _mm256_setr_epi8(-1, -1, 11, -1, -1, 12, -1, -1, 13, -1, -1, 14, -1, -1, 15, -1, -1, -1, 11, -1, -1, 12, -1, -1, 13, -1, -1, 14, -1, -1, 15, -1);
  const __m256i ss2 = _mm256_setr_epi8(10, -1, -1, 11, -1, -1, 12, -1, -1, 13, -1, -1, 14, -1, -1, 15, 10, -1, -1, 11, -1, -1, 12, -1, -1, 13, -1, -1, 14, -1, -1, 15);
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    const __m256i w = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i));
    const __m256i t = _mm256_and_si256(w, m3);
    const __m256i c = _mm256_and_si256(_mm256_srli_epi16(w, 2), m15);
    const __m256i s = _mm256_andnot_si256(_mm256_cmpeq_epi8(t, two), _mm256_and_si256(_mm256_srli_epi16(w, 6), m3));
    const __m256i a = _mm256_or_si256(_mm256_or_si256(_mm256_shuffle_epi8(t, st0), _mm256_shuffle_epi8(c, sc0)), _mm256_shuffle_epi8(s, ss0));
    const __m256i b = _mm256_or_si256(_mm256_or_si256(_mm256_shuffle_epi8(t, st1), _mm256_shuffle_epi8(c, sc1)), _mm256_shuffle_epi8(s, ss1));
    const __m256i d = _mm256_or_si256(_mm256_or_si256(_mm256_shuffle_epi8(t, st2), _mm256_shuffle_epi8(c, sc2)), _mm256_shuffle_epi8(s, ss2));
    // lanes: a = [A0 | A1], b = [B0 | B1], d = [C0 | C1]; memory order is A0 B0 C0 A1 B1 C1
    uint8_t* o = out + 3 * i;
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(o), _mm256_permute2x128_si256(a, b, 0x20));
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(o + 32), _mm256_permute2x128_si256(d, a, 0x30));
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(o + 64), _mm256_permute2x128_si256(b, d, 0x31));
  }
  expand_scalar(in + i, out + 3 * i, n - i);
}

static bool have_avx2() {
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}

static inline void expand_cells(const uint8_t* in, uint8_t* out, size_t n) {
  if (have_avx2()) expand_avx2(in, out, n); else expand_scalar(in, out, n);
}

// units per chunk: ~10 us of work each, so that claiming a chunk is noise and a straggler holds little
constexpr size_t kExpandGrain = 16384, kDeltaGrain = 1024, kRowsGrain = 128;

static inline void expand_range(const uint8_t* grid, uint8_t* obs, size_t lo, size_t hi) { expand_cells(grid + lo, obs + 3 * lo, hi - lo); }

// records of envs [lo, hi): flags, rewards (table lookup), patches of the mirror
static const size_t g_prefetch_dist = [] { const char* v = getenv("MG_DECODE_PREFETCH_DIST"); return v ? (size_t)atoi(v) : (size_t)24; }();   // 0 = off
static void delta_range(const HostDeltaJob& j, size_t lo, size_t hi) {
  const int A = j.A, R = j.stride;
  const size_t row = (size_t)3 * j.cells;
  const size_t ncell = (size_t)j.cells;   // indices are checked: a corrupt record must not write outside its env's row
  const bool patch = j.obs && !j.skip_patches;
  {
    for (size_t e = lo; e < hi; ++e) {
      const uint8_t* rec = j.records + e * R;
      const uint8_t b0_ = rec[0];
      const int n = b0_ & 31;
      if (j.terminated) j.terminated[e] = (b0_ >> 5) & 1;
      if (j.truncated) j.truncated[e] = (b0_ >> 6) & 1;
      if (j.rewards)
        for (int i = 0; i < A; ++i) j.rewards[e * A + i] = j.reward_table[rec[1 + i] <= 32 ? rec[1 + i] : 0];
      if (!patch) continue;
      uint8_t* o = j.obs + e * row;
      const uint8_t* ent = rec + 1 + A;
      const int nmax = n <= 3 * A ? n : 3 * A;
      if (g_prefetch_dist && e + g_prefetch_dist < hi) {
        // the mirror of a large batch does not stay in the caches between steps (several env batches, several ranks per host): ask for
        // the line of the first cell a later env's record names, and of its third (the other agent's) - unconditionally, the index
        // is only clamped: a stale entry costs a useless prefetch, a branch would cost more
        const uint8_t* r2 = rec + g_prefetch_dist * R;
        const uint8_t* o2 = o + g_prefetch_dist * row;
        if (!j.wide) {
          const size_t i0 = r2[1 + A], i2 = r2[1 + A + 4];
          __builtin_prefetch(o2 + 3 * (i0 < ncell ? i0 : 0), 1, 3);
          __builtin_prefetch(o2 + 3 * (i2 < ncell ? i2 : 0), 1, 3);
        } else {
          const size_t i0 = (size_t)r2[1 + A] | ((size_t)r2[2 + A] << 8);
          __builtin_prefetch(o2 + 3 * (i0 < ncell ? i0 : 0), 1, 3);
        }
      }
      if (!j.wide) {
        for (int q = 0; q < nmax; ++q) { const size_t idx = ent[2 * q]; if (idx < ncell) put3(o + 3 * idx, ent[2 * q + 1]); }
      } else {
        for (int q = 0; q < nmax; ++q) { const size_t idx = (size_t)ent[3 * q] | ((size_t)ent[3 * q + 1] << 8); if (idx < ncell) put3(o + 3 * idx, ent[3 * q + 2]); }
      }
      if ((b0_ & 0x80) && j.final_obs) std::memcpy(j.final_obs + e * row, o, row);   // terminal observation, before the fresh row lands
    }
  }
}

static void rows_range(const uint8_t* rows, size_t stride, int cells, uint8_t* obs, size_t num_envs, size_t lo, size_t hi) {
  for (size_t s = lo; s < hi; ++s) {
    const uint8_t* r = rows + s * stride;
    int32_t e;
    std::memcpy(&e, r, 4);
    if (e < 0 || (size_t)e >= num_envs) continue;
    expand_cells(r + 4, obs + (size_t)e * 3 * cells, (size_t)cells);
  }
}

void host_expand_plane(const uint8_t* grid, uint8_t* obs, size_t n_cells, int threads) {
  HostPool::get().run(threads, n_cells, kExpandGrain, [=](size_t lo, size_t hi) { expand_range(grid, obs, lo, hi); });
}

void host_apply_delta(const HostDeltaJob& j, int threads) {
  HostPool::get().run(threads, j.n, kDeltaGrain, [&j](size_t lo, size_t hi) { delta_range(j, lo, hi); });
}

void host_apply_rows(const uint8_t* rows, size_t stride, size_t count, int cells, uint8_t* obs, size_t num_envs, int threads) {
  HostPool::get().run(threads, count, kRowsGrain, [=](size_t lo, size_t hi) { rows_range(rows, stride, cells, obs, num_envs, lo, hi); });
}

// the same three as tasks for host_submit (the caller keeps `job` alive until the task's `then` has run)
void host_expand_task(HostTask* t, const uint8_t* grid, uint8_t* obs, size_t n_cells, int threads) {
  t->body = [=](size_t lo, size_t hi) { expand_range(grid, obs, lo, hi); };
  t->n = n_cells; t->grain = kExpandGrain; t->cap = threads;
}
void host_delta_task(HostTask* t, const HostDeltaJob* job, int threads) {
  t->body = [job](size_t lo, size_t hi) { delta_range(*job, lo, hi); };
  t->n = job->n; t->grain = kDeltaGrain; t->cap = threads;
}
void host_rows_task(HostTask* t, const uint8_t* rows, size_t stride, size_t count, int cells, uint8_t* obs, size_t num_envs, int threads) {
  t->body = [=](size_t lo, size_t hi) { rows_range(rows, stride, cells, obs, num_envs, lo, hi); };
  t->n = count; t->grain = kRowsGrain; t->cap = threads;
}

}  // namespace mg
