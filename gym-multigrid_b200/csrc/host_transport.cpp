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
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace mg {

// ------------------------------------------------------------------------------------------------ thread pool
// Persistent workers; a job is a function of (worker index, worker count).  Workers spin ~1 ms for the next job (steps
// arrive back to back in an RL loop) and then sleep on a condition variable, so an idle env costs no CPU.
class HostPool {
 public:
  static HostPool& get() { static HostPool p; return p; }

  void ensure(int threads) {
    std::lock_guard<std::mutex> g(run_mu_);
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if (threads <= size_) return;
    std::lock_guard<std::mutex> lk(mu_);
    const uint64_t now = gen_.load(std::memory_order_acquire);   // no job can start before this returns (run_mu_ is held)
    for (int i = size_; i < threads; ++i) workers_.emplace_back([this, i, now] { loop(i, now); });   // worker 0 is the caller's thread
    size_ = threads;
  }
  int size() const { return size_; }

  // runs f(k, n) for k in [0, n) with n = min(threads, size()); returns when all are done
  void run(int threads, const std::function<void(int, int)>& f) {
    std::lock_guard<std::mutex> g(run_mu_);
    int n = threads < size_ ? threads : size_;
    if (n < 1) n = 1;
    if (n == 1) { f(0, 1); return; }
    job_ = &f; job_n_ = n;
    pending_.store(size_ - 1, std::memory_order_release);   // EVERY worker acknowledges every generation (those beyond n without running)
    {
      std::lock_guard<std::mutex> lk(mu_);
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    f(0, n);
    int spins = 0;
    while (pending_.load(std::memory_order_acquire) != 0) {
      if (++spins < 4096) _mm_pause(); else std::this_thread::yield();
    }
    job_ = nullptr;
  }

 private:
  HostPool() = default;
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    for (auto& t : workers_) if (t.joinable()) t.join();
  }
  void loop(int id, uint64_t seen) {
    for (;;) {
      int spins = 0;
      while (gen_.load(std::memory_order_acquire) == seen && spins < 20000) { _mm_pause(); ++spins; }
      if (gen_.load(std::memory_order_acquire) == seen) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
      }
      seen = gen_.load(std::memory_order_acquire);
      if (stop_) return;
      // workers beyond the job's width sit this one out; worker `id` is index id of the job (0 is the caller)
      if (id < job_n_) (*job_)(id, job_n_);
      pending_.fetch_sub(1, std::memory_order_acq_rel);
    }
  }

  std::mutex run_mu_, mu_;
  std::condition_variable cv_;
  std::vector<std::thread> workers_;
  std::atomic<uint64_t> gen_{0};
  std::atomic<int> pending_{0};
  const std::function<void(int, int)>* job_ = nullptr;
  int job_n_ = 0, size_ = 1;
  bool stop_ = false;
};

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

static inline void split(size_t n, int k, int parts, size_t granule, size_t& lo, size_t& hi) {
  const size_t units = (n + granule - 1) / granule;
  lo = units * (size_t)k / (size_t)parts * granule;
  hi = units * (size_t)(k + 1) / (size_t)parts * granule;
  if (lo > n) lo = n;
  if (hi > n) hi = n;
}

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

void host_expand_plane(const uint8_t* grid, uint8_t* obs, size_t n_cells, int threads) {
  HostPool::get().run(threads, [&](int k, int parts) {
    size_t lo, hi;
    split(n_cells, k, parts, 4096, lo, hi);
    if (hi > lo) expand_cells(grid + lo, obs + 3 * lo, hi - lo);
  });
}

void host_apply_delta(const HostDeltaJob& j, int threads) {
  HostPool::get().run(threads, [&](int k, int parts) {
    size_t lo, hi;
    split(j.n, k, parts, 64, lo, hi);
    const int A = j.A, R = j.stride;
    const size_t row = (size_t)3 * j.cells;
    for (size_t e = lo; e < hi; ++e) {
      const uint8_t* rec = j.records + e * R;
      const uint8_t b0 = rec[0];
      const int n = b0 & 31;
      if (j.terminated) j.terminated[e] = (b0 >> 5) & 1;
      if (j.truncated) j.truncated[e] = (b0 >> 6) & 1;
      if (j.rewards)
        for (int i = 0; i < A; ++i) j.rewards[e * A + i] = j.reward_table[rec[1 + i] <= 32 ? rec[1 + i] : 0];
      if (!j.obs || j.skip_patches) continue;
      uint8_t* o = j.obs + e * row;
      const uint8_t* ent = rec + 1 + A;
      const size_t ncell = (size_t)j.cells;   // indices are checked: a corrupt record must not write outside its env's row
      const int nmax = n <= 3 * A ? n : 3 * A;
      if (!j.wide) {
        for (int q = 0; q < nmax; ++q) { const size_t idx = ent[2 * q]; if (idx < ncell) put3(o + 3 * idx, ent[2 * q + 1]); }
      } else {
        for (int q = 0; q < nmax; ++q) { const size_t idx = (size_t)ent[3 * q] | ((size_t)ent[3 * q + 1] << 8); if (idx < ncell) put3(o + 3 * idx, ent[3 * q + 2]); }
      }
      if ((b0 & 0x80) && j.final_obs) std::memcpy(j.final_obs + e * row, o, row);   // terminal observation, before the fresh row lands
    }
  });
}

void host_apply_rows(const uint8_t* rows, size_t stride, size_t count, int cells, uint8_t* obs, size_t num_envs, int threads) {
  HostPool::get().run(threads, [&](int k, int parts) {
    size_t lo, hi;
    split(count, k, parts, 16, lo, hi);
    for (size_t s = lo; s < hi; ++s) {
      const uint8_t* r = rows + s * stride;
      int32_t e;
      std::memcpy(&e, r, 4);
      if (e < 0 || (size_t)e >= num_envs) continue;
      expand_cells(r + 4, obs + (size_t)e * 3 * cells, (size_t)cells);
    }
  });
}

}  // namespace mg
