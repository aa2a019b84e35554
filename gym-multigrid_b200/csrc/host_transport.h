// host_transport.h -- host-side decoders of the compact result transports (see host_transport.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <atomic>
#include <functional>

namespace mg {

// A unit of work for the host pool: body(lo, hi) over [0, n) in chunks of `grain` units on at most `cap` threads at a time; `then` runs
// on the thread that finished the last chunk (it may submit a follow-up task or raise a flag somebody waits on in host_help_until).
struct HostTask {
  std::function<void(size_t, size_t)> body;
  std::function<void()> then;
  size_t n = 0, grain = 1;
  int cap = 1 << 30;
  // pool-owned
  size_t chunks = 0, next = 0;
  std::atomic<size_t> left{0};
  std::atomic<int> active{0};
};
void host_submit(HostTask* t);                          // returns at once; the pool's workers (and helpers) run it
void host_help_until(const std::atomic<int>& done);     // the caller works on queued chunks until `done` is non-zero
// steps still on the device: while host_poll_add's sum is positive, ONE idle pool thread at a time calls probe(); a non-null
// result is passed to start() (outside the watcher lock), which submits that step's decode tasks
void host_set_poller(void* (*probe)(), void (*start)(void*));
void host_poll_add(int delta);

int host_default_threads();          // host cores this process may run on (affinity mask), capped at 32
void host_pool_ensure(int threads);  // grow the process-wide worker pool to `threads` (the caller's thread is one of them)

// packed plane -> Grid.encode(): obs[3k .. 3k+2] = (type, colour, state) of cell k, for n_cells cells
void host_expand_plane(const uint8_t* grid, uint8_t* obs, size_t n_cells, int threads);

struct HostDeltaJob {
  const uint8_t* records;  // [n][stride], layout in mg_device.cuh (CollectParams::delta)
  size_t n;
  int stride, wide, cells, A;
  const double* reward_table;  // [33]: entry 0 = 0.0, entry 1 + (colour | respawned << 4) = that ball's reward
  uint8_t* obs;                // [n][cells][3] persistent mirror, patched in place (NULL = skip)
  int skip_patches;            // 1 = the mirror is being refreshed in full: decode rewards / flags only
  double* rewards;             // [n][A]
  uint8_t* terminated;         // [n]
  uint8_t* truncated;          // [n]
  uint8_t* final_obs;          // [n][cells][3] or NULL: rows of the envs that autoreset receive their terminal observation
};
void host_apply_delta(const HostDeltaJob& job, int threads);

// fresh rows of the envs that autoreset: each entry = int32 env index followed by `cells` packed cells (stride bytes apart)
void host_apply_rows(const uint8_t* rows, size_t stride, size_t count, int cells, uint8_t* obs, size_t num_envs, int threads);


// task forms of the three decoders for host_submit (asynchronous decode of a step that is still in flight)
void host_expand_task(HostTask* t, const uint8_t* grid, uint8_t* obs, size_t n_cells, int threads);
void host_delta_task(HostTask* t, const HostDeltaJob* job, int threads);
void host_rows_task(HostTask* t, const uint8_t* rows, size_t stride, size_t count, int cells, uint8_t* obs, size_t num_envs, int threads);

}  // namespace mg
