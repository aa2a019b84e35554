// host_transport.h -- host-side decoders of the compact result transports (see host_transport.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace mg {

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

}  // namespace mg
