// wildfire_params.cuh -- kernel parameter block of the Wildfire extension; shared by wildfire_kernels.cu and mg_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/multigrid_b200.h"

namespace mg {

struct WildfireParams {
  int W, H, cells, A, num_fires, max_steps, autoreset, op;   // op: 0 = reset(mask), 1 = step
  uint32_t ignite_threshold[5], burnout_threshold;
  uint32_t rv_magic;       // same for 16-byte vectors per row (H / 16), used when H % 16 == 0; else = rw_magic
  uint32_t rw_magic;       // floor(2^32 / (H / 4)) + 1: word-in-row index without a division (fast kernel)
  uint8_t agent_colour[MG_MAX_WILDFIRE_AGENTS];
  long long N;
  unsigned long long env_id_base, seed;
  uint8_t* terrain;        // [N_pad][cells]
  uint8_t* agents;         // [N_pad][A][4]
  int4* hdr;               // [N_pad] step_count, tick, Philox block counter, episodes
  const int8_t* actions;   // [N][A]
  uint8_t* obs;            // [N][cells][3]
  double* rewards;         // [N][A]
  uint8_t* terminated; uint8_t* truncated; uint8_t* final_obs;
  const uint8_t* reset_mask;
  const uint8_t* order;    // [N][A] replayed agent order, or null (Philox)
};

}  // namespace mg
