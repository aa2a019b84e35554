// map_params.cuh -- kernel parameter block of the static-map families (Maze, CtF); shared by
// map_kernels.cu (device) and mg_api.cu (host launcher).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

struct MapParams {
  int S, cells, nb, nr, n, family, max_steps, autoreset, obs_dtype, variant_1v1;
  double flag_reward, obstacle_penalty, step_penalty, battle_reward, battle_range, randomness;
  int n_background, len_blue, len_red, blue_flag, red_flag;
  long long N;
  unsigned long long env_id_base, seed;
  // handle-owned tables
  const uint8_t* field_map;    // [cells] x*S+y
  const uint8_t* obs_period;   // [L] obs base repeated to a multiple of 16 bytes (CtF: transposed map)
  int L;                       // lcm(cells, 16)
  const uint16_t* background;  // Maze: cells with code background, np.where order
  const uint16_t* blue_terr;   // CtF: blue territory cells + blue flag (ctf.py:765-769)
  const uint16_t* red_terr;
  // state planes
  uint8_t* pos; uint8_t* dir; uint8_t* flags; int4* hdr;
  // io
  const int8_t* actions; void* obs; double* rewards; uint8_t* terminated; uint8_t* truncated; void* final_obs;
  const uint8_t* reset_mask;
  // trace replay
  int rng_mode;
  const int32_t* start_index; const int32_t* blue_place; const int32_t* red_place;
  const int8_t* red_actions; const uint8_t* order; const uint8_t* blue_win; int KB; int32_t* battles_used;
  int32_t* status;
  int op;  // 0 = reset(mask), 1 = step
};

}  // namespace mg
