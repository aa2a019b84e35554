// policy_params.cuh -- parameter block of the CtF scripted-opponent kernel (policy_kernels.cu <-> mg_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

struct PolicyParams {
  int S, cells, nb, nr, row_bytes, n_along, blue_flag_cell;
  long long N;
  unsigned long long seed, env_id_base;
  const uint8_t* agents;        // state plane: [N_pad][row_bytes], agent i at bytes 4i..4i+3 = x, y, dir, flags
  const int4* hdr;              // state plane: step_count, stats, rng counter, episode count
  const uint8_t* field_map;     // [cells] x*S+y, CtfWorld codes
  const uint8_t* first_move;    // [cells][cells] action of the first move of the reference's A* route, start-major
  const uint16_t* patrol_goal;  // [cells] border cell closest to each cell (cell index)
  const uint8_t* on_border;     // [cells]
  const uint16_t* along;        // [n_along] patrol candidates (cell indices, duplicates kept)
  int kind[16];                 // per red agent: MG_POLICY_*
  unsigned long long thr[16];   // follow the route iff u32 < thr = ceil(randomness * 2^32)
  int8_t* out;                  // [N][nr]
  // validation mode (mg_set_policy_trace): the reference generator's recorded outputs per (env, red agent) instead of Philox draws
  const uint16_t* tr_patrol;    // [N][nr] cell drawn by PatrolPolicy on the border (np_random.choice, heuristic.py:334); unused entries ignored
  const uint8_t* tr_follow;     // [N][nr] 1 = follow the route (np_random.choice([True, False], p=...), heuristic.py:150-151)
  const int8_t* tr_action;      // [N][nr] the uniform action (np_random.integers, :72 / :175) where one was drawn
};

}  // namespace mg
