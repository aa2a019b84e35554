// map_params.cuh -- kernel parameter block of the static-map families (Maze, CtF); shared by
// map_kernels.cu (device) and mg_api.cu (host launcher).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

struct MapParams {
  int S, cells, nb, nr, n, family, max_steps, autoreset, obs_dtype, variant_1v1;
  double flag_reward, obstacle_penalty, step_penalty, battle_reward, battle_range, randomness;
  int n_background, len_blue, len_red;
  int blue_flag, red_flag;     // packed x | y << 8
  int L16, tile_mod_L16;       // L / 16, (envs per tile) mod L16
  unsigned L16_magic;          // floor(2^32 / L16) + 1: t mod L16 = t - umulhi(t, magic) * L16 for t < 2^16
  long long N;
  unsigned long long env_id_base, seed;
  // handle-owned tables
  const uint8_t* field_map;    // [cells] x*S+y
  const uint8_t* obs_period;   // [L] obs base repeated to a multiple of 16 bytes (CtF: transposed map)
  int L;                       // lcm(cells, 16)
  const uint8_t* obs_tile;     // staged u8 tiles: [envs per tile][cells] the obs base once per env of a tile (one bulk load), else null
  const uint16_t* background;  // Maze: cells with code background, np.where order; every list entry is packed x | y << 8
  const uint16_t* blue_terr;   // CtF: blue territory cells + blue flag (ctf.py:765-769)
  const uint16_t* red_terr;
  const int32_t* d2_tables;    // [3][cells] min squared distance from each cell to: Maze flag / obstacle / -, CtF blue territory / red territory / obstacle (-1 = empty list)
  // battle tests restated in integers (bit-exact, computed on the host with the same double arithmetic):
  int d2_max;                              // largest squared distance d2 with sqrt((double)d2) <= battle_range (-1: none)
  unsigned long long thr_blue_home, thr_red_home, thr_even;  // blue wins iff u32 < ceil(p * 2^32), p = randomness / 1 - randomness / 0.5
  int tma_reps;                            // copies of the period per bulk store (~32 KB chunks)
  int obs_tma;                             // 1 = (not staged) the period streams out through TMA bulk stores
  // Maze partial-observation mode (mg_set_partial_obs): the step / reset write gen_obs views [N][1][V][V][3] instead of the map
  int view_V, view_see_through;            // 0 = off; 3 / 5 / 7
  const uint8_t* map_padded; int pad, pitch, map_padded_bytes;   // packed static map surrounded by the out-of-map filler
  uint8_t view_oob, view_agent;
  // memoised views: on a static map a Maze view is a pure function of (x, y, dir), so mg_set_partial_obs runs the view kernel once
  // over all S*S*4 agent states and the step copies its env's row: [((x*S + y) << 2) | dir][view_row16] uint4, 3*V*V bytes used
  const uint4* view_table; int view_row16;
  int obs_staged;                          // 1 = the tile's u8 obs slab is assembled in shared memory (small maps)
  // state planes
  uint8_t* agents;   // [N_pad][row_bytes]: agent i at bytes 4i..4i+3 = x, y, dir, flags (bit0 terminated, bit1 collided)
  int row_bytes;     // 4 * (n rounded up to a power of two)
  int4* hdr;
  // io
  const int8_t* actions; void* obs; double* rewards; uint8_t* terminated; uint8_t* truncated; void* final_obs;
  const uint8_t* reset_mask;
  // trace replay
  int rng_mode;
  const int32_t* start_index; const int32_t* blue_place; const int32_t* red_place;
  const int8_t* red_actions; const uint8_t* order; const uint8_t* blue_win; int KB; int32_t* battles_used;
  int32_t* status;
  int op;  // 0 = reset(mask), 1 = step
  // scripted opponents decided inside the step kernel (mg_set_red_policy_fusion; the 2v2 lean kernel): the tables of
  // policy_params.cuh, the kind / follow threshold of the two red agents, and where the decided actions are written
  int carry_flags;   // CtF: reset keeps the agents' flag byte (mg_set_carry_agent_flags: one env instance through several episodes)
  int pol_on, pol_n_along;
  const uint8_t* pol_first_move; const uint16_t* pol_goal; const uint8_t* pol_border; const uint16_t* pol_along;
  int pol_kind[2]; unsigned long long pol_thr[2];
  int8_t* pol_out;
};

}  // namespace mg
