// view_params.cuh -- kernel parameter block of the partial-view kernel; shared by view_kernels.cu and mg_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

struct ViewParams {
  int W, H, cells, A, V, see_through, family;
  long long N;
  const uint8_t* grid;      // Collect: packed cells [N_pad][cells]
  const uint8_t* pos;       // (x, y) of view g = e*A + k at pos[g * pos_stride]
  int pos_stride;
  const uint8_t* dirs;      // dir of view g at dirs[g * dir_stride], or null (= 3: Collect agents never turn, multigrid.py:371-374)
  int dir_stride;
  const uint8_t* map_codes; // Maze: packed static map [cells] (type | colour << 2), handle-owned
  const uint8_t* map_padded; // Maze: the same map surrounded by `pad` cells of the out-of-map filler, row pitch `pitch`
  int pad, pitch, map_padded_bytes;
  uint8_t oob_code;         // cell shown outside the grid: Wall grey (Collect, grid.py:124-127)
  uint8_t agent_code;       // Maze: packed agent cell without the dir bits
  uint8_t* out;             // [N][A][V][V][3]
  int out_bulk_ok;
};

// MultiGridEnv.gen_obs for the generic (DefaultWorld, encode_dim 6) family
struct View6Params {
  int W, H, cells, A, V, see_through;
  long long N;
  const uint8_t* gcell;   // [N_pad][cells] type | colour << 4
  const uint8_t* gstate;  // [N_pad][cells] door state / agent dir
  const uint8_t* pos;     // [N_pad][A][2]
  const uint8_t* dirs;    // [N][A] override, or null = the dir stored with the agent cell
  uint8_t* out;           // [N][A][V][V][6]
};

}  // namespace mg
