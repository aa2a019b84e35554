// view_params.cuh -- kernel parameter block of the partial-view kernel; shared by view_kernels.cu and mg_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

struct ViewParams {
  int W, H, cells, A, V, see_through, family;
  long long N;
  const uint8_t* grid;      // Collect: packed cells [N_pad][cells]
  const uint8_t* pos;       // [N_pad][A][2]
  const uint8_t* dirs;      // [N][A] or null (= 3: Collect agents never turn, multigrid.py:371-374)
  const uint8_t* map_codes; // Maze: packed static map [cells] (type | colour << 2), handle-owned
  uint8_t oob_code;         // cell shown outside the grid: Wall grey (Collect, grid.py:124-127)
  uint8_t agent_code;       // Maze: packed agent cell without the dir bits
  uint8_t* out;             // [N][A][V][V][3]
  int out_bulk_ok;
};

}  // namespace mg
