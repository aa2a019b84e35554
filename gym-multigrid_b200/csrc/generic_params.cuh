// generic_params.cuh -- kernel parameter block of the generic MultiGridEnv.step family (DefaultWorld).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

struct GenericParams {
  uint32_t cells_magic, per_env_magic;   // floor(2^32 / cells) + 1, floor(2^32 / (A * cells)) + 1 (tile-local indices < 2^16)
  uint32_t half_magic;                   // floor(2^32 / (cells / 2)) + 1 (even cell counts: the encode handles two cells per thread)
  int W, H, cells, A, max_steps, autoreset, op;   // op: 0 = reset(mask) from the snapshot planes, 1 = step
  long long N;
  unsigned long long env_id_base, seed;
  uint8_t* gcell; uint8_t* gstate; uint8_t* pos; int4* hdr;               // live state
  const uint8_t* icell; const uint8_t* istate; const uint8_t* ipos;       // episode-start snapshot (reset source)
  const int8_t* actions; uint8_t* obs; double* rewards; uint8_t* terminated; uint8_t* truncated; uint8_t* final_obs;
  const uint8_t* reset_mask;
  const uint8_t* order;   // replayed np.random.permutation outputs, or null (Philox Fisher-Yates)
  int32_t* status;
};

}  // namespace mg
