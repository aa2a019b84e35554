// policy_device.cuh -- device-side pieces shared by ctf_policy_kernel (policy_kernels.cu) and the fused policy prologue of the
// 2v2 CtF step kernel (map_kernels.cu).
#pragma once
#include "mg_device.cuh"

namespace mg {

// The opponents' draws: Philox blocks of their own - counter (env id, 16 * step_count + block, 2^31 | episode) - so they never meet
// the step's blocks (4th counter word 0).  Shared by ctf_policy_kernel and the fused prologue of the 2v2 step kernel.
struct PolicyRng {
  uint32_t k0, k1, id0, id1, c2, c3, b0, b1, b2, b3;
  int have;
  __device__ __forceinline__ void open(unsigned long long seed, unsigned long long env_id, int step_count, int episode) {
    k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); id0 = (uint32_t)env_id; id1 = (uint32_t)(env_id >> 32);
    c2 = (uint32_t)step_count * 16u; c3 = 0x80000000u | (uint32_t)episode; have = 0;
  }
  __device__ __forceinline__ uint32_t u32() {
    if (have == 0) {
      uint32_t o[4];
      philox4x32_10(id0, id1, c2, c3, k0, k1, o);
      b0 = o[0]; b1 = o[1]; b2 = o[2]; b3 = o[3];
      ++c2; have = 4;
    }
    const uint32_t v = b0;   // words in order, a shift register (no dynamically indexed array)
    b0 = b1; b1 = b2; b2 = b3; --have;
    return v;
  }
  __device__ __forceinline__ int below(int n) { return (int)__umulhi(u32(), (uint32_t)n); }
};

}  // namespace mg
