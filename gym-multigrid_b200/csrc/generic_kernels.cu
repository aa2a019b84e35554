// generic_kernels.cu -- the base-class MultiGridEnv.step (multigrid.py:397-483) with DefaultWorld
// (world.py:33-52, encode_dim 6) and per-agent full-grid observations Grid.encode_for_agents (grid.py:254-284).
// No shipped env reaches this path; it is here for API completeness, built for correctness first:
// one thread per env walks the agents in the given order on the env's two cell planes (type|colour, state), then
// the whole tile encodes [env][agent][cell] -> 6 bytes with consecutive threads on consecutive cells.
// Only still / left / right / forward are defined: any other action makes the reference evaluate
// `self.actions.available` (multigrid.py:447), which no action enum defines, and raise.
#include <cstdlib>

#include "mg_device.cuh"
#include "generic_params.cuh"

namespace mg {

constexpr int kGenE = 32, kGenThreads = 128;
constexpr int G_EMPTY = 1, G_DOOR = 4, G_GOAL = 8, G_AGENT = 10;  // DefaultWorld.OBJECT_TO_IDX (world.py:37-51)

__device__ __forceinline__ void generic_encode_tile(const GenericParams& p, long long e0, int n_here, uint8_t* out, const uint8_t* done_only,
                                                    int tid) {
  const int cells = p.cells, A = p.A;
  const long long total = (long long)n_here * A * cells;
  for (long long idx = tid; idx < total; idx += kGenThreads) {
    const int el = (int)(idx / ((long long)A * cells));
    if (done_only && !done_only[el]) continue;
    const int rem = (int)(idx - (long long)el * A * cells), k = rem / cells, i = rem - k * cells;
    const long long e = e0 + el;
    const uint8_t c = p.gcell[e * cells + i], s = p.gstate[e * cells + i];
    const int type = c & 15;
    uint8_t o2 = 0, o4 = 0, o5 = 0;
    if (type == G_DOOR) o2 = s;                                   // Door.encode object.py:238-259
    else if (type == G_AGENT) {                                   // Agent.encode agent.py:127-165 (carrying is always None here)
      o4 = s & 3;
      o5 = (i == p.pos[(e * A + k) * 2] * p.H + p.pos[(e * A + k) * 2 + 1]);
    }
    uint16_t* o = reinterpret_cast<uint16_t*>(out + ((e * A + k) * cells + i) * 6);
    o[0] = (uint16_t)(type | ((c >> 4) << 8)); o[1] = o2; o[2] = (uint16_t)(o4 | (o5 << 8));
  }
}

__device__ __forceinline__ void generic_reset_env(const GenericParams& p, long long e, int4& h) {
  for (int i = 0; i < p.cells; ++i) { p.gcell[e * p.cells + i] = p.icell[e * p.cells + i]; p.gstate[e * p.cells + i] = p.istate[e * p.cells + i]; }
  for (int i = 0; i < p.A * 2; ++i) p.pos[e * p.A * 2 + i] = p.ipos[e * p.A * 2 + i];
  h.x = 0; h.w += 1;
}

__global__ void __launch_bounds__(kGenThreads) generic_kernel(const __grid_constant__ GenericParams p) {
  __shared__ uint8_t s_done[kGenE];
  const int tid = threadIdx.x, A = p.A, cells = p.cells, H = p.H;
  const long long e0 = (long long)blockIdx.x * kGenE;
  const int n_here = (int)min((long long)kGenE, p.N - e0);
  pdl_launch_dependents();
  pdl_wait();
  bool done = false;
  int4 h = make_int4(0, 0, 0, 0);
  if (tid < n_here) {
    const long long e = e0 + tid;
    h = p.hdr[e];
    if (p.op == 0) {
      if (!p.reset_mask || p.reset_mask[e]) generic_reset_env(p, e, h);
    } else {
      uint8_t* gc = p.gcell + e * cells; uint8_t* gs = p.gstate + e * cells; uint8_t* pos = p.pos + e * A * 2;
      Rng<1> r;
      r.open_philox(p.seed, p.env_id_base + (unsigned long long)e, (uint32_t)h.z);
      uint8_t order[8];
      if (p.order) { for (int i = 0; i < A; ++i) order[i] = p.order[e * A + i]; }
      else {
        for (int i = 0; i < A; ++i) order[i] = (uint8_t)i;
        for (int i = A - 1; i > 0; --i) { const int j = (int)__umulhi(r.u32(), (uint32_t)(i + 1)); const uint8_t t = order[i]; order[i] = order[j]; order[j] = t; }
      }
      h.x += 1;  // multigrid.py:400
      bool term = false;
      int err = 0;
      for (int i = 0; i < A; ++i) p.rewards[e * A + i] = 0.0;
      for (int k = 0; k < A; ++k) {  // for i in order :408
        const int i = order[k], a = p.actions[e * A + i];
        if (a == 0) continue;  // still :413
        const int x = pos[2 * i], y = pos[2 * i + 1], here = x * H + y, dir = gs[here] & 3;
        const int fx = x + (dir == 0) - (dir == 2), fy = y + (dir == 1) - (dir == 3);  // DIR_TO_VEC constants.py:65-74
        if (a == 1) gs[here] = (uint8_t)((dir + 3) & 3);       // left :424-427
        else if (a == 2) gs[here] = (uint8_t)((dir + 1) & 3);  // right :430-431
        else if (a == 3) {                                     // forward :434-445
          if (fx < 0 || fy < 0 || fx >= p.W || fy >= H) { err |= MG_ERR_OOB; continue; }  // reference: bounds assert
          const int f = fx * H + fy, ftype = gc[f] & 15;
          if (ftype != G_EMPTY) {
            if (ftype == G_GOAL) {  // terminated + _reward(i, rewards, 1) :436-438, :218-223 -- mul, div, sub as separate roundings
              term = true;
              p.rewards[e * A + i] = __dadd_rn(p.rewards[e * A + i],
                                               __dsub_rn(1.0, __dmul_rn(0.9, __ddiv_rn((double)h.x, (double)p.max_steps))));
            }  // switch: empty hook (:439-440); every other object: nothing
          } else {  // an agent only ever advances into an EMPTY cell (:441-444)
            gc[f] = gc[here]; gs[f] = gs[here];
            gc[here] = G_EMPTY; gs[here] = 0;
            pos[2 * i] = (uint8_t)fx; pos[2 * i + 1] = (uint8_t)fy;
          }
        } else err |= MG_ERR_BAD_ACTION;  // the reference raises for pickup/drop/toggle/done (multigrid.py:447)
      }
      const bool trunc = h.x >= p.max_steps;  // :470-471
      p.terminated[e] = term; p.truncated[e] = trunc;
      done = p.autoreset && (term || trunc);
      h.z = (int)r.ctr;
      if (err) atomicOr(p.status, err);
    }
  }
  if (tid < kGenE) s_done[tid] = done;
  const int any_done = __syncthreads_or(done);
  if (any_done) {  // same-step autoreset: terminal observation first, then restore the episode-start snapshot
    if (p.final_obs) generic_encode_tile(p, e0, n_here, p.final_obs, s_done, tid);
    __syncthreads();
    if (done) generic_reset_env(p, e0 + tid, h);
  }
  if (tid < n_here) p.hdr[e0 + tid] = h;
  __syncthreads();
  if (p.obs) generic_encode_tile(p, e0, n_here, p.obs, nullptr, tid);
}

int generic_tile_envs() { return kGenE; }

cudaError_t launch_generic(const GenericParams& p, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((p.N + kGenE - 1) / kGenE)); cfg.blockDim = dim3(kGenThreads); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = [] { const char* v = std::getenv("MG_PDL"); return !(v && v[0] == '0'); }();
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, generic_kernel, p);
}

}  // namespace mg
