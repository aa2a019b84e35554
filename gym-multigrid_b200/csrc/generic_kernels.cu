// generic_kernels.cu -- the base-class MultiGridEnv.step (multigrid.py:397-483) with DefaultWorld
// (world.py:33-52, encode_dim 6) and per-agent full-grid observations Grid.encode_for_agents (grid.py:254-284).
// No shipped env reaches this path; it is here for API completeness.
// Only still / left / right / forward are defined: any other action makes the reference evaluate
// `self.actions.available` (multigrid.py:447), which no action enum defines, and raise.
#include <cstdlib>

#include "mg_device.cuh"
#include "smem_config.h"
#include "generic_params.cuh"

namespace mg {

// One CTA = a tile of kGenE = 4 envs, 128 threads (8 x 256 measured 3 % slower: more warps idle behind warp 0's step phase).  The tile's two cell planes and agent positions are staged in shared memory (coalesced
// loads); warp 0 steps the envs, one lane per env, in the given agent order; then all threads encode
// [env][agent][cell] -> 6 bytes into shared memory (two cells per thread: three 32-bit stores per agent) and the tile's contiguous observation slab leaves as ONE TMA bulk
// store (full-line writes); the cell planes are written back coalesced.
constexpr int kGenE = 4, kGenThreads = 128;
constexpr int G_EMPTY = 1, G_DOOR = 4, G_GOAL = 8, G_AGENT = 10;  // DefaultWorld.OBJECT_TO_IDX (world.py:37-51)

struct GenSmem {
  uint8_t* obs;     // [kGenE][A][cells][6]
  uint8_t* cell;    // [kGenE][cells]
  uint8_t* state;   // [kGenE][cells]
  uint8_t* pos;     // [kGenE][A][2]
};
__host__ __device__ inline size_t gen_obs_bytes(int A, int cells) { return ((size_t)kGenE * A * cells * 6 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t gen_smem_bytes(int A, int cells) {
  return gen_obs_bytes(A, cells) + 2 * (((size_t)kGenE * cells + 15) & ~(size_t)15) + (size_t)kGenE * A * 2 + 16;
}

// all threads: encode_for_agents (grid.py:254-284) of the tile's envs (those with sel[el] != 0 when sel is given) into s.obs
__device__ __forceinline__ void generic_encode_tile(const GenericParams& p, const GenSmem& s, int n_here, const uint8_t* sel, int tid) {
  const int cells = p.cells, A = p.A;
  auto enc = [&](int idx, uint32_t& w0, uint32_t& w1, uint32_t& w2, bool& agent) {   // the three 16-bit halves of one cell's 6 bytes
    const uint32_t c = s.cell[idx], st = s.state[idx];
    const uint32_t type = c & 15u;
    w0 = type | ((c >> 4) << 8);
    w1 = type == G_DOOR ? st : 0u;                     // Door.encode object.py:238-259
    agent = type == G_AGENT;
    w2 = agent ? (st & 3u) : 0u;                       // Agent.encode agent.py:127-165 (carrying is always None here)
  };
  if ((cells & 1) == 0) {   // two cells per thread: 12 bytes per agent as three 32-bit stores (6-byte records, 4-byte aligned pairs)
    const int half = cells / 2, total = n_here * half;
    for (int q = tid; q < total; q += kGenThreads) {
      const int el = (int)__umulhi((uint32_t)q, p.half_magic);
      if (sel && !sel[el]) continue;
      const int i = 2 * (q - el * half), idx = el * cells + i;
      uint32_t a0, a1, a2, b0, b1, b2;
      bool aga, agb;
      enc(idx, a0, a1, a2, aga); enc(idx + 1, b0, b1, b2, agb);
      uint32_t* o = reinterpret_cast<uint32_t*>(s.obs + ((size_t)el * A * cells + i) * 6);
      for (int k = 0; k < A; ++k) {
        const int self = s.pos[(el * A + k) * 2] * p.H + s.pos[(el * A + k) * 2 + 1];                // the is_self plane
        const uint32_t sa = (aga && i == self) ? 0x100u : 0u, sb = (agb && i + 1 == self) ? 0x100u : 0u;
        o[0] = a0 | (a1 << 16); o[1] = (a2 | sa) | (b0 << 16); o[2] = b1 | ((b2 | sb) << 16);
        o += (cells * 6) / 4;
      }
    }
    return;
  }
  const int total = n_here * cells;
  for (int idx = tid; idx < total; idx += kGenThreads) {   // one thread per (env, cell): decoded once, written for every agent
    const int el = (int)__umulhi((uint32_t)idx, p.cells_magic);
    if (sel && !sel[el]) continue;
    const int i = idx - el * cells;
    uint32_t w0, w1, w2;
    bool agent;
    enc(idx, w0, w1, w2, agent);
    uint16_t* o = reinterpret_cast<uint16_t*>(s.obs + ((size_t)el * A * cells + i) * 6);
    for (int k = 0; k < A; ++k) {
      const bool self = agent && i == s.pos[(el * A + k) * 2] * p.H + s.pos[(el * A + k) * 2 + 1];   // the is_self plane
      o[0] = (uint16_t)w0; o[1] = (uint16_t)w1; o[2] = (uint16_t)(w2 | (self ? 0x100u : 0u));
      o += cells * 3;
    }
  }
}

// all threads: the tile's slab (or the selected envs of it) -> global; full tiles with an aligned destination use one TMA bulk store
__device__ __forceinline__ void generic_store_tile(const GenericParams& p, const GenSmem& s, uint8_t* dst_base, long long e0, int n_here,
                                                   const uint8_t* sel, int tid) {
  const size_t per_env = (size_t)p.A * p.cells * 6;
  uint8_t* dst = dst_base + (size_t)e0 * per_env;
  if (!sel && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
    const uint32_t bytes = (uint32_t)(n_here * per_env), bulk = bytes & ~15u;
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0 && bulk) { tma_store_1d(dst, s.obs, bulk); tma_commit(); }
    for (uint32_t i = bulk + tid; i < bytes; i += kGenThreads) dst[i] = s.obs[i];
    if (tid == 0) tma_wait_read_all();
    __syncthreads();
    return;
  }
  __syncthreads();
  for (int el = 0; el < n_here; ++el) {
    if (sel && !sel[el]) continue;
    const uint16_t* src = reinterpret_cast<const uint16_t*>(s.obs + el * per_env);
    uint16_t* d = reinterpret_cast<uint16_t*>(dst + el * per_env);
    for (int i = tid; i < (int)(per_env / 2); i += kGenThreads) d[i] = src[i];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kGenThreads) generic_kernel(const __grid_constant__ GenericParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint8_t s_done[kGenE];
  const int tid = threadIdx.x, A = p.A, cells = p.cells, H = p.H;
  const long long e0 = (long long)blockIdx.x * kGenE;
  const int n_here = (int)min((long long)kGenE, p.N - e0);
  GenSmem s;
  s.obs = smem_raw;
  s.cell = smem_raw + gen_obs_bytes(A, cells);
  s.state = s.cell + (((size_t)kGenE * cells + 15) & ~(size_t)15);
  s.pos = s.state + (((size_t)kGenE * cells + 15) & ~(size_t)15);
  pdl_launch_dependents();
  pdl_wait();
  // ---- stage the tile (reset: from the episode-start snapshot planes of the selected envs)
  // full step tiles whose plane slabs are 16-byte multiples move as 128-bit words (e0 * cells is then a multiple of 16 too)
  const bool vec = p.op == 1 && n_here == kGenE && ((kGenE * cells) & 15) == 0 &&
                   ((reinterpret_cast<uintptr_t>(p.gcell) | reinterpret_cast<uintptr_t>(p.gstate)) & 15) == 0;
  if (vec) {
    const int n16 = kGenE * cells / 16;
    const uint4* gc4 = reinterpret_cast<const uint4*>(p.gcell + e0 * cells);
    const uint4* gs4 = reinterpret_cast<const uint4*>(p.gstate + e0 * cells);
    for (int i = tid; i < 2 * n16; i += kGenThreads) {
      if (i < n16) reinterpret_cast<uint4*>(s.cell)[i] = gc4[i];
      else reinterpret_cast<uint4*>(s.state)[i - n16] = gs4[i - n16];
    }
  } else
  for (int i = tid; i < n_here * cells; i += kGenThreads) {
    const int el = (int)__umulhi((uint32_t)i, p.cells_magic);
    const bool from_init = p.op == 0 && (!p.reset_mask || p.reset_mask[e0 + el]);
    s.cell[i] = from_init ? p.icell[e0 * cells + i] : p.gcell[e0 * cells + i];
    s.state[i] = from_init ? p.istate[e0 * cells + i] : p.gstate[e0 * cells + i];
  }
  for (int i = tid; i < n_here * A * 2; i += kGenThreads) {
    const int el = i / (A * 2);
    const bool from_init = p.op == 0 && (!p.reset_mask || p.reset_mask[e0 + el]);
    s.pos[i] = from_init ? p.ipos[e0 * A * 2 + i] : p.pos[e0 * A * 2 + i];
  }
  if (tid < kGenE) s_done[tid] = 0;
  __syncthreads();

  bool done = false;
  int4 h = make_int4(0, 0, 0, 0);
  if (tid < n_here) {
    const long long e = e0 + tid;
    h = p.hdr[e];
    if (p.op == 0) {
      if (!p.reset_mask || p.reset_mask[e]) { h.x = 0; h.w += 1; }
    } else {
      uint8_t* gc = s.cell + tid * cells; uint8_t* gs = s.state + tid * cells; uint8_t* pos = s.pos + tid * A * 2;
      Rng<1> r;
      r.open_philox(p.seed, p.env_id_base + (unsigned long long)e, (uint32_t)h.z);
      uint32_t order = 0x76543210u;   // one nibble per agent (A <= 8)
      if (p.order) { order = 0; for (int i = 0; i < A; ++i) order |= (uint32_t)(p.order[e * A + i] & 7) << (4 * i); }
      else
        for (int i = A - 1; i > 0; --i) {
          const int j = (int)__umulhi(r.u32(), (uint32_t)(i + 1));
          const uint32_t x = ((order >> (4 * i)) ^ (order >> (4 * j))) & 15u;
          order ^= (x << (4 * i)) | (x << (4 * j));
        }
      h.x += 1;  // multigrid.py:400
      bool term = false;
      int err = 0;
      for (int i = 0; i < A; ++i) p.rewards[e * A + i] = 0.0;
      for (int k = 0; k < A; ++k) {  // for i in order :408
        const int i = (int)((order >> (4 * k)) & 15u), a = p.actions[e * A + i];
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
      s_done[tid] = done;
    }
  }
  const int any_done = __syncthreads_or(done);
  if (any_done) {  // same-step autoreset: terminal observation first, then restore the episode-start snapshot
    if (p.final_obs) {
      generic_encode_tile(p, s, n_here, s_done, tid);
      generic_store_tile(p, s, p.final_obs, e0, n_here, s_done, tid);
    }
    for (int el = 0; el < n_here; ++el) {
      if (!s_done[el]) continue;
      for (int i = tid; i < cells; i += kGenThreads) { s.cell[el * cells + i] = p.icell[(e0 + el) * cells + i]; s.state[el * cells + i] = p.istate[(e0 + el) * cells + i]; }
      for (int i = tid; i < A * 2; i += kGenThreads) s.pos[el * A * 2 + i] = p.ipos[(e0 + el) * A * 2 + i];
    }
    if (done) { h.x = 0; h.w += 1; }
    __syncthreads();
  }
  if (tid < n_here) p.hdr[e0 + tid] = h;
  // ---- state write-back (coalesced) and the observation
  if (vec) {
    const int n16 = kGenE * cells / 16;
    uint4* gc4 = reinterpret_cast<uint4*>(p.gcell + e0 * cells);
    uint4* gs4 = reinterpret_cast<uint4*>(p.gstate + e0 * cells);
    for (int i = tid; i < 2 * n16; i += kGenThreads) {
      if (i < n16) gc4[i] = reinterpret_cast<const uint4*>(s.cell)[i];
      else gs4[i - n16] = reinterpret_cast<const uint4*>(s.state)[i - n16];
    }
  } else
  for (int i = tid; i < n_here * cells; i += kGenThreads) { p.gcell[e0 * cells + i] = s.cell[i]; p.gstate[e0 * cells + i] = s.state[i]; }
  for (int i = tid; i < n_here * A * 2; i += kGenThreads) p.pos[e0 * A * 2 + i] = s.pos[i];
  if (p.obs) {
    generic_encode_tile(p, s, n_here, nullptr, tid);
    generic_store_tile(p, s, p.obs, e0, n_here, nullptr, tid);
  }
}

int generic_tile_envs() { return kGenE; }
size_t generic_smem_bytes(int A, int cells) { return gen_smem_bytes(A, cells); }

cudaError_t configure_generic_kernel(int A, int cells) {
  return raise_smem_limit((const void*)generic_kernel, (size_t)gen_smem_bytes(A, cells));
}

cudaError_t launch_generic(const GenericParams& p, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((p.N + kGenE - 1) / kGenE)); cfg.blockDim = dim3(kGenThreads); cfg.stream = st;
  cfg.dynamicSmemBytes = gen_smem_bytes(p.A, p.cells);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = [] { const char* v = std::getenv("MG_PDL"); return !(v && v[0] == '0'); }();
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, generic_kernel, p);
}

}  // namespace mg
