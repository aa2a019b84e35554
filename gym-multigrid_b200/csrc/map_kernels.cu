// map_kernels.cu -- sm_100a kernels for the static-map families:
//   MazeSingleAgentEnv.step / reset  (envs/maze.py:180-219, 245-260, 271-377)
//   CtFMvNEnv.step / reset           (envs/ctf.py:998-1075, 1137-1163, 1184-1251, 1292-1433)
//
// Per-env state is a few bytes (agent positions, dirs, flags, a 16-byte header); the map is shared and
// lives in handle-owned device tables.  One thread per env runs the (sequential, order-dependent) agent
// loop; the observation - the static map with the agents drawn on top - is then written for the whole
// tile of 128 envs cooperatively: every thread streams 16-byte chunks of the map's "super-period"
// (lcm(cells, 16) bytes, staged in shared memory by one TMA bulk copy) into the tile's contiguous obs slab,
// and after a block barrier each env's thread patches its agents' cells.
#include <cstdlib>

#include "mg_device.cuh"
#include "map_params.cuh"

namespace mg {

constexpr int kMapE = 128;  // envs per CTA = threads per CTA


// MazeWorld / CtfWorld codes (world.py:66-91)
constexpr int MZ_AGENT = 1, MZ_FLAG = 2, MZ_OBSTACLE = 3;
constexpr int CT_BLUE_TERR = 0, CT_RED_TERR = 1, CT_BLUE_AGENT = 2, CT_RED_AGENT = 3, CT_BLUE_FLAG = 4, CT_RED_FLAG = 5,
              CT_OBSTACLE = 6;

// CtfActions / MazeActions: 0 stay, 1 left (0,-1), 2 down (-1,0), 3 right (0,+1), 4 up (+1,0)  (agent.py:54-67)
__device__ __forceinline__ void action_delta(int a, int& dx, int& dy) {
  dx = (a == 4) - (a == 2);
  dy = (a == 3) - (a == 1);
}
// DIR_TO_VEC (constants.py:65-74); Agent.move leaves dir alone when no vector matches (agent.py:176-183)
__device__ __forceinline__ int dir_of(int dx, int dy, int old) {
  if (dx == 1 && dy == 0) return 0;
  if (dx == 0 && dy == 1) return 1;
  if (dx == -1 && dy == 0) return 2;
  if (dx == 0 && dy == -1) return 3;
  return old;
}

template <int MODE>
__device__ __forceinline__ int below(Rng<MODE>& r, int n) { return (int)__umulhi(r.u32(), (uint32_t)n); }

// our Philox-mode stand-in for np_random.choice(len, k, replace=False)
template <int MODE>
__device__ __forceinline__ void sample_distinct(Rng<MODE>& r, int len, int k, int* out) {
  for (int i = 0; i < k; ++i) {
    for (;;) {
      const int v = below(r, len);
      bool dup = false;
      for (int j = 0; j < i; ++j) dup |= out[j] == v;
      if (!dup) { out[i] = v; break; }
    }
  }
}

// per-thread view of one env's agents in shared memory
struct Agents {
  uint8_t* x; uint8_t* y; uint8_t* dir; uint8_t* fl;  // each [n], stride 1
};

template <int FAMILY, int MODE>
__device__ __forceinline__ void reset_one(const MapParams& p, long long e, Agents ag, int4& h, Rng<MODE>& r) {
  const int S = p.S;
  if (FAMILY == MG_FAMILY_MAZE) {  // maze.py:202-205: agent on a random background cell, dir 3
    const int idx = MODE == 0 ? p.start_index[e] : below(r, p.n_background);
    const int cell = p.background[idx];
    ag.x[0] = (uint8_t)(cell / S); ag.y[0] = (uint8_t)(cell % S); ag.dir[0] = 3; ag.fl[0] = 0;
  } else {  // ctf.py:1033-1048
    int bp[MG_MAX_MAP_AGENTS], rp[MG_MAX_MAP_AGENTS];
    if (MODE == 0) {
      for (int i = 0; i < p.nb; ++i) bp[i] = p.blue_place[e * p.nb + i];
      for (int i = 0; i < p.nr; ++i) rp[i] = p.red_place[e * p.nr + i];
    } else {
      sample_distinct(r, p.len_blue, p.nb, bp);
      sample_distinct(r, p.len_red, p.nr, rp);
    }
    for (int i = 0; i < p.n; ++i) {
      const int cell = i < p.nb ? p.blue_terr[bp[i]] : p.red_terr[rp[i - p.nb]];
      ag.x[i] = (uint8_t)(cell / S); ag.y[i] = (uint8_t)(cell % S); ag.dir[i] = 3;
      ag.fl[i] = 0;  // a fresh env instance (the reference never clears terminated/collided on reset)
    }
  }
  h.x = 0; h.w += 1;  // step_count = 0 (multigrid.py:141); episode counter
}

template <int MODE>
__device__ __forceinline__ void maze_step_one(const MapParams& p, int a, Agents ag, int4& h, double& rew, bool& term,
                                              bool& trunc, int& err) {
  const int S = p.S;
  h.x += 1;  // maze.py:334
  if (a < 0 || a > 4) err |= MG_ERR_BAD_ACTION;  // reference: ValueError (maze.py:286)
  else {  // _move_agent maze.py:271-307
    int dx, dy;
    action_delta(a, dx, dy);
    const int ox = ag.x[0], oy = ag.y[0], nx = ox + dx, ny = oy + dy;
    if (!(nx < 0 || ny < 0 || nx >= S || ny >= S)) {
      // every cell holds an object: background Floor / Flag overlap, Obstacle overlaps iff penalty != 0
      // (object.py:200-201), the agent's own cell (stay) does not (object.py:38-40)
      const int code = p.field_map[nx * S + ny];
      const bool self = (dx == 0 && dy == 0);
      if (!self && (code != MZ_OBSTACLE || p.obstacle_penalty != 0)) {
        ag.dir[0] = (uint8_t)dir_of(dx, dy, ag.dir[0]);
        ag.x[0] = (uint8_t)nx; ag.y[0] = (uint8_t)ny;
      }
    }
  }
  term = false; trunc = h.x >= p.max_steps;  // :346-347
  rew = 0.0;
  const int here = p.field_map[ag.x[0] * S + ag.y[0]];
  if (here == MZ_FLAG) { rew += p.flag_reward; term = true; }                                        // :354-356
  if (p.obstacle_penalty != 0 && here == MZ_OBSTACLE) { rew -= p.obstacle_penalty; term = true; }    // :360-363
  rew -= p.step_penalty;                                                                             // :371
}

template <int MODE>
__device__ __forceinline__ void ctf_step_one(const MapParams& p, long long e, const int8_t* blue_act, Agents ag, int4& h,
                                             Rng<MODE>& r, double& rew, bool& term, bool& trunc, int& err) {
  const int S = p.S, nb = p.nb, nr = p.nr, n = p.n;
  h.x += 1;  // ctf.py:1295
  int act[MG_MAX_MAP_AGENTS], order[MG_MAX_MAP_AGENTS];
  for (int i = 0; i < nb; ++i) act[i] = blue_act[i];
  for (int k = 0; k < nr; ++k)  // RwPolicy.act for EVERY red agent, defeated or not (:1297-1301)
    act[nb + k] = MODE == 0 ? p.red_actions[e * nr + k] : below(r, 5);
  if (p.variant_1v1) {  // Ctf1v1Env._move_agents: blue, then red (ctf.py:503-510)
    order[0] = 0; order[1] = 1;
  } else if (MODE == 0) {
    for (int i = 0; i < n; ++i) order[i] = p.order[e * n + i];
  } else {  // np_random.shuffle stand-in: Fisher-Yates
    for (int i = 0; i < n; ++i) order[i] = i;
    for (int i = n - 1; i > 0; --i) { const int j = below(r, i + 1), t = order[i]; order[i] = order[j]; order[j] = t; }
  }
  for (int k = 0; k < n; ++k) {  // _move_agents :1240-1251
    const int i = order[k];
    if (ag.fl[i] & 1) continue;  // "Defeated agent doesn't move, sadly."
    const int a = act[i];
    if (a < 0 || a > 4) { err |= MG_ERR_BAD_ACTION; continue; }
    int dx, dy;
    action_delta(a, dx, dy);
    const int nx = ag.x[i] + dx, ny = ag.y[i] + dy;  // _move_agent :1184-1238
    if (nx < 0 || ny < 0 || nx >= S || ny >= S) continue;
    bool occupied = false;  // an agent object (alive, defeated, or itself when staying) sits on the cell
    for (int j = 0; j < n; ++j) occupied |= (ag.x[j] == nx && ag.y[j] == ny);
    if (occupied) { if (p.obstacle_penalty != 0 && !p.variant_1v1) ag.fl[i] |= 2; continue; }  // :1231-1236 (1v1 has no collided logic, :498-501)
    if (p.field_map[nx * S + ny] == CT_OBSTACLE && p.obstacle_penalty == 0) continue;  // Obstacle.can_overlap()
    ag.dir[i] = (uint8_t)dir_of(dx, dy, ag.dir[i]);  // Agent.move agent.py:167-200
    ag.x[i] = (uint8_t)nx; ag.y[i] = (uint8_t)ny;
  }
  term = false; trunc = h.x >= p.max_steps;  // :1310-1311
  rew = 0.0;
  if (p.obstacle_penalty != 0) {  // :1316-1332 (collided is never cleared)
    for (int i = 0; i < nb; ++i) if (ag.fl[i] & 2) { rew -= p.obstacle_penalty; ag.fl[i] |= 1; }
    for (int i = nb; i < n; ++i) if (ag.fl[i] & 2) ag.fl[i] |= 1;
  }
  for (int i = 0; i < nb; ++i) if (ag.x[i] * S + ag.y[i] == p.red_flag) { rew += p.flag_reward; term = true; }   // :1335-1344
  for (int i = nb; i < n; ++i) if (ag.x[i] * S + ag.y[i] == p.blue_flag) { rew -= p.flag_reward; term = true; }  // :1347-1356
  int nbattle = 0;
  for (int b = 0; b < nb; ++b)  // np.where(distances <= battle_range): row-major, blue-major (:1368-1377)
    for (int q = 0; q < nr; ++q) {
      const int ddx = (int)ag.x[b] - (int)ag.x[nb + q], ddy = (int)ag.y[b] - (int)ag.y[nb + q];
      if (!(sqrt((double)(ddx * ddx + ddy * ddy)) <= p.battle_range)) continue;  // np.linalg.norm of an int vector
      if ((ag.fl[b] & 1) || (ag.fl[nb + q] & 1)) continue;                        // :1380-1383
      const int cb = p.field_map[ag.x[b] * S + ag.y[b]], cr = p.field_map[ag.x[nb + q] * S + ag.y[nb + q]];
      const bool bh = (cb == CT_BLUE_TERR || cb == CT_BLUE_FLAG), rh = (cr == CT_RED_TERR || cr == CT_RED_FLAG);
      bool blue_win;
      if (MODE == 0) {
        blue_win = nbattle < p.KB ? p.blue_win[e * p.KB + nbattle] != 0 : false;
        if (nbattle >= p.KB) err |= MG_ERR_TRACE_OVERFLOW;
      } else {  // :1392-1407
        const double pb = (bh && !rh) ? p.randomness : ((!bh && rh) ? 1.0 - p.randomness : 0.5);
        blue_win = (double)r.u32() * (1.0 / 4294967296.0) < pb;
      }
      ++nbattle;
      if (blue_win) { rew += p.battle_reward; ag.fl[nb + q] |= 1; }                 // :1409-1418
      else if (p.variant_1v1) { rew -= p.battle_reward; term = true; }              // 1v1: losing ends the episode (ctf.py:629-632)
      else { rew -= p.battle_reward; ag.fl[b] |= 1; }
    }
  if (MODE == 0 && p.battles_used) p.battles_used[e] = nbattle;
  bool all_dead = true;
  for (int i = 0; i < nb; ++i) all_dead &= (ag.fl[i] & 1) != 0;
  if (all_dead) term = true;           // :1423
  rew = __dsub_rn(rew, __dmul_rn(p.step_penalty, (double)nb));  // :1428 -- two roundings like the reference, never an FMA
}

// value an agent shows in the "map" observation
template <int FAMILY>
__device__ __forceinline__ int agent_code(const MapParams& p, int i, int fl) {
  if (FAMILY == MG_FAMILY_MAZE) return MZ_AGENT;                                      // maze.py:256-258
  return (fl & 1) ? CT_OBSTACLE : (i < p.nb ? CT_BLUE_AGENT : CT_RED_AGENT);          // ctf.py:1157-1161
}
template <int FAMILY>
__device__ __forceinline__ int obs_index(const MapParams& p, int x, int y) {
  return FAMILY == MG_FAMILY_MAZE ? x * p.S + y : y * p.S + x;  // Maze [x][y]; CtF returns encoded_map.T
}

template <typename T>
__device__ __forceinline__ void put(void* base, long long idx, int v) { static_cast<T*>(base)[idx] = (T)v; }
__device__ __forceinline__ void put_obs(const MapParams& p, void* base, long long idx, int v) {
  if (p.obs_dtype == MG_OBS_U8) put<uint8_t>(base, idx, v);
  else if (p.family == MG_FAMILY_MAZE) put<double>(base, idx, v);
  else put<long long>(base, idx, v);
}

template <int FAMILY, int MODE>
__global__ void __launch_bounds__(kMapE) map_kernel(const __grid_constant__ MapParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint8_t s_done[kMapE];
  const int tid = threadIdx.x, n = p.n, cells = p.cells;
  uint8_t* s_period = smem_raw;                                 // [L]
  uint8_t* s_ag = smem_raw + p.L;                               // [4][kMapE][n]: x, y, dir, flags
  const long long e0 = (long long)blockIdx.x * kMapE;
  const int n_here = (int)min((long long)kMapE, p.N - e0);
  const long long e = e0 + tid;

  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();
  if (tid == 0) { mbar_expect_tx(&bar, (uint32_t)p.L); tma_load_1d(s_period, p.obs_period, (uint32_t)p.L, &bar); }

  Agents ag;
  ag.x = s_ag + (size_t)tid * n; ag.y = ag.x + (size_t)kMapE * n; ag.dir = ag.y + (size_t)kMapE * n; ag.fl = ag.dir + (size_t)kMapE * n;
  bool done = false;
  int err = 0;
  int4 h = make_int4(0, 0, 0, 0);
  Rng<MODE> r;
  r.open_trace(nullptr, 0);
  if (tid < n_here) {
    h = p.hdr[e];
    for (int i = 0; i < n; ++i) {
      ag.x[i] = p.pos[(e * n + i) * 2]; ag.y[i] = p.pos[(e * n + i) * 2 + 1];
      ag.dir[i] = p.dir[e * n + i]; ag.fl[i] = p.flags[e * n + i];
    }
    if (MODE == 1) r.open_philox(p.seed, p.env_id_base + (unsigned long long)e, (uint32_t)h.z);
    if (p.op == 0) {
      if (!p.reset_mask || p.reset_mask[e]) reset_one<FAMILY, MODE>(p, e, ag, h, r);
    } else {
      double rew; bool term, trunc;
      if (FAMILY == MG_FAMILY_MAZE) maze_step_one<MODE>(p, p.actions[e], ag, h, rew, term, trunc, err);
      else ctf_step_one<MODE>(p, e, p.actions + e * p.nb, ag, h, r, rew, term, trunc, err);
      p.rewards[e] = rew; p.terminated[e] = term; p.truncated[e] = trunc;
      done = p.autoreset && (term || trunc);
    }
    // same-step autoreset; when the caller wants final_observation the reset waits until it has been drawn
    if (done && !p.final_obs) reset_one<FAMILY, MODE>(p, e, ag, h, r);
  }
  s_done[tid] = done;
  const int any_final = p.final_obs ? __syncthreads_or(done) : 0;
  mbar_wait(&bar, 0);

  // ---- rare path: terminal observations of finished envs, then their reset
  if (any_final) {
    for (int j = 0; j < n_here; ++j) {
      if (!s_done[j]) continue;
      for (int i = tid; i < cells; i += kMapE) put_obs(p, p.final_obs, (e0 + j) * cells + i, s_period[i]);
    }
    __syncthreads();
    if (done) {
      for (int i = 0; i < n; ++i)
        put_obs(p, p.final_obs, e * cells + obs_index<FAMILY>(p, ag.x[i], ag.y[i]), agent_code<FAMILY>(p, i, ag.fl[i]));
      reset_one<FAMILY, MODE>(p, e, ag, h, r);
    }
  }
  if (tid < n_here) {
    if (MODE == 1) h.z = (int)r.ctr;
    p.hdr[e] = h;
    if (err) atomicOr(p.status, err);
  }

  // ---- state write-back
  if (tid < n_here) {
    for (int i = 0; i < n; ++i) {
      p.pos[(e * n + i) * 2] = ag.x[i]; p.pos[(e * n + i) * 2 + 1] = ag.y[i];
      p.dir[e * n + i] = ag.dir[i]; p.flags[e * n + i] = ag.fl[i];
    }
  }

  // ---- observation: static map for the whole tile (coalesced 16-byte stores), then the agents on top
  if (p.obs) {
    const long long slab = (long long)n_here * cells;  // elements in this tile's obs slab
    if (p.obs_dtype == MG_OBS_U8) {
      const int L16 = p.L / 16;
      const long long chunks = slab / 16;
      uint4* dst = reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.obs) + e0 * cells);  // e0*cells is a multiple of L
      const uint4* src = reinterpret_cast<const uint4*>(s_period);
      int m = tid % L16;
      const int step = kMapE % L16;
      for (long long c = tid; c < chunks; c += kMapE) {
        dst[c] = src[m];
        m += step; if (m >= L16) m -= L16;
      }
      for (long long k = chunks * 16 + tid; k < slab; k += kMapE)  // ragged tail of the last tile
        static_cast<uint8_t*>(p.obs)[e0 * cells + k] = s_period[k % p.L];
    } else {
      int m = (2 * tid) % p.L;
      const int step = (2 * kMapE) % p.L;
      for (long long k = 2 * tid; k + 1 < slab + 1; k += 2 * kMapE) {
        const int v0 = s_period[m], v1 = s_period[m + 1 < p.L ? m + 1 : 0];
        if (k + 1 < slab) {
          if (p.family == MG_FAMILY_MAZE) reinterpret_cast<double2*>(static_cast<double*>(p.obs) + e0 * cells)[k / 2] = make_double2(v0, v1);
          else reinterpret_cast<longlong2*>(static_cast<long long*>(p.obs) + e0 * cells)[k / 2] = make_longlong2(v0, v1);
        } else if (k < slab) {
          put_obs(p, p.obs, e0 * cells + k, v0);
        }
        m += step; if (m >= p.L) m -= p.L;
      }
    }
    __syncthreads();  // the tile's static fill is ordered before the per-env patches
    if (tid < n_here)
      for (int i = 0; i < n; ++i)  // agents in index order: later agents overwrite earlier ones (ctf.py:1157-1161)
        put_obs(p, p.obs, e * cells + obs_index<FAMILY>(p, ag.x[i], ag.y[i]), agent_code<FAMILY>(p, i, ag.fl[i]));
  }
}

static bool map_pdl_enabled() {
  static const bool on = [] { const char* v = std::getenv("MG_PDL"); return !(v && v[0] == '0'); }();
  return on;
}

size_t map_smem_bytes(int L, int n) { return (size_t)L + (size_t)4 * kMapE * n + 16; }
int map_tile_envs() { return kMapE; }

template <int FAMILY, int MODE>
static cudaError_t launch_one(const MapParams& p, cudaStream_t st) {
  const size_t smem = map_smem_bytes(p.L, p.n);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((p.N + kMapE - 1) / kMapE)); cfg.blockDim = dim3(kMapE);
  cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = map_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, map_kernel<FAMILY, MODE>, p);
}

cudaError_t configure_map_kernels(int L, int n) {
  const int smem = (int)map_smem_bytes(L, n);
  cudaError_t e;
  if ((e = cudaFuncSetAttribute((const void*)map_kernel<MG_FAMILY_MAZE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute((const void*)map_kernel<MG_FAMILY_MAZE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute((const void*)map_kernel<MG_FAMILY_CTF, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
  return cudaFuncSetAttribute((const void*)map_kernel<MG_FAMILY_CTF, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

cudaError_t launch_map(const MapParams& p, cudaStream_t st) {
  if (p.family == MG_FAMILY_MAZE) return p.rng_mode == 0 ? launch_one<MG_FAMILY_MAZE, 0>(p, st) : launch_one<MG_FAMILY_MAZE, 1>(p, st);
  return p.rng_mode == 0 ? launch_one<MG_FAMILY_CTF, 0>(p, st) : launch_one<MG_FAMILY_CTF, 1>(p, st);
}

}  // namespace mg
