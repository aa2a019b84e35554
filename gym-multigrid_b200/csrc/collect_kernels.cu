// collect_kernels.cu -- sm_100a kernels for the Collect family hot path:
//   collect_step_kernel   CollectGameEnv.step + Grid.encode fused   (collect_game.py:183-214, grid.py:223-252)
//   collect_reset_kernel  MultiGridEnv.reset + _gen_grid variants   (multigrid.py:114-153, collect_game.py:107-119,236-399)
//   encode3_kernel        Grid.encode alone                         (grid.py:223-252)
//
// Shape of every kernel: one CTA owns a tile of E consecutive envs.  The tile's packed grids
// (E * W*H bytes, contiguous in HBM) are pulled into shared memory with one TMA bulk copy
// (cp.async.bulk + mbarrier); one thread per env then walks that env's agents IN THE GIVEN
// ORDER on the shared-memory grid (the reference's sequential semantics: order-dependent
// blocking, respawn between pickup and move); finally all threads expand the packed cells to
// the 3-byte (OBJECT_IDX, COLOR_IDX, STATE) encoding in shared memory and one thread issues
// two TMA bulk stores: the obs slab (E * 3*W*H contiguous bytes) and the updated grid slab.
#include "mg_device.cuh"

namespace mg {

#define GCELL(g, H, x, y) (g)[(x) * (H) + (y)]

// MultiGridEnv.place_obj (multigrid.py:282-339): rejection-sample an EMPTY cell in
// [top, min(top + size, dim - 1)] (inclusive), x drawn before y.
template <int MODE>
__device__ __forceinline__ void place_obj(const CollectParams& p, uint8_t* g, Rng<MODE>& r, uint8_t code, int tx, int ty,
                                          int sx, int sy, int& ox, int& oy) {
  const int hx = min(tx + sx, p.W - 1), hy = min(ty + sy, p.H - 1);
  for (;;) {
    const int x = r.rand_int(tx, hx);
    const int y = r.rand_int(ty, hy);
    ox = x; oy = y;
    if (MODE == 0 && (r.err & MG_ERR_TRACE_OVERFLOW)) return;  // trace exhausted: leave the grid untouched
    if (GCELL(g, p.H, x, y) != 0) continue;
    GCELL(g, p.H, x, y) = code;
    return;
  }
}

// CollectGameEnv._respawn (collect_game.py:129-130) / CollectGameQuadrantsRespawn._respawn (:401-409)
template <int MODE>
__device__ __noinline__ void respawn(const CollectParams& p, uint8_t* g, Rng<MODE>& r, int colour) {
  int x, y;
  if (p.layout == MG_LAYOUT_QUADRANTS_RESPAWN) {
    const int q = colour < 3 ? colour : 0;
    const int tx = q == 0 ? 0 : p.W / 2 - 1, ty = q == 1 ? p.H / 2 - 1 : 0;
    place_obj<MODE>(p, g, r, cell(T_BALL, colour, 0), tx, ty, p.W / 2 + 1, p.H / 2 + 1, x, y);
  } else {
    place_obj<MODE>(p, g, r, cell(T_BALL, colour, 0), 0, 0, p.W, p.H, x, y);
  }
}

// CollectGameEnv.reset (collect_game.py:107-119) + the layout's _gen_grid.  `g`, `pos` live in smem.
template <int MODE>
__device__ __noinline__ void reset_env(const CollectParams& p, uint8_t* g, uint8_t* pos, Rng<MODE>& r) {
  const int W = p.W, H = p.H, A = p.A, nb = p.nb;
  for (int i = 0; i < p.cells; ++i) g[i] = 0;
  for (int i = 0; i < W; ++i) { GCELL(g, H, i, 0) = WALL_GREY; GCELL(g, H, i, H - 1) = WALL_GREY; }  // horz_wall grid.py:66-78
  for (int j = 0; j < H; ++j) { GCELL(g, H, 0, j) = WALL_GREY; GCELL(g, H, W - 1, j) = WALL_GREY; }  // vert_wall grid.py:80-89
  int x, y;
  if (p.layout == MG_LAYOUT_EVEN_DIST) {  // collect_game.py:236-259
    const int per = p.num_balls / nb;
    for (int t = 0; t < nb; ++t)
      for (int b = 0; b < per; ++b) place_obj<MODE>(p, g, r, cell(T_BALL, p.ball_colour[t], 0), 0, 0, W, H, x, y);
    for (int i = 0; i < A; ++i) {  // place_agent(a): anywhere empty (multigrid.py:364-369)
      place_obj<MODE>(p, g, r, p.agent_code[i], 0, 0, W, H, x, y);
      pos[2 * i] = (uint8_t)x; pos[2 * i + 1] = (uint8_t)y;
    }
  } else if (p.layout == MG_LAYOUT_QUADRANTS) {  // collect_game.py:266-300
    const int per = p.num_balls / nb;
    for (int t = 0; t < nb; ++t) {
      const int tx = (t == 1 || t == 2) ? W / 2 - 1 : 0, ty = t == 1 ? H / 2 - 1 : (t == 3 ? H / 2 : 0);
      for (int b = 0; b < per; ++b)
        place_obj<MODE>(p, g, r, cell(T_BALL, p.ball_colour[t], 0), tx, ty, W / 2 - 1, H / 2 - 1, x, y);
    }
    for (int i = 0; i < A; ++i) {  // place_agent(a, pos): overwrites (put_obj multigrid.py:341-348)
      GCELL(g, H, 1 + i, H - 2) = p.agent_code[i];
      pos[2 * i] = (uint8_t)(1 + i); pos[2 * i + 1] = (uint8_t)(H - 2);
    }
  } else if (p.layout == MG_LAYOUT_ROOMS) {  // collect_game.py:306-362 (`width` on both axes)
    const int ws = W / 2 - 1, m = W / 2;
    for (int i = 0; i < ws; ++i) {
      GCELL(g, H, i, m) = WALL_GREY; GCELL(g, H, W - ws + i, m) = WALL_GREY;
      GCELL(g, H, m, i) = WALL_GREY; GCELL(g, H, m, W - ws + i) = WALL_GREY;
    }
    for (int i = 0; i < A; ++i) {  // _rand_elem(possible_coords) -> _rand_int(0, 4)
      const int k = r.rand_int(0, 4);
      const int cx = k == 0 ? m : (k <= 2 ? m - 1 : m + 1);
      const int cy = k == 0 ? m : ((k == 1 || k == 4) ? m - 1 : m + 1);
      GCELL(g, H, cx, cy) = p.agent_code[i];  // a second agent on the same cell overwrites the first
      pos[2 * i] = (uint8_t)cx; pos[2 * i + 1] = (uint8_t)cy;
    }
    const int ps = W / 2 - 1;
    const int num_ball = (int)nearbyint((double)p.num_balls / nb);  // python round(): half-to-even
    int index = 0, tx = 0, ty = 0;
    for (int ball = 0; ball < p.num_balls; ++ball) {
      if (ball % num_ball == 0) {
        index = ball / num_ball;
        tx = (index == 1 || index == 2) ? m + 1 : 0;
        ty = (index == 1 || index == 3) ? m + 1 : 0;
        // the extra ball of this colour in partition 3 (:349-355)
        place_obj<MODE>(p, g, r, cell(T_BALL, p.ball_colour[index], 0), 0, m + 1, ps, ps, x, y);
      }
      place_obj<MODE>(p, g, r, cell(T_BALL, p.ball_colour[index], 0), tx, ty, ps, ps, x, y);
    }
  } else {  // MG_LAYOUT_QUADRANTS_RESPAWN, collect_game.py:376-399
    const int per = p.num_balls / 3;
    int index = 0, tx = 0, ty = 0;
    for (int ball = 0; ball < p.num_balls; ++ball) {
      if (ball % per == 0) {
        index = ball / per;
        tx = index == 0 ? 0 : W / 2 - 1;
        ty = index == 1 ? H / 2 - 1 : 0;
      }
      // Ball(self.world, index, 1): the colour IS the partition index (:391)
      place_obj<MODE>(p, g, r, cell(T_BALL, index, 0), tx, ty, W / 2 + 1, H / 2 + 1, x, y);
    }
    for (int i = 0; i < A; ++i) {
      GCELL(g, H, 1 + i, H - 2) = p.agent_code[i];
      pos[2 * i] = (uint8_t)(1 + i); pos[2 * i + 1] = (uint8_t)(H - 2);
    }
  }
}

// shared-memory carve-up of one tile (all offsets 16-byte aligned because E % 16 == 0)
struct TileSmem {
  uint8_t* grid;   // [E][cells]
  uint8_t* obs;    // [E][cells][3]
  double* rew;     // [E][A]
  uint8_t* pos;    // [E][A][2]
  int8_t* act;     // [E][A]
  uint8_t* ord;    // [E][A]
  uint8_t* done;   // [E]
};
__host__ __device__ inline size_t tile_smem_bytes(int E, int cells, int A) {
  return (size_t)E * cells * 4 + (size_t)E * A * 8 + (size_t)E * A * 4 + E + 16;
}
__device__ __forceinline__ TileSmem carve(uint8_t* base, int E, int cells, int A) {
  TileSmem s;
  s.grid = base;
  s.obs = base + (size_t)E * cells;
  s.rew = reinterpret_cast<double*>(base + (size_t)E * cells * 4);
  s.pos = reinterpret_cast<uint8_t*>(s.rew + (size_t)E * A);
  s.act = reinterpret_cast<int8_t*>(s.pos + (size_t)E * A * 2);
  s.ord = reinterpret_cast<uint8_t*>(s.act + (size_t)E * A);
  s.done = s.ord + (size_t)E * A;
  return s;
}

// all threads: packed grid tile -> 3-byte encoding, both in shared memory
template <int THREADS>
__device__ __forceinline__ void expand_tile(const uint8_t* s_grid, uint8_t* s_obs, int n16, int tid) {
  const uint4* in = reinterpret_cast<const uint4*>(s_grid);
  uint4* out = reinterpret_cast<uint4*>(s_obs);
  for (int g = tid; g < n16; g += THREADS) {
    uint4 a, b, c;
    expand16(in[g], a, b, c);
    out[3 * g + 0] = a; out[3 * g + 1] = b; out[3 * g + 2] = c;
  }
}

// store `bytes` of the tile's obs from smem to global: TMA bulk for the 16-byte multiple, plain
// byte stores for a ragged tail (last tile only) or when the caller's pointer is unaligned.
template <int THREADS>
__device__ __forceinline__ void store_obs_tail(uint8_t* gdst, const uint8_t* s_obs, uint32_t bulk, uint32_t bytes, int tid) {
  for (uint32_t i = bulk + tid; i < bytes; i += THREADS) gdst[i] = s_obs[i];
}

template <int MODE, int E, int THREADS>
__global__ void __launch_bounds__(THREADS) collect_step_kernel(const __grid_constant__ CollectParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, A = p.A, cells = p.cells;
  const TileSmem s = carve(smem_raw, E, cells, A);
  const long long e0 = (long long)blockIdx.x * E;
  const int n_here = (int)min((long long)E, p.N - e0);
  const uint32_t grid_bytes = (uint32_t)E * cells;

  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, grid_bytes);
    tma_load_1d(s.grid, p.grid + e0 * cells, grid_bytes, &bar);  // grid plane is padded to whole tiles
  }
  // small per-env arrays: coalesced cooperative loads while the TMA copy is in flight
  for (int i = tid; i < n_here * A * 2; i += THREADS) s.pos[i] = p.agent_pos[e0 * A * 2 + i];
  for (int i = tid; i < n_here * A; i += THREADS) {
    s.act[i] = p.actions[e0 * A + i];
    if (MODE == 0) s.ord[i] = p.order[e0 * A + i];
  }
  int4 h = make_int4(0, 0, 0, 0);
  if (tid < n_here) h = p.hdr[e0 + tid];
  __syncthreads();
  mbar_wait(&bar, 0);

  // ---- one thread per env: the ordered agent loop (collect_game.py:183-211)
  bool done = false;
  int err = 0;
  Rng<MODE> r;
  if (tid < n_here) {
    const long long e = e0 + tid;
    uint8_t* g = s.grid + (size_t)tid * cells;
    uint8_t* pos = s.pos + tid * A * 2;
    uint8_t* ord = s.ord + tid * A;
    double* rew = s.rew + tid * A;
    if (MODE == 0) {
      r.open_trace(p.draws ? p.draws + e * p.K : nullptr, p.draws ? (p.n_draws ? p.n_draws[e] : p.K) : 0);
    } else {
      r.open_philox(p.seed, p.env_id_base + (unsigned long long)e, (uint32_t)h.z);
      for (int i = 0; i < A; ++i) ord[i] = (uint8_t)i;  // Fisher-Yates over Philox draws
      for (int i = A - 1; i > 0; --i) {
        const int j = (int)__umulhi(r.u32(), (uint32_t)(i + 1));
        const uint8_t t = ord[i]; ord[i] = ord[j]; ord[j] = t;
      }
    }
    for (int i = 0; i < A; ++i) rew[i] = 0.0;  // :187
    h.x += 1;                                  // step_count += 1 :190
    for (int k = 0; k < A; ++k) {              // for i in order :191
      const int i = ord[k];
      const int a = s.act[tid * A + i];
      if (a < 0 || a > 3) continue;  // no branch of :192-207 matches: silently ignored
      const int ox = pos[2 * i], oy = pos[2 * i + 1];
      // north (0,-1) east (+1,0) south (0,+1) west (-1,0)  agent.py:230-264
      const int nx = ox + (a == 1) - (a == 3), ny = oy + (a == 2) - (a == 0);
      if (nx < 0 || ny < 0 || nx >= p.W || ny >= p.H) { err |= MG_ERR_OOB; continue; }
      const uint8_t c = GCELL(g, p.H, nx, ny);
      bool enter = (c == 0);                       // :178-181
      if ((c & 3) == T_BALL) {                     // move_agent :169-177 -> _handle_pickup :132-147
        const int colour = (c >> 2) & 15;
        GCELL(g, p.H, nx, ny) = 0;                 // grid.set(*fwd_pos, None) :141
        if (p.respawn) respawn<MODE>(p, g, r, colour);  // :142-143 -- may land on (nx, ny)
        h.y += 1;                                  // collected_balls += 1 :144
        rew[i] += p.reward_of_colour[colour];      // _reward(i, rewards, fwd_cell.reward) :145
        const int t = p.type_of_colour[colour];
        if (t >= 0) p.info[e * (A * p.nb) + p.nb * i + t] += 1;  // info[keys[nb*i + ball_idx]] :147 (rare RMW)
        enter = true;
      }
      if (enter) {  // wall / other agent: neither ball nor None -> blocked (:169-171)
        GCELL(g, p.H, nx, ny) = p.agent_code[i];  // overwrites a respawn that landed here (ball lost)
        GCELL(g, p.H, ox, oy) = 0;                // also erases a co-located partner from the grid
        pos[2 * i] = (uint8_t)nx; pos[2 * i + 1] = (uint8_t)ny;
      }
    }
    bool term = !p.respawn && h.y == p.num_balls;  // :208-209
    if (p.fixed_horizon) term = false;             // CollectGameRoomsFixedHorizon.step :368-370
    bool trunc = h.x >= p.max_steps;               // :210-211
    if (p.time_limit > 0 && h.x >= p.time_limit) trunc = true;  // gymnasium TimeLimit of the registration
    p.terminated[e] = term; p.truncated[e] = trunc;
    if (MODE == 0 && p.draws_used) p.draws_used[e] = r.k;
    err |= r.err;
    done = p.autoreset && (term || trunc);
  }

  // ---- rare path: same-step autoreset (gymnasium 0.29.1 VectorEnv semantics)
  if (p.autoreset) {
    if (tid < E) s.done[tid] = done;
    if (__syncthreads_or(done)) {
      if (p.final_obs) {  // terminal observation of the finished envs
        expand_tile<THREADS>(s.grid, s.obs, (int)(grid_bytes / 16), tid);
        __syncthreads();
        for (int j = 0; j < n_here; ++j) {
          if (!s.done[j]) continue;
          uint8_t* dst = p.final_obs + (e0 + j) * 3 * cells;
          const uint8_t* src = s.obs + (size_t)j * 3 * cells;
          for (int i = tid; i < 3 * cells; i += THREADS) dst[i] = src[i];
        }
        __syncthreads();
      }
      if (done) {
        const long long e = e0 + tid;
        if (MODE == 0) {
          Rng<MODE> rr;
          rr.open_trace(p.reset_draws ? p.reset_draws + e * p.R : nullptr,
                        p.reset_draws ? (p.n_reset_draws ? p.n_reset_draws[e] : p.R) : 0);
          reset_env<MODE>(p, s.grid + (size_t)tid * cells, s.pos + tid * A * 2, rr);
          if (p.reset_draws_used) p.reset_draws_used[e] = rr.k;
          err |= rr.err;
        } else {
          reset_env<MODE>(p, s.grid + (size_t)tid * cells, s.pos + tid * A * 2, r);
        }
        h.x = 0; h.y = 0; h.w += 1;  // step_count, collected_balls (:108, multigrid.py:141); episode counter
        for (int k = 0; k < A * p.nb; ++k) p.info[e * (A * p.nb) + k] = 0;  // :109-116
      }
    }
  }
  if (tid < n_here) {
    if (MODE == 1) h.z = (int)r.ctr;
    p.hdr[e0 + tid] = h;
    if (err) atomicOr(p.status, err);
  }
  __syncthreads();

  // ---- all threads: Grid.encode of the tile, then TMA bulk stores
  expand_tile<THREADS>(s.grid, s.obs, (int)(grid_bytes / 16), tid);
  fence_proxy_async_smem();
  __syncthreads();
  const uint32_t obs_bytes = (uint32_t)n_here * 3 * cells;
  const uint32_t bulk = (p.obs && p.obs_bulk_ok) ? (obs_bytes & ~15u) : 0u;
  if (tid == 0) {
    tma_store_1d(p.grid + e0 * cells, s.grid, grid_bytes);
    if (bulk) tma_store_1d(p.obs + e0 * 3 * cells, s.obs, bulk);
    tma_commit();
  }
  if (p.obs) store_obs_tail<THREADS>(p.obs + e0 * 3 * cells, s.obs, bulk, obs_bytes, tid);
  for (int i = tid; i < n_here * A; i += THREADS) p.rewards[e0 * A + i] = s.rew[i];
  for (int i = tid; i < n_here * A * 2; i += THREADS) p.agent_pos[e0 * A * 2 + i] = s.pos[i];
  if (tid == 0) tma_wait_read_all();  // shared memory must outlive the bulk reads
}

// reset(mask): envs with mask[e] != 0 (or all when mask == NULL) are re-generated; obs (if given)
// receives Grid.encode() of every env of the tile.
template <int MODE, int E, int THREADS>
__global__ void __launch_bounds__(THREADS) collect_reset_kernel(const __grid_constant__ CollectParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, A = p.A, cells = p.cells;
  const TileSmem s = carve(smem_raw, E, cells, A);
  const long long e0 = (long long)blockIdx.x * E;
  const int n_here = (int)min((long long)E, p.N - e0);
  const uint32_t grid_bytes = (uint32_t)E * cells;

  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, grid_bytes);
    tma_load_1d(s.grid, p.grid + e0 * cells, grid_bytes, &bar);
  }
  for (int i = tid; i < n_here * A * 2; i += THREADS) s.pos[i] = p.agent_pos[e0 * A * 2 + i];
  __syncthreads();
  mbar_wait(&bar, 0);

  if (tid < n_here) {
    const long long e = e0 + tid;
    if (!p.reset_mask || p.reset_mask[e]) {
      int4 h = p.hdr[e];
      Rng<MODE> r;
      if (MODE == 0)
        r.open_trace(p.reset_draws ? p.reset_draws + e * p.R : nullptr,
                     p.reset_draws ? (p.n_reset_draws ? p.n_reset_draws[e] : p.R) : 0);
      else
        r.open_philox(p.seed, p.env_id_base + (unsigned long long)e, (uint32_t)h.z);
      reset_env<MODE>(p, s.grid + (size_t)tid * cells, s.pos + tid * A * 2, r);
      h.x = 0; h.y = 0; h.w += 1;
      if (MODE == 1) h.z = (int)r.ctr;
      p.hdr[e] = h;
      for (int k = 0; k < A * p.nb; ++k) p.info[e * (A * p.nb) + k] = 0;
      if (MODE == 0 && p.reset_draws_used) p.reset_draws_used[e] = r.k;
      if (r.err) atomicOr(p.status, r.err);
    }
  }
  __syncthreads();
  if (p.obs) expand_tile<THREADS>(s.grid, s.obs, (int)(grid_bytes / 16), tid);
  fence_proxy_async_smem();
  __syncthreads();
  const uint32_t obs_bytes = (uint32_t)n_here * 3 * cells;
  const uint32_t bulk = (p.obs && p.obs_bulk_ok) ? (obs_bytes & ~15u) : 0u;
  if (tid == 0) {
    tma_store_1d(p.grid + e0 * cells, s.grid, grid_bytes);
    if (bulk) tma_store_1d(p.obs + e0 * 3 * cells, s.obs, bulk);
    tma_commit();
  }
  if (p.obs) store_obs_tail<THREADS>(p.obs + e0 * 3 * cells, s.obs, bulk, obs_bytes, tid);
  for (int i = tid; i < n_here * A * 2; i += THREADS) p.agent_pos[e0 * A * 2 + i] = s.pos[i];
  if (tid == 0) tma_wait_read_all();
}

// Grid.encode alone (grid.py:223-252): packed grid plane -> obs.
template <int E, int THREADS>
__global__ void __launch_bounds__(THREADS) encode3_kernel(const uint8_t* __restrict__ grid, uint8_t* __restrict__ obs,
                                                         long long N, int cells, int obs_bulk_ok) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  uint8_t* s_grid = smem_raw;
  uint8_t* s_obs = smem_raw + (size_t)E * cells;
  const long long e0 = (long long)blockIdx.x * E;
  const int n_here = (int)min((long long)E, N - e0);
  const uint32_t grid_bytes = (uint32_t)E * cells;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, grid_bytes);
    tma_load_1d(s_grid, grid + e0 * cells, grid_bytes, &bar);
  }
  mbar_wait(&bar, 0);
  expand_tile<THREADS>(s_grid, s_obs, (int)(grid_bytes / 16), tid);
  fence_proxy_async_smem();
  __syncthreads();
  const uint32_t obs_bytes = (uint32_t)n_here * 3 * cells;
  const uint32_t bulk = obs_bulk_ok ? (obs_bytes & ~15u) : 0u;
  if (tid == 0 && bulk) { tma_store_1d(obs + e0 * 3 * cells, s_obs, bulk); tma_commit(); }
  store_obs_tail<THREADS>(obs + e0 * 3 * cells, s_obs, bulk, obs_bytes, tid);
  if (tid == 0) tma_wait_read_all();
}

// ------------------------------------------------------------------------------ launchers
constexpr int kE = 64, kThreads = 128;

static cudaError_t set_smem(const void* fn, size_t bytes) {
  return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// once per handle (mg_create): opt every kernel in to the tile's dynamic shared memory size
cudaError_t configure_kernels(int cells, int A) {
  const size_t smem = tile_smem_bytes(kE, cells, A);
  cudaError_t e;
  if ((e = set_smem((const void*)collect_step_kernel<0, kE, kThreads>, smem)) != cudaSuccess) return e;
  if ((e = set_smem((const void*)collect_step_kernel<1, kE, kThreads>, smem)) != cudaSuccess) return e;
  if ((e = set_smem((const void*)collect_reset_kernel<0, kE, kThreads>, smem)) != cudaSuccess) return e;
  if ((e = set_smem((const void*)collect_reset_kernel<1, kE, kThreads>, smem)) != cudaSuccess) return e;
  return set_smem((const void*)encode3_kernel<kE, kThreads>, (size_t)kE * cells * 4);
}

cudaError_t launch_collect_step(const CollectParams& p, cudaStream_t st) {
  const size_t smem = tile_smem_bytes(kE, p.cells, p.A);
  const unsigned blocks = (unsigned)((p.N + kE - 1) / kE);
  if (p.rng_mode == 0) collect_step_kernel<0, kE, kThreads><<<blocks, kThreads, smem, st>>>(p);
  else collect_step_kernel<1, kE, kThreads><<<blocks, kThreads, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_collect_reset(const CollectParams& p, cudaStream_t st) {
  const size_t smem = tile_smem_bytes(kE, p.cells, p.A);
  const unsigned blocks = (unsigned)((p.N + kE - 1) / kE);
  if (p.rng_mode == 0) collect_reset_kernel<0, kE, kThreads><<<blocks, kThreads, smem, st>>>(p);
  else collect_reset_kernel<1, kE, kThreads><<<blocks, kThreads, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_encode3(const uint8_t* grid, uint8_t* obs, long long N, int cells, int obs_bulk_ok, cudaStream_t st) {
  const size_t smem = (size_t)kE * cells * 4;
  const unsigned blocks = (unsigned)((N + kE - 1) / kE);
  encode3_kernel<kE, kThreads><<<blocks, kThreads, smem, st>>>(grid, obs, N, cells, obs_bulk_ok);
  return cudaGetLastError();
}

int tile_envs() { return kE; }
size_t tile_smem(int cells, int A) { return tile_smem_bytes(kE, cells, A); }

}  // namespace mg
