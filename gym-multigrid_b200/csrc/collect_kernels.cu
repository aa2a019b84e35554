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
#include <cstdlib>
#include <type_traits>

#include "collect_device.cuh"
#include "mg_device.cuh"
#include "smem_config.h"

namespace mg {

// shared-memory carve-up of one tile.  Every array starts 16-byte aligned (E % 16 == 0) so that each
// one can be the source / destination of a TMA bulk copy.
struct TileSmem {
  uint8_t* grid;   // [E][cells]          in+out
  uint8_t* obs;    // [E][cells][3]       out
  int4* hdr;       // [E]                 in+out
  double* rew;     // [E][A]              out
  uint8_t* pos;    // [E][A][2]           in+out
  int8_t* act;     // [E][A]              in
  uint8_t* ord;    // [E][A]              in (trace) / scratch (Philox)
  uint8_t* term;   // [E]                 out
  uint8_t* trunc;  // [E]                 out
  uint8_t* done;   // [E]
  uint16_t* chg;   // [E][3A] cells written by the step (index x*H+y), for patching the pre-expanded obs
  uint8_t* delta;  // [E][delta_record_bytes] compact host-transport records (bytes 1..A double as the per-agent pickup scratch)
};
__host__ __device__ inline size_t tile_smem_bytes(int E, int cells, int A) {
  return (size_t)E * cells * 4 + (size_t)E * 16 + (size_t)E * A * 8 + (size_t)E * A * 4 + (size_t)E * 3 + (size_t)E * A * 6 + 16 +
         (size_t)E * delta_record_bytes(cells, A) + 16;
}
__device__ __forceinline__ TileSmem carve(uint8_t* base, int E, int cells, int A) {
  TileSmem s;
  s.grid = base;
  s.obs = base + (size_t)E * cells;
  s.hdr = reinterpret_cast<int4*>(base + (size_t)E * cells * 4);
  s.rew = reinterpret_cast<double*>(s.hdr + E);
  s.pos = reinterpret_cast<uint8_t*>(s.rew + (size_t)E * A);
  s.act = reinterpret_cast<int8_t*>(s.pos + (size_t)E * A * 2);
  s.ord = reinterpret_cast<uint8_t*>(s.act + (size_t)E * A);
  s.term = s.ord + (size_t)E * A;
  s.trunc = s.term + E;
  s.done = s.trunc + E;
  s.chg = reinterpret_cast<uint16_t*>(s.done + E + (E & 1));
  s.delta = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s.chg + (size_t)E * 3 * A) + 15) & ~(uintptr_t)15);
  return s;
}

// all threads: packed grid tile -> 3-byte encoding, both in shared memory
template <int THREADS>
__device__ __forceinline__ void expand_tile(const uint8_t* s_grid, uint8_t* s_obs, int n16, int tid) {
  const uint4* in = reinterpret_cast<const uint4*>(s_grid);
  uint4* out = reinterpret_cast<uint4*>(s_obs);
#pragma unroll 1
  for (int g = tid; g < n16; g += THREADS) {
    uint4 a, b, c;
    expand16(in[g], a, b, c);
    out[3 * g + 0] = a; out[3 * g + 1] = b; out[3 * g + 2] = c;
  }
}

// plain (non-TMA) copy of `bytes` starting at offset `from`: ragged last tile / unaligned caller pointers
template <int THREADS>
__device__ __forceinline__ void copy_out_tail(uint8_t* gdst, const uint8_t* src, uint32_t from, uint32_t bytes, int tid) {
  for (uint32_t i = from + tid; i < bytes; i += THREADS) gdst[i] = src[i];
}

// CTAs per SM the register allocation must allow: what the ~455 B/env shared-memory tile permits
// (10x10 grid, A = 2), so that 65 536 envs are resident in a single wave.
__host__ __device__ constexpr int min_blocks(int E, int threads) {
  const int by_smem = (227 * 1024) / (E * 456 + 1024), by_threads = 2048 / threads;
  return by_smem < by_threads ? (by_smem < 1 ? 1 : by_smem) : by_threads;
}

template <int MODE, int E, int THREADS>
__global__ void __launch_bounds__(THREADS, min_blocks(E, THREADS)) collect_step_kernel(const __grid_constant__ CollectParams p) {
  static_assert(E % 16 == 0 && E <= THREADS, "tile must be a multiple of 16 envs and fit one thread per env");
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, A = p.A, cells = p.cells;
  const TileSmem s = carve(smem_raw, E, cells, A);
  const long long e0 = (long long)blockIdx.x * E;
  const int n_here = (int)min((long long)E, p.N - e0);
  const uint32_t grid_bytes = (uint32_t)E * cells;
  const int R = delta_record_bytes(cells, A);
  // the caller's per-env arrays are not padded: bulk copies only for full tiles with aligned pointers
  const bool io_bulk = p.io_bulk_ok && n_here == E;
  unsigned long long* tl = p.timeline ? p.timeline + (size_t)blockIdx.x * 8 : nullptr;
  if (tl && tid == 0) tl[0] = globaltimer_ns();

  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  pdl_launch_dependents();  // let the next launch in the stream get scheduled behind this one
  __syncthreads();
  pdl_wait();               // ... and do not touch global memory before the previous launch has fully completed
  if (tid == 0) {  // every input of the tile arrives by TMA on one mbarrier (state planes are padded to whole tiles)
    const uint32_t act_bytes = io_bulk ? (uint32_t)E * A : 0u;
    mbar_expect_tx(&bar, grid_bytes + (uint32_t)E * 16 + (uint32_t)E * A * 2 + act_bytes);
    tma_load_1d(s.grid, p.grid + e0 * cells, grid_bytes, &bar);
    tma_load_1d(s.hdr, p.hdr + e0, (uint32_t)E * 16, &bar);
    tma_load_1d(s.pos, p.agent_pos + e0 * A * 2, (uint32_t)E * A * 2, &bar);
    if (io_bulk) tma_load_1d(s.act, p.actions + e0 * A, act_bytes, &bar);
  }
  if (!io_bulk)
    for (int i = tid; i < n_here * A; i += THREADS) s.act[i] = p.actions[e0 * A + i];
  if (MODE == 0)
    for (int i = tid; i < n_here * A; i += THREADS) s.ord[i] = p.order[e0 * A + i];
  if (!io_bulk || MODE == 0) __syncthreads();   // only the hand-copied inputs need it: everything else arrives on the mbarrier
  mbar_wait(&bar, 0);
  if (tl && tid == 0) tl[1] = globaltimer_ns();

  // ---- warps [0, E/32): one thread per env walks the ordered agent loop.
  //      warps [E/32, ..): meanwhile expand the PRE-step grid (only the <= 3A cells a step writes can change;
  //      they are patched below), which takes Grid.encode off the critical path of the tile.
  constexpr bool OVERLAP = (THREADS - E) >= 64;
  // early observation store: 70 % of a tile's output bytes (the pre-step encoding, produced by warps [E/32, ..) while the env threads
  // walk their agents) go out during the agent loop instead of after it; full tiles with a 16-byte aligned obs pointer only
  const bool early = OVERLAP && p.early_obs && p.obs && p.obs_bulk_ok && n_here == E && !p.timeline;
  const int n16 = (int)(grid_bytes / 16);
  bool done = false;
  int err = 0, nchg = 0;
  int4 h = make_int4(0, 0, 0, 0);
  Rng<MODE> r;
  uint16_t* chg = s.chg + (size_t)(tid < E ? tid : 0) * 3 * A;
  if (tid < n_here) {
    const long long e = e0 + tid;
    h = s.hdr[tid];
    if (MODE == 0) r.open_trace(p.draws ? p.draws + e * p.K : nullptr, p.draws ? (p.n_draws ? p.n_draws[e] : p.K) : 0);
    else r.open_philox(p.seed, p.env_id_base + (unsigned long long)e, (uint32_t)h.z);
    bool term, trunc;
    err = step_one_env<MODE>(p, e, s.grid + (size_t)tid * cells, s.pos + tid * A * 2, s.ord + tid * A, s.act + tid * A,
                             s.rew + tid * A, h, r, term, trunc, chg, nchg, s.delta + (size_t)tid * R + 1);
    s.term[tid] = term; s.trunc[tid] = trunc;
    if (MODE == 0 && p.draws_used) p.draws_used[e] = r.k;
    done = p.autoreset && (term || trunc);
  } else if (OVERLAP && tid >= E && p.obs) {
    expand_tile<THREADS - E>(s.grid, s.obs, n16, tid - E);
    if (early) {  // the slab leaves NOW, under the agent loop; cells the step changes are patched in place below
      fence_proxy_async_smem();
      named_bar_sync(1, THREADS - E);
      if (tid == E) {
        tma_store_1d(p.obs + e0 * 3 * cells, s.obs, (uint32_t)E * 3 * cells);
        tma_commit();
        tma_wait_all();   // performed, not just read: the patches below go to the same addresses
      }
    }
  }
  if (tl && tid == 0) tl[2] = globaltimer_ns();
  if (tid < E) s.done[tid] = done;
  const int any_done = __syncthreads_or(done);

  if (p.obs) {
    if (!OVERLAP) {
      expand_tile<THREADS>(s.grid, s.obs, n16, tid);
    } else if (tid < n_here) {  // patch the cells this env's step wrote
      const uint8_t* g = s.grid + (size_t)tid * cells;
      uint8_t* o = s.obs + (size_t)tid * cells * 3;
      // early store: the same three bytes also go to the slab already in global memory - unless the tile is about to be
      // re-encoded and stored again as a whole (autoreset), which must not race with these stores
      uint8_t* go = (early && !any_done) ? p.obs + (e0 + tid) * 3 * cells : nullptr;
      for (int k = 0; k < nchg; ++k) {
        const int idx = chg[k];
        const uint8_t c = g[idx];
        o[3 * idx] = c & 3; o[3 * idx + 1] = (c >> 2) & 15; o[3 * idx + 2] = state_of(c);
        if (go) { go[3 * idx] = c & 3; go[3 * idx + 1] = (c >> 2) & 15; go[3 * idx + 2] = state_of(c); }
      }
    }
  }

  if (p.delta && tid < n_here) {  // compact host transport: the cells this env's step wrote, with their post-step codes
    const uint8_t* g = s.grid + (size_t)tid * cells;
    uint8_t* rec = s.delta + (size_t)tid * R;
    rec[0] = (uint8_t)(nchg | (s.term[tid] << 5) | (s.trunc[tid] << 6) | ((int)done << 7));
    uint8_t* ent = rec + 1 + A;
    if (!p.delta_wide) {
      for (int k = 0; k < nchg; ++k) { const int idx = chg[k]; ent[2 * k] = (uint8_t)idx; ent[2 * k + 1] = g[idx]; }
    } else {
      for (int k = 0; k < nchg; ++k) { const int idx = chg[k]; ent[3 * k] = (uint8_t)idx; ent[3 * k + 1] = (uint8_t)(idx >> 8); ent[3 * k + 2] = g[idx]; }
    }
  }

  // ---- rare path: same-step autoreset (gymnasium 0.29.1 VectorEnv semantics)
  if (any_done) {
    __syncthreads();
    if (p.final_obs && p.obs) {  // terminal observation of the finished envs (s.obs holds the post-step encoding)
      for (int j = 0; j < n_here; ++j) {
        if (!s.done[j]) continue;
        uint8_t* dst = p.final_obs + (e0 + j) * 3 * cells;
        const uint8_t* src = s.obs + (size_t)j * 3 * cells;
        if (((3 * cells) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.final_obs) & 3) == 0) {   // 32-bit words when every env's slab is word-aligned
          for (int i = tid; i < 3 * cells / 4; i += THREADS) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
        } else {
          for (int i = tid; i < 3 * cells; i += THREADS) dst[i] = src[i];
        }
      }
      __syncthreads();
    }
    if (done) {
      const long long e = e0 + tid;
      Rng<MODE> rr;  // a copy: only this rarely-taken path hands the generator to a non-inlined function
      if (MODE == 0)
        rr.open_trace(p.reset_draws ? p.reset_draws + e * p.R : nullptr,
                      p.reset_draws ? (p.n_reset_draws ? p.n_reset_draws[e] : p.R) : 0);
      else
        rr = r;
      reset_env<MODE>(p, s.grid + (size_t)tid * cells, s.pos + tid * A * 2, rr);
      if (MODE == 0) { if (p.reset_draws_used) p.reset_draws_used[e] = rr.k; }
      else r.ctr = rr.ctr;
      err |= rr.err;
      h.x = 0; h.y = 0; h.w += 1;  // step_count, collected_balls (:108, multigrid.py:141); episode counter
      for (int k = 0; k < A * p.nb; ++k) p.info[e * (A * p.nb) + k] = 0;  // :109-116
      // compact host transport: claim a slot for this env's fresh row (every chg list has been consumed above)
      if (p.delta) reinterpret_cast<int*>(s.chg)[tid] = atomicAdd(p.reset_count, 1);
    }
    __syncthreads();
    if (p.obs) expand_tile<THREADS>(s.grid, s.obs, n16, tid);  // re-encode (the reset envs changed everywhere)
    if (p.delta) {
      const int* slots = reinterpret_cast<const int*>(s.chg);
      for (int j = 0; j < n_here; ++j) {
        if (!s.done[j]) continue;
        uint8_t* dst_row = p.reset_rows + (size_t)slots[j] * p.reset_stride;
        if (tid == 0) *reinterpret_cast<int32_t*>(dst_row) = (int32_t)(e0 + j);
        if ((cells & 3) == 0) {
          const uint32_t* src = reinterpret_cast<const uint32_t*>(s.grid + (size_t)j * cells);
          uint32_t* dst = reinterpret_cast<uint32_t*>(dst_row + 4);
          for (int i = tid; i < cells / 4; i += THREADS) dst[i] = src[i];
        } else {
          for (int i = tid; i < cells; i += THREADS) dst_row[4 + i] = s.grid[(size_t)j * cells + i];
        }
      }
    }
  }
  if (tid < n_here) {
    if (MODE == 1) h.z = (int)r.ctr;
    s.hdr[tid] = h;
    if (err) atomicOr(p.status, err);
  }
  if (tl && tid == 0) tl[3] = globaltimer_ns();

  // ---- TMA bulk stores of every output of the tile
  fence_proxy_async_smem();
  __syncthreads();
  if (tl && tid == 0) tl[4] = globaltimer_ns();
  const uint32_t obs_bytes = (uint32_t)n_here * 3 * cells;
  const bool obs_done = early && !any_done;   // the slab went out early and was patched in place
  const uint32_t obs_bulk = (p.obs && p.obs_bulk_ok && !obs_done) ? (obs_bytes & ~15u) : 0u;
  if (tid == 0) {
    if (obs_bulk) tma_store_1d(p.obs + e0 * 3 * cells, s.obs, obs_bulk);
    tma_store_1d(p.grid + e0 * cells, s.grid, grid_bytes);
    tma_store_1d(p.hdr + e0, s.hdr, (uint32_t)E * 16);
    tma_store_1d(p.agent_pos + e0 * A * 2, s.pos, (uint32_t)E * A * 2);
    if (io_bulk) {
      tma_store_1d(p.rewards + e0 * A, s.rew, (uint32_t)E * A * 8);
      tma_store_1d(p.terminated + e0, s.term, (uint32_t)E);
      tma_store_1d(p.truncated + e0, s.trunc, (uint32_t)E);
    }
    if (p.delta && n_here == E) tma_store_1d(p.delta + e0 * R, s.delta, (uint32_t)E * R);
    tma_commit();
  }
  if (p.delta && n_here != E) copy_out_tail<THREADS>(p.delta + e0 * R, s.delta, 0u, (uint32_t)n_here * R, tid);
  if (p.obs && !obs_done) copy_out_tail<THREADS>(p.obs + e0 * 3 * cells, s.obs, obs_bulk, obs_bytes, tid);
  if (!io_bulk) {
    for (int i = tid; i < n_here * A; i += THREADS) p.rewards[e0 * A + i] = s.rew[i];
    for (int i = tid; i < n_here; i += THREADS) { p.terminated[e0 + i] = s.term[i]; p.truncated[e0 + i] = s.trunc[i]; }
  }
  if (tid == 0) {
    tma_wait_read_all();  // shared memory must outlive the bulk reads
    if (tl) tl[5] = globaltimer_ns();
  }
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
    mbar_expect_tx(&bar, grid_bytes + (uint32_t)E * 16 + (uint32_t)E * A * 2);
    tma_load_1d(s.grid, p.grid + e0 * cells, grid_bytes, &bar);
    tma_load_1d(s.hdr, p.hdr + e0, (uint32_t)E * 16, &bar);
    tma_load_1d(s.pos, p.agent_pos + e0 * A * 2, (uint32_t)E * A * 2, &bar);
  }
  mbar_wait(&bar, 0);

  if (tid < n_here) {
    const long long e = e0 + tid;
    if (!p.reset_mask || p.reset_mask[e]) {
      int4 h = s.hdr[tid];
      Rng<MODE> r;
      if (MODE == 0)
        r.open_trace(p.reset_draws ? p.reset_draws + e * p.R : nullptr,
                     p.reset_draws ? (p.n_reset_draws ? p.n_reset_draws[e] : p.R) : 0);
      else
        r.open_philox(p.seed, p.env_id_base + (unsigned long long)e, (uint32_t)h.z);
      reset_env<MODE>(p, s.grid + (size_t)tid * cells, s.pos + tid * A * 2, r);
      h.x = 0; h.y = 0; h.w += 1;
      if (MODE == 1) h.z = (int)r.ctr;
      s.hdr[tid] = h;
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
  const uint32_t obs_bulk = (p.obs && p.obs_bulk_ok) ? (obs_bytes & ~15u) : 0u;
  if (tid == 0) {
    if (obs_bulk) tma_store_1d(p.obs + e0 * 3 * cells, s.obs, obs_bulk);
    tma_store_1d(p.grid + e0 * cells, s.grid, grid_bytes);
    tma_store_1d(p.hdr + e0, s.hdr, (uint32_t)E * 16);
    tma_store_1d(p.agent_pos + e0 * A * 2, s.pos, (uint32_t)E * A * 2);
    tma_commit();
  }
  if (p.obs) copy_out_tail<THREADS>(p.obs + e0 * 3 * cells, s.obs, obs_bulk, obs_bytes, tid);
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
  copy_out_tail<THREADS>(obs + e0 * 3 * cells, s_obs, bulk, obs_bytes, tid);
  if (tid == 0) tma_wait_read_all();
}

// ------------------------------------------------------------------------------ launchers
// Tile variants (envs per CTA x threads per CTA); MG_TILE=<index> selects one at mg_create time.
struct TileCfg { int E, threads; };
static const TileCfg kTiles[] = {{64, 128}, {32, 64}, {64, 64}, {128, 128}, {128, 256}, {32, 128}, {16, 64}, {64, 192}, {64, 256}, {32, 96}};
constexpr int kNumTiles = sizeof(kTiles) / sizeof(kTiles[0]);

template <typename F>
static cudaError_t for_tile(int v, F&& f) {
  switch (v) {
    case 0: return f(std::integral_constant<int, 64>{}, std::integral_constant<int, 128>{});
    case 1: return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 64>{});
    case 2: return f(std::integral_constant<int, 64>{}, std::integral_constant<int, 64>{});
    case 3: return f(std::integral_constant<int, 128>{}, std::integral_constant<int, 128>{});
    case 4: return f(std::integral_constant<int, 128>{}, std::integral_constant<int, 256>{});
    case 5: return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 128>{});
    case 6: return f(std::integral_constant<int, 16>{}, std::integral_constant<int, 64>{});
    case 7: return f(std::integral_constant<int, 64>{}, std::integral_constant<int, 192>{});
    case 8: return f(std::integral_constant<int, 64>{}, std::integral_constant<int, 256>{});
    case 9: return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 96>{});
  }
  return cudaErrorInvalidValue;
}

static bool pdl_enabled() {
  static const bool on = [] { const char* v = std::getenv("MG_PDL"); return !(v && v[0] == '0'); }();
  return on;
}

static cudaError_t set_smem(const void* fn, size_t bytes) {
  return raise_smem_limit(fn, bytes);
}

int num_tile_variants() { return kNumTiles; }
int tile_envs(int v) { return kTiles[v].E; }
size_t tile_smem(int v, int cells, int A) { return tile_smem_bytes(kTiles[v].E, cells, A); }

// once per handle (mg_create): opt every kernel in to the tile's dynamic shared memory size
cudaError_t configure_kernels(int v, int cells, int A) {
  return for_tile(v, [&](auto e, auto t) {
    constexpr int E = decltype(e)::value, T = decltype(t)::value;
    const size_t smem = tile_smem_bytes(E, cells, A);
    cudaError_t r;
    if ((r = set_smem((const void*)collect_step_kernel<0, E, T>, smem)) != cudaSuccess) return r;
    if ((r = set_smem((const void*)collect_step_kernel<1, E, T>, smem)) != cudaSuccess) return r;
    if ((r = set_smem((const void*)collect_reset_kernel<0, E, T>, smem)) != cudaSuccess) return r;
    if ((r = set_smem((const void*)collect_reset_kernel<1, E, T>, smem)) != cudaSuccess) return r;
    return set_smem((const void*)encode3_kernel<E, T>, (size_t)E * cells * 4);
  });
}

cudaError_t launch_collect_step(int v, const CollectParams& p, cudaStream_t st) {
  return for_tile(v, [&](auto e, auto t) {
    constexpr int E = decltype(e)::value, T = decltype(t)::value;
    const size_t smem = tile_smem_bytes(E, p.cells, p.A);
    const unsigned blocks = (unsigned)((p.N + E - 1) / E);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    if (p.rng_mode == 0) return cudaLaunchKernelEx(&cfg, collect_step_kernel<0, E, T>, p);
    return cudaLaunchKernelEx(&cfg, collect_step_kernel<1, E, T>, p);
  });
}

cudaError_t launch_collect_reset(int v, const CollectParams& p, cudaStream_t st) {
  return for_tile(v, [&](auto e, auto t) {
    constexpr int E = decltype(e)::value, T = decltype(t)::value;
    const size_t smem = tile_smem_bytes(E, p.cells, p.A);
    const unsigned blocks = (unsigned)((p.N + E - 1) / E);
    if (p.rng_mode == 0) collect_reset_kernel<0, E, T><<<blocks, T, smem, st>>>(p);
    else collect_reset_kernel<1, E, T><<<blocks, T, smem, st>>>(p);
    return cudaGetLastError();
  });
}

cudaError_t launch_encode3(int v, const uint8_t* grid, uint8_t* obs, long long N, int cells, int obs_bulk_ok, cudaStream_t st) {
  return for_tile(v, [&](auto e, auto t) {
    constexpr int E = decltype(e)::value, T = decltype(t)::value;
    const size_t smem = (size_t)E * cells * 4;
    const unsigned blocks = (unsigned)((N + E - 1) / E);
    encode3_kernel<E, T><<<blocks, T, smem, st>>>(grid, obs, N, cells, obs_bulk_ok);
    return cudaGetLastError();
  });
}

}  // namespace mg
