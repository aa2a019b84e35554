// collect_rollout_kernels.cu -- warp-tile kernel of the Collect family: T >= 1 env steps per launch with the env state
// resident in shared memory (mg_rollout; T = 1 is mg_step's launch-amortised sibling).
//
//   reference loop being replaced:  for t in range(T): obs, rew, term, trunc, info = env.step(actions[t])
//                                   (CollectGameEnv.step collect_game.py:183-214 + Grid.encode grid.py:223-252, T times)
//
// Shape: ONE WARP owns a tile of 32 consecutive envs for all T steps, lane = env.  There is no CTA-wide barrier anywhere: every
// warp is its own pipeline (own mbarriers, own shared-memory slice), so the warps of an SM drift apart and the load / step /
// store phases of different tiles overlap instead of running in lockstep.  Per tile:
//   - lane 0 pulls the tile's state (packed grids, header rows, agent positions) into shared memory with TMA bulk copies, once;
//   - per step: the step's actions arrive by TMA into a double buffer (step t+1's are in flight while step t runs) or are drawn
//     on the device (uniform policy); each lane walks its env's agents in order on the shared-memory grid (collect_device.cuh,
//     the same per-env function as the tile kernel); the warp expands the 32 grids to the 3-byte encoding; lane 0 issues the
//     TMA bulk stores of the step's observation slab, rewards and flags, which drain while the next step is computed;
//   - after the last step the state goes back to HBM.  State traffic (2 x 136 B per env) is paid once per launch, not per step.
// Same-step autoreset (final observation, _gen_grid, re-encode) is handled inside the warp with ballots.
#include <cstdlib>

#include "collect_device.cuh"
#include "mg_device.cuh"
#include "smem_config.h"

namespace mg {

constexpr int kRollTileMax = 32;   // envs per warp: CollectParams::roll_tile, a power of two <= 32 (lane = env; lanes beyond it only help encode)

struct WarpSmem {
  uint8_t* grid;      // [32][cells]
  uint8_t* obs;       // [32][cells][3]
  int4* hdr;          // [32]
  uint8_t* pos;       // [32][A][2]
  int8_t* act[2];     // [32][A]   double-buffered over steps
  double* rew[2];     // [32][A]   double-buffered: the TMA store of step t reads it while step t+1 is computed
  uint8_t* term[2];   // [32]
  uint8_t* trunc[2];  // [32]
  uint8_t* ord;       // [32][A]
  uint16_t* chg;      // [32][3A]
  uint8_t* delta;     // [32][R]
  uint64_t* bar;      // [3] state, actions (even steps), actions (odd steps)
};

__host__ __device__ inline size_t up16(size_t v) { return (v + 15) & ~(size_t)15; }

// Two-agent handles use the register path: the staging arrays of the generic path (rewards, actions, flags, order) are not carved,
// and the delta-record buffer only when the launch writes records (host transport).  That is what lets 8 two-warp CTAs - 16 tiles -
// share an SM at 32 envs per warp on a 10x10 grid (13.5 KB per warp) instead of 7.
__host__ __device__ inline size_t warp_smem_bytes(int cells, int A, int tile, bool with_delta) {
  const size_t E = tile;
  const bool lean = A == 2;
  size_t b = up16(E * cells) + 3 * E * cells + E * 16 + up16(E * A * 2) + up16(E * 3 * A * 2) + 32;
  if (!lean) b += 2 * up16(E * A) + 2 * E * A * 8 + 4 * E + up16(E * A);
  if (!lean || with_delta) b += up16(E * delta_record_bytes(cells, A));
  return b;
}

__device__ __forceinline__ WarpSmem carve_warp(uint8_t* b, int cells, int A, int tile, bool with_delta) {
  const size_t E = tile;
  const bool lean = A == 2;
  WarpSmem s;
  s.grid = b; b += up16(E * cells);
  s.obs = b; b += 3 * E * cells;
  s.hdr = reinterpret_cast<int4*>(b); b += E * 16;
  s.pos = b; b += up16(E * A * 2);
  s.chg = reinterpret_cast<uint16_t*>(b); b += up16(E * 3 * A * 2);
  s.bar = reinterpret_cast<uint64_t*>(b); b += 32;
  s.rew[0] = s.rew[1] = nullptr; s.act[0] = s.act[1] = nullptr; s.term[0] = s.term[1] = s.trunc[0] = s.trunc[1] = nullptr; s.ord = nullptr;
  if (!lean) {
    s.rew[0] = reinterpret_cast<double*>(b); b += E * A * 8;
    s.rew[1] = reinterpret_cast<double*>(b); b += E * A * 8;
    s.act[0] = reinterpret_cast<int8_t*>(b); b += up16(E * A);
    s.act[1] = reinterpret_cast<int8_t*>(b); b += up16(E * A);
    s.term[0] = b; b += E; s.term[1] = b; b += E;
    s.trunc[0] = b; b += E; s.trunc[1] = b; b += E;
    s.ord = b; b += up16(E * A);
  }
  s.delta = (!lean || with_delta) ? b : nullptr;
  return s;
}

template <bool MARK>
__device__ __forceinline__ void expand_warp(const uint8_t* s_grid, uint8_t* s_obs, int n16, int lane) {
  const uint4* in = reinterpret_cast<const uint4*>(s_grid);
  uint4* out = reinterpret_cast<uint4*>(s_obs);
#pragma unroll 2
  for (int g = lane; g < n16; g += 32) {
    uint4 a, b, c;
    expand16<MARK>(in[g], a, b, c);
    out[3 * g + 0] = a; out[3 * g + 1] = b; out[3 * g + 2] = c;
  }
}

// mark = CollectParams::mark_respawned (uniform): only then can a ball carry bit 6 and the encoding needs the four extra masking
// operations per word; no registered config does.  Two copies of this small loop instead of two copies of the whole tile body.
__device__ __forceinline__ void expand_tile_warp(const uint8_t* s_grid, uint8_t* s_obs, int n16, int lane, int mark) {
  if (mark) expand_warp<true>(s_grid, s_obs, n16, lane);
  else expand_warp<false>(s_grid, s_obs, n16, lane);
}

// `bytes` from shared to global memory by the warp (32-bit words when both sides allow it)
__device__ __forceinline__ void warp_copy(uint8_t* gdst, const uint8_t* src, uint32_t bytes, int lane) {
  if (((reinterpret_cast<uintptr_t>(gdst) | reinterpret_cast<uintptr_t>(src) | bytes) & 3u) == 0) {
    for (uint32_t i = lane; i < bytes / 4; i += 32) reinterpret_cast<uint32_t*>(gdst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  } else {
    for (uint32_t i = lane; i < bytes; i += 32) gdst[i] = src[i];
  }
}

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kRollThreads = 64;   // two independent warps per CTA; 8 CTAs / SM = 16 warps: exactly the register file at 128 registers per thread

// profiling (mg_debug_set_timeline): cycles a warp spent per phase, summed over the T steps of a tile.  Compiled in only with
// -DMG_ROLLOUT_CLOCK (tools/dev/rollout_timeline.py rebuilds the library with it): even a disabled clock costs 4 % of the instructions.
struct TileClock {
#ifdef MG_ROLLOUT_CLOCK
  unsigned long long* out;
  long long acc[6], c0, begin;
  __device__ __forceinline__ void open(unsigned long long* o) { out = o; for (int i = 0; i < 6; ++i) acc[i] = 0; c0 = begin = o ? clock64() : 0; }
  __device__ __forceinline__ void start() { if (out) c0 = clock64(); }
  __device__ __forceinline__ void lap(int k) { if (out) { const long long c = clock64(); acc[k] += c - c0; c0 = c; } }
  __device__ __forceinline__ void close(int lane, int T) {
    if (out && lane == 0) {
      for (int i = 0; i < 6; ++i) out[i] = (unsigned long long)acc[i];
      out[6] = (unsigned long long)(clock64() - begin); out[7] = (unsigned long long)T;
    }
  }
#else
  __device__ __forceinline__ void open(unsigned long long*) {}
  __device__ __forceinline__ void start() {}
  __device__ __forceinline__ void lap(int) {}
  __device__ __forceinline__ void close(int, int) {}
#endif
};

// One tile (32 envs, lane = env) for T steps.  AT = 2: the two-agent fast path - positions, actions, rewards and flags live in
// registers, the next step's actions are prefetched into a register, rewards / flags leave as plain coalesced stores and only the
// observation slab goes through shared memory and TMA.  AT = 0: any agent count, everything staged in shared memory.
template <int MODE, int AT>
__device__ __forceinline__ void rollout_tile(const CollectParams& p, const WarpSmem& s, long long tile, int lane, uint32_t& ph_state,
                                             uint32_t& ph_act0, uint32_t& ph_act1) {
  const int A = AT ? AT : p.A, cells = p.cells, T = p.T;
  const int E = p.roll_tile;
  const int R = delta_record_bytes(cells, A);
  const long long N = p.N;
  const uint32_t grid_bytes = (uint32_t)E * cells, obs_row = 3u * cells;
  const int n16 = (int)(grid_bytes / 16);
  const long long e0 = tile * E;
  const int n_here = (int)min((long long)E, N - e0);
  const bool full = n_here == E;
  const bool io_bulk = p.io_bulk_ok && full;       // rewards / flags / actions: aligned pointers and strides, whole tile
  const bool obs_bulk = p.obs && p.obs_bulk_ok && full && E >= 16;   // small tiles: a few coalesced 16-byte stores beat a TMA round trip
  const bool act_tma = AT == 0 && io_bulk && p.actions;

  __syncwarp();            // every lane is done with the previous tile's shared memory
  if (lane == 0) {
    tma_wait_read_all();   // ... and so are the previous tile's bulk stores
    mbar_expect_tx(&s.bar[0], grid_bytes + (uint32_t)E * 16u + (uint32_t)E * A * 2);
    tma_load_1d(s.grid, p.grid + e0 * cells, grid_bytes, &s.bar[0]);
    tma_load_1d(s.hdr, p.hdr + e0, (uint32_t)E * 16u, &s.bar[0]);
    tma_load_1d(s.pos, p.agent_pos + e0 * A * 2, (uint32_t)E * A * 2, &s.bar[0]);
    if (act_tma) {
      mbar_expect_tx(&s.bar[1], (uint32_t)E * A);
      tma_load_1d(s.act[0], p.actions + e0 * A, (uint32_t)E * A, &s.bar[1]);
    }
  }
  const long long e = e0 + lane;
  const bool live = lane < n_here;
  // AT = 2: step 0's actions travel while the state loads
  uint32_t act_next = 0;
  auto load_act2 = [&](long long idx) -> uint32_t {   // the two action bytes of env `idx` (one 16-bit load when the pointer allows it)
    const int8_t* a = p.actions + idx * 2;
    if (p.io_bulk_ok) return *reinterpret_cast<const uint16_t*>(a);
    return (uint32_t)(uint8_t)a[0] | ((uint32_t)(uint8_t)a[1] << 8);
  };
  if (AT == 2 && p.actions && live) act_next = load_act2(e);
  __syncwarp();
  mbar_wait(&s.bar[0], ph_state); ph_state ^= 1;

  TileClock clk;
  clk.open(p.timeline ? p.timeline + (size_t)tile * 8 : nullptr);
  int4 h = make_int4(0, 0, 0, 0);
  Rng<MODE> r;
  int err = 0;
  uint32_t posw = 0;
  if (live) {
    h = s.hdr[lane];
    if (MODE == 1) r.open_philox(p.seed, p.env_id_base + (unsigned long long)e, (uint32_t)h.z);
    if (AT == 2) posw = *reinterpret_cast<const uint32_t*>(s.pos + lane * 4);
  }
  uint16_t* chg = s.chg + (size_t)lane * 3 * A;
  uint8_t* g = s.grid + (size_t)lane * cells;
  uint8_t* o_row = s.obs + (size_t)lane * obs_row;
  int32_t* info_row = p.info + e * (A * p.nb);
  bool obs_valid = false;   // s.obs holds the encoding of the grids as they were before this step (patching is enough)

  for (int t = 0; t < T; ++t) {
    const int cur = t & 1;
    const long long te = (long long)t * N + e0;     // first env of this tile in the [T][N] output arrays
    double* const s_rew = cur ? s.rew[1] : s.rew[0];
    uint8_t* const s_term = cur ? s.term[1] : s.term[0];
    uint8_t* const s_trunc = cur ? s.trunc[1] : s.trunc[0];
    clk.start();
    // ---- actions of step t
    int8_t* act = (cur ? s.act[1] : s.act[0]) + lane * A;
    uint32_t actw = 0;
    if (AT == 2 && p.actions) {
      actw = act_next;
      if (t + 1 < T && live) act_next = load_act2((long long)(t + 1) * N + e);   // in flight during step t
    } else if (act_tma) {
      if (cur == 0) { mbar_wait(&s.bar[1], ph_act0); ph_act0 ^= 1; } else { mbar_wait(&s.bar[2], ph_act1); ph_act1 ^= 1; }
      if (t + 1 < T && lane == 0) {   // step t+1's actions fly in while step t is computed (its buffer was last read at step t-1)
        uint64_t* nbar = cur ? &s.bar[1] : &s.bar[2];
        mbar_expect_tx(nbar, (uint32_t)E * A);
        tma_load_1d(cur ? s.act[0] : s.act[1], p.actions + ((long long)(t + 1) * N + e0) * A, (uint32_t)E * A, nbar);
      }
    } else if (p.actions) {
      if (live) for (int i = 0; i < A; ++i) act[i] = p.actions[(te + lane) * A + i];
    } else if (live) {
      // uniform random policy drawn on the device: one Philox block per env and step, counter (env id, step_count, 2^30 | episode)
      // - disjoint from the env's own stream (word 3 = 0) - two bits per agent
      uint32_t o[4];
      const unsigned long long id = p.env_id_base + (unsigned long long)e;
      philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)h.x, 0x40000000u | ((uint32_t)h.w & 0x3FFFFFFFu), (uint32_t)p.seed,
                    (uint32_t)(p.seed >> 32), o);
      if (AT == 2) {
        actw = (o[0] & 3u) | (((o[0] >> 2) & 3u) << 8);
        if (p.actions_out) { p.actions_out[(te + lane) * 2] = (int8_t)(actw & 3u); p.actions_out[(te + lane) * 2 + 1] = (int8_t)(actw >> 8); }
      } else {
        for (int i = 0; i < A; ++i) act[i] = (int8_t)((o[0] >> (2 * i)) & 3u);
        if (p.actions_out) for (int i = 0; i < A; ++i) p.actions_out[(te + lane) * A + i] = act[i];
      }
    }
    clk.lap(0);

    // ---- step: lane = env, agents in order on the shared-memory grid
    bool done = false, term = false, trunc = false;
    int nchg = 0;
    double r0 = 0.0, r1 = 0.0;
    uint32_t pick = 0;
    if (live) {
      if (MODE == 0) r.open_trace(p.draws ? p.draws + e * p.K : nullptr, p.draws ? (p.n_draws ? p.n_draws[e] : p.K) : 0);
      else r.have = 0;    // every step starts on a fresh Philox block, exactly as T separate mg_step launches would
      if (AT == 2) {
        int i0;
        if (MODE == 0) i0 = p.order[e * 2] & 1;
        else i0 = __umulhi(r.u32(), 2u) == 0 ? 1 : 0;   // Fisher-Yates over [0, 1]: j = 0 swaps the pair
        err |= step_one_env_a2<MODE>(p, info_row, g, posw, actw, i0, r0, r1, pick, h, r, term, trunc, chg, nchg);
      } else {
        if (MODE == 0) for (int i = 0; i < A; ++i) s.ord[lane * A + i] = p.order[e * A + i];
        err |= step_one_env<MODE>(p, e, g, s.pos + lane * A * 2, s.ord + lane * A, act, s_rew + lane * A, h, r, term, trunc, chg, nchg,
                                  s.delta + (size_t)lane * R + 1);
        s_term[lane] = term; s_trunc[lane] = trunc;
      }
      if (MODE == 0 && p.draws_used) p.draws_used[e] = r.k;
      done = p.autoreset && (term || trunc);
    }
    const unsigned done_mask = __ballot_sync(kFull, done);   // also orders the lanes' grid writes before the expansion below
    clk.lap(1);

    // ---- Grid.encode of the post-step grids.  The observation buffer is the source of the previous step's TMA store: wait until
    //      that store has read it (it has had the whole agent loop to do so).  After the tile's first step the buffer already holds
    //      the pre-step encoding, so only the <= 3A cells this env's step wrote are re-encoded.
    if (lane == 0) tma_wait_read_all();
    __syncwarp();
    clk.lap(2);
    if (p.obs) {
      if (!obs_valid) {
        expand_tile_warp(s.grid, s.obs, n16, lane, p.mark_respawned);
      } else if (AT == 2 && live) {   // (lanes beyond the tile own no row: their pointers lie outside the warp's slice)
        // <= 6 cells; all index loads, then all cell loads, then the stores: the loads overlap instead of forming one dependent chain per cell
        int idx[6]; uint8_t cc[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) idx[k] = k < nchg ? chg[k] : 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) cc[k] = g[idx[k]];
#pragma unroll
        for (int k = 0; k < 6; ++k)
          if (k < nchg) {
            uint8_t* o = o_row + 3 * idx[k];
            o[0] = cc[k] & 3; o[1] = (cc[k] >> 2) & 15; o[2] = state_of(cc[k]);
          }
      } else if (AT != 2) {
        for (int k = 0; k < nchg; ++k) {
          const int idx = chg[k];
          const uint8_t c = g[idx];
          o_row[3 * idx] = c & 3; o_row[3 * idx + 1] = (c >> 2) & 15; o_row[3 * idx + 2] = state_of(c);
        }
      }
      obs_valid = true;
    }
    if (p.delta && live) {  // compact host transport: the cells this env's step wrote, with their post-step codes
      uint8_t* rec = s.delta + (size_t)lane * R;
      rec[0] = (uint8_t)(nchg | ((int)term << 5) | ((int)trunc << 6) | ((int)done << 7));
      if (AT == 2) { rec[1] = (uint8_t)pick; rec[2] = (uint8_t)(pick >> 8); }
      uint8_t* ent = rec + 1 + A;
      if (!p.delta_wide) {
        for (int k = 0; k < nchg; ++k) { const int idx = chg[k]; ent[2 * k] = (uint8_t)idx; ent[2 * k + 1] = g[idx]; }
      } else {
        for (int k = 0; k < nchg; ++k) { const int idx = chg[k]; ent[3 * k] = (uint8_t)idx; ent[3 * k + 1] = (uint8_t)(idx >> 8); ent[3 * k + 2] = g[idx]; }
      }
    }
    __syncwarp();
    clk.lap(3);

    // ---- same-step autoreset (gymnasium 0.29.1 VectorEnv semantics) of the lanes that finished
    if (done_mask) {
      if (p.final_obs && p.obs) {   // terminal observation: s.obs holds the post-step encoding
        for (unsigned m = done_mask; m; m &= m - 1) {
          const int j = __ffs(m) - 1;
          warp_copy(p.final_obs + (te + j) * obs_row, s.obs + (size_t)j * obs_row, obs_row, lane);
        }
        __syncwarp();
      }
      if (done) {
        uint8_t* pos = s.pos + lane * A * 2;
        if (MODE == 0) {
          Rng<MODE> rr;
          rr.open_trace(p.reset_draws ? p.reset_draws + e * p.R : nullptr, p.reset_draws ? (p.n_reset_draws ? p.n_reset_draws[e] : p.R) : 0);
          reset_env<MODE>(p, g, pos, rr);
          if (p.reset_draws_used) p.reset_draws_used[e] = rr.k;
          err |= rr.err;
        } else {
          reset_env<MODE>(p, g, pos, r);   // continues on the step's Philox block, as the tile kernel does
          err |= r.err;
        }
        if (AT == 2) posw = *reinterpret_cast<const uint32_t*>(pos);
        h.x = 0; h.y = 0; h.w += 1;  // step_count, collected_balls (:108, multigrid.py:141); episode counter
        for (int k = 0; k < A * p.nb; ++k) info_row[k] = 0;  // :109-116
      }
      int slot = 0;
      if (p.delta && done) slot = atomicAdd(p.reset_count, 1);
      __syncwarp();
      if (p.obs) expand_tile_warp(s.grid, s.obs, n16, lane, p.mark_respawned);  // re-encode (the reset envs changed everywhere)
      if (p.delta) {
        for (unsigned m = done_mask; m; m &= m - 1) {
          const int j = __ffs(m) - 1;
          uint8_t* dst_row = p.reset_rows + (size_t)__shfl_sync(kFull, slot, j) * p.reset_stride;
          if (lane == 0) *reinterpret_cast<int32_t*>(dst_row) = (int32_t)(e0 + j);
          warp_copy(dst_row + 4, s.grid + (size_t)j * cells, (uint32_t)cells, lane);
        }
      }
    }
    clk.lap(4);

    // ---- the step's outputs: the observation slab (and the delta records) leave by TMA bulk stores; the two-agent path writes
    //      rewards and flags straight from registers (one 16-byte and two 1-byte coalesced stores per lane)
    const bool any_tma = obs_bulk || (AT == 0 && io_bulk) || (p.delta && full);
    if (any_tma) fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && any_tma) {
      if (obs_bulk) tma_store_1d(p.obs + te * obs_row, s.obs, (uint32_t)E * obs_row);
      if (AT == 0 && io_bulk) {
        tma_store_1d(p.rewards + te * A, s_rew, (uint32_t)E * A * 8);
        tma_store_1d(p.terminated + te, s_term, (uint32_t)E);
        tma_store_1d(p.truncated + te, s_trunc, (uint32_t)E);
      }
      if (p.delta && full) tma_store_1d(p.delta + te * R, s.delta, (uint32_t)E * R);
      tma_commit();
    }
    if (p.obs && !obs_bulk) {
      uint8_t* dst = p.obs + te * obs_row;
      const uint32_t bytes = (uint32_t)n_here * obs_row;
      if (p.obs_bulk_ok && (bytes & 15u) == 0) {
        for (uint32_t i = lane; i < bytes / 16; i += 32) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(s.obs)[i];
      } else {
        warp_copy(dst, s.obs, bytes, lane);
      }
    }
    if (AT == 2) {
      if (live) {
        double* rw = p.rewards + (te + lane) * 2;
        if (p.io_bulk_ok) *reinterpret_cast<double2*>(rw) = make_double2(r0, r1); else { rw[0] = r0; rw[1] = r1; }
        p.terminated[te + lane] = term; p.truncated[te + lane] = trunc;
      }
    } else if (!io_bulk) {
      for (int i = lane; i < n_here * A; i += 32) p.rewards[te * A + i] = s_rew[i];
      if (live) { p.terminated[te + lane] = s_term[lane]; p.truncated[te + lane] = s_trunc[lane]; }
    }
    if (p.delta && !full) warp_copy(p.delta + te * R, s.delta, (uint32_t)n_here * R, lane);
    clk.lap(5);
  }
  clk.close(lane, T);

  // ---- state back to HBM
  if (live) {
    if (MODE == 1) h.z = (int)r.ctr;
    s.hdr[lane] = h;
    if (AT == 2) *reinterpret_cast<uint32_t*>(s.pos + lane * 4) = posw;
    if (err) atomicOr(p.status, err);
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_1d(p.grid + e0 * cells, s.grid, grid_bytes);
    tma_store_1d(p.hdr + e0, s.hdr, (uint32_t)E * 16u);
    tma_store_1d(p.agent_pos + e0 * A * 2, s.pos, (uint32_t)E * A * 2);
    tma_commit();
  }
}

template <int MODE>
__global__ void __launch_bounds__(kRollThreads, 8) collect_rollout_kernel(const __grid_constant__ CollectParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const bool with_delta = p.delta != nullptr;
  const WarpSmem s = carve_warp(smem_raw + (size_t)wib * warp_smem_bytes(p.cells, p.A, p.roll_tile, with_delta), p.cells, p.A, p.roll_tile, with_delta);
  const long long ntiles = (p.N + p.roll_tile - 1) / p.roll_tile;
  const long long gw = (long long)blockIdx.x * wpc + wib, nw = (long long)gridDim.x * wpc;

  if (lane == 0) { mbar_init(&s.bar[0], 1); mbar_init(&s.bar[1], 1); mbar_init(&s.bar[2], 1); fence_mbar_init(); }
  pdl_launch_dependents();
  __syncwarp();
  pdl_wait();
  uint32_t ph_state = 0, ph_act0 = 0, ph_act1 = 0;
  for (long long tile = gw; tile < ntiles; tile += nw) {
    if (p.A == 2) rollout_tile<MODE, 2>(p, s, tile, lane, ph_state, ph_act0, ph_act1);
    else rollout_tile<MODE, 0>(p, s, tile, lane, ph_state, ph_act0, ph_act1);
  }
  if (lane == 0) tma_wait_read_all();  // shared memory must outlive the bulk reads
}

// ------------------------------------------------------------------------------ launcher
constexpr int kRollWarps = kRollThreads / 32;

size_t rollout_smem_bytes(int cells, int A, int tile, bool with_delta) { return (size_t)kRollWarps * warp_smem_bytes(cells, A, tile, with_delta); }

cudaError_t configure_rollout_kernels(int cells, int A) {
  const size_t smem = rollout_smem_bytes(cells, A, kRollTileMax, true);
  cudaError_t r;
  if ((r = raise_smem_limit((const void*)collect_rollout_kernel<0>, smem)) != cudaSuccess) return r;
  return raise_smem_limit((const void*)collect_rollout_kernel<1>, smem);
}

// Envs per warp.  A warp executes the union of its lanes' paths (pickups, respawn rejection loops), so a step of a 32-env tile
// takes ~3x as long as a step of a 4-env tile: when the batch is small enough that every tile still gets its own resident warp,
// smaller tiles cut the per-step latency - which is all that matters for a launch-amortised rollout of a few thousand envs.
static int pick_roll_tile(const CollectParams& p, long long resident_warps) {
  if (const char* e = std::getenv("MG_ROLLOUT_TILE")) {
    const int v = std::atoi(e);
    if (v == 4 || v == 8 || v == 16 || v == 32) return v;
  }
  if (p.A != 2) return kRollTileMax;           // the generic path stages whole 32-env arrays for TMA
  for (int tile = 4; tile < kRollTileMax; tile *= 2) {
    if ((tile * p.cells) % 16 != 0) continue;  // every TMA piece of a tile must be a multiple of 16 bytes
    if ((p.N + tile - 1) / tile <= resident_warps) return tile;
  }
  return kRollTileMax;
}

cudaError_t launch_collect_rollout(const CollectParams& p_in, int num_sms, cudaStream_t st) {
  CollectParams p = p_in;
  int per_sm = 0;
  cudaError_t ce;
  // occupancy at the largest tile: 7 CTAs / SM for a 10x10 grid (the smaller tiles use less shared memory, registers then bound it)
  const bool with_delta = p.delta != nullptr;
  const size_t smem_max = rollout_smem_bytes(p.cells, p.A, kRollTileMax, with_delta);
  if (p.rng_mode == 0) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, collect_rollout_kernel<0>, kRollThreads, smem_max);
  else ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, collect_rollout_kernel<1>, kRollThreads, smem_max);
  if (ce != cudaSuccess) return ce;
  if (per_sm < 1) per_sm = 1;
  const long long resident = (long long)num_sms * per_sm;
  p.roll_tile = pick_roll_tile(p, resident * kRollWarps);
  size_t smem = rollout_smem_bytes(p.cells, p.A, p.roll_tile, with_delta);
  // Experiments: extra dynamic shared memory per CTA lowers the CTAs per SM.  At 8 per SM a 1024-CTA launch (65 536 envs) fills 128 of
  // the 148 SMs - the hardware places consecutive CTAs on one SM - and takes 10.5 us alone on a stream; padded to 7 per SM it spreads
  // over 146 SMs and takes 9.9 us, but independent launches on 4 streams then overlap less (6.8 against 6.5 us) and a rollout step takes
  // 4.3 instead of 4.1 us.  Always launching the full resident grid with the tile-less CTAs spread over the SMs was worse on every count.
  if (const char* e = std::getenv("MG_ROLLOUT_PAD")) smem += (size_t)std::atoi(e);
  const long long ntiles = (p.N + p.roll_tile - 1) / p.roll_tile;
  long long blocks = (ntiles + kRollWarps - 1) / kRollWarps;
  if (blocks > resident) blocks = resident;      // persistent: every warp walks its tiles with stride = warps in the grid
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(kRollThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  const char* v = std::getenv("MG_PDL");
  cfg.attrs = attr; cfg.numAttrs = (v && v[0] == '0') ? 0 : 1;
  if (p.rng_mode == 0) return cudaLaunchKernelEx(&cfg, collect_rollout_kernel<0>, p);
  return cudaLaunchKernelEx(&cfg, collect_rollout_kernel<1>, p);
}

}  // namespace mg
