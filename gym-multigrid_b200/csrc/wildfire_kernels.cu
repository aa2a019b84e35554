// wildfire_kernels.cu -- the Wildfire EXTENSION (no reference code exists; the rules are specified in
// include/multigrid_b200.h and restated by oracle/mg_oracle_wildfire.c).
//
// One CTA per env.  Two kernels with identical results:
//   wildfire_kernel       (any H): terrain staged in shared memory by a TMA bulk copy; warp 0 walks the agents in the step's
//                         random order with lane = agent (the acting lane broadcasts its target with __shfl_sync, the others
//                         vote with __ballot_sync whether they stand there); fire spread is a double-buffered per-cell
//                         4-neighbour stencil with one Philox block per group of four cells.
//   wildfire_fast_kernel  (H % 4 == 0, the one that runs for square power-of-two grids): word-parallel stencil and encode,
//                         Philox only on the queued fire front, per-lane rank tracking + __match_any_sync conflict detection
//                         so that only conflicting moves are resolved in order (see the comment above the kernel).
// The observation is expanded in shared memory and leaves, like the new terrain, as one TMA bulk store.
#include <cstdlib>

#include "mg_device.cuh"
#include "smem_config.h"
#include "wildfire_params.cuh"

namespace mg {

constexpr int kWfThreads = 256;
constexpr int WF_HEALTHY = 0, WF_BURNING = 1, WF_BURNT = 2;

__device__ __forceinline__ uint8_t wf_packed(int s) {  // healthy green, burning red, burnt grey (constants.py:8-19)
  return s == WF_HEALTHY ? cell(0, 3, 0) : (s == WF_BURNING ? cell(1, 0, 0) : cell(2, 7, 0));
}

struct WfSmem {
  uint8_t* told;   // [cells] terrain in / packed cells for the encode
  uint8_t* tnew;   // [cells]
  uint8_t* obs;    // [3*cells]
};

// all threads: terrain + agents -> packed cells (in `packed`) -> 3-byte encoding in s.obs
__device__ __forceinline__ void wf_encode(const WildfireParams& p, const uint8_t* terrain, uint8_t* packed, uint8_t* obs,
                                          const int* s_ax, const int* s_ay, const int* s_adir, int tid, int nthreads = kWfThreads) {
  for (int i = tid; i < p.cells; i += nthreads) packed[i] = wf_packed(terrain[i]);
  __syncthreads();
  if (tid < p.A) packed[s_ax[tid] * p.H + s_ay[tid]] = (uint8_t)(cell(3, p.agent_colour[tid], 0) | (s_adir[tid] << 6));
  __syncthreads();
  const uint4* in = reinterpret_cast<const uint4*>(packed);
  uint4* out = reinterpret_cast<uint4*>(obs);
  for (int g = tid; g < p.cells / 16; g += nthreads) {
    uint4 a, b, c;
    expand16(in[g], a, b, c);
    out[3 * g] = a; out[3 * g + 1] = b; out[3 * g + 2] = c;
  }
}

// thread 0: reset of one env on shared memory (all healthy was written by all threads before)
__device__ __noinline__ void wf_reset_agents_fires(const WildfireParams& p, uint8_t* t, int* s_ax, int* s_ay, int* s_adir,
                                                   Rng<1>& r) {
  for (int f = 0; f < p.num_fires; ++f)
    for (;;) { const int i = (int)__umulhi(r.u32(), (uint32_t)p.cells); if (t[i] == WF_HEALTHY) { t[i] = WF_BURNING; break; } }
  for (int k = 0; k < p.A; ++k)
    for (;;) {
      const int i = (int)__umulhi(r.u32(), (uint32_t)p.cells);
      bool taken = false;
      for (int j = 0; j < k; ++j) taken |= (s_ax[j] * p.H + s_ay[j] == i);
      if (taken) continue;
      s_ax[k] = i / p.H; s_ay[k] = i % p.H; s_adir[k] = 3;
      break;
    }
}

__global__ void __launch_bounds__(kWfThreads) wildfire_kernel(const __grid_constant__ WildfireParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_ax[MG_MAX_WILDFIRE_AGENTS], s_ay[MG_MAX_WILDFIRE_AGENTS], s_adir[MG_MAX_WILDFIRE_AGENTS];
  __shared__ int s_order[MG_MAX_WILDFIRE_AGENTS], s_burning, s_flag;
  __shared__ int4 s_hdr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, A = p.A, W = p.W, H = p.H, cells = p.cells;
  const long long e = blockIdx.x;
  WfSmem s;
  s.told = smem_raw; s.tnew = smem_raw + cells; s.obs = smem_raw + 2 * (size_t)cells;
  uint8_t* g_terrain = p.terrain + e * cells;

  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); s_burning = 0; s_flag = 0; }
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();
  if (tid == 0) { mbar_expect_tx(&bar, (uint32_t)cells); tma_load_1d(s.told, g_terrain, (uint32_t)cells, &bar); }
  Rng<1> r;
  int4 h = make_int4(0, 0, 0, 0);
  if (tid < A) {
    const uchar4 a = reinterpret_cast<const uchar4*>(p.agents)[e * A + tid];
    s_ax[tid] = a.x; s_ay[tid] = a.y; s_adir[tid] = a.z;
  }
  if (tid == 0) {
    h = p.hdr[e];
    r.open_philox(p.seed, p.env_id_base + (unsigned long long)e, (uint32_t)h.z);
  }
  const bool do_reset = p.op == 0 && (!p.reset_mask || p.reset_mask[e]);
  mbar_wait(&bar, 0);
  __syncthreads();

  if (p.op == 1) {
    // ---- 1/2. ordered agent moves: warp 0, lane = agent
    if (tid == 0) {
      h.x += 1; h.y += 1;  // step_count, tick
      if (p.order) {
        for (int i = 0; i < A; ++i) s_order[i] = p.order[e * A + i];
      } else {  // Fisher-Yates over the env's Philox stream
        for (int i = 0; i < A; ++i) s_order[i] = i;
        for (int i = A - 1; i > 0; --i) { const int j = (int)__umulhi(r.u32(), (uint32_t)(i + 1)), t = s_order[i]; s_order[i] = s_order[j]; s_order[j] = t; }
      }
    }
    if (warp == 0) {
      __syncwarp();
      int x = lane < A ? s_ax[lane] : -1, y = lane < A ? s_ay[lane] : -1, dir = lane < A ? s_adir[lane] : 0;
      const int a = lane < A ? p.actions[e * A + lane] : 0;
      double rew = 0.0;
      for (int k = 0; k < A; ++k) {
        const int i = s_order[k];  // the acting agent; warp-uniform
        int dx = 0, dy = 0;
        if (lane == i && a >= 1 && a <= 4) { dx = (a == 4) - (a == 2); dy = (a == 3) - (a == 1); }
        const int nx = __shfl_sync(0xffffffffu, x + dx, i), ny = __shfl_sync(0xffffffffu, y + dy, i);
        const unsigned occupied = __ballot_sync(0xffffffffu, lane != i && lane < A && x == nx && y == ny);
        if (lane == i) {
          if ((dx | dy) && nx >= 0 && ny >= 0 && nx < W && ny < H && !occupied) {
            dir = dx == 1 ? 0 : (dy == 1 ? 1 : (dx == -1 ? 2 : 3));  // DIR_TO_VEC (constants.py:65-74)
            x = nx; y = ny;
          }
          if (s.told[x * H + y] == WF_BURNING) { s.told[x * H + y] = WF_BURNT; rew += 1.0; }  // extinguish
        }
        __syncwarp();
      }
      if (lane < A) { s_ax[lane] = x; s_ay[lane] = y; s_adir[lane] = dir; p.rewards[e * A + lane] = rew; }
    }
    __syncthreads();

    // ---- 3. fire dynamics: double-buffered 4-neighbour stencil, one Philox block per group of 4 cells
    const unsigned long long env_id = p.env_id_base + (unsigned long long)e;
    if (tid == 0) s_hdr = h;
    __syncthreads();
    const uint32_t tk = (uint32_t)s_hdr.y;
    int burning = 0;
    for (int g = tid; g < (cells + 3) / 4; g += kWfThreads) {
      int st[4], kk[4];
      bool active = false;
      for (int j = 0; j < 4; ++j) {
        const int i = 4 * g + j;
        st[j] = -1; kk[j] = 0;
        if (i >= cells) continue;
        const int x = i / H, y = i - x * H, sv = s.told[i];
        st[j] = sv;
        if (sv == WF_HEALTHY) {
          int k = 0;
          if (x > 0) k += s.told[i - H] == WF_BURNING;
          if (x < W - 1) k += s.told[i + H] == WF_BURNING;
          if (y > 0) k += s.told[i - 1] == WF_BURNING;
          if (y < H - 1) k += s.told[i + 1] == WF_BURNING;
          kk[j] = k;
          active |= k > 0;
        } else if (sv == WF_BURNING) {
          active = true;
        }
      }
      uint32_t u[4] = {0, 0, 0, 0};
      if (active) philox4x32_10((uint32_t)env_id, (uint32_t)(env_id >> 32), tk, 1u + (uint32_t)g, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), u);
      for (int j = 0; j < 4; ++j) {
        if (st[j] < 0) continue;
        int ns = st[j];
        if (st[j] == WF_BURNING) ns = u[j] < p.burnout_threshold ? WF_BURNT : WF_BURNING;
        else if (st[j] == WF_HEALTHY && kk[j] > 0) ns = u[j] < p.ignite_threshold[kk[j]] ? WF_BURNING : WF_HEALTHY;
        s.tnew[4 * g + j] = (uint8_t)ns;
        burning += ns == WF_BURNING;
      }
    }
    for (int o = 16; o > 0; o >>= 1) burning += __shfl_xor_sync(0xffffffffu, burning, o);
    if (lane == 0 && burning) atomicAdd(&s_burning, burning);
    __syncthreads();
    // ---- 4. termination, same-step autoreset
    const bool term = s_burning == 0, trunc = s_hdr.x >= p.max_steps;
    if (tid == 0) { p.terminated[e] = term; p.truncated[e] = trunc; }
    if (p.autoreset && (term || trunc)) {
      if (p.final_obs) {
        wf_encode(p, s.tnew, s.told, s.obs, s_ax, s_ay, s_adir, tid);
        __syncthreads();
        uint4* dst = reinterpret_cast<uint4*>(p.final_obs + e * 3 * cells);
        for (int i = tid; i < 3 * cells / 16; i += kWfThreads) dst[i] = reinterpret_cast<const uint4*>(s.obs)[i];
        __syncthreads();
      }
      for (int i = tid; i < cells; i += kWfThreads) s.tnew[i] = WF_HEALTHY;
      __syncthreads();
      if (tid == 0) { wf_reset_agents_fires(p, s.tnew, s_ax, s_ay, s_adir, r); h.x = 0; h.w += 1; }
      __syncthreads();
    }
  } else {
    // ---- reset(mask)
    for (int i = tid; i < cells; i += kWfThreads) s.tnew[i] = do_reset ? (uint8_t)WF_HEALTHY : s.told[i];
    __syncthreads();
    if (tid == 0 && do_reset) { wf_reset_agents_fires(p, s.tnew, s_ax, s_ay, s_adir, r); h.x = 0; h.w += 1; }
    __syncthreads();
  }

  // ---- 5. observation + write-back
  if (p.obs) wf_encode(p, s.tnew, s.told, s.obs, s_ax, s_ay, s_adir, tid);
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    tma_store_1d(g_terrain, s.tnew, (uint32_t)cells);
    if (p.obs) tma_store_1d(p.obs + e * 3 * cells, s.obs, (uint32_t)(3 * cells));
    tma_commit();
    h.z = (int)r.ctr;
    p.hdr[e] = h;
  }
  if (tid < A) reinterpret_cast<uchar4*>(p.agents)[e * A + tid] = make_uchar4((uint8_t)s_ax[tid], (uint8_t)s_ay[tid], (uint8_t)s_adir[tid], 0);
  if (tid == 0) tma_wait_read_all();
}

// ------------------------------------------------------------------------------------------------------------------
// Fast path (H % 4 == 0): the same step with the fire dynamics and the encoding done four cells per 32-bit word (SWAR).
// States are 0 / 1 / 2, so "burning" is bit 0 of each byte: the four neighbour flags of a word are the words one row
// up / down and the word itself shifted by one byte with the carry byte of the adjacent word (zero at the row ends; zero
// guard rows above and below the staged terrain stand in for the x bounds tests).  A Philox block is computed only for
// words that hold a burning cell or a healthy cell with a burning neighbour - the fire front - which is what makes the
// step memory-bound instead of RNG-bound.  The agent order is drawn by warp 0 in parallel (lane d computes the Philox
// block of draw d; the Fisher-Yates swaps are register shuffles), so the serial section per env is the ordered move loop
// only.  Same Philox counters and word assignment as the generic kernel and the oracle: results are bit-identical.
// all threads: tnew + agents -> the observation tile (type byte = state; colour green / red / grey), word-parallel
template <int T, int VEC>
__device__ __forceinline__ void wf_encode_tile(const WildfireParams& p, const uint32_t* n32, uint8_t* s_obs, const int* s_ax, const int* s_ay,
                                               const int* s_adir, int nwords, int tid) {
  const int A = p.A, H = p.H;
  uint32_t* o32 = reinterpret_cast<uint32_t*>(s_obs);
  for (int q = tid; q < nwords / VEC; q += T) {
    uint32_t w[VEC], o[3 * VEC];
    if (VEC == 4) {
      const uint4 a = reinterpret_cast<const uint4*>(n32)[q];
      w[0] = a.x; w[VEC > 1 ? 1 : 0] = a.y; w[VEC > 2 ? 2 : 0] = a.z; w[VEC > 3 ? 3 : 0] = a.w;
    } else {
      w[0] = n32[q];
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const uint32_t burnt = (w[i] >> 1) & 0x01010101u, healthy = ~(w[i] | (w[i] >> 1)) & 0x01010101u;
      interleave3_zero_state(w[i], healthy * 3u + burnt * 7u, o[3 * i], o[3 * i + 1], o[3 * i + 2]);
    }
    if (VEC == 4) {
      uint4* d = reinterpret_cast<uint4*>(o32) + 3 * q;
      d[0] = make_uint4(o[0], o[1], o[2], o[3 % (3 * VEC)]);
      d[1] = make_uint4(o[4 % (3 * VEC)], o[5 % (3 * VEC)], o[6 % (3 * VEC)], o[7 % (3 * VEC)]);
      d[2] = make_uint4(o[8 % (3 * VEC)], o[9 % (3 * VEC)], o[10 % (3 * VEC)], o[11 % (3 * VEC)]);
    } else {
      o32[3 * q] = o[0]; o32[3 * q + 1] = o[1]; o32[3 * q + 2] = o[2];
    }
  }
  __syncthreads();
  if (tid < A) {
    uint8_t* o = s_obs + 3 * (s_ax[tid] * H + s_ay[tid]);
    o[0] = 3; o[1] = p.agent_colour[tid]; o[2] = (uint8_t)s_adir[tid];
  }
}

// the same, out of line: the terminal observation of an env that resets (rare) must not cost the hot path registers
template <int T, int VEC>
__device__ __noinline__ void wf_encode_tile_cold(const WildfireParams& p, const uint32_t* n32, uint8_t* s_obs, const int* s_ax, const int* s_ay,
                                                 const int* s_adir, int nwords, int tid) {
  wf_encode_tile<T, VEC>(p, n32, s_obs, s_ax, s_ay, s_adir, nwords, tid);
}

// Shared memory (4 bytes per cell): guard | told [cells] | guard | queue [2 * cells] | tnew [cells].  The observation tile
// (3 bytes per cell) is assembled over told + guard + queue once pass 2 has consumed them: 66 KB instead of 82 KB for a 128x128
// env, i.e. 3 CTAs per SM instead of 2 (956 -> 760 us per step of 32 768 envs).  Not capped at 40 registers for 12 CTAs of a 64x64
// env per SM: measured, 798 against 793 us at 10 CTAs - the step is issue-bound there, not latency-bound - and 32x32 loses 4 %.
// The bound below asks for 1 280 threads per SM = 10 CTAs of 128: 47 registers without spills (uncapped, the second inlined copy of
// the encoder takes the allocation to 53-56 registers and 64x64 to 9 CTAs per SM: 827 us).
template <int T, int VEC>
__global__ void __launch_bounds__(T, 1280 / T) wildfire_fast_kernel(const __grid_constant__ WildfireParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_ax[MG_MAX_WILDFIRE_AGENTS], s_ay[MG_MAX_WILDFIRE_AGENTS], s_adir[MG_MAX_WILDFIRE_AGENTS];
  __shared__ int s_count;   // length of the fire-front queue
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, A = p.A, W = p.W, H = p.H, cells = p.cells;
  const int rw = H >> 2, nwords = cells >> 2, guard = (H + 15) & ~15;
  const long long e = blockIdx.x;
  uint8_t* s_told = smem_raw + guard;                                      // guard | [cells] | guard
  uint8_t* s_queue = smem_raw + 2 * (size_t)guard + cells;                 // [2 * cells] fire-front queue of passes 1 / 2
  uint8_t* s_tnew = s_queue + 2 * (size_t)cells;                           // [cells]
  uint8_t* s_obs = s_told;                                                 // [3 * cells] over told | guard | queue, after pass 2
  uint32_t* t32 = reinterpret_cast<uint32_t*>(s_told);
  uint32_t* n32 = reinterpret_cast<uint32_t*>(s_tnew);
  uint8_t* g_terrain = p.terrain + e * cells;

  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); s_count = 0; }
  pdl_launch_dependents();
  for (int i = tid; i < guard / 4; i += T) { reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u; t32[nwords + i] = 0u; }
  __syncthreads();
  pdl_wait();
  if (tid == 0) { mbar_expect_tx(&bar, (uint32_t)cells); tma_load_1d(s_told, g_terrain, (uint32_t)cells, &bar); }
  int4 h = p.hdr[e];                        // every thread: one broadcast load
  const uint32_t ctr0 = (uint32_t)h.z;
  const unsigned long long env_id = p.env_id_base + (unsigned long long)e;
  const uint32_t id0 = (uint32_t)env_id, id1 = (uint32_t)(env_id >> 32), k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
  if (tid < A) {
    const uchar4 a = reinterpret_cast<const uchar4*>(p.agents)[e * A + tid];
    s_ax[tid] = a.x; s_ay[tid] = a.y; s_adir[tid] = a.z;
  }
  const bool do_reset = p.op == 0 && (!p.reset_mask || p.reset_mask[e]);
  bool need_reset = do_reset;
  uint32_t order_blocks = 0;                // Philox blocks the agent order consumed
  // Warp 0's share of the step that needs no terrain runs UNDER the terrain load: the agents' actions (a global load) and the
  // step's agent order (a Philox block and A - 1 dependent shuffles) - the other warps wait for warp 0's agent phase at a barrier,
  // so every cycle taken off it is taken off the whole CTA (A/B in one run: 64x64 772.4 against 774.3 us, 128x128 657 against 667).
  int act_pre = 0, rank_pre = lane;
  if (p.op == 1 && warp == 0) {
    act_pre = lane < A ? p.actions[e * A + lane] : 0;
    if (!p.order && A > 1) {
      // rank = this agent's position in the step's order.  Fisher-Yates: draw d (for i = A-1-d) is word d % 4 of Philox
      // block ctr0 + d / 4; instead of permuting an array serially every lane tracks ITS OWN element through the swaps
      // (p == i -> j, p == j -> i): the draws are broadcast by independent shuffles, no lane waits for another.
      uint32_t u[4];
      philox4x32_10(id0, id1, ctr0 + (uint32_t)(lane >> 2), 0u, k0, k1, u);
      const uint32_t word = (lane & 2) ? ((lane & 1) ? u[3] : u[2]) : ((lane & 1) ? u[1] : u[0]);
      const int jl = (int)__umulhi(word, (uint32_t)(A - lane));   // below(i + 1), i = A - 1 - lane
      for (int d = 0; d < A - 1; ++d) {
        const int i = A - 1 - d, j = __shfl_sync(0xffffffffu, jl, d);
        rank_pre = rank_pre == i ? j : (rank_pre == j ? i : rank_pre);
      }
    }
  }
  mbar_wait(&bar, 0);
  __syncthreads();

  if (p.op == 1) {
    h.x += 1; h.y += 1;  // step_count, tick
    // ---- 1/2. ordered agent moves: warp 0, lane = agent
    order_blocks = (!p.order && A > 1) ? (uint32_t)(A - 1 + 3) / 4u : 0u;
    if (warp == 0) {
      int rank = rank_pre;   // (computed above, under the terrain load; a replayed order is read here)
      if (p.order) {
        int* s_rank = s_adir;                // scratch: s_adir is read into `dir` first
        const int d0 = lane < A ? s_adir[lane] : 0;
        __syncwarp();
        if (lane < A) s_rank[p.order[e * A + lane]] = lane;
        __syncwarp();
        if (lane < A) rank = s_rank[lane];
        __syncwarp();
        if (lane < A) s_adir[lane] = d0;
        __syncwarp();
      }
      // an agent acts once per step, so its target cell is known up front; positions travel packed as x | y << 8.
      // `tgt` = the cell the agent will stand on if nothing blocks it (its own cell when it stays).
      const int a = act_pre;
      int x = lane < A ? s_ax[lane] : 0, y = lane < A ? s_ay[lane] : 0, dir = lane < A ? s_adir[lane] : 0;
      const int dx = (a == 4) - (a == 2), dy = (a == 3) - (a == 1);
      const int nx = x + dx, ny = y + dy;
      const bool wants = lane < A && a >= 1 && a <= 4 && nx >= 0 && ny >= 0 && nx < W && ny < H;
      const uint32_t cur = lane < A ? ((uint32_t)x | ((uint32_t)y << 8)) : (0xFFFF0000u | (uint32_t)lane);   // idle lanes: unique, never a cell
      const uint32_t tgt = wants ? ((uint32_t)nx | ((uint32_t)ny << 8)) : cur;
      // A move can only be blocked (or depend on the order at all) if another agent stands on the target now, stays on it,
      // or wants the same target: those few lanes are resolved in rank order below, everybody else just moves.
      // (every lane executes every warp collective: no short-circuit in front of a *_sync call)
      const unsigned same_tgt = __match_any_sync(0xffffffffu, tgt);             // same target, or the target of a stayer
      bool conflict = __popc(same_tgt) > 1;
      {  // "someone stands on my target now": one bit per cell in the (not yet used) queue area instead of A shuffles - every lane
         // clears the words it will touch, the agents mark their cells, and a mover's own cell is never its target
        uint32_t* bm = reinterpret_cast<uint32_t*>(s_queue);
        const int cc = lane < A ? x * H + y : 0, ct = wants ? nx * H + ny : cc;
        bm[cc >> 5] = 0u; bm[ct >> 5] = 0u;
        __syncwarp();
        if (lane < A) atomicOr(&bm[cc >> 5], 1u << (cc & 31));
        __syncwarp();
        conflict |= ((bm[ct >> 5] >> (ct & 31)) & 1u) != 0;
        __syncwarp();
      }
      conflict &= wants;
      uint32_t fin = (wants && !conflict) ? tgt : cur;   // position after this agent's own turn
      unsigned rem = __ballot_sync(0xffffffffu, conflict);
      while (rem) {  // sequential semantics for the conflicting lanes: at k's turn agent j is at (rank_j < rank_k ? fin_j : cur_j)
        const int rk = (int)__reduce_min_sync(0xffffffffu, (unsigned)(((rem >> lane) & 1u) ? rank : 255));
        const int k = __ffs(__ballot_sync(0xffffffffu, ((rem >> lane) & 1u) && rank == rk)) - 1;
        const uint32_t t = __shfl_sync(0xffffffffu, tgt, k);
        const unsigned occupied = __ballot_sync(0xffffffffu, lane != k && (rank < rk ? fin : cur) == t);
        if (lane == k && !occupied) fin = tgt;
        rem &= ~(1u << k);
      }
      if (fin != cur) dir = dx == 1 ? 0 : (dy == 1 ? 1 : (dx == -1 ? 2 : 3));  // DIR_TO_VEC (constants.py:65-74): dir follows an actual move
      const uint32_t cur_after = fin;
      // extinguish the burning cell under each agent: agents stand on distinct cells and never act twice, so this
      // does not depend on the order
      if (lane < A) {
        x = (int)(cur_after & 255u); y = (int)(cur_after >> 8);
        double rew = 0.0;
        if (s_told[x * H + y] == WF_BURNING) { s_told[x * H + y] = WF_BURNT; rew = 1.0; }
        s_ax[lane] = x; s_ay[lane] = y; s_adir[lane] = dir; p.rewards[e * A + lane] = rew;
      }
    }
    __syncthreads();

    // ---- 3. fire dynamics, four cells per word, VEC words per thread and iteration
    //      pass 1: neighbour counts for every word; quiet words are copied, words on the fire front are queued
    //      (word index + packed state / neighbour-count bytes) in the not-yet-used obs area
    const uint32_t tick = (uint32_t)h.y;
    uint2* s_list = reinterpret_cast<uint2*>(s_queue);
    const int rv = rw / VEC;   // vectors per row (VEC == 4 only when H % 16 == 0)
    for (int q = tid; q < nwords / VEC; q += T) {
      const int j0 = q * VEC;
      const int vy = rv == 1 ? 0 : q - (int)__umulhi((uint32_t)q, p.rv_magic) * rv;   // vector index inside its row
      uint32_t w[VEC], up[VEC], dn[VEC];
      if (VEC == 4) {
        const uint4 a = reinterpret_cast<const uint4*>(t32)[q], u4 = *reinterpret_cast<const uint4*>(t32 + j0 - rw),
                    d4 = *reinterpret_cast<const uint4*>(t32 + j0 + rw);
        w[0] = a.x; w[VEC > 1 ? 1 : 0] = a.y; w[VEC > 2 ? 2 : 0] = a.z; w[VEC > 3 ? 3 : 0] = a.w;
        up[0] = u4.x; up[VEC > 1 ? 1 : 0] = u4.y; up[VEC > 2 ? 2 : 0] = u4.z; up[VEC > 3 ? 3 : 0] = u4.w;
        dn[0] = d4.x; dn[VEC > 1 ? 1 : 0] = d4.y; dn[VEC > 2 ? 2 : 0] = d4.z; dn[VEC > 3 ? 3 : 0] = d4.w;
      } else {
        w[0] = t32[j0]; up[0] = t32[j0 - rw]; dn[0] = t32[j0 + rw];
      }
      const uint32_t pv = vy > 0 ? (t32[j0 - 1] & 0x01010101u) : 0u, nx = vy < rv - 1 ? (t32[j0 + VEC] & 0x01010101u) : 0u;
      uint32_t b[VEC], code[VEC];
      int n_act = 0;
#pragma unroll
      for (int i = 0; i < VEC; ++i) b[i] = w[i] & 0x01010101u;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const uint32_t left = i > 0 ? b[i > 0 ? i - 1 : 0] : pv, right = i < VEC - 1 ? b[i + 1 < VEC ? i + 1 : 0] : nx;
        const uint32_t k4 = (up[i] & 0x01010101u) + (dn[i] & 0x01010101u) + ((b[i] << 8) | (left >> 24)) + ((b[i] >> 8) | (right << 24));
        const uint32_t healthy = ~(w[i] | (w[i] >> 1)) & 0x01010101u;
        const uint32_t kh = k4 & (healthy * 7u);           // burning 4-neighbours of the healthy cells (<= 4 per byte)
        code[i] = (kh | b[i]) ? (w[i] | (kh << 2)) : 0u;   // != 0 <=> some cell of the word can change
        n_act += code[i] != 0;
      }
      if (VEC == 4) reinterpret_cast<uint4*>(n32)[q] = make_uint4(w[0], w[VEC > 1 ? 1 : 0], w[VEC > 2 ? 2 : 0], w[VEC > 3 ? 3 : 0]);
      else n32[j0] = w[0];
      if (n_act) {
        int pos = atomicAdd(&s_count, n_act);
#pragma unroll
        for (int i = 0; i < VEC; ++i)
          if (code[i]) s_list[pos++] = make_uint2((uint32_t)(j0 + i), code[i]);
      }
    }
    __syncthreads();
    //      pass 2: one Philox block per queued word, all lanes busy
    bool any_burning = false;
    const int n_list = s_count;
    for (int idx = tid; idx < n_list; idx += T) {
      const uint2 ent = s_list[idx];
      uint32_t u[4];
      philox4x32_10(id0, id1, tick, 1u + ent.x, k0, k1, u);
      uint32_t nw = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t s = (ent.y >> (8 * i)) & 3u, k = (ent.y >> (8 * i + 2)) & 7u;
        uint32_t ns = s;
        if (s == WF_BURNING) ns = u[i] < p.burnout_threshold ? WF_BURNT : WF_BURNING;
        else if (k) ns = u[i] < p.ignite_threshold[k] ? WF_BURNING : WF_HEALTHY;
        nw |= ns << (8 * i);
      }
      n32[ent.x] = nw;
      any_burning |= (nw & 0x01010101u) != 0;
    }
    // ---- 4. termination, same-step autoreset
    const bool term = !__syncthreads_or(any_burning), trunc = h.x >= p.max_steps;
    if (tid == 0) { p.terminated[e] = term; p.truncated[e] = trunc; }
    need_reset = p.autoreset && (term || trunc);
    if (need_reset && p.final_obs) {   // (the queue is consumed: every thread has passed the barrier above)
      wf_encode_tile_cold<T, VEC>(p, n32, s_obs, s_ax, s_ay, s_adir, nwords, tid);
      __syncthreads();
      uint4* dst = reinterpret_cast<uint4*>(p.final_obs + e * 3 * cells);
      for (int i = tid; i < 3 * cells / 16; i += T) dst[i] = reinterpret_cast<const uint4*>(s_obs)[i];
      __syncthreads();
    }
  } else if (!do_reset) {
    for (int j = tid; j < nwords; j += T) n32[j] = t32[j];
  }
  // ---- reset(mask) / autoreset (rare): all healthy, then thread 0 places fires and agents from the env's Philox stream
  uint32_t ctr_end = ctr0 + order_blocks;
  if (need_reset) {
    for (int j = tid; j < nwords; j += T) n32[j] = 0u;
    __syncthreads();
    if (tid == 0) {
      Rng<1> r;
      r.open_philox(p.seed, env_id, ctr_end);
      const int used = (A - 1) & 3;   // words of the last order block already consumed (0 = block boundary)
      if (order_blocks && used) {     // the sequential stream continues inside that block
        uint32_t u[4];
        philox4x32_10(id0, id1, ctr_end - 1u, 0u, k0, k1, u);
        r.have = 4 - used;
        r.b0 = u[used]; r.b1 = used + 1 < 4 ? u[used + 1] : 0u; r.b2 = used + 2 < 4 ? u[used + 2] : 0u; r.b3 = 0u;
      }
      wf_reset_agents_fires(p, s_tnew, s_ax, s_ay, s_adir, r);
      ctr_end = r.ctr;
    }
    h.x = 0; h.w += 1;
  }
  __syncthreads();

  // ---- 5. observation (type byte = state; colour green / red / grey) + write-back
  if (p.obs) wf_encode_tile<T, VEC>(p, n32, s_obs, s_ax, s_ay, s_adir, nwords, tid);
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    tma_store_1d(g_terrain, s_tnew, (uint32_t)cells);
    if (p.obs) tma_store_1d(p.obs + e * 3 * cells, s_obs, (uint32_t)(3 * cells));
    tma_commit();
    h.z = (int)ctr_end;
    p.hdr[e] = h;
  }
  if (tid < A) reinterpret_cast<uchar4*>(p.agents)[e * A + tid] = make_uchar4((uint8_t)s_ax[tid], (uint8_t)s_ay[tid], (uint8_t)s_adir[tid], 0);
  if (tid == 0) tma_wait_read_all();
}

static bool wf_fast(int H) {
  static const bool off = [] { const char* v = std::getenv("MG_WF_GENERIC"); return v && v[0] == '1'; }();
  return !off && (H & 3) == 0;
}
static size_t wf_fast_smem(int cells, int H) { return (size_t)cells * 4 + 2 * (size_t)((H + 15) & ~15) + 16; }

size_t wildfire_smem_bytes(int cells, int H) { return wf_fast(H) ? wf_fast_smem(cells, H) : (size_t)cells * 5 + 64; }

cudaError_t configure_wildfire_kernel(int cells, int H) {
  const int bytes = (int)wildfire_smem_bytes(cells, H);
  cudaError_t e;
  if ((e = raise_smem_limit((const void*)wildfire_kernel, (size_t)bytes)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)wildfire_fast_kernel<64, 1>, (size_t)bytes)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)wildfire_fast_kernel<128, 1>, (size_t)bytes)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)wildfire_fast_kernel<64, 4>, (size_t)bytes)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)wildfire_fast_kernel<256, 4>, (size_t)bytes)) != cudaSuccess) return e;
  return raise_smem_limit((const void*)wildfire_fast_kernel<128, 4>, (size_t)bytes);
}

cudaError_t launch_wildfire(const WildfireParams& p, cudaStream_t st) {
  const bool fast = wf_fast(p.H);
  const bool vec = (p.H & 15) == 0;   // rows are whole 16-byte vectors
  const int items = p.cells / (vec ? 16 : 4);   // words / vectors one CTA sweeps
  int threads = fast ? (vec && items >= 1024 ? 256 : (items >= 256 ? 128 : 64)) : kWfThreads;
  static const int forced = [] { const char* v = std::getenv("MG_WF_THREADS"); return v ? std::atoi(v) : 0; }();   // experiment switch
  if (fast && (forced == 64 || forced == 128 || (forced == 256 && vec))) threads = forced;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)p.N); cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = wildfire_smem_bytes(p.cells, p.H); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = [] { const char* v = std::getenv("MG_PDL"); return !(v && v[0] == '0'); }();
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  if (!fast) return cudaLaunchKernelEx(&cfg, wildfire_kernel, p);
  if (vec && threads == 256) return cudaLaunchKernelEx(&cfg, wildfire_fast_kernel<256, 4>, p);
  if (vec) return threads == 128 ? cudaLaunchKernelEx(&cfg, wildfire_fast_kernel<128, 4>, p) : cudaLaunchKernelEx(&cfg, wildfire_fast_kernel<64, 4>, p);
  return threads == 128 ? cudaLaunchKernelEx(&cfg, wildfire_fast_kernel<128, 1>, p) : cudaLaunchKernelEx(&cfg, wildfire_fast_kernel<64, 1>, p);
}

}  // namespace mg
