// map_kernels.cu -- sm_100a kernels for the static-map families:
//   MazeSingleAgentEnv.step / reset  (envs/maze.py:180-219, 245-260, 271-377)
//   CtFMvNEnv.step / reset           (envs/ctf.py:998-1075, 1137-1163, 1184-1251, 1292-1433)
//   Ctf1v1Env.step                   (envs/ctf.py:503-510, 551-654)
//
// Per-env state is one packed word per agent (x | y << 8 | dir << 16 | flags << 24) plus a 16-byte header; the
// map is shared and lives in handle-owned device tables.  One CTA = a tile of 128 envs, one thread per env:
//   * the thread's agent row and header arrive as 128-bit loads (rows are padded to a power of two of agents, so
//     a warp reads one contiguous run); the words sit in shared memory TRANSPOSED ([agent][env]) so that the
//     order-dependent agent loop indexes them dynamically without bank conflicts and without local memory;
//   * the per-step agent order and the actions are nibble-packed in one 64-bit register each (Fisher-Yates swaps
//     are three xors), the battle test is an integer compare of the squared distance against the largest d2 with
//     sqrt(d2) <= battle_range (host-computed, exact), battle draws compare the Philox word with
//     ceil(p * 2^32) - both are bit-exact restatements of the reference's float tests;
//   * the observation - the static map with the agents drawn on top - is assembled in shared memory for the whole
//     tile (16-byte copies of the map's "super-period" = lcm(cells, 16) bytes, staged by one TMA bulk load, then each
//     env's thread patches its agents' cells) and leaves as ONE TMA bulk store; tiles too large for shared memory
//     (64x64 maps, the reference's 8-byte dtypes) stream the period straight to global memory instead.
#include <cstdlib>

#include "mg_device.cuh"
#include "smem_config.h"
#include "map_params.cuh"
#include "view_device.cuh"
#include "policy_device.cuh"

namespace mg {

constexpr int kMapE = 128;  // envs per CTA = threads per CTA

// MazeWorld / CtfWorld codes (world.py:66-91)
constexpr int MZ_AGENT = 1, MZ_FLAG = 2, MZ_OBSTACLE = 3;
constexpr int CT_BLUE_TERR = 0, CT_RED_TERR = 1, CT_BLUE_AGENT = 2, CT_RED_AGENT = 3, CT_BLUE_FLAG = 4, CT_RED_FLAG = 5,
              CT_OBSTACLE = 6;
constexpr uint32_t FL_DEAD = 1u << 24, FL_COLLIDED = 2u << 24;  // Agent.terminated / Agent.collided (agent.py:97-100)
// flags bits 2-3: the CtF agent's sticky background colour, 0 as constructed (the team's), 1 light_blue, 2 light_red: set by a move onto
// a territory / flag cell, left alone elsewhere (ctf.py:1214-1230, agent.py:197-200); only render() reads it
constexpr uint32_t FL_BG_MASK = 12u << 24, FL_BG_BLUE = 4u << 24, FL_BG_RED = 8u << 24;
__device__ __forceinline__ uint32_t bg_after_move(uint32_t w, int terrain_code) {
  // nibble table over the CtfWorld codes: blue territory 0 / blue flag 4 -> 1, red territory 1 / red flag 5 -> 2, anything else 0 = keep
  const uint32_t v = (0x0210021u >> (4 * terrain_code)) & 3u;
  return v ? ((w & ~FL_BG_MASK) | (v << 26)) : w;
}

__device__ __forceinline__ uint32_t ag_pack(int x, int y, int dir, int fl) {
  return (uint32_t)x | ((uint32_t)y << 8) | ((uint32_t)dir << 16) | ((uint32_t)fl << 24);
}
__device__ __forceinline__ int ag_x(uint32_t w) { return (int)(w & 255u); }
__device__ __forceinline__ int ag_y(uint32_t w) { return (int)((w >> 8) & 255u); }

// CtfActions / MazeActions: 0 stay, 1 left (0,-1), 2 down (-1,0), 3 right (0,+1), 4 up (+1,0)  (agent.py:54-67)
__device__ __forceinline__ void action_delta(int a, int& dx, int& dy) {  // two-bit lookup tables holding delta + 1, a in [0, 4]
  dx = (int)((0x245u >> (2 * a)) & 3u) - 1;
  dy = (int)((0x191u >> (2 * a)) & 3u) - 1;
}
// DIR_TO_VEC (constants.py:65-74) for the four unit moves; Agent.move leaves dir alone when no vector matches (agent.py:176-183)
__device__ __forceinline__ int dir_of_action(int a, int old) { return a == 0 ? old : ((4 - a) & 3); }

template <int MODE>
__device__ __forceinline__ int below(Rng<MODE>& r, int n) { return (int)__umulhi(r.u32(), (uint32_t)n); }
// the step's small draws (n <= 16): 16 bits each, multiply-shift
template <int MODE>
__device__ __forceinline__ int below16(Rng<MODE>& r, int n) { return (int)((r.u16() * (uint32_t)n) >> 16); }

// `ag` = this env's column of the transposed agent words: agent i at ag[i * kMapE].  Cell lists hold packed x | y << 8.
template <int FAMILY, int MODE>
__device__ __forceinline__ void reset_one(const MapParams& p, long long e, uint32_t* ag, Rng<MODE>& r) {
  if (FAMILY == MG_FAMILY_MAZE) {  // maze.py:202-205: agent on a random background cell, dir 3
    const int idx = MODE == 0 ? p.start_index[e] : below(r, p.n_background);
    ag[0] = (uint32_t)p.background[idx] | (3u << 16);
  } else {  // ctf.py:1033-1048; flags 0: a fresh env instance - or, with carry_flags, the SAME instance going on: the reference never
            // clears terminated / collided / bg_color on reset (agent.py:97-100 are their only assignments outside step)
    const uint32_t keep = p.carry_flags ? 0xFF000000u : 0u;
    for (int team = 0; team < 2; ++team) {
      const int k = team ? p.nr : p.nb, len = team ? p.len_red : p.len_blue, base = team ? p.nb : 0;
      const uint16_t* terr = team ? p.red_terr : p.blue_terr;
      const int32_t* place = team ? p.red_place : p.blue_place;
      unsigned long long lo = 0, hi = 0;   // np_random.choice(len, k, replace=False) stand-in: rejection over a bitmap of <= 128 entries
      for (int i = 0; i < k; ++i) {
        int v;
        if (MODE == 0) v = place[e * k + i];
        else if (len <= 128) {
          for (;;) {
            v = below(r, len);
            const unsigned long long bit = 1ull << (v & 63);
            if (v < 64) { if (lo & bit) continue; lo |= bit; }
            else { if (hi & bit) continue; hi |= bit; }
            break;
          }
        } else {  // large territories: compare with the agents already placed
          for (;;) {
            v = below(r, len);
            const uint32_t cand = terr[v];
            bool dup = false;
            for (int j = 0; j < i; ++j) dup |= (ag[(base + j) * kMapE] & 0xFFFFu) == cand;
            if (!dup) break;
          }
        }
        ag[(base + i) * kMapE] = (uint32_t)terr[v] | (3u << 16) | (ag[(base + i) * kMapE] & keep);
      }
    }
  }
}

// MazeSingleAgentEnv.step on the env's single agent word
__device__ __forceinline__ void maze_step_one(const MapParams& p, int a, uint32_t& w, int4& h, double& rew, bool& term,
                                              bool& trunc, int& err) {
  const int S = p.S;
  h.x += 1;  // maze.py:334
  if (a < 0 || a > 4) err |= MG_ERR_BAD_ACTION;  // reference: ValueError (maze.py:286)
  else if (a != 0) {  // _move_agent maze.py:271-307; staying = moving onto the agent's own cell, which never overlaps (object.py:38-40)
    int dx, dy;
    action_delta(a, dx, dy);
    const int nx = ag_x(w) + dx, ny = ag_y(w) + dy;
    if (!(nx < 0 || ny < 0 || nx >= S || ny >= S)) {
      // every cell holds an object: background Floor / Flag overlap, Obstacle overlaps iff penalty != 0 (object.py:200-201)
      const int code = __ldg(p.field_map + nx * S + ny);
      if (code != MZ_OBSTACLE || p.obstacle_penalty != 0) w = ag_pack(nx, ny, dir_of_action(a, 0), (int)(w >> 24));
    }
  }
  term = false; trunc = h.x >= p.max_steps;  // :346-347
  rew = 0.0;
  const int here = __ldg(p.field_map + ag_x(w) * S + ag_y(w));
  if (here == MZ_FLAG) { rew += p.flag_reward; term = true; }                                        // :354-356
  if (p.obstacle_penalty != 0 && here == MZ_OBSTACLE) { rew -= p.obstacle_penalty; term = true; }    // :360-363
  rew -= p.step_penalty;                                                                             // :371
}

// CtFMvNEnv.step / Ctf1v1Env.step for ONE env.  `terr` = the map in observation order (s_period), `ag` as in reset_one.
// NB / NR: compile-time team sizes of the common configurations (all loops unroll), 0 = read them from the parameters.
template <int MODE, typename NIB, int NB, int NR>
__device__ __forceinline__ void ctf_step_one(const MapParams& p, long long e, const int8_t* blue_act, const uint8_t* terr,
                                             uint32_t* ag, int4& h, Rng<MODE>& r, double& rew, bool& term, bool& trunc,
                                             int& err) {
  const int S = p.S, nb = NB ? NB : p.nb, nr = NB ? NR : p.nr, n = nb + nr;
  h.x += 1;  // ctf.py:1295
  // actions and order, one nibble per agent (n <= 16); action nibble 15 = outside the action set
  NIB acts = 0, order = 0;   // NIB = uint32_t when n <= 8, else 64 bits
  for (int i = 0; i < nb; ++i) {
    const int a = blue_act[i];
    if (a < 0 || a > 4) err |= MG_ERR_BAD_ACTION;  // reference: ValueError (ctf.py:1200-1201)
    acts |= (NIB)((a < 0 || a > 4) ? 15 : a) << (4 * i);
  }
  for (int k = 0; k < nr; ++k) {  // RwPolicy.act for EVERY red agent, defeated or not (:1297-1301)
    // Philox mode with red_actions given: an external enemy policy (the reference's `enemy_policies`, ctf.py:666) - no draw
    const int a = (MODE == 0 || p.red_actions) ? p.red_actions[e * nr + k] : below16(r, 5);
    if (a < 0 || a > 4) err |= MG_ERR_BAD_ACTION;
    acts |= (NIB)((a < 0 || a > 4) ? 15 : a) << (4 * (nb + k));
  }
  if (p.variant_1v1) {  // Ctf1v1Env._move_agents: blue, then red (ctf.py:503-510)
    order = (NIB)0x10;
  } else if (MODE == 0) {
    for (int i = 0; i < n; ++i) order |= (NIB)(p.order[e * n + i] & 15) << (4 * i);
  } else {  // np_random.shuffle stand-in: Fisher-Yates
    order = (NIB)0xFEDCBA9876543210ull;
    for (int i = n - 1; i > 0; --i) {
      const int j = below16(r, i + 1);
      const NIB x = ((order >> (4 * i)) ^ (order >> (4 * j))) & (NIB)15;
      order ^= (x << (4 * i)) | (x << (4 * j));
    }
  }
  for (int k = 0; k < n; ++k) {  // _move_agents :1240-1251
    const int i = (int)((order >> (4 * k)) & (NIB)15);
    const uint32_t w = ag[i * kMapE];
    if (w & FL_DEAD) continue;  // "Defeated agent doesn't move, sadly."
    const int a = (int)((acts >> (4 * i)) & (NIB)15);
    if (a == 15) continue;
    int dx, dy;
    action_delta(a, dx, dy);
    const int nx = ag_x(w) + dx, ny = ag_y(w) + dy;  // _move_agent :1184-1238
    if (nx < 0 || ny < 0 || nx >= S || ny >= S) continue;
    const uint32_t target = (uint32_t)nx | ((uint32_t)ny << 8);
    bool occupied = false;  // an agent object (alive, defeated, or itself when staying) sits on the cell
    for (int j = 0; j < n; ++j) occupied |= ((ag[j * kMapE] ^ target) & 0xFFFFu) == 0;
    if (occupied) { if (p.obstacle_penalty != 0 && !p.variant_1v1) ag[i * kMapE] = w | FL_COLLIDED; continue; }  // :1231-1236 (1v1 has no collided logic, :498-501)
    const int tc = terr[ny * S + nx];
    if (tc == CT_OBSTACLE && p.obstacle_penalty == 0) continue;  // Obstacle.can_overlap()
    ag[i * kMapE] = (bg_after_move(w, tc) & 0xFF000000u) | target | ((uint32_t)dir_of_action(a, (int)((w >> 16) & 255u)) << 16);  // Agent.move agent.py:167-200
  }
  term = false; trunc = h.x >= p.max_steps;  // :1310-1311
  rew = 0.0;
  if (p.obstacle_penalty != 0) {  // :1316-1332 (collided is never cleared)
    for (int i = 0; i < n; ++i) {
      const uint32_t w = ag[i * kMapE];
      if (w & FL_COLLIDED) { if (i < nb) rew -= p.obstacle_penalty; ag[i * kMapE] = w | FL_DEAD; }
    }
  }
  const uint32_t red_flag = (uint32_t)p.red_flag, blue_flag = (uint32_t)p.blue_flag;
  // h.y = game_stats (ctf.py:1068-1073): bit0 blue_flag_captured, bit1 red_flag_captured, bit 8+i agent i defeated in a battle
  for (int i = 0; i < nb; ++i) if (((ag[i * kMapE] ^ red_flag) & 0xFFFFu) == 0) { rew += p.flag_reward; term = true; h.y |= 2; }   // :1335-1344
  for (int i = nb; i < n; ++i) if (((ag[i * kMapE] ^ blue_flag) & 0xFFFFu) == 0) { rew -= p.flag_reward; term = true; h.y |= 1; }  // :1347-1356
  int nbattle = 0;
  bool all_dead = true;
  for (int b = 0; b < nb; ++b) {  // np.where(distances <= battle_range): row-major, blue-major (:1368-1377)
    uint32_t wb = ag[b * kMapE];
    for (int q = 0; q < nr; ++q) {
      const uint32_t wr = ag[(nb + q) * kMapE];
      const int ddx = ag_x(wb) - ag_x(wr), ddy = ag_y(wb) - ag_y(wr);
      if (ddx * ddx + ddy * ddy > p.d2_max) continue;  // == !(np.linalg.norm(int vector) <= battle_range), see MapParams::d2_max
      if ((wb | wr) & FL_DEAD) continue;               // :1380-1383
      const int cb = terr[ag_y(wb) * S + ag_x(wb)], cr = terr[ag_y(wr) * S + ag_x(wr)];
      const bool bh = (cb == CT_BLUE_TERR || cb == CT_BLUE_FLAG), rh = (cr == CT_RED_TERR || cr == CT_RED_FLAG);
      bool blue_win;
      if (MODE == 0) {
        blue_win = nbattle < p.KB ? p.blue_win[e * p.KB + nbattle] != 0 : false;
        if (nbattle >= p.KB) err |= MG_ERR_TRACE_OVERFLOW;
      } else {  // :1392-1407; (double)u / 2^32 < p  <=>  u < ceil(p * 2^32)
        const unsigned long long thr = (bh && !rh) ? p.thr_blue_home : ((!bh && rh) ? p.thr_red_home : p.thr_even);
        blue_win = (unsigned long long)r.u32() < thr;
      }
      ++nbattle;
      if (blue_win) { rew += p.battle_reward; ag[(nb + q) * kMapE] = wr | FL_DEAD; h.y |= 1 << (8 + nb + q); }   // :1409-1418
      else if (p.variant_1v1) { rew -= p.battle_reward; term = true; h.y |= 1 << 8; }  // 1v1: losing ends the episode (ctf.py:629-636)
      else { rew -= p.battle_reward; wb |= FL_DEAD; ag[b * kMapE] = wb; h.y |= 1 << (8 + b); }
    }
    all_dead &= (wb & FL_DEAD) != 0;
  }
  if (MODE == 0 && p.battles_used) p.battles_used[e] = nbattle;
  if (all_dead) term = true;           // :1423
  rew = __dsub_rn(rew, __dmul_rn(p.step_penalty, (double)nb));  // :1428 -- two roundings like the reference, never an FMA
}

// The general step (run-time team sizes) for maps of <= 256 cells - STEPV 3.  Same statements as ctf_step_one, two changes in HOW:
//   * "does an agent stand on the target cell" is one bit of a per-env occupancy bitmap (`occ`: kOccWords words per env,
//     transposed like the agent words; built from the agent rows, updated by every move) instead of a scan over all agents -
//     the scan was a quarter of an 8v8 step's instructions (16 moves x 16 compares);
//   * positions do not change during the battle phase, so the in-range test of a blue agent against all reds runs first,
//     branch-free and two reds per instruction (positions byte-packed in `occ`'s second half: __vabsdiffu4 + two dp4a give
//     both squared distances), and only the pairs in range - visited in the reference's row-major order - take the path
//     with the dead-flag test, the terrain look-ups and the draw.
constexpr int kOccWords = 8, kOccBytes = 2 * kOccWords * kMapE * 4;   // bitmap words + packed red pairs (<= 8 words) per env
template <int MODE, typename NIB>
__device__ __forceinline__ void ctf_step_occ(const MapParams& p, long long e, const int8_t* blue_act, const uint8_t* terr,
                                             uint32_t* ag, uint32_t* occ, int4& h, Rng<MODE>& r, double& rew, bool& term, bool& trunc,
                                             int& err) {
  const int S = p.S, nb = p.nb, nr = p.nr, n = nb + nr;
  uint32_t* pr = occ + kOccWords * kMapE;
  h.x += 1;  // ctf.py:1295
  NIB acts = 0, order = 0;
  for (int i = 0; i < nb; ++i) {
    const int a = blue_act[i];
    const bool bad = a < 0 || a > 4;   // reference: ValueError (ctf.py:1200-1201)
    if (bad) err |= MG_ERR_BAD_ACTION;
    acts |= (NIB)(bad ? 15 : a) << (4 * i);
  }
  for (int k = 0; k < nr; ++k) {  // RwPolicy.act for EVERY red agent, defeated or not (:1297-1301), or the external enemy policy
    const int a = (MODE == 0 || p.red_actions) ? p.red_actions[e * nr + k] : below16(r, 5);
    const bool bad = a < 0 || a > 4;
    if (bad) err |= MG_ERR_BAD_ACTION;
    acts |= (NIB)(bad ? 15 : a) << (4 * (nb + k));
  }
  if (p.variant_1v1) {  // Ctf1v1Env._move_agents: blue, then red (ctf.py:503-510)
    order = (NIB)0x10;
  } else if (MODE == 0) {
    for (int i = 0; i < n; ++i) order |= (NIB)(p.order[e * n + i] & 15) << (4 * i);
  } else {  // np_random.shuffle stand-in: Fisher-Yates
    order = (NIB)0xFEDCBA9876543210ull;
    for (int i = n - 1; i > 0; --i) {
      const int j = below16(r, i + 1);
      const NIB x = ((order >> (4 * i)) ^ (order >> (4 * j))) & (NIB)15;
      order ^= (x << (4 * i)) | (x << (4 * j));
    }
  }
  // occupancy bitmap over the cells in observation order (y * S + x): every agent object, alive or defeated
#pragma unroll
  for (int k = 0; k < kOccWords; ++k) occ[k * kMapE] = 0u;
  for (int j = 0; j < n; ++j) {
    const uint32_t w = ag[j * kMapE];
    const int c = ag_y(w) * S + ag_x(w);
    occ[(c >> 5) * kMapE] |= 1u << (c & 31);
  }
  const bool pen = p.obstacle_penalty != 0;
  for (int k = 0; k < n; ++k) {  // _move_agents :1240-1251
    const int i = (int)((order >> (4 * k)) & (NIB)15);
    const uint32_t w = ag[i * kMapE];
    const int a = (int)((acts >> (4 * i)) & (NIB)15);
    int dx, dy;
    action_delta(a & 7, dx, dy);
    const int nx = ag_x(w) + dx, ny = ag_y(w) + dy;  // _move_agent :1184-1238
    const bool live = !(w & FL_DEAD) && a != 15;     // "Defeated agent doesn't move, sadly."
    const bool inb = (unsigned)nx < (unsigned)S && (unsigned)ny < (unsigned)S;
    const int c = inb ? ny * S + nx : 0;
    const uint32_t bit = 1u << (c & 31);
    uint32_t* ow = occ + (c >> 5) * kMapE;
    const bool occupied = (*ow & bit) != 0;   // an agent object (alive, defeated, or itself when staying) sits on the cell
    const int tc = terr[c];
    if (live && inb) {
      if (occupied) {
        if (pen && !p.variant_1v1) ag[i * kMapE] = w | FL_COLLIDED;   // :1231-1236 (1v1 has no collided logic, :498-501)
      } else if (!(tc == CT_OBSTACLE && !pen)) {                      // Obstacle.can_overlap()
        const int c0 = ag_y(w) * S + ag_x(w);
        occ[(c0 >> 5) * kMapE] &= ~(1u << (c0 & 31));
        *ow |= bit;
        ag[i * kMapE] = (bg_after_move(w, tc) & 0xFF000000u) | (uint32_t)nx | ((uint32_t)ny << 8) |
                        ((uint32_t)dir_of_action(a, (int)((w >> 16) & 255u)) << 16);  // Agent.move agent.py:167-200
      }
    }
  }
  term = false; trunc = h.x >= p.max_steps;  // :1310-1311
  rew = 0.0;
  if (pen) {  // :1316-1332 (collided is never cleared)
    for (int i = 0; i < n; ++i) {
      const uint32_t w = ag[i * kMapE];
      if (w & FL_COLLIDED) { if (i < nb) rew -= p.obstacle_penalty; ag[i * kMapE] = w | FL_DEAD; }
    }
  }
  const uint32_t red_flag = (uint32_t)p.red_flag, blue_flag = (uint32_t)p.blue_flag;
  // h.y = game_stats (ctf.py:1068-1073): bit0 blue_flag_captured, bit1 red_flag_captured, bit 8+i agent i defeated in a battle
  for (int i = 0; i < nb; ++i) if (((ag[i * kMapE] ^ red_flag) & 0xFFFFu) == 0) { rew += p.flag_reward; term = true; h.y |= 2; }   // :1335-1344
  // red positions two per word for the range tests; the flag test of the reds rides along (:1347-1356)
  const int nr2 = (nr + 1) >> 1;
  for (int j = 0; j < nr2; ++j) {
    const uint32_t a0 = ag[(nb + 2 * j) * kMapE] & 0xFFFFu;
    const uint32_t a1 = 2 * j + 1 < nr ? (ag[(nb + 2 * j + 1) * kMapE] & 0xFFFFu) : a0;
    if (a0 == blue_flag || (2 * j + 1 < nr && a1 == blue_flag)) { term = true; h.y |= 1; }
    if (a0 == blue_flag) rew -= p.flag_reward;
    if (2 * j + 1 < nr && a1 == blue_flag) rew -= p.flag_reward;
    pr[j * kMapE] = a0 | (a1 << 16);
  }
  const uint32_t row_valid = (1u << nr) - 1u;
  int nbattle = 0;
  bool all_dead = true;
  for (int b = 0; b < nb; ++b) {  // np.where(distances <= battle_range): row-major, blue-major (:1368-1377)
    uint32_t wb = ag[b * kMapE];
    const uint32_t bb = (wb & 0xFFFFu) * 0x10001u;
    uint32_t row = 0;   // bit q: red q within battle range of blue b
    for (int j = 0; j < nr2; ++j) {
      const uint32_t d = __vabsdiffu4(bb, pr[j * kMapE]);   // |dx0| |dy0| |dx1| |dy1|
      const int d2a = (int)__dp4a(d, d & 0x0000FFFFu, 0u), d2b = (int)__dp4a(d, d & 0xFFFF0000u, 0u);
      // == (np.linalg.norm(int vector) <= battle_range), see MapParams::d2_max
      row |= ((d2a <= p.d2_max ? 1u : 0u) | (d2b <= p.d2_max ? 2u : 0u)) << (2 * j);
    }
    row &= row_valid;
    while (row) {
      const int q = __ffs((int)row) - 1;
      row &= row - 1;
      const uint32_t wr = ag[(nb + q) * kMapE];
      if ((wb | wr) & FL_DEAD) continue;               // :1380-1383
      const int cb = terr[ag_y(wb) * S + ag_x(wb)], cr = terr[ag_y(wr) * S + ag_x(wr)];
      const bool bh = (cb == CT_BLUE_TERR || cb == CT_BLUE_FLAG), rh = (cr == CT_RED_TERR || cr == CT_RED_FLAG);
      bool blue_win;
      if (MODE == 0) {
        blue_win = nbattle < p.KB ? p.blue_win[e * p.KB + nbattle] != 0 : false;
        if (nbattle >= p.KB) err |= MG_ERR_TRACE_OVERFLOW;
      } else {  // :1392-1407; (double)u / 2^32 < p  <=>  u < ceil(p * 2^32)
        const unsigned long long thr = (bh && !rh) ? p.thr_blue_home : ((!bh && rh) ? p.thr_red_home : p.thr_even);
        blue_win = (unsigned long long)r.u32() < thr;
      }
      ++nbattle;
      if (blue_win) { rew += p.battle_reward; ag[(nb + q) * kMapE] = wr | FL_DEAD; h.y |= 1 << (8 + nb + q); }   // :1409-1418
      else if (p.variant_1v1) { rew -= p.battle_reward; term = true; h.y |= 1 << 8; }  // 1v1: losing ends the episode (ctf.py:629-636)
      else { rew -= p.battle_reward; wb |= FL_DEAD; ag[b * kMapE] = wb; h.y |= 1 << (8 + b); }
    }
    all_dead &= (wb & FL_DEAD) != 0;
  }
  if (MODE == 0 && p.battles_used) p.battles_used[e] = nbattle;
  if (all_dead) term = true;           // :1423
  rew = __dsub_rn(rew, __dmul_rn(p.step_penalty, (double)nb));  // :1428 -- two roundings like the reference, never an FMA
}

// The scripted opponents of policy_kernels.cu decided inside the step kernel (mg_set_red_policy_fusion), on the agent words the
// step has just loaded: statement for statement ctf_policy_kernel (Philox draws; the validation trace keeps the separate kernel),
// for compile-time team sizes.  `w` = the PRE-step agent words, `h` the pre-step header.  Returns the red actions one byte each
// and writes them where mg_red_policy_actions would.
template <int NB, int NR>
__device__ __forceinline__ uint32_t policy_red_actions(const MapParams& p, long long e, const uint32_t* w, const int4& h) {
  PolicyRng r;
  r.open(p.seed, p.env_id_base + (unsigned long long)e, h.x, h.w);
  const int S = p.S;
  bool intruder = false;   // "a blue agent stands on red ground", shared by every red agent of the env
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int code = __ldg(p.field_map + (w[i] & 255u) * S + ((w[i] >> 8) & 255u));
    intruder |= code == CT_RED_TERR || code == CT_RED_FLAG;   // observation["red_territory"] = red territory cells + the red flag (ctf.py:765-769)
  }
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    const int kind = p.pol_kind[k];
    int a;
    if (kind == MG_POLICY_RW) {
      a = r.below(5);
    } else {
      const uint32_t me = w[NB + k];
      const int x = me & 255u, y = (me >> 8) & 255u, cell = x * S + y;
      int target;
      if (kind == MG_POLICY_CAPTURE) {
        target = (p.blue_flag & 255) * S + (p.blue_flag >> 8);
      } else if (kind == MG_POLICY_FIGHT || (kind == MG_POLICY_PATROL_FIGHT && intruder)) {
        int best = 0x7fffffff;
        target = cell;
#pragma unroll
        for (int i = 0; i < NB; ++i) {   // the closest blue agent, first of equals, defeated ones included (heuristic.py:216-226)
          const int bx = w[i] & 255u, by = (w[i] >> 8) & 255u, dx = bx - x, dy = by - y, d2 = dx * dx + dy * dy;
          if (d2 < best) { best = d2; target = bx * S + by; }
        }
      } else if (__ldg(p.pol_border + cell)) {
        target = (int)__ldg(p.pol_along + r.below(p.pol_n_along));
      } else {
        target = __ldg(p.pol_goal + cell);
      }
      const int mv = __ldg(p.pol_first_move + (size_t)cell * p.cells + target);
      const bool follow = (unsigned long long)r.u32() < p.pol_thr[k];
      a = mv;
      if (!follow) a = r.below(5);
    }
    out |= (uint32_t)a << (8 * k);
  }
  if (NR == 2) { p.pol_out[e * 2] = (int8_t)(out & 255u); p.pol_out[e * 2 + 1] = (int8_t)(out >> 8); }
  else
#pragma unroll
    for (int k = 0; k < NR; ++k) p.pol_out[e * NR + k] = (int8_t)((out >> (8 * k)) & 255u);
  return out;
}

// The same step for the compile-time team sizes with every agent word in a REGISTER: the order-dependent loop picks and
// updates "agent order[k]" with select chains instead of dynamically indexed shared memory, the body is branch-free
// (one commit per agent), and occupancy / flag / battle tests run on registers.  Statement for statement the semantics of
// ctf_step_one above (which stays the general path for run-time team sizes); tested against it through the oracle.
template <int MODE, int NB, int NR, int POL = 0>
__device__ __forceinline__ void ctf_step_regs(const MapParams& p, long long e, uint32_t blue_raw, const uint8_t* terr,
                                              uint32_t* ag, int4& h, Rng<MODE>& r, double& rew, bool& term, bool& trunc,
                                              int& err) {
  constexpr int n = NB + NR;
  const int S = p.S;
  uint32_t w[n];
#pragma unroll
  for (int i = 0; i < n; ++i) w[i] = ag[i * kMapE];
  uint32_t red_raw = 0;   // POL: the scripted opponents decide on the pre-step state (the reference calls them first, ctf.py:1297-1301)
  if (POL) red_raw = policy_red_actions<NB, NR>(p, e, w, h);
  h.x += 1;  // ctf.py:1295
  uint32_t acts = 0, order;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int a = (int)(int8_t)(blue_raw >> (8 * i));   // int8 actions, fetched by the caller together with the state rows
    const bool bad = a < 0 || a > 4;   // reference: ValueError (ctf.py:1200-1201)
    if (bad) err |= MG_ERR_BAD_ACTION;
    acts |= (uint32_t)(bad ? 15 : a) << (4 * i);
  }
#pragma unroll
  for (int k = 0; k < NR; ++k) {  // RwPolicy.act for EVERY red agent, defeated or not (:1297-1301), or the external enemy policy
    const int a = POL ? (int)((red_raw >> (8 * k)) & 255u) : ((MODE == 0 || p.red_actions) ? p.red_actions[e * NR + k] : below16(r, 5));
    const bool bad = a < 0 || a > 4;
    if (bad) err |= MG_ERR_BAD_ACTION;
    acts |= (uint32_t)(bad ? 15 : a) << (4 * (NB + k));
  }
  if (p.variant_1v1) {  // Ctf1v1Env._move_agents: blue, then red (ctf.py:503-510)
    order = 0x10u;
  } else if (MODE == 0) {
    order = 0;
#pragma unroll
    for (int i = 0; i < n; ++i) order |= (uint32_t)(p.order[e * n + i] & 15) << (4 * i);
  } else {  // np_random.shuffle stand-in: Fisher-Yates
    order = 0x76543210u;
#pragma unroll
    for (int i = n - 1; i > 0; --i) {
      const int j = below16(r, i + 1);
      const uint32_t x = ((order >> (4 * i)) ^ (order >> (4 * j))) & 15u;
      order ^= (x << (4 * i)) | (x << (4 * j));
    }
  }
  const bool pen = p.obstacle_penalty != 0;
#pragma unroll
  for (int k = 0; k < n; ++k) {  // _move_agents :1240-1251
    const int i = (int)((order >> (4 * k)) & 15u);
    uint32_t wi = w[0];
#pragma unroll
    for (int j = 1; j < n; ++j) wi = (i == j) ? w[j] : wi;
    const int a = (int)((acts >> (4 * i)) & 15u);
    int dx, dy;
    action_delta(a & 7, dx, dy);
    const int nx = ag_x(wi) + dx, ny = ag_y(wi) + dy;  // _move_agent :1184-1238
    const bool live = !(wi & FL_DEAD) && a != 15;      // "Defeated agent doesn't move, sadly."
    const bool inb = (unsigned)nx < (unsigned)S && (unsigned)ny < (unsigned)S;
    const uint32_t target = ((uint32_t)nx & 255u) | (((uint32_t)ny & 255u) << 8);
    bool occupied = false;  // an agent object (alive, defeated, or itself when staying) sits on the cell
#pragma unroll
    for (int j = 0; j < n; ++j) occupied |= ((w[j] ^ target) & 0xFFFFu) == 0;
    const int tc = terr[inb ? ny * S + nx : 0];
    const bool go = live && inb && !occupied && !(tc == CT_OBSTACLE && !pen);  // Obstacle.can_overlap()
    const bool hit = live && inb && occupied && pen && !p.variant_1v1;        // :1231-1236 (1v1 has no collided logic, :498-501)
    const uint32_t moved = (bg_after_move(wi, tc) & 0xFF000000u) | target | ((uint32_t)dir_of_action(a, (int)((wi >> 16) & 255u)) << 16);  // Agent.move agent.py:167-200
    const uint32_t nw = go ? moved : (hit ? (wi | FL_COLLIDED) : wi);
#pragma unroll
    for (int j = 0; j < n; ++j) w[j] = (i == j) ? nw : w[j];
  }
  term = false; trunc = h.x >= p.max_steps;  // :1310-1311
  rew = 0.0;
  if (pen) {  // :1316-1332 (collided is never cleared)
#pragma unroll
    for (int i = 0; i < n; ++i)
      if (w[i] & FL_COLLIDED) { if (i < NB) rew -= p.obstacle_penalty; w[i] |= FL_DEAD; }
  }
  const uint32_t red_flag = (uint32_t)p.red_flag, blue_flag = (uint32_t)p.blue_flag;
#pragma unroll
  for (int i = 0; i < NB; ++i) if (((w[i] ^ red_flag) & 0xFFFFu) == 0) { rew += p.flag_reward; term = true; h.y |= 2; }   // :1335-1344
#pragma unroll
  for (int i = NB; i < n; ++i) if (((w[i] ^ blue_flag) & 0xFFFFu) == 0) { rew -= p.flag_reward; term = true; h.y |= 1; }  // :1347-1356
  int nbattle = 0;
  bool all_dead = true;
#pragma unroll
  for (int b = 0; b < NB; ++b) {  // np.where(distances <= battle_range): row-major, blue-major (:1368-1377)
#pragma unroll
    for (int q = 0; q < NR; ++q) {
      const uint32_t wb = w[b], wr = w[NB + q];
      const int ddx = ag_x(wb) - ag_x(wr), ddy = ag_y(wb) - ag_y(wr);
      if (ddx * ddx + ddy * ddy > p.d2_max) continue;  // == !(np.linalg.norm(int vector) <= battle_range), see MapParams::d2_max
      if ((wb | wr) & FL_DEAD) continue;               // :1380-1383
      const int cb = terr[ag_y(wb) * S + ag_x(wb)], cr = terr[ag_y(wr) * S + ag_x(wr)];
      const bool bh = (cb == CT_BLUE_TERR || cb == CT_BLUE_FLAG), rh = (cr == CT_RED_TERR || cr == CT_RED_FLAG);
      bool blue_win;
      if (MODE == 0) {
        blue_win = nbattle < p.KB ? p.blue_win[e * p.KB + nbattle] != 0 : false;
        if (nbattle >= p.KB) err |= MG_ERR_TRACE_OVERFLOW;
      } else {  // :1392-1407; (double)u / 2^32 < p  <=>  u < ceil(p * 2^32)
        const unsigned long long thr = (bh && !rh) ? p.thr_blue_home : ((!bh && rh) ? p.thr_red_home : p.thr_even);
        blue_win = (unsigned long long)r.u32() < thr;
      }
      ++nbattle;
      if (blue_win) { rew += p.battle_reward; w[NB + q] = wr | FL_DEAD; h.y |= 1 << (8 + NB + q); }   // :1409-1418
      else if (p.variant_1v1) { rew -= p.battle_reward; term = true; h.y |= 1 << 8; }  // 1v1: losing ends the episode (ctf.py:629-636)
      else { rew -= p.battle_reward; w[b] = wb | FL_DEAD; h.y |= 1 << (8 + b); }
    }
    all_dead &= (w[b] & FL_DEAD) != 0;
  }
  if (MODE == 0 && p.battles_used) p.battles_used[e] = nbattle;
  if (all_dead) term = true;           // :1423
  rew = __dsub_rn(rew, __dmul_rn(p.step_penalty, (double)NB));  // :1428 -- two roundings like the reference, never an FMA
#pragma unroll
  for (int i = 0; i < n; ++i) ag[i * kMapE] = w[i];
}

// value an agent shows in the "map" observation
template <int FAMILY>
__device__ __forceinline__ int agent_code(const MapParams& p, int i, uint32_t w) {
  if (FAMILY == MG_FAMILY_MAZE) return MZ_AGENT;                                           // maze.py:256-258
  return (w & FL_DEAD) ? CT_OBSTACLE : (i < p.nb ? CT_BLUE_AGENT : CT_RED_AGENT);          // ctf.py:1157-1161
}
template <int FAMILY>
__device__ __forceinline__ int obs_index(const MapParams& p, uint32_t w) {
  return FAMILY == MG_FAMILY_MAZE ? ag_x(w) * p.S + ag_y(w) : ag_y(w) * p.S + ag_x(w);  // Maze [x][y]; CtF returns encoded_map.T
}

template <typename T>
__device__ __forceinline__ void put(void* base, long long idx, int v) { static_cast<T*>(base)[idx] = (T)v; }
__device__ __forceinline__ void put_obs(const MapParams& p, void* base, long long idx, int v) {
  if (p.obs_dtype == MG_OBS_U8) put<uint8_t>(base, idx, v);
  else if (p.family == MG_FAMILY_MAZE) put<double>(base, idx, v);
  else put<long long>(base, idx, v);
}

// MINB: minimum resident CTAs per SM the register allocation must allow.  8 (64 registers, a few spills) wins when many
// waves of tiles keep every SM full (>= 256 K envs); 1 (no cap, no spills) has the shorter dependent chain and wins for
// launches of a wave or two, which are latency-bound.
// STEPV: which CtF step body the kernel carries - 0 the general one (run-time team sizes), 1 the 2v2 and 2 the 1v1 register
// bodies, 3 the general one with the occupancy bitmap (maps of <= 256 cells).  One body per kernel keeps the instruction stream of the hot path inside the instruction cache (with all of them
// inlined into one kernel `no_instruction` became the top stall reason).
// LEAN: the kernel of ONE hot configuration with everything else compiled out (the general kernel is ~6 500 SASS instructions of
// which a step of the common case executes ~1 000: `no_instruction` stalls).  1 = step (op 1) of a CtF handle with the staged u8
// tile image, obs given, no final_obs (the 2v2 register body, or the general body with the occupancy bitmap); 2 = step of a Maze handle in partial-view mode computed from the padded map (no memoised
// table), obs given, no final_obs.  The launcher checks those conditions; 0 = the general kernel.
// POL: (2v2 lean kernel only) the scripted opponents' decisions are part of the step (policy_red_actions above).
template <int FAMILY, int MODE, int MINB, int STEPV = 0, int LEAN = 0, int POL = 0>
__global__ void __launch_bounds__(kMapE, MINB) map_kernel(const __grid_constant__ MapParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, bar_tile;
  __shared__ uint8_t s_done[kMapE];
  const int tid = threadIdx.x, n = (LEAN == 1 && STEPV == 1) ? 4 : (LEAN == 2 ? 1 : p.n), cells = p.cells;
  const bool view_mode = LEAN == 2 || (LEAN == 0 && FAMILY == MG_FAMILY_MAZE && p.view_V != 0);   // obs = partial views (gen_obs) instead of the map
  const bool view_table = LEAN == 0 && view_mode && p.view_table != nullptr;        // memoised views: no map needed in shared memory
  const int head = view_table ? 0 : (view_mode ? p.map_padded_bytes : p.L);
  uint8_t* s_period = smem_raw;                                                     // [L], or the padded packed map in view mode
  uint32_t* s_ag = reinterpret_cast<uint32_t*>(smem_raw + head);                    // [n][kMapE] agent words, transposed
  uint32_t* s_occ = reinterpret_cast<uint32_t*>(smem_raw + head + (size_t)n * kMapE * 4);   // STEPV 3: occupancy bitmap + packed red pairs
  uint8_t* s_obs = smem_raw + head + (size_t)n * kMapE * 4 + (STEPV == 3 ? kOccBytes : 0);  // [kMapE][cells] (staged tiles) / the tile's views
  const long long e0 = (long long)blockIdx.x * kMapE;
  const int n_here = (int)min((long long)kMapE, p.N - e0);
  const long long e = e0 + tid;

  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar_tile, 1); fence_mbar_init(); }
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();
  // staged u8 tiles: the static part of the whole tile's observation slab (the map repeated once per env) arrives as ONE bulk
  // load of a host-built image while the envs step - no per-thread replication of the period, no barrier in front of the patches
  const bool tile_img = LEAN == 1 || (LEAN == 0 && !view_mode && p.obs_tile && p.obs);
  if (tid == 0 && head) {
    mbar_expect_tx(&bar, (uint32_t)head);
    tma_load_1d(s_period, view_mode ? p.map_padded : p.obs_period, (uint32_t)head, &bar);
    if (tile_img) {   // own barrier: the step only waits for the period
      mbar_expect_tx(&bar_tile, (uint32_t)(kMapE * cells));
      tma_load_1d(s_obs, p.obs_tile, (uint32_t)(kMapE * cells), &bar_tile);
    }
  }

  // ---- the env's agent row (padded to row_bytes = 4 * 2^k) and header; state planes are padded to whole tiles
  uint32_t* ag = s_ag + tid;
  const uint8_t* row = p.agents + e * p.row_bytes;
  // every global load that does not depend on another is issued here, ahead of the first shared-memory store of a loaded value:
  // issue is in order, so a store waiting for its row would hold the header / action loads back by one memory latency
  int maze_act = 0;
  if (FAMILY == MG_FAMILY_MAZE && (LEAN == 2 || p.op == 1) && tid < n_here) maze_act = p.actions[e];
  int4 h = p.hdr[e];
  uint32_t blue_raw = 0;   // 2v2 / 1v1 step: the blue actions travel with the state loads, ahead of the wait for the staged period
  if (FAMILY == MG_FAMILY_CTF && (LEAN || p.op == 1) && tid < n_here) {
    if (STEPV == 1)
      blue_raw = (reinterpret_cast<uintptr_t>(p.actions) & 1) ? ((uint32_t)(uint8_t)p.actions[e * 2] | ((uint32_t)(uint8_t)p.actions[e * 2 + 1] << 8))
                                                                : *reinterpret_cast<const uint16_t*>(p.actions + e * 2);
    else if (STEPV == 2) blue_raw = (uint8_t)p.actions[e];
  }
  if (LEAN == 1 && STEPV == 1) {   // 2v2: the row is one 16-byte word
    const uint4 v = *reinterpret_cast<const uint4*>(row);
    ag[0] = v.x; ag[kMapE] = v.y; ag[2 * kMapE] = v.z; ag[3 * kMapE] = v.w;
  } else if (LEAN == 2) {
    ag[0] = *reinterpret_cast<const uint32_t*>(row);
  } else if (p.row_bytes >= 16) {
    for (int c = 0; c < p.row_bytes / 16; ++c) {
      const uint4 v = *reinterpret_cast<const uint4*>(row + 16 * c);
      if (4 * c + 0 < n) ag[(4 * c + 0) * kMapE] = v.x;
      if (4 * c + 1 < n) ag[(4 * c + 1) * kMapE] = v.y;
      if (4 * c + 2 < n) ag[(4 * c + 2) * kMapE] = v.z;
      if (4 * c + 3 < n) ag[(4 * c + 3) * kMapE] = v.w;
    }
  } else if (p.row_bytes == 8) {
    const uint2 v = *reinterpret_cast<const uint2*>(row);
    ag[0] = v.x; if (n > 1) ag[kMapE] = v.y;
  } else {
    ag[0] = *reinterpret_cast<const uint32_t*>(row);
  }
  bool done = false, want_reset = false;
  int err = 0;
  Rng<MODE> r;
  r.open_trace(nullptr, 0);
  if (FAMILY == MG_FAMILY_CTF && (LEAN || p.op == 1)) mbar_wait(&bar, 0);  // the CtF step reads terrain codes from the staged period
  if (tid < n_here) {
    if (MODE == 1) r.open_philox(p.seed, p.env_id_base + (unsigned long long)e, (uint32_t)h.z);
    if (LEAN == 0 && p.op == 0) {
      want_reset = !p.reset_mask || p.reset_mask[e];
    } else {
      double rew; bool term, trunc;
      if (FAMILY == MG_FAMILY_MAZE) { uint32_t w = ag[0]; maze_step_one(p, maze_act, w, h, rew, term, trunc, err); ag[0] = w; }
      else if (STEPV == 1) ctf_step_regs<MODE, 2, 2, POL>(p, e, blue_raw, s_period, ag, h, r, rew, term, trunc, err);
      else if (STEPV == 2) ctf_step_regs<MODE, 1, 1>(p, e, blue_raw, s_period, ag, h, r, rew, term, trunc, err);
      else if (STEPV == 3 && n <= 8) ctf_step_occ<MODE, uint32_t>(p, e, p.actions + e * p.nb, s_period, ag, s_occ + tid, h, r, rew, term, trunc, err);
      else if (STEPV == 3) ctf_step_occ<MODE, unsigned long long>(p, e, p.actions + e * p.nb, s_period, ag, s_occ + tid, h, r, rew, term, trunc, err);
      else if (n <= 8) ctf_step_one<MODE, uint32_t, 0, 0>(p, e, p.actions + e * p.nb, s_period, ag, h, r, rew, term, trunc, err);
      else ctf_step_one<MODE, unsigned long long, 0, 0>(p, e, p.actions + e * p.nb, s_period, ag, h, r, rew, term, trunc, err);
      p.rewards[e] = rew; p.terminated[e] = term; p.truncated[e] = trunc;
      want_reset = done = p.autoreset && (term || trunc);  // same-step autoreset
    }
  }
  if (LEAN != 1 && head) mbar_wait(&bar, 0);   // (LEAN 1 has waited above)

  // ---- terminal observations of the finished envs are drawn before their reset (only when the caller asked for them)
  if (LEAN == 0 && p.final_obs) {
    s_done[tid] = done;
    if (__syncthreads_or(done)) {
      for (int j = 0; j < n_here; ++j) {
        if (!s_done[j]) continue;
        for (int i = tid; i < cells; i += kMapE) put_obs(p, p.final_obs, (e0 + j) * cells + i, s_period[i]);
      }
      __syncthreads();
      if (done)
        for (int i = 0; i < n; ++i) {
          const uint32_t w = ag[i * kMapE];
          put_obs(p, p.final_obs, e * cells + obs_index<FAMILY>(p, w), agent_code<FAMILY>(p, i, w));
        }
    }
  }
  // ---- reset(mask) / autoreset: ONE call site (a random CtF episode lasts ~20 steps, so most warps take it every step)
  if (want_reset) {
    reset_one<FAMILY, MODE>(p, e, ag, r);
    h.x = 0; h.y = 0; h.w += 1;  // step_count = 0 (multigrid.py:141); game_stats cleared (ctf.py:1068-1073); episode counter
  }

  // ---- state write-back (rows of padded envs of the last tile are written too: the planes are padded)
  if (MODE == 1) h.z = (int)r.ctr;
  p.hdr[e] = h;
  if (err) atomicOr(p.status, err);
  {
    uint8_t* wrow = p.agents + e * p.row_bytes;
    if (LEAN == 1 && STEPV == 1) {
      *reinterpret_cast<uint4*>(wrow) = make_uint4(ag[0], ag[kMapE], ag[2 * kMapE], ag[3 * kMapE]);
    } else if (LEAN == 2) {
      *reinterpret_cast<uint32_t*>(wrow) = ag[0];
    } else if (p.row_bytes >= 16) {
      for (int c = 0; c < p.row_bytes / 16; ++c) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (4 * c + 0 < n) v.x = ag[(4 * c + 0) * kMapE];
        if (4 * c + 1 < n) v.y = ag[(4 * c + 1) * kMapE];
        if (4 * c + 2 < n) v.z = ag[(4 * c + 2) * kMapE];
        if (4 * c + 3 < n) v.w = ag[(4 * c + 3) * kMapE];
        *reinterpret_cast<uint4*>(wrow + 16 * c) = v;
      }
    } else if (p.row_bytes == 8) {
      *reinterpret_cast<uint2*>(wrow) = make_uint2(ag[0], n > 1 ? ag[kMapE] : 0u);
    } else {
      *reinterpret_cast<uint32_t*>(wrow) = ag[0];
    }
  }

  if (LEAN == 0 && !p.obs) return;
  // ---- Maze partial-observation mode: this env's egocentric view (fused MazeSingleAgentEnv.step + MultiGridEnv.gen_obs)
  if (view_mode) {
    const int V = p.view_V, VV3 = V * V * 3;
    if (tid < n_here) {
      const uint32_t w = ag[0];
      const int x = ag_x(w), y = ag_y(w), dir = (int)((w >> 16) & 3u);
      const uint32_t agent_cell = (uint32_t)p.view_agent | ((uint32_t)dir << 6);
      int x0, y0, sa, sb;
      if (view_table) {
        const uint4* row = p.view_table + (size_t)((((x * p.S + y) << 2) | dir)) * p.view_row16;
        if (V == 7) view_copy_store<7>(row, s_obs, tid);
        else if (V == 5) view_copy_store<5>(row, s_obs, tid);
        else view_copy_store<3>(row, s_obs, tid);
      } else {
        // the view cells inside the map, over a and over b (everything else shows the filler and blocks sight)
        auto masks = [&](int VV_, uint32_t& mA, uint32_t& mB) {
          const bool a_is_x = dir & 1;
          mA = a_is_x ? range_mask(x0, dir == 3 ? 1 : -1, p.S, VV_) : range_mask(y0, dir == 0 ? 1 : -1, p.S, VV_);
          mB = a_is_x ? range_mask(y0, dir == 3 ? 1 : -1, p.S, VV_) : range_mask(x0, dir == 2 ? 1 : -1, p.S, VV_);
        };
        uint32_t mA, mB;
        if (V == 7) {
          view_geometry<7>(x, y, dir, p.pitch, x0, y0, sa, sb);
          masks(7, mA, mB);
          view_compute_store<false, 7>(s_period + (x0 + p.pad) * p.pitch + (y0 + p.pad), sa, sb, mA, mB, p.view_oob, agent_cell, p.view_see_through != 0, s_obs, tid);
        } else if (V == 5) {
          view_geometry<5>(x, y, dir, p.pitch, x0, y0, sa, sb);
          masks(5, mA, mB);
          view_compute_store<false, 5>(s_period + (x0 + p.pad) * p.pitch + (y0 + p.pad), sa, sb, mA, mB, p.view_oob, agent_cell, p.view_see_through != 0, s_obs, tid);
        } else {
          view_geometry<3>(x, y, dir, p.pitch, x0, y0, sa, sb);
          masks(3, mA, mB);
          view_compute_store<false, 3>(s_period + (x0 + p.pad) * p.pitch + (y0 + p.pad), sa, sb, mA, mB, p.view_oob, agent_cell, p.view_see_through != 0, s_obs, tid);
        }
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    const uint32_t bytes = (uint32_t)n_here * VV3, bulk = bytes & ~15u;
    uint8_t* g = static_cast<uint8_t*>(p.obs) + (size_t)e0 * VV3;   // 128 * 3 * V * V is a multiple of 16
    if (tid == 0 && bulk) { tma_store_1d(g, s_obs, bulk); tma_commit(); }
    for (uint32_t k = bulk + tid; k < bytes; k += kMapE) g[k] = s_obs[k];
    if (tid == 0) tma_wait_read_all();
    return;
  }
  // ---- observation: static map for the whole tile, then the agents on top
  const long long slab = (long long)n_here * cells;  // elements in this tile's obs slab
  if (LEAN == 1 || p.obs_staged) {  // u8, small map: assemble the tile in shared memory, one TMA bulk store
    if (LEAN == 0 && !tile_img) {   // handle without the image table: replicate the period by hand
      const int L16 = p.L16, chunks = kMapE * cells / 16;
      uint4* dst = reinterpret_cast<uint4*>(s_obs);
      const uint4* src = reinterpret_cast<const uint4*>(s_period);
      int m = L16 == 1 ? 0 : tid - (int)__umulhi((uint32_t)tid, p.L16_magic) * L16;   // tid mod L16
      const int step = p.tile_mod_L16;
      for (int c = tid; c < chunks; c += kMapE) {
        dst[c] = src[m];
        m += step; if (m >= L16) m -= L16;
      }
      __syncthreads();
    } else {
      mbar_wait(&bar_tile, 0);
    }
    if (tid < n_here)
      for (int i = 0; i < n; ++i) {  // agents in index order: later agents overwrite earlier ones (ctf.py:1157-1161)
        const uint32_t w = ag[i * kMapE];
        s_obs[tid * cells + obs_index<FAMILY>(p, w)] = (uint8_t)agent_code<FAMILY>(p, i, w);
      }
    fence_proxy_async_smem();
    __syncthreads();
    uint8_t* g = static_cast<uint8_t*>(p.obs) + e0 * cells;
    const uint32_t bulk = (uint32_t)slab & ~15u;
    if (tid == 0 && bulk) { tma_store_1d(g, s_obs, bulk); tma_commit(); }
    for (uint32_t k = bulk + tid; k < (uint32_t)slab; k += kMapE) g[k] = s_obs[k];
    if (tid == 0) tma_wait_read_all();
    return;
  }
  if (LEAN != 0) return;   // (not reached: the lean kernels have returned above)
  if (p.obs_tma) {
    // large tiles: the static map leaves straight from the staged period, one TMA bulk store per period (full-line
    // writes, no per-thread store instructions); the stores are spread over the CTA's threads, each waits for its own
    // to be performed, and after the barrier every env's thread patches its agents' cells on top
    const uint32_t elem = p.obs_dtype == MG_OBS_U8 ? 1u : 8u, reps = (uint32_t)p.tma_reps, Lb = (uint32_t)p.L * elem * reps;
    uint8_t* src = s_period;
    if (elem == 8 || reps > 1) {  // `reps` copies of the period in the observation dtype (float64 Maze / int64 CtF for the reference's dtypes):
      src = s_obs;                // bulk stores of ~32 KB run closer to the write bandwidth than 4 KB ones
      if (elem == 1) {
        const int total16 = p.L16 * (int)reps;
        int m = p.L16 == 1 ? 0 : tid - (int)__umulhi((uint32_t)tid, p.L16_magic) * p.L16;   // tid mod L16
        for (int i = tid; i < total16; i += kMapE) {
          reinterpret_cast<uint4*>(src)[i] = reinterpret_cast<const uint4*>(s_period)[m];
          m += p.tile_mod_L16; if (m >= p.L16) m -= p.L16;
        }
      } else {
        const int total = p.L * (int)reps;
        int m = tid;                // tid < L is not guaranteed for tiny maps
        while (m >= p.L) m -= p.L;
        const int step = kMapE % p.L;
        for (int i = tid; i < total; i += kMapE) {
          const uint8_t v = s_period[m];
          if (FAMILY == MG_FAMILY_MAZE) reinterpret_cast<double*>(src)[i] = (double)v;
          else reinterpret_cast<long long*>(src)[i] = (long long)v;
          m += step; if (m >= p.L) m -= p.L;
        }
      }
      fence_proxy_async_smem();
      __syncthreads();
    }
    uint8_t* g = static_cast<uint8_t*>(p.obs) + (size_t)e0 * cells * elem;
    const size_t slab_bytes = (size_t)slab * elem;
    const int nch = (int)(slab_bytes / Lb);
    for (int c = tid; c < nch; c += kMapE) tma_store_1d(g + (size_t)c * Lb, src, Lb);
    if (tid == 0 && (slab_bytes - (size_t)nch * Lb) >= (size_t)p.L * elem)   // whole periods left over by the big chunks
      tma_store_1d(g + (size_t)nch * Lb, src, (uint32_t)((slab_bytes - (size_t)nch * Lb) / ((size_t)p.L * elem) * ((size_t)p.L * elem)));
    tma_commit();
    for (size_t k = (size_t)slab / p.L * p.L + tid; k < (size_t)slab; k += kMapE) put_obs(p, p.obs, e0 * cells + (long long)k, s_period[k % p.L]);  // ragged last tile
    tma_wait_all();
    __syncthreads();
  } else if (p.obs_dtype == MG_OBS_U8) {
    const int L16 = p.L16;
    const long long chunks = slab / 16;
    uint4* dst = reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.obs) + e0 * cells);  // e0*cells is a multiple of L
    const uint4* src = reinterpret_cast<const uint4*>(s_period);
    int m = L16 == 1 ? 0 : tid - (int)__umulhi((uint32_t)tid, p.L16_magic) * L16;   // tid mod L16
    const int step = p.tile_mod_L16;
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
  if (!p.obs_tma) __syncthreads();  // the tile's static fill is ordered before the per-env patches
  if (tid < n_here)
    for (int i = 0; i < n; ++i) {
      const uint32_t w = ag[i * kMapE];
      put_obs(p, p.obs, e * cells + obs_index<FAMILY>(p, w), agent_code<FAMILY>(p, i, w));
    }
}

// _get_info of the map families (maze.py:262-269; ctf.py:1165-1182, 434-452): distances from agents[0] / agents[1] to the
// flags and to the nearest cell of the territory / obstacle lists.  distance_points is np.linalg.norm of an integer
// vector = sqrt((double)d2) (correctly rounded on both sides); distance_area_point is the minimum of such norms =
// sqrt(min d2), looked up in host-built per-cell tables.  out: float64 [N][2] (Maze) or [N][11] (CtF, dict key order).
template <int FAMILY>
__global__ void __launch_bounds__(256) map_info_kernel(const __grid_constant__ MapParams p, double* __restrict__ out) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= p.N) return;
  const int S = p.S, cells = p.cells;
  const uint32_t* row = reinterpret_cast<const uint32_t*>(p.agents + e * p.row_bytes);
  auto area = [&](int table, uint32_t w) {
    const int d2 = __ldg(p.d2_tables + table * cells + ag_x(w) * S + ag_y(w));
    return d2 < 0 ? __longlong_as_double(0x7FF0000000000000ll) : sqrt((double)d2);   // empty list: inf (the reference raises)
  };
  auto points = [](uint32_t a, uint32_t b) {
    const int dx = ag_x(a) - ag_x(b), dy = ag_y(a) - ag_y(b);
    return sqrt((double)(dx * dx + dy * dy));
  };
  const uint32_t a0 = row[0];
  if (FAMILY == MG_FAMILY_MAZE) {
    out[e * 2] = area(0, a0); out[e * 2 + 1] = area(1, a0);
    return;
  }
  const uint32_t a1 = row[1], bf = (uint32_t)p.blue_flag, rf = (uint32_t)p.red_flag;   // agents[1]: the second agent of the list
  double* o = out + e * 11;
  o[0] = points(a0, a1); o[1] = points(a0, bf); o[2] = points(a0, rf); o[3] = points(a1, bf); o[4] = points(a1, rf);
  o[5] = points(bf, rf);
  o[6] = area(0, a0); o[7] = area(1, a0); o[8] = area(0, a1); o[9] = area(1, a1); o[10] = area(2, a0);
}

cudaError_t launch_map_info(const MapParams& p, double* out, cudaStream_t st) {
  const unsigned blocks = (unsigned)((p.N + 255) / 256);
  if (p.family == MG_FAMILY_MAZE) map_info_kernel<MG_FAMILY_MAZE><<<blocks, 256, 0, st>>>(p, out);
  else map_info_kernel<MG_FAMILY_CTF><<<blocks, 256, 0, st>>>(p, out);
  return cudaGetLastError();
}

// CtFMvNEnv._get_obs, observation_option="flattened" (ctf.py:1084-1104): int64 [N][L] = the host-built template (flags,
// territory and obstacle cell lists in the reference's order; zeros where per-env values go) with each env's agent positions
// (first 2 n entries) and terminated flags (last n) filled in.  A CTA assembles kFlatE env rows in shared memory and writes them
// with one TMA bulk store: the output is ~1.7 KB per env of which 3 n values differ, so the kernel is a pure HBM write stream.
constexpr int kFlatThreads = 128;
template <typename T>   // long long = the reference's dtype; uint8_t = the compact form (every entry is a coordinate < 256 or a flag)
__global__ void __launch_bounds__(kFlatThreads) ctf_flat_kernel(const __grid_constant__ MapParams p, const long long* __restrict__ tmpl, int L,
                                                                int E, T* __restrict__ out, int bulk_ok) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  T* s = reinterpret_cast<T*>(smem_raw);   // [E][L]
  const int tid = threadIdx.x, n = p.n;
  const long long e0 = (long long)blockIdx.x * E;
  const int n_here = (int)min((long long)E, p.N - e0);
  for (int c = tid; c < L; c += kFlatThreads) {
    const T v = (T)__ldg(tmpl + c);
#pragma unroll 4
    for (int r = 0; r < E; ++r) s[r * L + c] = v;
  }
  __syncthreads();
  for (int t = tid; t < n_here * n; t += kFlatThreads) {
    const int r = t / n, i = t - r * n;
    const uint32_t w = *reinterpret_cast<const uint32_t*>(p.agents + (e0 + r) * p.row_bytes + 4 * i);
    s[r * L + 2 * i] = (T)ag_x(w); s[r * L + 2 * i + 1] = (T)ag_y(w);
    if (!p.variant_1v1) s[r * L + L - n + i] = (w & FL_DEAD) ? 1 : 0;
    else if (i == 1) s[r * L + L - 1] = (w & FL_DEAD) ? 1 : 0;   // Ctf1v1Env: the tail is int(_is_red_agent_defeated) alone (ctf.py:359-371)
  }
  fence_proxy_async_smem();
  __syncthreads();
  T* g = out + e0 * L;
  const uint32_t bytes = (uint32_t)n_here * (uint32_t)L * (uint32_t)sizeof(T);
  if (bulk_ok && bytes % 16 == 0) {
    if (tid == 0) { tma_store_1d(g, s, bytes); tma_commit(); tma_wait_read_all(); }
  } else {
    for (int k = tid; k < n_here * L; k += kFlatThreads) g[k] = s[k];
  }
}

// envs per CTA: a multiple of 16 / sizeof(T) rows (tile bases stay 16-byte aligned) that fits in ~96 KB of shared memory and moves
// ~28 KB per bulk store; 0 = the map's lists are too long
int ctf_flat_tile_envs(int L, int elem) {
  const int unit = elem == 8 ? 2 : 16;
  int E = (int)((96 * 1024) / ((size_t)L * elem)) / unit * unit;
  const int cap = elem == 8 ? 16 : 128;
  return E > cap ? cap : E;
}

cudaError_t launch_ctf_flat(const MapParams& p, const long long* tmpl, int L, void* out, int elem, cudaStream_t st) {
  const int E = ctf_flat_tile_envs(L, elem);
  if (E < 2) return cudaErrorInvalidValue;
  const size_t smem = (size_t)E * L * elem;
  const unsigned blocks = (unsigned)((p.N + E - 1) / E);
  const int bulk_ok = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  cudaError_t e;
  if (elem == 8) {
    if ((e = raise_smem_limit((const void*)ctf_flat_kernel<long long>, smem)) != cudaSuccess) return e;
    ctf_flat_kernel<long long><<<blocks, kFlatThreads, smem, st>>>(p, tmpl, L, E, static_cast<long long*>(out), bulk_ok);
  } else {
    if ((e = raise_smem_limit((const void*)ctf_flat_kernel<uint8_t>, smem)) != cudaSuccess) return e;
    ctf_flat_kernel<uint8_t><<<blocks, kFlatThreads, smem, st>>>(p, tmpl, L, E, static_cast<uint8_t*>(out), bulk_ok);
  }
  return cudaGetLastError();
}

static bool map_pdl_enabled() {
  static const bool on = [] { const char* v = std::getenv("MG_PDL"); return !(v && v[0] == '0'); }();
  return on;
}

// the tile's u8 observation slab is staged in shared memory when it (plus period and agent words) leaves room for
// at least four CTAs per SM
bool map_obs_staged(int cells, int obs_dtype) {
  static const bool off = [] { const char* v = std::getenv("MG_MAP_DIRECT"); return v && v[0] == '1'; }();
  return !off && obs_dtype == MG_OBS_U8 && (size_t)kMapE * cells <= 48 * 1024 && (kMapE * cells) % 16 == 0;
}
// tiles that are not staged stream the period with TMA bulk stores; the 8-byte dtypes need an 8-byte copy of it
bool map_obs_tma(int L, int cells, int obs_dtype) {
  static const bool off = [] { const char* v = std::getenv("MG_MAP_NO_TMA"); return v && v[0] == '1'; }();
  if (off || map_obs_staged(cells, obs_dtype)) return false;
  return obs_dtype == MG_OBS_U8 || (size_t)L * 8 <= 64 * 1024;
}
// copies of the period per bulk store: ~32 KB chunks, but never more than 1/16 of a tile (the copies must amortise)
int map_tma_reps(int L, int cells, int obs_dtype) {
  const size_t Lb = (size_t)L * (obs_dtype == MG_OBS_U8 ? 1 : 8), tile = (size_t)kMapE * cells * (obs_dtype == MG_OBS_U8 ? 1 : 8);
  size_t r = 32768 / Lb;
  if (r > tile / (16 * Lb)) r = tile / (16 * Lb);
  return r < 1 ? 1 : (int)r;
}
// padded_bytes = 0 with the memoised view table (the map is then not staged)
size_t map_view_smem_bytes(int padded_bytes, int V) { return (size_t)padded_bytes + (size_t)4 * kMapE + (size_t)kMapE * V * V * 3 + 16; }
cudaError_t configure_map_view_mode(size_t smem) {
  cudaError_t e;
  if ((e = raise_smem_limit((const void*)map_kernel<MG_FAMILY_MAZE, 0, 1>, (size_t)smem)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)map_kernel<MG_FAMILY_MAZE, 0, 8>, (size_t)smem)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)map_kernel<MG_FAMILY_MAZE, 1, 1>, (size_t)smem)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)map_kernel<MG_FAMILY_MAZE, 1, 1, 0, 2>, (size_t)smem)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)map_kernel<MG_FAMILY_MAZE, 1, 8, 0, 2>, (size_t)smem)) != cudaSuccess) return e;
  return raise_smem_limit((const void*)map_kernel<MG_FAMILY_MAZE, 1, 8>, (size_t)smem);
}
// the general CtF body of small maps keeps an occupancy bitmap per env (STEPV 3)
bool map_ctf_occ(int family, int nb, int nr, int cells) {
  static const bool off = [] { const char* v = std::getenv("MG_CTF_NO_OCC"); return v && v[0] == '1'; }();
  return !off && family == MG_FAMILY_CTF && cells <= 32 * kOccWords && !(nb == 2 && nr == 2) && !(nb == 1 && nr == 1);
}
size_t map_smem_bytes(int L, int n, int cells, int obs_dtype, bool occ) {
  size_t extra = occ ? (size_t)kOccBytes : 0;
  if (map_obs_staged(cells, obs_dtype)) extra += (size_t)kMapE * cells;
  else if (map_obs_tma(L, cells, obs_dtype)) {
    const int reps = map_tma_reps(L, cells, obs_dtype);
    if (obs_dtype != MG_OBS_U8 || reps > 1) extra += (size_t)L * (obs_dtype == MG_OBS_U8 ? 1 : 8) * reps;
  }
  return (size_t)L + (size_t)4 * kMapE * n + extra + 16;
}
int map_tile_envs() { return kMapE; }

template <int FAMILY, int MODE, int MINB, int STEPV, int LEAN = 0, int POL = 0>
static cudaError_t launch_one(const MapParams& p, cudaStream_t st) {
  const size_t smem = (p.family == MG_FAMILY_MAZE && p.view_V) ? map_view_smem_bytes(p.view_table ? 0 : p.map_padded_bytes, p.view_V)
                                                               : map_smem_bytes(p.L, p.n, p.cells, p.obs_dtype, STEPV == 3);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((p.N + kMapE - 1) / kMapE)); cfg.blockDim = dim3(kMapE);
  cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = map_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, map_kernel<FAMILY, MODE, MINB, STEPV, LEAN, POL>, p);
}

static bool map_lean_enabled() {
  static const bool on = [] { const char* v = std::getenv("MG_MAP_LEAN"); return !(v && v[0] == '0'); }();
  return on;
}

template <int FAMILY, int MODE, int STEPV>
static cudaError_t configure_pair(int smem) {
  cudaError_t e = raise_smem_limit((const void*)map_kernel<FAMILY, MODE, 1, STEPV>, (size_t)smem);
  if (e != cudaSuccess) return e;
  return raise_smem_limit((const void*)map_kernel<FAMILY, MODE, 8, STEPV>, (size_t)smem);
}

cudaError_t configure_map_kernels(int L, int n, int cells, int obs_dtype, bool occ) {
  const int smem = (int)map_smem_bytes(L, n, cells, obs_dtype, occ);
  cudaError_t e;
  if ((e = configure_pair<MG_FAMILY_CTF, 0, 3>(smem)) != cudaSuccess) return e;
  if ((e = configure_pair<MG_FAMILY_CTF, 1, 3>(smem)) != cudaSuccess) return e;
  if ((e = configure_pair<MG_FAMILY_MAZE, 0, 0>(smem)) != cudaSuccess) return e;
  if ((e = configure_pair<MG_FAMILY_MAZE, 1, 0>(smem)) != cudaSuccess) return e;
  if ((e = configure_pair<MG_FAMILY_CTF, 0, 0>(smem)) != cudaSuccess) return e;
  if ((e = configure_pair<MG_FAMILY_CTF, 0, 1>(smem)) != cudaSuccess) return e;
  if ((e = configure_pair<MG_FAMILY_CTF, 0, 2>(smem)) != cudaSuccess) return e;
  if ((e = configure_pair<MG_FAMILY_CTF, 1, 0>(smem)) != cudaSuccess) return e;
  if ((e = configure_pair<MG_FAMILY_CTF, 1, 1>(smem)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)map_kernel<MG_FAMILY_CTF, 1, 1, 1, 1>, (size_t)smem)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)map_kernel<MG_FAMILY_CTF, 1, 8, 1, 1>, (size_t)smem)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)map_kernel<MG_FAMILY_CTF, 1, 10, 1, 1>, (size_t)smem)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)map_kernel<MG_FAMILY_CTF, 1, 12, 1, 1>, (size_t)smem)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)map_kernel<MG_FAMILY_CTF, 1, 1, 1, 1, 1>, (size_t)smem)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)map_kernel<MG_FAMILY_CTF, 1, 8, 1, 1, 1>, (size_t)smem)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)map_kernel<MG_FAMILY_CTF, 1, 1, 3, 1>, (size_t)smem)) != cudaSuccess) return e;
  if ((e = raise_smem_limit((const void*)map_kernel<MG_FAMILY_CTF, 1, 8, 3, 1>, (size_t)smem)) != cudaSuccess) return e;
  return configure_pair<MG_FAMILY_CTF, 1, 2>(smem);
}

// env count from which the 64-register allocation (8 CTAs per SM) is used; MG_MAP_MINB8_FROM overrides it for experiments
static long long minb8_from() {
  static const long long v = [] { const char* e = std::getenv("MG_MAP_MINB8_FROM"); return e ? std::atoll(e) : 262144ll; }();
  return v;
}

template <int FAMILY, int MODE, int STEPV>
static cudaError_t launch_by_size(const MapParams& p, cudaStream_t st) {
  // Maze partial-view mode is issue-bound: as soon as the uncapped allocation (5 CTAs per SM) would need a second wave
  // (148 * 5 * 128 envs) the 8-CTA one wins (131 072 envs: 11.3 -> 10.1 us); the other modes switch at 256 K envs
  const long long from = (FAMILY == MG_FAMILY_MAZE && p.view_V && !std::getenv("MG_MAP_MINB8_FROM")) ? 148ll * 5 * kMapE + 1 : minb8_from();
  if constexpr (MODE == 1 && ((FAMILY == MG_FAMILY_CTF && (STEPV == 1 || STEPV == 3)) || FAMILY == MG_FAMILY_MAZE)) {   // the hot configurations have kernels of their own
    constexpr int LEAN = FAMILY == MG_FAMILY_CTF ? 1 : 2;
    const bool fits = FAMILY == MG_FAMILY_CTF ? (p.obs_staged && p.obs_tile && (STEPV == 3 || p.row_bytes == 16)) : (p.view_V && !p.view_table && p.row_bytes == 4);
    if (map_lean_enabled() && p.op == 1 && p.obs && !p.final_obs && fits) {
      if constexpr (FAMILY == MG_FAMILY_CTF && STEPV == 1) {
        if (p.pol_on) return p.N >= from ? launch_one<FAMILY, MODE, 8, STEPV, LEAN, 1>(p, st) : launch_one<FAMILY, MODE, 1, STEPV, LEAN, 1>(p, st);
      }
      if constexpr (FAMILY == MG_FAMILY_CTF && STEPV == 1) {   // experiment switch: more resident CTAs at fewer registers
        static const int minb = [] { const char* v = std::getenv("MG_MAP_MINB"); return v ? std::atoi(v) : 0; }();
        if (minb == 10 && p.N >= from) return launch_one<FAMILY, MODE, 10, STEPV, LEAN>(p, st);
        if (minb == 12 && p.N >= from) return launch_one<FAMILY, MODE, 12, STEPV, LEAN>(p, st);
      }
      return p.N >= from ? launch_one<FAMILY, MODE, 8, STEPV, LEAN>(p, st) : launch_one<FAMILY, MODE, 1, STEPV, LEAN>(p, st);
    }
  }
  if (p.pol_on) return cudaErrorInvalidValue;   // map_can_fuse_policy said no: the caller launches ctf_policy_kernel itself
  return p.N >= from ? launch_one<FAMILY, MODE, 8, STEPV>(p, st) : launch_one<FAMILY, MODE, 1, STEPV>(p, st);
}

template <int MODE>
static cudaError_t launch_ctf(const MapParams& p, cudaStream_t st) {
  if (p.nb == 2 && p.nr == 2) return launch_by_size<MG_FAMILY_CTF, MODE, 1>(p, st);
  if (p.nb == 1 && p.nr == 1) return launch_by_size<MG_FAMILY_CTF, MODE, 2>(p, st);
  if (map_ctf_occ(p.family, p.nb, p.nr, p.cells)) return launch_by_size<MG_FAMILY_CTF, MODE, 3>(p, st);
  return launch_by_size<MG_FAMILY_CTF, MODE, 0>(p, st);
}

// mg_set_red_policy_fusion: can this launch carry the opponents' decisions (the 2v2 lean step kernel), or does the caller run
// ctf_policy_kernel ahead of it?
bool map_can_fuse_policy(const MapParams& p) {
  return map_lean_enabled() && p.family == MG_FAMILY_CTF && p.rng_mode == 1 && p.nb == 2 && p.nr == 2 && !p.variant_1v1 && p.op == 1 &&
         p.obs && !p.final_obs && p.obs_staged && p.obs_tile && p.row_bytes == 16;
}

cudaError_t launch_map(const MapParams& p, cudaStream_t st) {
  if (p.family == MG_FAMILY_MAZE) return p.rng_mode == 0 ? launch_by_size<MG_FAMILY_MAZE, 0, 0>(p, st) : launch_by_size<MG_FAMILY_MAZE, 1, 0>(p, st);
  return p.rng_mode == 0 ? launch_ctf<0>(p, st) : launch_ctf<1>(p, st);
}

}  // namespace mg
