// collect_device.cuh -- per-env device functions of the Collect family shared by the tile kernels (collect_kernels.cu) and the
// warp-tile step / rollout kernel (collect_rollout_kernels.cu): placement by rejection sampling, reset + the _gen_grid layouts,
// and CollectGameEnv.step for ONE env on its shared-memory grid.  Citations are file:line under the reference checkout.
#pragma once
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
    int x, y;
    r.rand_pair(tx, hx, ty, hy, x, y);
    ox = x; oy = y;
    if (MODE == 0 && (r.err & MG_ERR_TRACE_OVERFLOW)) return;  // trace exhausted: leave the grid untouched
    if (GCELL(g, p.H, x, y) != 0) continue;
    GCELL(g, p.H, x, y) = code;
    return;
  }
}

// CollectGameEnv._respawn (collect_game.py:129-130) / CollectGameQuadrantsRespawn._respawn (:401-409)
template <int MODE>
__device__ __forceinline__ int respawn(const CollectParams& p, uint8_t* g, Rng<MODE>& r, int colour) {
  int x, y;
  if (p.layout == MG_LAYOUT_QUADRANTS_RESPAWN) {
    const int q = colour < 3 ? colour : 0;
    const int tx = q == 0 ? 0 : p.W / 2 - 1, ty = q == 1 ? p.H / 2 - 1 : 0;
    place_obj<MODE>(p, g, r, cell(T_BALL, colour, p.mark_respawned), tx, ty, p.W / 2 + 1, p.H / 2 + 1, x, y);
  } else {
    place_obj<MODE>(p, g, r, cell(T_BALL, colour, p.mark_respawned), 0, 0, p.W, p.H, x, y);
  }
  return x * p.H + y;  // the cell that received the ball
}

// One object of a _gen_grid placement sequence: its cell code and the inclusive box it is rejection-sampled in
// (place_obj, multigrid.py:282-339: [top, min(top + size, dim - 1)]).
struct Placement { uint8_t code; int tx, ty, hx, hy; };
__device__ __forceinline__ Placement boxed(const CollectParams& p, uint8_t code, int tx, int ty, int sx, int sy) {
  Placement q;
  q.code = code; q.tx = tx; q.ty = ty; q.hx = min(tx + sx, p.W - 1); q.hy = min(ty + sy, p.H - 1);
  return q;
}

// Places objects 0 .. count-1 of a sequence, `spec(k)` describing the k-th one, in ONE loop over candidate draws: a lane
// whose candidate was accepted moves on to its next object instead of idling until the slowest lane of the warp has placed
// the current one (nested per-object rejection loops cost the sum of per-object maxima over the warp; this costs the
// maximum of per-lane totals).  The env's draw sequence is the same as with nested loops, so results are unchanged.
template <int MODE, typename Spec, typename Placed>
__device__ __forceinline__ void place_sequence(const CollectParams& p, uint8_t* g, Rng<MODE>& r, int count, Spec&& spec, Placed&& placed) {
  if (count <= 0) return;
  int k = 0;
  Placement q = spec(0);
  while (k < count) {
    int x, y;
    r.rand_pair(q.tx, q.hx, q.ty, q.hy, x, y);
    if (MODE == 0 && (r.err & MG_ERR_TRACE_OVERFLOW)) return;  // trace exhausted: leave the rest unplaced
    if (GCELL(g, p.H, x, y) != 0) continue;
    GCELL(g, p.H, x, y) = q.code;
    placed(k, x, y);
    if (++k < count) q = spec(k);
  }
}

// CollectGameEnv.reset (collect_game.py:107-119) + the layout's _gen_grid.  `g`, `pos` live in smem.
template <int MODE>
__device__ __forceinline__ void reset_env(const CollectParams& p, uint8_t* g, uint8_t* pos, Rng<MODE>& r) {
  const int W = p.W, H = p.H, A = p.A, nb = p.nb;
  // Grid(width, height) + border walls (+ the Rooms inner walls): copied from the handle's template
  // (grid.py:66-89; collect_game.py:239-243, 269-273, 309-320, 379-382)
  if ((p.cells & 3) == 0) {
    const uint32_t* t32 = reinterpret_cast<const uint32_t*>(p.wall_template);
    uint32_t* g32 = reinterpret_cast<uint32_t*>(g);
    for (int i = 0; i < p.cells / 4; ++i) g32[i] = __ldg(t32 + i);
  } else {
    for (int i = 0; i < p.cells; ++i) g[i] = __ldg(p.wall_template + i);
  }
  auto nothing = [](int, int, int) {};
  if (p.layout == MG_LAYOUT_EVEN_DIST) {  // collect_game.py:236-259: balls anywhere, then place_agent(a) anywhere empty (multigrid.py:364-369)
    const int per = p.num_balls / nb, balls = per * nb;
    place_sequence<MODE>(p, g, r, balls + A,
                         [&](int k) { return boxed(p, k < balls ? cell(T_BALL, p.ball_colour[k / per], 0) : p.agent_code[k - balls], 0, 0, W, H); },
                         [&](int k, int x, int y) { if (k >= balls) { pos[2 * (k - balls)] = (uint8_t)x; pos[2 * (k - balls) + 1] = (uint8_t)y; } });
  } else if (p.layout == MG_LAYOUT_QUADRANTS) {  // collect_game.py:266-300
    const int per = p.num_balls / nb;
    place_sequence<MODE>(p, g, r, per * nb, [&](int k) {
      const int t = k / per;
      return boxed(p, cell(T_BALL, p.ball_colour[t], 0), (t == 1 || t == 2) ? W / 2 - 1 : 0, t == 1 ? H / 2 - 1 : (t == 3 ? H / 2 : 0), W / 2 - 1, H / 2 - 1);
    }, nothing);
    for (int i = 0; i < A; ++i) {  // place_agent(a, pos): overwrites (put_obj multigrid.py:341-348)
      GCELL(g, H, 1 + i, H - 2) = p.agent_code[i];
      pos[2 * i] = (uint8_t)(1 + i); pos[2 * i + 1] = (uint8_t)(H - 2);
    }
  } else if (p.layout == MG_LAYOUT_ROOMS) {  // collect_game.py:306-362 (`width` on both axes)
    const int m = W / 2;
    for (int i = 0; i < A; ++i) {  // _rand_elem(possible_coords) -> _rand_int(0, 4)
      const int k = r.rand_int(0, 4);
      const int cx = k == 0 ? m : (k <= 2 ? m - 1 : m + 1);
      const int cy = k == 0 ? m : ((k == 1 || k == 4) ? m - 1 : m + 1);
      GCELL(g, H, cx, cy) = p.agent_code[i];  // a second agent on the same cell overwrites the first
      pos[2 * i] = (uint8_t)cx; pos[2 * i + 1] = (uint8_t)cy;
    }
    const int ps = W / 2 - 1;
    const int num_ball = (int)nearbyint((double)p.num_balls / nb);  // python round(): half-to-even
    // per colour index: one extra ball in partition 3 (:349-355), then that colour's balls in its own partition
    // (num_ball == 0: the reference's countdown never returns to zero, so colour 0 gets one extra ball and every ball)
    const int group = num_ball > 0 ? num_ball + 1 : 0x7fffffff;
    const int groups = num_ball > 0 ? (p.num_balls + num_ball - 1) / num_ball : (p.num_balls > 0 ? 1 : 0);
    place_sequence<MODE>(p, g, r, p.num_balls + groups, [&](int k) {
      const int index = k / group, j = k - index * group;
      const uint8_t code = cell(T_BALL, p.ball_colour[index], 0);
      if (j == 0) return boxed(p, code, 0, m + 1, ps, ps);
      return boxed(p, code, (index == 1 || index == 2) ? m + 1 : 0, (index == 1 || index == 3) ? m + 1 : 0, ps, ps);
    }, nothing);
  } else {  // MG_LAYOUT_QUADRANTS_RESPAWN, collect_game.py:376-399
    const int per = p.num_balls / 3;
    place_sequence<MODE>(p, g, r, p.num_balls, [&](int k) {
      const int index = per > 0 ? k / per : 0;  // Ball(self.world, index, 1): the colour IS the partition index (:391)
      return boxed(p, cell(T_BALL, index, 0), index == 0 ? 0 : W / 2 - 1, index == 1 ? H / 2 - 1 : 0, W / 2 + 1, H / 2 + 1);
    }, nothing);
    for (int i = 0; i < A; ++i) {
      GCELL(g, H, 1 + i, H - 2) = p.agent_code[i];
      pos[2 * i] = (uint8_t)(1 + i); pos[2 * i + 1] = (uint8_t)(H - 2);
    }
  }
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// CollectGameEnv.step for ONE env on its shared-memory grid (collect_game.py:183-211).
// Returns the device error bits; h = {step_count, collected, rng_ctr, episodes}.
template <int MODE>
__device__ __forceinline__ int step_one_env(const CollectParams& p, long long e, uint8_t* g, uint8_t* pos, uint8_t* ord,
                                            const int8_t* act, double* rew, int4& h, Rng<MODE>& r, bool& term, bool& trunc,
                                            uint16_t* chg, int& nchg, uint8_t* pick) {
  const int A = p.A;
  int err = 0;
  if (MODE == 1) {  // production order: Fisher-Yates over Philox draws (trace mode replays np.random.permutation)
    for (int i = 0; i < A; ++i) ord[i] = (uint8_t)i;
    for (int i = A - 1; i > 0; --i) {
      const int j = (int)__umulhi(r.u32(), (uint32_t)(i + 1));
      const uint8_t t = ord[i]; ord[i] = ord[j]; ord[j] = t;
    }
  }
  for (int i = 0; i < A; ++i) { rew[i] = 0.0; pick[i] = 0; }  // :187
  h.x += 1;                                  // step_count += 1 :190
  for (int k = 0; k < A; ++k) {              // for i in order :191
    const int i = ord[k];
    const int a = act[i];
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
      if (p.respawn) chg[nchg++] = (uint16_t)respawn<MODE>(p, g, r, colour);  // :142-143 -- may land on (nx, ny)
      h.y += 1;                                  // collected_balls += 1 :144
      const int resp = (c >> 6) & 1;             // placed by _respawn (only marked when its reward differs, mg_create)
      rew[i] += resp ? p.reward_respawned[colour] : p.reward_initial[colour];  // _reward(i, rewards, fwd_cell.reward) :145
      pick[i] = (uint8_t)(1 + (colour | (resp << 4)));
      const int t = p.type_of_colour[colour];
      if (t >= 0) atomicAdd(&p.info[e * (A * p.nb) + p.nb * i + t], 1);  // info[keys[nb*i + ball_idx]] += 1 :147 (fire-and-forget RED)
      enter = true;
    }
    if (enter) {  // wall / other agent: neither ball nor None -> blocked (:169-171)
      GCELL(g, p.H, nx, ny) = p.agent_code[i];  // overwrites a respawn that landed here (ball lost)
      GCELL(g, p.H, ox, oy) = 0;                // also erases a co-located partner from the grid
      pos[2 * i] = (uint8_t)nx; pos[2 * i + 1] = (uint8_t)ny;
      chg[nchg++] = (uint16_t)(nx * p.H + ny); chg[nchg++] = (uint16_t)(ox * p.H + oy);
    }
  }
  term = !p.respawn && h.y == p.num_balls;  // :208-209
  if (p.fixed_horizon) term = false;        // CollectGameRoomsFixedHorizon.step :368-370
  trunc = h.x >= p.max_steps;               // :210-211
  if (p.time_limit > 0 && h.x >= p.time_limit) trunc = true;  // gymnasium TimeLimit of the registration
  return err | r.err;
}


// The same step for the common two-agent case with the env's small state in REGISTERS: `posw` = x0 | y0 << 8 | x1 << 16 | y1 << 24,
// `actw` = the two action bytes, `i0` = the agent that moves first (ord[0]); rewards come back in r0 / r1 and the pickup codes
// (CollectParams::delta, bytes 1..2) in `pick`.  Statement for statement the loop of step_one_env - it only spares the
// dynamically indexed shared-memory traffic, which is most of the dependent-latency chain of a step.
template <int MODE>
__device__ __forceinline__ int step_one_env_a2(const CollectParams& p, int32_t* info_row, uint8_t* g, uint32_t& posw, uint32_t actw, int i0,
                                               double& r0, double& r1, uint32_t& pick, int4& h, Rng<MODE>& r, bool& term, bool& trunc,
                                               uint16_t* chg, int& nchg) {
  int err = 0;
  r0 = 0.0; r1 = 0.0; pick = 0;                 // :187
  h.x += 1;                                     // step_count += 1 :190
  const int H = p.H;
  const uint32_t code0 = p.agent_code[0], code1 = p.agent_code[1];
#pragma unroll
  for (int k = 0; k < 2; ++k) {                 // for i in order :191
    const int i = k == 0 ? i0 : (i0 ^ 1);
    const int a = (int8_t)(i ? (actw >> 8) : actw);
    if (a < 0 || a > 3) continue;               // no branch of :192-207 matches: silently ignored
    const int sh = i * 16;
    const int ox = (posw >> sh) & 255, oy = (posw >> (sh + 8)) & 255;
    const int nx = ox + (a == 1) - (a == 3), ny = oy + (a == 2) - (a == 0);   // agent.py:230-264
    if (nx < 0 || ny < 0 || nx >= p.W || ny >= H) { err |= MG_ERR_OOB; continue; }
    const int n = nx * H + ny;
    const uint8_t c = g[n];
    bool enter = (c == 0);                      // :178-181
    if ((c & 3) == T_BALL) {                    // move_agent :169-177 -> _handle_pickup :132-147
      const int colour = (c >> 2) & 15;
      g[n] = 0;                                 // grid.set(*fwd_pos, None) :141
      if (p.respawn) chg[nchg++] = (uint16_t)respawn<MODE>(p, g, r, colour);  // :142-143 -- may land on (nx, ny)
      h.y += 1;                                 // collected_balls += 1 :144
      const int resp = (c >> 6) & 1;
      const double rw = resp ? p.reward_respawned[colour] : p.reward_initial[colour];   // _reward(i, rewards, fwd_cell.reward) :145
      if (i) r1 += rw; else r0 += rw;
      pick |= (uint32_t)(1 + (colour | (resp << 4))) << (8 * i);
      const int t = p.type_of_colour[colour];
      if (t >= 0) atomicAdd(info_row + p.nb * i + t, 1);   // info[keys[nb*i + ball_idx]] += 1 :147 (fire-and-forget RED)
      enter = true;
    }
    if (enter) {  // wall / other agent: neither ball nor None -> blocked (:169-171)
      g[n] = (uint8_t)(i ? code1 : code0);      // overwrites a respawn that landed here (ball lost)
      g[ox * H + oy] = 0;                       // also erases a co-located partner from the grid
      posw = (posw & ~(0xFFFFu << sh)) | ((uint32_t)(nx | (ny << 8)) << sh);
      chg[nchg++] = (uint16_t)n; chg[nchg++] = (uint16_t)(ox * H + oy);
    }
  }
  term = !p.respawn && h.y == p.num_balls;  // :208-209
  if (p.fixed_horizon) term = false;        // CollectGameRoomsFixedHorizon.step :368-370
  trunc = h.x >= p.max_steps;               // :210-211
  if (p.time_limit > 0 && h.x >= p.time_limit) trunc = true;  // gymnasium TimeLimit of the registration
  return err | r.err;
}

}  // namespace mg
