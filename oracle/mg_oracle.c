/* mg_oracle.c -- CPU restatement of the gym-multigrid Collect hot path (plain C, scalar loops).
 *
 * TEST INFRASTRUCTURE ONLY (see mg_oracle.h).  Written for clarity, one env at a time, in
 * the order the reference executes; OpenMP only spreads independent envs over host cores
 * so that bench.py can quote an all-cores CPU baseline.
 *
 * Citations are file:line under /root/reference.
 */
#include "mg_oracle.h"

#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------------- Philox4x32-10
 * Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11).  Not part of the
 * reference (which uses python `random` / legacy numpy MT19937, neither reproducible on a
 * device): this is the production-mode generator both this oracle and the CUDA kernels use. */
void oc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* per-env cursor over the random source */
typedef struct {
  int mode;
  const uint8_t* draws; int n, k;          /* trace */
  uint64_t seed, env_id; uint32_t ctr;     /* philox */
  uint32_t buf[4]; int have;
  int32_t err;
} rng_t;

static void rng_open(rng_t* r, const oc_rng_src* src, int64_t e, uint32_t ctr) {
  memset(r, 0, sizeof *r);
  r->mode = src ? src->mode : 0;
  if (r->mode == 0) {
    if (src && src->draws) { r->draws = src->draws + (int64_t)e * src->K; r->n = src->n_draws ? src->n_draws[e] : src->K; }
  } else {
    r->seed = src->seed; r->env_id = src->env_id_base + (uint64_t)e; r->ctr = ctr;
  }
}

static uint32_t rng_u32(rng_t* r) {
  if (!r->have) {
    uint32_t c[4] = {(uint32_t)r->env_id, (uint32_t)(r->env_id >> 32), r->ctr, 0u};
    uint32_t k[2] = {(uint32_t)r->seed, (uint32_t)(r->seed >> 32)};
    oc_philox4x32_10(c, k, r->buf);
    r->ctr++; r->have = 4;
  }
  return r->buf[4 - r->have--];
}

/* MultiGridEnv._rand_int = random.randint(low, high), INCLUSIVE bounds (multigrid.py:225-230) */
static int rng_int(rng_t* r, int lo, int hi) {
  if (r->mode == 0) {
    if (r->k >= r->n) { r->err |= OC_ERR_TRACE_OVERFLOW; return lo; }
    int v = r->draws[r->k++];
    if (v < lo || v > hi) r->err |= OC_ERR_TRACE_RANGE;
    return v;
  }
  return lo + (int)(((uint64_t)rng_u32(r) * (uint32_t)(hi - lo + 1)) >> 32);
}

/* One placement candidate of place_obj (multigrid.py:316-321, x before y).  Trace mode replays the two
 * recorded randint outputs; Philox mode spends one 32-bit word per candidate (x low half, y high half). */
static void rng_pair(rng_t* r, int lox, int hix, int loy, int hiy, int* x, int* y) {
  if (r->mode == 0) {
    *x = rng_int(r, lox, hix);
    *y = rng_int(r, loy, hiy);
  } else {
    uint32_t w = rng_u32(r);
    *x = lox + (int)(((w & 0xFFFFu) * (uint32_t)(hix - lox + 1)) >> 16);
    *y = loy + (int)(((w >> 16) * (uint32_t)(hiy - loy + 1)) >> 16);
  }
}

/* -------------------------------------------------------------------------------- encode */
void oc_encode3(const uint8_t* cells, int64_t n, uint8_t* obs) {
  /* Grid.encode (grid.py:223-252): None -> (empty=0,0,0); WorldObj.encode (object.py:58-74) ->
   * (OBJECT_TO_IDX[type], COLOR_TO_IDX[color], 0); Agent.encode (agent.py:119-126) -> state=dir */
  for (int64_t i = 0; i < n; ++i) {
    uint8_t c = cells[i];
    obs[3 * i + 0] = c & 3;
    obs[3 * i + 1] = (c >> 2) & 15;
    obs[3 * i + 2] = (c & 3) == OC_T_BALL ? 0 : c >> 6; /* a ball's bit 6 is the "placed by _respawn" mark below, not its STATE */
  }
}

/* ----------------------------------------------------------------------------- placement */
#define CELL(g, H, x, y) (g)[(x) * (H) + (y)]

/* MultiGridEnv.place_obj (multigrid.py:282-339): rejection-sample an EMPTY cell in
 * [top, min(top+size, dim-1)] inclusive, x drawn before y. */
static void place_obj(const oc_collect_cfg* c, uint8_t* g, rng_t* r, uint8_t code, int tx, int ty, int sx,
                      int sy, int* ox, int* oy) {
  const int W = c->width, H = c->height;
  if (tx < 0) tx = 0;
  if (ty < 0) ty = 0;
  int hx = tx + sx < W - 1 ? tx + sx : W - 1, hy = ty + sy < H - 1 ? ty + sy : H - 1;
  for (;;) {
    int x, y;
    rng_pair(r, tx, hx, ty, hy, &x, &y);
    if (r->err & OC_ERR_TRACE_OVERFLOW) { *ox = x; *oy = y; return; } /* leave the grid untouched */
    if (CELL(g, H, x, y) != 0) continue;
    CELL(g, H, x, y) = code;
    *ox = x; *oy = y;
    return;
  }
}

static void horz_wall(uint8_t* g, int H, int x, int y, int len) { /* grid.py:66-78 */
  for (int i = 0; i < len; ++i) CELL(g, H, x + i, y) = OC_WALL_GREY;
}
static void vert_wall(uint8_t* g, int H, int x, int y, int len) { /* grid.py:80-89 */
  for (int j = 0; j < len; ++j) CELL(g, H, x, y + j) = OC_WALL_GREY;
}

static uint8_t agent_code(const oc_collect_cfg* c, int i) {
  return OC_CELL(OC_T_AGENT, c->agent_colour[i], 3); /* dir = 3 after place_agent (multigrid.py:371-374) */
}

static int reset_env(const oc_collect_cfg* c, uint8_t* g, uint8_t* pos, int32_t* step, int32_t* collected,
                     int32_t* info, rng_t* r) {
  const int W = c->width, H = c->height, A = c->num_agents, nb = c->num_ball_types;
  /* CollectGameEnv.reset (collect_game.py:107-119) */
  *collected = 0;
  for (int k = 0; k < A * nb; ++k) info[k] = 0;
  *step = 0; /* multigrid.py:141 */
  memset(g, 0, (size_t)W * H);
  horz_wall(g, H, 0, 0, W); horz_wall(g, H, 0, H - 1, W);
  vert_wall(g, H, 0, 0, H); vert_wall(g, H, W - 1, 0, H);
  int x, y;
  switch (c->layout) {
  case OC_LAYOUT_EVEN_DIST: { /* collect_game.py:236-259 */
    int per = c->num_balls / nb; /* :234 */
    for (int t = 0; t < nb; ++t)
      for (int b = 0; b < per; ++b) place_obj(c, g, r, OC_CELL(OC_T_BALL, c->ball_colour[t], 0), 0, 0, W, H, &x, &y);
    for (int i = 0; i < A; ++i) { /* place_agent(a) -> place_obj anywhere (multigrid.py:364-369) */
      place_obj(c, g, r, agent_code(c, i), 0, 0, W, H, &x, &y);
      pos[2 * i] = (uint8_t)x; pos[2 * i + 1] = (uint8_t)y;
    }
    break;
  }
  case OC_LAYOUT_QUADRANTS: { /* collect_game.py:266-300 */
    if (nb > 4) return -1;
    int per = c->num_balls / nb;
    int px[4] = {0, W / 2 - 1, W / 2 - 1, 0}, py[4] = {0, H / 2 - 1, 0, H / 2};
    for (int t = 0; t < nb; ++t)
      for (int b = 0; b < per; ++b)
        place_obj(c, g, r, OC_CELL(OC_T_BALL, c->ball_colour[t], 0), px[t], py[t], W / 2 - 1, H / 2 - 1, &x, &y);
    for (int i = 0; i < A; ++i) { /* place_agent(a, pos) overwrites (put_obj, multigrid.py:341-348) */
      x = 1 + i; y = H - 2;
      if (x >= W) return -1;
      CELL(g, H, x, y) = agent_code(c, i);
      pos[2 * i] = (uint8_t)x; pos[2 * i + 1] = (uint8_t)y;
    }
    break;
  }
  case OC_LAYOUT_ROOMS: { /* collect_game.py:306-362 (uses `width` on both axes) */
    int ws = W / 2 - 1;
    horz_wall(g, H, 0, W / 2, ws); horz_wall(g, H, W - ws, W / 2, ws);
    vert_wall(g, H, W / 2, 0, ws); vert_wall(g, H, W / 2, W - ws, ws);
    int cx[5] = {W / 2, W / 2 - 1, W / 2 - 1, W / 2 + 1, W / 2 + 1};
    int cy[5] = {W / 2, W / 2 - 1, W / 2 + 1, W / 2 + 1, W / 2 - 1};
    for (int i = 0; i < A; ++i) { /* _rand_elem -> _rand_int(0, 4) (multigrid.py:246-253) */
      int k = rng_int(r, 0, 4);
      CELL(g, H, cx[k], cy[k]) = agent_code(c, i); /* overwrites a previously placed agent */
      pos[2 * i] = (uint8_t)cx[k]; pos[2 * i + 1] = (uint8_t)cy[k];
    }
    int px[4] = {0, W / 2 + 1, W / 2 + 1, 0}, py[4] = {0, W / 2 + 1, 0, W / 2 + 1};
    int ps = W / 2 - 1;
    int num_ball = (int)nearbyint((double)c->num_balls / nb); /* python round(), half-to-even */
    if (num_ball <= 0) return -1;
    int index = 0, tx = 0, ty = 0;
    for (int ball = 0; ball < c->num_balls; ++ball) {
      if (ball % num_ball == 0) {
        index = ball / num_ball;
        if (index >= nb || index >= 4) return -1; /* reference: IndexError */
        tx = px[index]; ty = py[index];
        /* the extra ball of this colour in partition 3 (:349-355) */
        place_obj(c, g, r, OC_CELL(OC_T_BALL, c->ball_colour[index], 0), px[3], py[3], ps, ps, &x, &y);
      }
      place_obj(c, g, r, OC_CELL(OC_T_BALL, c->ball_colour[index], 0), tx, ty, ps, ps, &x, &y);
    }
    break;
  }
  case OC_LAYOUT_QUADRANTS_RESPAWN: { /* collect_game.py:376-399 */
    int px[3] = {0, W / 2 - 1, W / 2 - 1}, py[3] = {0, H / 2 - 1, 0};
    int per = c->num_balls / 3;
    if (per <= 0) return -1;
    int index = 0, tx = 0, ty = 0;
    for (int ball = 0; ball < c->num_balls; ++ball) {
      if (ball % per == 0) {
        index = ball / per;
        if (index >= 3) return -1; /* reference: IndexError */
        tx = px[index]; ty = py[index];
      }
      /* Ball(self.world, index, 1): the colour IS the partition index (:391) */
      place_obj(c, g, r, OC_CELL(OC_T_BALL, index, 0), tx, ty, W / 2 + 1, H / 2 + 1, &x, &y);
    }
    for (int i = 0; i < A; ++i) {
      x = 1 + i; y = H - 2;
      if (x >= W) return -1;
      CELL(g, H, x, y) = agent_code(c, i);
      pos[2 * i] = (uint8_t)x; pos[2 * i + 1] = (uint8_t)y;
    }
    break;
  }
  default: return -1;
  }
  return 0;
}

static int type_of_colour(const oc_collect_cfg* c, int colour);

/* Ball.reward (object.py:308-321) is an attribute of the ball OBJECT, and the reference gives it two different values:
 *   balls placed by _gen_grid:  balls_reward[type] (collect_game.py:98-101, :252, :287, :354, :359); QuadrantsRespawn: the literal 1 (:393)
 *   balls placed by _respawn:   balls_reward[colour index] (:130, :409 -- indexed by COLOUR, not by type)
 * One byte per cell keeps (type, colour); where the two values differ for some colour the env can hold, a respawned ball also
 * carries bit 6 of its cell (never shown: oc_encode3).  Where the reference raises (colour >= len(balls_reward): IndexError)
 * the initial value is kept. */
static double reward_initial(const oc_collect_cfg* c, int colour) {
  if (c->layout == OC_LAYOUT_QUADRANTS_RESPAWN) return 1.0;
  int t = type_of_colour(c, colour);
  return t >= 0 ? c->ball_reward[t] : 1.0;
}
static double reward_respawned(const oc_collect_cfg* c, int colour) {
  return colour < c->num_ball_types ? c->ball_reward[colour] : reward_initial(c, colour);
}
int oc_collect_marks_respawned(const oc_collect_cfg* c) {
  if (!c->respawn) return 0;
  if (c->layout == OC_LAYOUT_QUADRANTS_RESPAWN) {
    for (int k = 0; k < 3; ++k) if (reward_initial(c, k) != reward_respawned(c, k)) return 1;
    return 0;
  }
  for (int t = 0; t < c->num_ball_types; ++t)
    if (reward_initial(c, c->ball_colour[t]) != reward_respawned(c, c->ball_colour[t])) return 1;
  return 0;
}

/* CollectGameEnv._respawn (:129-130) / CollectGameQuadrantsRespawn._respawn (:401-409) */
static void respawn(const oc_collect_cfg* c, uint8_t* g, rng_t* r, int colour) {
  const int W = c->width, H = c->height;
  const int mark = oc_collect_marks_respawned(c);
  int x, y;
  if (c->layout == OC_LAYOUT_QUADRANTS_RESPAWN) {
    int px[3] = {0, W / 2 - 1, W / 2 - 1}, py[3] = {0, H / 2 - 1, 0};
    int p = colour < 3 ? colour : 0; /* reference: IndexError for colour >= 3 */
    place_obj(c, g, r, OC_CELL(OC_T_BALL, colour, mark), px[p], py[p], W / 2 + 1, H / 2 + 1, &x, &y);
  } else {
    place_obj(c, g, r, OC_CELL(OC_T_BALL, colour, mark), 0, 0, W, H, &x, &y);
  }
}

static int type_of_colour(const oc_collect_cfg* c, int colour) {
  for (int t = 0; t < c->num_ball_types; ++t)
    if (c->ball_colour[t] == colour) return t;
  return -1;
}

static void step_env(const oc_collect_cfg* c, uint8_t* g, uint8_t* pos, int32_t* step, int32_t* collected,
                     int32_t* info, const int8_t* act, const uint8_t* order, rng_t* r, double* rew,
                     uint8_t* term, uint8_t* trunc) {
  const int W = c->width, H = c->height, A = c->num_agents, nb = c->num_ball_types;
  static const int DX[4] = {0, 1, 0, -1}, DY[4] = {-1, 0, 1, 0}; /* north east south west, agent.py:230-264 */
  for (int i = 0; i < A; ++i) rew[i] = 0.0; /* :187 */
  *step += 1;                                /* :190 */
  for (int k = 0; k < A; ++k) {              /* for i in order :191 */
    int i = order[k];
    int a = act[i];
    if (a < 0 || a > 3) continue; /* no branch matches: silently ignored :192-207 */
    int ox = pos[2 * i], oy = pos[2 * i + 1];
    int nx = ox + DX[a], ny = oy + DY[a];
    if (nx < 0 || ny < 0 || nx >= W || ny >= H) { r->err |= OC_ERR_OOB; continue; }
    uint8_t cell = CELL(g, H, nx, ny);
    int enter = 0;
    if ((cell & 3) == OC_T_BALL) { /* move_agent :169-177 -> _handle_pickup :132-147 */
      int colour = (cell >> 2) & 15;
      CELL(g, H, nx, ny) = 0;                       /* grid.set(*fwd_pos, None) :141 */
      if (c->respawn) respawn(c, g, r, colour);     /* :142-143, may land on (nx, ny) */
      *collected += 1;                              /* :144 */
      rew[i] += (cell >> 6) & 1 ? reward_respawned(c, colour) : reward_initial(c, colour); /* fwd_cell.reward :145 */
      if (colour < nb) info[nb * i + colour] += 1;  /* :147 info[keys[num_ball_types * i + ball_idx]], ball_idx = the COLOUR index (:139) */
      enter = 1;
    } else if (cell == 0) { /* :178-181 */
      enter = 1;
    } /* wall / agent: not None and not ball -> nothing happens :169-171 */
    if (enter) {
      CELL(g, H, nx, ny) = agent_code(c, i); /* grid.set(*next_pos, agent) - overwrites a respawn that landed here */
      CELL(g, H, ox, oy) = 0;                /* grid.set(*agent.pos, None) - also erases a co-located "ghost" partner */
      pos[2 * i] = (uint8_t)nx; pos[2 * i + 1] = (uint8_t)ny;
    }
  }
  *term = (!c->respawn && *collected == c->num_balls) ? 1 : 0; /* :208-209 */
  if (c->fixed_horizon) *term = 0;                             /* :368-370 */
  *trunc = (*step >= c->max_steps) ? 1 : 0;                    /* :210-211 */
  if (c->time_limit > 0 && *step >= c->time_limit) *trunc = 1; /* gymnasium TimeLimit (registration) */
}

/* production-mode agent order: Fisher-Yates over Philox draws (the reference's legacy
 * np.random.permutation is replayed, not re-implemented, in trace mode). */
static void philox_order(rng_t* r, int A, uint8_t* order) {
  for (int i = 0; i < A; ++i) order[i] = (uint8_t)i;
  for (int i = A - 1; i > 0; --i) {
    int j = (int)(((uint64_t)rng_u32(r) * (uint32_t)(i + 1)) >> 32);
    uint8_t t = order[i]; order[i] = order[j]; order[j] = t;
  }
}

int oc_collect_reset(const oc_collect_cfg* c, int64_t N, oc_collect_state* st, const uint8_t* mask,
                     const oc_rng_src* rng, uint8_t* obs, int32_t* status, int nthreads) {
  const int cells = c->width * c->height, A = c->num_agents, nb = c->num_ball_types;
  int32_t err = 0;
  int rc = 0;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 512) num_threads(nthreads) reduction(| : err) reduction(| : rc)
  for (int64_t e = 0; e < N; ++e) {
    if (mask && !mask[e]) continue;
    rng_t r;
    rng_open(&r, rng, e, st->rng_ctr ? st->rng_ctr[e] : 0);
    uint8_t* g = st->grid + e * cells;
    if (reset_env(c, g, st->agent_pos + e * 2 * A, st->step_count + e, st->collected + e, st->info + e * A * nb, &r))
      rc |= 1;
    if (st->rng_ctr && r.mode == 1) st->rng_ctr[e] = r.ctr;
    if (rng && rng->draws_used && r.mode == 0) rng->draws_used[e] = r.k;
    err |= r.err;
    if (obs) oc_encode3(g, cells, obs + e * 3 * cells);
  }
  if (status) *status |= err;
  return rc ? -1 : 0;
}

int oc_collect_step(const oc_collect_cfg* c, int64_t N, oc_collect_state* st, const int8_t* actions,
                    const oc_rng_src* rng, uint8_t* obs, double* rewards, uint8_t* terminated,
                    uint8_t* truncated, int autoreset, const oc_rng_src* reset_rng, uint8_t* final_obs,
                    int32_t* status, int nthreads) {
  const int cells = c->width * c->height, A = c->num_agents, nb = c->num_ball_types;
  int32_t err = 0;
  int rc = 0;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 512) num_threads(nthreads) reduction(| : err) reduction(| : rc)
  for (int64_t e = 0; e < N; ++e) {
    rng_t r;
    rng_open(&r, rng, e, st->rng_ctr ? st->rng_ctr[e] : 0);
    uint8_t* g = st->grid + e * cells;
    uint8_t* pos = st->agent_pos + e * 2 * A;
    uint8_t ord[OC_MAX_AGENTS];
    if (r.mode == 0) memcpy(ord, rng->order + e * A, (size_t)A);
    else philox_order(&r, A, ord);
    uint8_t term, trunc;
    step_env(c, g, pos, st->step_count + e, st->collected + e, st->info + e * A * nb, actions + e * A, ord, &r,
             rewards + e * A, &term, &trunc);
    if (rng->draws_used && r.mode == 0) rng->draws_used[e] = r.k;
    terminated[e] = term; truncated[e] = trunc;
    if (autoreset && (term || trunc)) {
      if (final_obs) oc_encode3(g, cells, final_obs + e * 3 * cells);
      rng_t rr;
      rng_t* pr = &r;
      if (r.mode == 0) { rng_open(&rr, reset_rng, e, 0); pr = &rr; }
      if (reset_env(c, g, pos, st->step_count + e, st->collected + e, st->info + e * A * nb, pr)) rc |= 1;
      if (r.mode == 0) { err |= rr.err; if (reset_rng && reset_rng->draws_used) reset_rng->draws_used[e] = rr.k; }
    }
    if (st->rng_ctr && r.mode == 1) st->rng_ctr[e] = r.ctr;
    err |= r.err;
    if (obs) oc_encode3(g, cells, obs + e * 3 * cells);
  }
  if (status) *status |= err;
  return rc ? -1 : 0;
}

/* ----------------------------------------------------------------------------- partial views */
static uint8_t view_cell(const uint8_t* g, int W, int H, int x, int y, int oob) {
  if (x < 0 || y < 0 || x >= W || y >= H) return (uint8_t)oob; /* Grid.slice: v = Wall(self.world) grid.py:124-127 */
  return g[x * H + y];
}

void oc_partial_view3(const uint8_t* grid, const uint8_t* pos, const uint8_t* dirs, int64_t N, int W, int H, int A,
                      int V, int see_through_walls, int oob_code, int opaque_rule, uint8_t* out) {
  enum { VMAX = 32 };
  if (V > VMAX || V < 1) return;
  for (int64_t e = 0; e < N; ++e)
    for (int k = 0; k < A; ++k) {
      const uint8_t* g = grid + e * W * H;
      const int x = pos[(e * A + k) * 2], y = pos[(e * A + k) * 2 + 1];
      const int dir = dirs ? dirs[e * A + k] : 3;
      uint8_t cell[VMAX][VMAX];
      /* get_view_exts (agent.py:294-324) + slice + (dir+1) x rotate_left, folded into direct indices */
      const int hs = V / 2;
      for (int a = 0; a < V; ++a)
        for (int b = 0; b < V; ++b) {
          int wx, wy;
          switch (dir) {
          case 0: wx = x + V - 1 - b;      wy = y - hs + a;          break; /* facing right: one rotation   */
          case 1: wx = x - hs + V - 1 - a; wy = y + V - 1 - b;       break; /* facing down:  two rotations  */
          case 2: wx = x - V + 1 + b;      wy = y - hs + V - 1 - a;  break; /* facing left:  three rotations */
          default: wx = x - hs + a;        wy = y - V + 1 + b;       break; /* facing up:    four = identity */
          }
          cell[a][b] = view_cell(g, W, H, wx, wy, oob_code);
        }
      uint8_t mask[VMAX][VMAX];
      memset(mask, see_through_walls ? 1 : 0, sizeof mask);
      if (!see_through_walls) { /* process_vis grid.py:286-323; only Wall blocks sight in CollectWorld (object.py:174-179) */
        mask[hs][V - 1] = 1;
        for (int j = V - 1; j >= 0; --j) {
          for (int i = 0; i < V - 1; ++i) {
            if (!mask[i][j]) continue;
            if (opaque_rule == 0 ? (cell[i][j] & 3) == OC_T_WALL : cell[i][j] == oob_code) continue;
            mask[i + 1][j] = 1;
            if (j > 0) { mask[i + 1][j - 1] = 1; mask[i][j - 1] = 1; }
          }
          for (int i = V - 1; i >= 1; --i) {
            if (!mask[i][j]) continue;
            if (opaque_rule == 0 ? (cell[i][j] & 3) == OC_T_WALL : cell[i][j] == oob_code) continue;
            mask[i - 1][j] = 1;
            if (j > 0) { mask[i - 1][j - 1] = 1; mask[i][j - 1] = 1; }
          }
        }
      }
      uint8_t* o = out + ((e * A + k) * V * V) * 3;
      for (int a = 0; a < V; ++a)
        for (int b = 0; b < V; ++b) { /* encode_for_agents grid.py:254-284: unseen cells stay (0,0,0) */
          const uint8_t c = mask[a][b] ? cell[a][b] : 0;
          o[(a * V + b) * 3 + 0] = c & 3; o[(a * V + b) * 3 + 1] = (c >> 2) & 15; o[(a * V + b) * 3 + 2] = (c & 3) == OC_T_BALL ? 0 : c >> 6;  /* type 2 = Collect ball (bit 6: respawn mark) or Maze flag, STATE 0 either way */
        }
    }
}

/* ------------------------------------------------------------------------------ Toroid wrapper */
void oc_toroid(const uint8_t* grid, const uint8_t* pos, int64_t N, int W, int A, int nb, float* out) {
  const int depth = nb + A; /* toroid.py:24 */
  memset(out, 0, (size_t)N * A * W * W * depth * sizeof(float));
  for (int64_t e = 0; e < N; ++e)
    for (int k = 0; k < A; ++k) {
      const int px = pos[(e * A + k) * 2], py = pos[(e * A + k) * 2 + 1];
      float* tor = out + ((e * A + k) * W * W) * depth;
      for (int i = 0; i < W; ++i)
        for (int j = 0; j < W; ++j) { /* toroid.py:46-66 */
          int nx = i - px, ny = j - py;
          if (nx < 0) nx += W;
          if (ny < 0) ny += W;
          const uint8_t c = grid[e * W * W + i * W + j];
          const int type = c & 3, colour = (c >> 2) & 15;
          if (c == 0) continue;
          int ch = -1;
          if (type == OC_T_WALL) ch = depth - 1;
          else if (type == OC_T_BALL) ch = colour;
          else if (type == OC_T_AGENT && !(i == px && j == py)) ch = depth - 2; /* obj.pos != this agent's pos */
          if (ch >= 0 && ch < depth) tor[(ny * W + nx) * depth + ch] = 1.0f;
        }
    }
}
