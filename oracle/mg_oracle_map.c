/* mg_oracle_map.c -- CPU restatement of MazeSingleAgentEnv and CtFMvNEnv (plain C, scalar).
 * TEST INFRASTRUCTURE ONLY (see mg_oracle.h).  Citations are file:line under /root/reference. */
#include <math.h>
#include <string.h>

#include "mg_oracle.h"

/* MazeWorld codes (world.py:81-91) / CtfWorld codes (world.py:66-79) */
enum { MZ_BACKGROUND = 0, MZ_AGENT = 1, MZ_FLAG = 2, MZ_OBSTACLE = 3 };
enum { CT_BLUE_TERR = 0, CT_RED_TERR = 1, CT_BLUE_AGENT = 2, CT_RED_AGENT = 3, CT_BLUE_FLAG = 4, CT_RED_FLAG = 5, CT_OBSTACLE = 6 };

typedef struct { uint64_t seed, env_id; uint32_t ctr, buf[4]; int have; uint32_t h16; int nh16; } prng_t;
static uint32_t p_u32(prng_t* r) {
  if (!r->have) {
    uint32_t c[4] = {(uint32_t)r->env_id, (uint32_t)(r->env_id >> 32), r->ctr, 0u};
    uint32_t k[2] = {(uint32_t)r->seed, (uint32_t)(r->seed >> 32)};
    oc_philox4x32_10(c, k, r->buf);
    r->ctr++; r->have = 4;
  }
  return r->buf[4 - r->have--];
}
static int p_below(prng_t* r, int n) { return (int)(((uint64_t)p_u32(r) * (uint32_t)n) >> 32); }
/* the step's small draws (an action out of 5, a Fisher-Yates index) take 16 bits each: low half of a Philox word first */
static int p_below16(prng_t* r, int n) {
  if (!r->nh16) { r->h16 = p_u32(r); r->nh16 = 2; }
  uint32_t v = r->h16 & 0xFFFFu;
  r->h16 >>= 16; r->nh16--;
  return (int)((v * (uint32_t)n) >> 16);
}

/* CtfActions / MazeActions deltas (agent.py:54-67; ctf.py:1189-1199; maze.py:276-285) */
static const int ADX[5] = {0, 0, -1, 0, 1}, ADY[5] = {0, -1, 0, 1, 0};
/* DIR_TO_VEC (constants.py:65-74): 0 (1,0), 1 (0,1), 2 (-1,0), 3 (0,-1) */
static int dir_of(int dx, int dy, int old) {
  if (dx == 1 && dy == 0) return 0;
  if (dx == 0 && dy == 1) return 1;
  if (dx == -1 && dy == 0) return 2;
  if (dx == 0 && dy == -1) return 3;
  return old; /* Agent.move: no vector matches (0,0) -> dir unchanged (agent.py:176-183) */
}

/* k-th cell (row-major over field_map[x][y], as np.where returns them) whose code == want; -1 if none */
static int nth_cell(const oc_map_cfg* c, int want, int k) {
  const int n = c->size * c->size;
  for (int i = 0; i < n; ++i)
    if (c->field_map[i] == want && k-- == 0) return i;
  return -1;
}
static int count_cells(const oc_map_cfg* c, int want) {
  int n = 0;
  for (int i = 0; i < c->size * c->size; ++i) n += c->field_map[i] == want;
  return n;
}

/* ------------------------------------------------------------------------------------ Maze */
static void maze_encode(const oc_map_cfg* c, const uint8_t* pos, uint8_t* obs) { /* _encode_map maze.py:245-260 */
  const int S = c->size;
  memcpy(obs, c->field_map, (size_t)S * S);
  obs[pos[0] * S + pos[1]] = MZ_AGENT;
}

static void maze_reset_env(const oc_map_cfg* c, uint8_t* pos, uint8_t* dir, int32_t* step, int idx) {
  const int cell = nth_cell(c, MZ_BACKGROUND, idx); /* self.background[np.random.randint(...)] maze.py:202-205 */
  pos[0] = (uint8_t)(cell / c->size); pos[1] = (uint8_t)(cell % c->size);
  *dir = 3;  /* place_agent: dir = 3 (multigrid.py:371-374) */
  *step = 0; /* multigrid.py:141 */
}

int oc_maze_reset(const oc_map_cfg* c, int64_t N, oc_map_state* st, const uint8_t* mask, const oc_map_rng* rng,
                  uint8_t* obs, int32_t* status) {
  const int S = c->size, nbg = count_cells(c, MZ_BACKGROUND);
  (void)status;
  for (int64_t e = 0; e < N; ++e) {
    if (!mask || mask[e]) {
      int idx;
      if (rng->mode == 0) idx = rng->start_index[e];
      else {
        prng_t r = {rng->seed, rng->env_id_base + (uint64_t)e, st->rng_ctr[e], {0}, 0};
        idx = p_below(&r, nbg);
        st->rng_ctr[e] = r.ctr;
      }
      maze_reset_env(c, st->pos + e * 2, st->dir + e, st->step_count + e, idx);
    }
    if (obs) maze_encode(c, st->pos + e * 2, obs + e * S * S);
  }
  return 0;
}

int oc_maze_step(const oc_map_cfg* c, int64_t N, oc_map_state* st, const int8_t* actions, const oc_map_rng* rng,
                 uint8_t* obs, double* reward, uint8_t* terminated, uint8_t* truncated, int autoreset,
                 uint8_t* final_obs, int32_t* status) {
  const int S = c->size, nbg = count_cells(c, MZ_BACKGROUND);
  for (int64_t e = 0; e < N; ++e) {
    uint8_t* pos = st->pos + e * 2;
    st->step_count[e] += 1; /* maze.py:334 */
    const int a = actions[e];
    if (a < 0 || a > 4) { if (status) *status |= OC_ERR_BAD_ACTION; }
    else { /* _move_agent maze.py:271-307 */
      const int nx = pos[0] + ADX[a], ny = pos[1] + ADY[a];
      if (!(nx < 0 || ny < 0 || nx >= S || ny >= S)) {
        /* the grid holds an object on every cell: Floor "background" / Flag overlap, Obstacle overlaps iff
         * penalty != 0 (object.py:200-201), the agent's own cell (action stay) does not (object.py:38-40) */
        const int code = c->field_map[nx * S + ny];
        const int self = (nx == pos[0] && ny == pos[1]);
        if (!self && (code != MZ_OBSTACLE || c->obstacle_penalty != 0)) {
          st->dir[e] = (uint8_t)dir_of(nx - pos[0], ny - pos[1], st->dir[e]);
          pos[0] = (uint8_t)nx; pos[1] = (uint8_t)ny;
        }
      }
    }
    uint8_t term = 0, trunc = st->step_count[e] >= c->max_steps; /* :346-347 */
    double r = 0.0;
    const int here = c->field_map[pos[0] * S + pos[1]];
    if (here == MZ_FLAG) { r += c->flag_reward; term = 1; }                                  /* :354-356 */
    if (c->obstacle_penalty != 0 && here == MZ_OBSTACLE) { r -= c->obstacle_penalty; term = 1; } /* :360-363 */
    r -= c->step_penalty;                                                                     /* :371 */
    reward[e] = r; terminated[e] = term; truncated[e] = trunc;
    if (autoreset && (term || trunc)) {
      if (final_obs) maze_encode(c, pos, final_obs + e * S * S);
      prng_t pr = {rng->seed, rng->env_id_base + (uint64_t)e, st->rng_ctr[e], {0}, 0};
      const int idx = rng->mode == 0 ? rng->start_index[e] : p_below(&pr, nbg);
      if (rng->mode == 1) st->rng_ctr[e] = pr.ctr;
      maze_reset_env(c, pos, st->dir + e, st->step_count + e, idx);
    }
    if (obs) maze_encode(c, pos, obs + e * S * S);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------- CtF */
static int terr_cell(const oc_map_cfg* c, int blue, int k) { /* self.blue_territory = where(map == terr) + [flag] ctf.py:765-773 */
  const int cnt = count_cells(c, blue ? CT_BLUE_TERR : CT_RED_TERR);
  if (k < cnt) return nth_cell(c, blue ? CT_BLUE_TERR : CT_RED_TERR, k);
  return nth_cell(c, blue ? CT_BLUE_FLAG : CT_RED_FLAG, 0);
}
static int in_territory(const oc_map_cfg* c, int blue, int x, int y) { /* _is_agent_in_territory ctf.py:1253-1290 */
  const int code = c->field_map[x * c->size + y];
  return blue ? (code == CT_BLUE_TERR || code == CT_BLUE_FLAG) : (code == CT_RED_TERR || code == CT_RED_FLAG);
}

static void ctf_encode(const oc_map_cfg* c, const uint8_t* pos, const uint8_t* flags, uint8_t* obs) {
  /* _encode_map ctf.py:1137-1163: agents drawn in index order, defeated ones as obstacle; returns .T -> obs[y][x] */
  const int S = c->size, n = c->num_blue + c->num_red;
  for (int x = 0; x < S; ++x)
    for (int y = 0; y < S; ++y) obs[y * S + x] = c->field_map[x * S + y];
  for (int i = 0; i < n; ++i)
    obs[pos[2 * i + 1] * S + pos[2 * i]] = (flags[i] & 1) ? CT_OBSTACLE : (i < c->num_blue ? CT_BLUE_AGENT : CT_RED_AGENT);
}

static void sample_distinct(prng_t* r, int len, int k, int* out) { /* our Philox-mode stand-in for choice(len, k, replace=False) */
  for (int i = 0; i < k; ++i) {
    for (;;) {
      int v = p_below(r, len), dup = 0;
      for (int j = 0; j < i; ++j) dup |= out[j] == v;
      if (!dup) { out[i] = v; break; }
    }
  }
}

static void ctf_reset_env(const oc_map_cfg* c, uint8_t* pos, uint8_t* dir, uint8_t* flags, int32_t* step,
                          const int* bplace, const int* rplace) {
  const int nb = c->num_blue, nr = c->num_red;
  for (int i = 0; i < nb + nr; ++i) {
    const int cell = i < nb ? terr_cell(c, 1, bplace[i]) : terr_cell(c, 0, rplace[i - nb]); /* ctf.py:1033-1048 */
    pos[2 * i] = (uint8_t)(cell / c->size); pos[2 * i + 1] = (uint8_t)(cell % c->size);
    dir[i] = 3;
    if (!c->carry_agent_flags) flags[i] = 0; /* a FRESH env instance; with carry_agent_flags the same instance goes on: the
                                                reference never clears terminated / collided / bg_color in reset */
  }
  *step = 0;
}

int oc_ctf_reset(const oc_map_cfg* c, int64_t N, oc_map_state* st, const uint8_t* mask, const oc_map_rng* rng,
                 uint8_t* obs, int32_t* status) {
  const int S = c->size, nb = c->num_blue, nr = c->num_red, n = nb + nr;
  const int lb = count_cells(c, CT_BLUE_TERR) + 1, lr = count_cells(c, CT_RED_TERR) + 1;
  (void)status;
  for (int64_t e = 0; e < N; ++e) {
    if (!mask || mask[e]) {
      int bp[OC_MAX_CTF_AGENTS], rp[OC_MAX_CTF_AGENTS];
      if (rng->mode == 0) {
        for (int i = 0; i < nb; ++i) bp[i] = rng->blue_place[e * nb + i];
        for (int i = 0; i < nr; ++i) rp[i] = rng->red_place[e * nr + i];
      } else {
        prng_t r = {rng->seed, rng->env_id_base + (uint64_t)e, st->rng_ctr[e], {0}, 0};
        sample_distinct(&r, lb, nb, bp); /* 1v1: one np_random.integers(0, len) each (ctf.py:317,322) = one draw */
        sample_distinct(&r, lr, nr, rp);
        st->rng_ctr[e] = r.ctr;
      }
      ctf_reset_env(c, st->pos + e * n * 2, st->dir + e * n, st->flags + e * n, st->step_count + e, bp, rp);
      if (st->stats) st->stats[e] = 0;
    }
    if (obs) ctf_encode(c, st->pos + e * n * 2, st->flags + e * n, obs + e * S * S);
  }
  return 0;
}

int oc_ctf_step(const oc_map_cfg* c, int64_t N, oc_map_state* st, const int8_t* blue_actions, const oc_map_rng* rng,
                uint8_t* obs, double* reward, uint8_t* terminated, uint8_t* truncated, int autoreset,
                uint8_t* final_obs, int32_t* status) {
  const int S = c->size, nb = c->num_blue, nr = c->num_red, n = nb + nr;
  const int lb = count_cells(c, CT_BLUE_TERR) + 1, lr = count_cells(c, CT_RED_TERR) + 1;
  const int bflag = nth_cell(c, CT_BLUE_FLAG, 0), rflag = nth_cell(c, CT_RED_FLAG, 0);
  for (int64_t e = 0; e < N; ++e) {
    uint8_t* pos = st->pos + e * n * 2; uint8_t* dir = st->dir + e * n; uint8_t* fl = st->flags + e * n;
    prng_t r = {rng->seed, rng->env_id_base + (uint64_t)e, st->rng_ctr ? st->rng_ctr[e] : 0, {0}, 0};
    st->step_count[e] += 1; /* ctf.py:1295 */
    int act[OC_MAX_CTF_AGENTS], order[OC_MAX_CTF_AGENTS];
    for (int i = 0; i < nb; ++i) act[i] = blue_actions[e * nb + i];
    for (int k = 0; k < nr; ++k) /* RwPolicy.act for EVERY red agent, defeated or not (:1297-1301) */
      act[nb + k] = (rng->mode == 0 || rng->red_actions) ? rng->red_actions[e * nr + k] : p_below16(&r, 5); /* mode 1 + red_actions: an external enemy policy (enemy_policies, ctf.py:666), no draw */
    if (c->variant_1v1) { order[0] = 0; order[1] = 1; } /* Ctf1v1Env._move_agents: blue, then red (ctf.py:503-510) */
    else if (rng->mode == 0) for (int i = 0; i < n; ++i) order[i] = rng->order[e * n + i];
    else { /* np_random.shuffle stand-in: Fisher-Yates */
      for (int i = 0; i < n; ++i) order[i] = i;
      for (int i = n - 1; i > 0; --i) { int j = p_below16(&r, i + 1), t = order[i]; order[i] = order[j]; order[j] = t; }
    }
    for (int k = 0; k < n; ++k) { /* _move_agents :1240-1251 */
      const int i = order[k];
      if (fl[i] & 1) continue;
      const int a = act[i];
      if (a < 0 || a > 4) { if (status) *status |= OC_ERR_BAD_ACTION; continue; }
      const int nx = pos[2 * i] + ADX[a], ny = pos[2 * i + 1] + ADY[a]; /* _move_agent :1184-1238 */
      if (nx < 0 || ny < 0 || nx >= S || ny >= S) continue;
      int occupied = 0; /* an agent object (alive or defeated, or itself when staying) sits on the cell */
      for (int j = 0; j < n; ++j) occupied |= (pos[2 * j] == nx && pos[2 * j + 1] == ny);
      if (occupied) { if (c->obstacle_penalty != 0 && !c->variant_1v1) fl[i] |= 2; continue; } /* :1231-1236; the 1v1 env has no collided logic (:498-501) */
      const int code = c->field_map[nx * S + ny];
      if (code == CT_OBSTACLE && c->obstacle_penalty == 0) continue; /* Obstacle.can_overlap() <=> penalty != 0 */
      dir[i] = (uint8_t)dir_of(nx - pos[2 * i], ny - pos[2 * i + 1], dir[i]); /* Agent.move agent.py:167-200 */
      pos[2 * i] = (uint8_t)nx; pos[2 * i + 1] = (uint8_t)ny;
      /* the agent's background colour follows the territory it moved onto and is left alone elsewhere (ctf.py:1214-1230,
       * agent.py:197-200): flags bits 2-3 = 0 as constructed (the team's colour), 1 light_blue, 2 light_red; only render() reads it */
      if (code == CT_BLUE_TERR || code == CT_BLUE_FLAG) fl[i] = (uint8_t)((fl[i] & ~12) | 4);
      else if (code == CT_RED_TERR || code == CT_RED_FLAG) fl[i] = (uint8_t)((fl[i] & ~12) | 8);
    }
    uint8_t term = 0, trunc = st->step_count[e] >= c->max_steps; /* :1310-1311 */
    double rew = 0.0;
    if (c->obstacle_penalty != 0) { /* :1316-1332 (collided is never cleared) */
      for (int i = 0; i < nb; ++i) if (fl[i] & 2) { rew -= c->obstacle_penalty; fl[i] |= 1; }
      for (int i = nb; i < n; ++i) if (fl[i] & 2) fl[i] |= 1;
    }
    int32_t gs = st->stats ? st->stats[e] : 0; /* game_stats (ctf.py:1068-1073): bit0 blue_flag_captured, bit1 red_flag_captured, bit 8+i agent i defeated in a battle */
    for (int i = 0; i < nb; ++i) if (pos[2 * i] * S + pos[2 * i + 1] == rflag) { rew += c->flag_reward; term = 1; gs |= 2; } /* :1335-1344 */
    for (int i = nb; i < n; ++i) if (pos[2 * i] * S + pos[2 * i + 1] == bflag) { rew -= c->flag_reward; term = 1; gs |= 1; } /* :1347-1356 */
    int nbattle = 0;
    for (int b = 0; b < nb; ++b) /* np.where(distances <= battle_range): row-major, blue-major (:1368-1377) */
      for (int q = 0; q < nr; ++q) {
        const int dx = pos[2 * b] - pos[2 * (nb + q)], dy = pos[2 * b + 1] - pos[2 * (nb + q) + 1];
        if (!(sqrt((double)(dx * dx + dy * dy)) <= c->battle_range)) continue;
        if ((fl[b] & 1) || (fl[nb + q] & 1)) continue; /* :1380-1383 */
        const int bh = in_territory(c, 1, pos[2 * b], pos[2 * b + 1]);
        const int rh = in_territory(c, 0, pos[2 * (nb + q)], pos[2 * (nb + q) + 1]);
        int blue_win;
        if (rng->mode == 0) {
          blue_win = nbattle < rng->KB ? rng->blue_win[e * rng->KB + nbattle] : 0;
          if (nbattle >= rng->KB && status) *status |= OC_ERR_TRACE_OVERFLOW;
        } else { /* :1392-1407 */
          const double pb = (bh && !rh) ? c->randomness : ((!bh && rh) ? 1.0 - c->randomness : 0.5);
          blue_win = (double)p_u32(&r) * (1.0 / 4294967296.0) < pb;
        }
        ++nbattle;
        if (blue_win) { rew += c->battle_reward; fl[nb + q] |= 1; gs |= 1 << (8 + nb + q); } /* :1409-1418 */
        else if (c->variant_1v1) { rew -= c->battle_reward; term = 1; gs |= 1 << 8; } /* 1v1: blue losing ends the episode (ctf.py:629-636) */
        else { rew -= c->battle_reward; fl[b] |= 1; gs |= 1 << (8 + b); }
      }
    if (rng->mode == 0 && rng->battles_used) rng->battles_used[e] = nbattle;
    int all_dead = 1;
    for (int i = 0; i < nb; ++i) all_dead &= (fl[i] & 1);
    if (all_dead) term = 1;                  /* :1423 */
    rew -= c->step_penalty * nb;             /* :1428; 1v1: reward -= step_penalty (:646), nb == 1 */
    reward[e] = rew; terminated[e] = term; truncated[e] = trunc;
    if (st->stats) st->stats[e] = gs;
    if (autoreset && (term || trunc)) {
      if (st->stats) st->stats[e] = 0;
      if (final_obs) ctf_encode(c, pos, fl, final_obs + e * S * S);
      int bp[OC_MAX_CTF_AGENTS], rp[OC_MAX_CTF_AGENTS];
      if (rng->mode == 0) {
        for (int i = 0; i < nb; ++i) bp[i] = rng->blue_place[e * nb + i];
        for (int i = 0; i < nr; ++i) rp[i] = rng->red_place[e * nr + i];
      } else { sample_distinct(&r, lb, nb, bp); sample_distinct(&r, lr, nr, rp); }
      ctf_reset_env(c, pos, dir, fl, st->step_count + e, bp, rp);
    }
    if (st->rng_ctr && rng->mode == 1) st->rng_ctr[e] = r.ctr;
    if (obs) ctf_encode(c, pos, fl, obs + e * S * S);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------ infos
 * MazeSingleAgentEnv._get_info (maze.py:262-269) -> [N][2] = d_a_f, d_a_ob
 * CtFMvNEnv._get_info / Ctf1v1Env._get_info (ctf.py:1165-1182, 434-452) -> [N][11] in the dict's key order.
 * distance_points = np.linalg.norm of an int vector; distance_area_point = min over the listed cells (utils/map.py:7-19).
 * Note agents[1] is the SECOND agent of the env's list: a red agent only when there is a single blue one. */
#include <math.h>
static double d_points(int x0, int y0, int x1, int y1) { const int dx = x0 - x1, dy = y0 - y1; return sqrt((double)(dx * dx + dy * dy)); }
static double d_area(const oc_map_cfg* c, int x, int y, int code_a, int code_b) { /* min over cells whose code is code_a or code_b */
  const int S = c->size;
  double best = INFINITY; /* the reference raises on an empty list (np.min of []); inf marks that case */
  for (int i = 0; i < S * S; ++i)
    if (c->field_map[i] == code_a || c->field_map[i] == code_b) { const double d = d_points(x, y, i / S, i % S); if (d < best) best = d; }
  return best;
}
int oc_map_info(const oc_map_cfg* c, int is_maze, int64_t N, const oc_map_state* st, double* out) {
  const int S = c->size, n = is_maze ? 1 : c->num_blue + c->num_red;
  for (int64_t e = 0; e < N; ++e) {
    const uint8_t* pos = st->pos + e * n * 2;
    if (is_maze) {
      out[e * 2 + 0] = d_area(c, pos[0], pos[1], MZ_FLAG, MZ_FLAG);
      out[e * 2 + 1] = d_area(c, pos[0], pos[1], MZ_OBSTACLE, MZ_OBSTACLE);
      continue;
    }
    const int bf = nth_cell(c, CT_BLUE_FLAG, 0), rf = nth_cell(c, CT_RED_FLAG, 0);
    const int ax = pos[0], ay = pos[1], bx = pos[2], by = pos[3]; /* agents[0], agents[1] */
    double* o = out + e * 11;
    o[0] = d_points(ax, ay, bx, by);
    o[1] = d_points(ax, ay, bf / S, bf % S); o[2] = d_points(ax, ay, rf / S, rf % S);
    o[3] = d_points(bx, by, bf / S, bf % S); o[4] = d_points(bx, by, rf / S, rf % S);
    o[5] = d_points(bf / S, bf % S, rf / S, rf % S);
    o[6] = d_area(c, ax, ay, CT_BLUE_TERR, CT_BLUE_FLAG); o[7] = d_area(c, ax, ay, CT_RED_TERR, CT_RED_FLAG); /* territory lists include the flag cell (ctf.py:765-773) */
    o[8] = d_area(c, bx, by, CT_BLUE_TERR, CT_BLUE_FLAG); o[9] = d_area(c, bx, by, CT_RED_TERR, CT_RED_FLAG);
    o[10] = d_area(c, ax, ay, CT_OBSTACLE, CT_OBSTACLE);
  }
  return 0;
}

/* CtFMvNEnv._get_obs, observation_option="flattened" (ctf.py:1084-1104; the dict of :1112-1135 concatenated in key order):
 * blue agent (x, y) pairs, red agent pairs, blue flag, red flag, blue_territory pairs (np.where order, then the blue flag,
 * ctf.py:765-769), red_territory pairs (+ red flag), obstacle pairs, int(agent.terminated) per agent.  out int64 [N][L];
 * returns L = 3 n + 4 + 2 (|blue_territory| + |red_territory| + |obstacle|) (1v1: 2 n + 1 + ...). */
int oc_ctf_flattened(const oc_map_cfg* c, int64_t N, const oc_map_state* st, int64_t* out) {
  const int S = c->size, n = c->num_blue + c->num_red;
  const int nbt = count_cells(c, CT_BLUE_TERR) + 1, nrt = count_cells(c, CT_RED_TERR) + 1, nob = count_cells(c, CT_OBSTACLE);
  const int tail = c->variant_1v1 ? 1 : n; /* Ctf1v1Env ends with int(self._is_red_agent_defeated) alone (ctf.py:359-371) */
  const int L = 2 * n + tail + 4 + 2 * (nbt + nrt + nob);
  if (!out) return L;
  for (int64_t e = 0; e < N; ++e) {
    int64_t* o = out + e * L;
    int k = 0;
    for (int i = 0; i < 2 * n; ++i) o[k++] = st->pos[e * n * 2 + i];
    const int bf = nth_cell(c, CT_BLUE_FLAG, 0), rf = nth_cell(c, CT_RED_FLAG, 0);
    o[k++] = bf / S; o[k++] = bf % S; o[k++] = rf / S; o[k++] = rf % S;
    for (int i = 0; i < nbt; ++i) { const int cell = terr_cell(c, 1, i); o[k++] = cell / S; o[k++] = cell % S; }
    for (int i = 0; i < nrt; ++i) { const int cell = terr_cell(c, 0, i); o[k++] = cell / S; o[k++] = cell % S; }
    for (int i = 0; i < nob; ++i) { const int cell = nth_cell(c, CT_OBSTACLE, i); o[k++] = cell / S; o[k++] = cell % S; }
    for (int i = n - tail; i < n; ++i) o[k++] = st->flags[e * n + i] & 1;
  }
  return L;
}

/* ------------------------------------------------------------------------------------ scripted CtF opponents
 * CPU restatement of csrc/policy_kernels.cu's decision rule (targets: heuristic.py:216-226, 265-272, 321-338, 434-463;
 * follow-or-random: :150-175; the first move of the A* route comes from the caller's [cell][target] table, which
 * tests/test_policy_device_gpu.py fills with the reference-pinned host A*).  Philox blocks: counter (env id,
 * 16 * step_count + block, 2^31 | episode), key = seed; draw order per red agent: patrol target, follow-or-not, random action. */
typedef struct { uint32_t c[4], k[2], buf[4]; int have; } pol_rng_t;
static uint32_t pol_u32(pol_rng_t* r) {
  if (!r->have) { oc_philox4x32_10(r->c, r->k, r->buf); r->c[2]++; r->have = 4; }
  return r->buf[4 - r->have--];
}
static int pol_below(pol_rng_t* r, int n) { return (int)(((uint64_t)pol_u32(r) * (uint32_t)n) >> 32); }

int oc_ctf_policy_actions(const oc_map_cfg* c, int64_t N, const oc_map_state* st, const int32_t* episode, const int32_t* kind,
                          const double* randomness, const uint8_t* first_move, const uint16_t* patrol_goal,
                          const uint8_t* on_border, const uint16_t* along, int32_t n_along, uint64_t seed, uint64_t env_id_base,
                          int8_t* out) {
  const int S = c->size, cells = S * S, nb = c->num_blue, nr = c->num_red, n = nb + nr;
  int blue_flag = -1;
  for (int i = 0; i < cells && blue_flag < 0; ++i) if (c->field_map[i] == CT_BLUE_FLAG) blue_flag = i;
  for (int64_t e = 0; e < N; ++e) {
    const uint8_t* pos = st->pos + (size_t)e * n * 2;
    const uint64_t id = env_id_base + (uint64_t)e;
    pol_rng_t r = {{(uint32_t)id, (uint32_t)(id >> 32), (uint32_t)st->step_count[e] * 16u, 0x80000000u | (uint32_t)episode[e]},
                   {(uint32_t)seed, (uint32_t)(seed >> 32)}, {0}, 0};
    int intruder = 0;
    for (int i = 0; i < nb; ++i) {
      const int code = c->field_map[pos[2 * i] * S + pos[2 * i + 1]];
      intruder |= code == CT_RED_TERR || code == CT_RED_FLAG;
    }
    for (int k = 0; k < nr; ++k) {
      int a;
      if (kind[k] == 0) a = pol_below(&r, 5);
      else {
        const int x = pos[2 * (nb + k)], y = pos[2 * (nb + k) + 1], cell = x * S + y;
        int target;
        if (kind[k] == 2) target = blue_flag;
        else if (kind[k] == 1 || (kind[k] == 4 && intruder)) {
          long best = 1L << 40;
          target = cell;
          for (int i = 0; i < nb; ++i) {
            const long dx = pos[2 * i] - x, dy = pos[2 * i + 1] - y, d2 = dx * dx + dy * dy;
            if (d2 < best) { best = d2; target = pos[2 * i] * S + pos[2 * i + 1]; }
          }
        } else if (on_border[cell]) target = along[pol_below(&r, n_along)];
        else target = patrol_goal[cell];
        const int mv = first_move[(size_t)cell * cells + target];
        const double u = (double)pol_u32(&r) / 4294967296.0;
        a = u < randomness[k] ? mv : pol_below(&r, 5);
      }
      out[e * nr + k] = (int8_t)a;
    }
  }
  return 0;
}
