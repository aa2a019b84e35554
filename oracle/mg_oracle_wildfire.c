/* mg_oracle_wildfire.c -- CPU restatement of the Wildfire extension's specification (see mg_oracle.h).
 * TEST INFRASTRUCTURE ONLY.  There is no reference Wildfire code; "parity unpinned". */
#include <stdlib.h>
#include <string.h>

#include "mg_oracle.h"

enum { WF_HEALTHY = 0, WF_BURNING = 1, WF_BURNT = 2 };

typedef struct { uint64_t seed, env_id; uint32_t ctr, buf[4]; int have; } wrng_t;
static uint32_t w_u32(wrng_t* r) {
  if (!r->have) {
    uint32_t c[4] = {(uint32_t)r->env_id, (uint32_t)(r->env_id >> 32), r->ctr, 0u};
    uint32_t k[2] = {(uint32_t)r->seed, (uint32_t)(r->seed >> 32)};
    oc_philox4x32_10(c, k, r->buf);
    r->ctr++; r->have = 4;
  }
  return r->buf[4 - r->have--];
}
static int w_below(wrng_t* r, int n) { return (int)(((uint64_t)w_u32(r) * (uint32_t)n) >> 32); }

static void wf_encode(const oc_wf_cfg* c, const uint8_t* t, const uint8_t* ag, uint8_t* obs) {
  static const uint8_t COL[3] = {3, 0, 7}; /* green, red, grey (constants.py:8-19) */
  const int cells = c->width * c->height;
  for (int i = 0; i < cells; ++i) { obs[3 * i] = t[i]; obs[3 * i + 1] = COL[t[i]]; obs[3 * i + 2] = 0; }
  for (int k = 0; k < c->num_agents; ++k) { /* agents in index order (distinct cells) */
    const int i = ag[4 * k] * c->height + ag[4 * k + 1];
    obs[3 * i] = 3; obs[3 * i + 1] = (uint8_t)c->agent_colour[k]; obs[3 * i + 2] = ag[4 * k + 2];
  }
}

static void wf_reset_env(const oc_wf_cfg* c, uint8_t* t, uint8_t* ag, int32_t* h, wrng_t* r) {
  const int cells = c->width * c->height;
  memset(t, WF_HEALTHY, (size_t)cells);
  for (int f = 0; f < c->num_fires; ++f)
    for (;;) { const int i = w_below(r, cells); if (t[i] == WF_HEALTHY) { t[i] = WF_BURNING; break; } }
  for (int k = 0; k < c->num_agents; ++k)
    for (;;) {
      const int i = w_below(r, cells);
      int taken = 0;
      for (int j = 0; j < k; ++j) taken |= (ag[4 * j] * c->height + ag[4 * j + 1] == i);
      if (taken) continue;
      ag[4 * k] = (uint8_t)(i / c->height); ag[4 * k + 1] = (uint8_t)(i % c->height); ag[4 * k + 2] = 3; ag[4 * k + 3] = 0;
      break;
    }
  h[0] = 0; h[3] += 1;
}

int oc_wf_reset(const oc_wf_cfg* c, int64_t N, oc_wf_state* st, const uint8_t* mask, uint64_t seed, uint64_t env_id_base,
                uint8_t* obs) {
  const int cells = c->width * c->height, A = c->num_agents;
  for (int64_t e = 0; e < N; ++e) {
    int32_t* h = st->hdr + e * 4;
    if (!mask || mask[e]) {
      wrng_t r = {seed, env_id_base + (uint64_t)e, (uint32_t)h[2], {0}, 0};
      wf_reset_env(c, st->terrain + e * cells, st->agents + e * A * 4, h, &r);
      h[2] = (int32_t)r.ctr;
    }
    if (obs) wf_encode(c, st->terrain + e * cells, st->agents + e * A * 4, obs + e * cells * 3);
  }
  return 0;
}

int oc_wf_step(const oc_wf_cfg* c, int64_t N, oc_wf_state* st, const int8_t* actions, const uint8_t* order_in,
               uint64_t seed, uint64_t env_id_base, uint8_t* obs, double* rewards, uint8_t* terminated, uint8_t* truncated,
               int autoreset, uint8_t* final_obs) {
  const int W = c->width, H = c->height, cells = W * H, A = c->num_agents;
  static const int ADX[5] = {0, 0, -1, 0, 1}, ADY[5] = {0, -1, 0, 1, 0};
  uint8_t* nt = (uint8_t*)malloc((size_t)cells);
  for (int64_t e = 0; e < N; ++e) {
    uint8_t* t = st->terrain + e * cells; uint8_t* ag = st->agents + e * A * 4; int32_t* h = st->hdr + e * 4;
    wrng_t r = {seed, env_id_base + (uint64_t)e, (uint32_t)h[2], {0}, 0};
    h[0] += 1; h[1] += 1; /* step_count, tick */
    int order[OC_MAX_WF_AGENTS];
    if (order_in) for (int i = 0; i < A; ++i) order[i] = order_in[e * A + i];
    else {
      for (int i = 0; i < A; ++i) order[i] = i;
      for (int i = A - 1; i > 0; --i) { int j = w_below(&r, i + 1), tmp = order[i]; order[i] = order[j]; order[j] = tmp; }
    }
    for (int i = 0; i < A; ++i) rewards[e * A + i] = 0.0;
    for (int k = 0; k < A; ++k) { /* 2. ordered agent moves */
      const int i = order[k], a = actions[e * A + i];
      int x = ag[4 * i], y = ag[4 * i + 1];
      if (a >= 1 && a <= 4) {
        const int nx = x + ADX[a], ny = y + ADY[a];
        if (nx >= 0 && ny >= 0 && nx < W && ny < H) {
          int occ = 0;
          for (int j = 0; j < A; ++j) occ |= (j != i && ag[4 * j] == nx && ag[4 * j + 1] == ny);
          if (!occ) {
            ag[4 * i + 2] = (uint8_t)(ADX[a] == 1 ? 0 : ADY[a] == 1 ? 1 : ADX[a] == -1 ? 2 : 3); /* DIR_TO_VEC */
            ag[4 * i] = (uint8_t)nx; ag[4 * i + 1] = (uint8_t)ny; x = nx; y = ny;
          }
        }
      }
      if (t[x * H + y] == WF_BURNING) { t[x * H + y] = WF_BURNT; rewards[e * A + i] += 1.0; }
    }
    int burning = 0; /* 3. fire dynamics, double buffered */
    const uint32_t kk[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int i = 0; i < cells; ++i) {
      const int x = i / H, y = i % H, s = t[i];
      int ns = s;
      int k = 0;
      if (s == WF_HEALTHY) {
        if (x > 0) k += t[i - H] == WF_BURNING;
        if (x < W - 1) k += t[i + H] == WF_BURNING;
        if (y > 0) k += t[i - 1] == WF_BURNING;
        if (y < H - 1) k += t[i + 1] == WF_BURNING;
      }
      if (s == WF_BURNING || k > 0) {
        const uint32_t ctr[4] = {(uint32_t)(env_id_base + e), (uint32_t)((env_id_base + e) >> 32), (uint32_t)h[1], 1u + (uint32_t)(i / 4)};
        uint32_t o[4];
        oc_philox4x32_10(ctr, kk, o);
        const uint32_t u = o[i & 3];
        if (s == WF_BURNING) ns = u < c->burnout_threshold ? WF_BURNT : WF_BURNING;
        else ns = u < c->ignite_threshold[k] ? WF_BURNING : WF_HEALTHY;
      }
      nt[i] = (uint8_t)ns;
      burning += ns == WF_BURNING;
    }
    memcpy(t, nt, (size_t)cells);
    const uint8_t term = burning == 0, trunc = h[0] >= c->max_steps;
    terminated[e] = term; truncated[e] = trunc;
    if (autoreset && (term || trunc)) {
      if (final_obs) wf_encode(c, t, ag, final_obs + e * cells * 3);
      wf_reset_env(c, t, ag, h, &r);
    }
    h[2] = (int32_t)r.ctr;
    if (obs) wf_encode(c, t, ag, obs + e * cells * 3);
  }
  free(nt);
  return 0;
}
