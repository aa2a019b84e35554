/* mg_oracle_generic.c -- CPU restatement of the base-class MultiGridEnv.step with DefaultWorld.
 * TEST INFRASTRUCTURE ONLY (see mg_oracle.h). */
#include <string.h>

#include "mg_oracle.h"

enum { G_EMPTY = 1, G_DOOR = 4, G_GOAL = 8, G_AGENT = 10 }; /* DefaultWorld.OBJECT_TO_IDX world.py:37-51 */

void oc_generic_encode(int64_t N, int W, int H, int A, const uint8_t* gcell, const uint8_t* gstate, const uint8_t* pos, uint8_t* obs) {
  const int cells = W * H;
  for (int64_t e = 0; e < N; ++e)
    for (int k = 0; k < A; ++k) {
      const int px = pos[(e * A + k) * 2], py = pos[(e * A + k) * 2 + 1];
      for (int i = 0; i < cells; ++i) { /* encode_for_agents grid.py:254-284 */
        const uint8_t c = gcell[e * cells + i], s = gstate[e * cells + i];
        const int type = c & 15;
        uint8_t* o = obs + ((e * A + k) * cells + i) * 6;
        o[0] = (uint8_t)type; o[1] = c >> 4; o[2] = 0; o[3] = 0; o[4] = 0; o[5] = 0;
        if (type == G_DOOR) o[2] = s;                                            /* Door.encode object.py:238-259 */
        else if (type == G_AGENT) { o[4] = s & 3; o[5] = (i == px * H + py); }   /* Agent.encode agent.py:127-165 (never carrying) */
      }
    }
}

int oc_generic_step(int64_t N, int W, int H, int A, int max_steps, uint8_t* gcell, uint8_t* gstate, uint8_t* pos,
                    int32_t* step_count, const int8_t* actions, const uint8_t* order, uint8_t* obs, double* rewards,
                    uint8_t* terminated, uint8_t* truncated, int32_t* status) {
  static const int DX[4] = {1, 0, -1, 0}, DY[4] = {0, 1, 0, -1}; /* DIR_TO_VEC constants.py:65-74 */
  const int cells = W * H;
  for (int64_t e = 0; e < N; ++e) {
    uint8_t* gc = gcell + e * cells; uint8_t* gs = gstate + e * cells; uint8_t* p = pos + e * A * 2;
    step_count[e] += 1; /* multigrid.py:400 */
    uint8_t term = 0;
    for (int i = 0; i < A; ++i) rewards[e * A + i] = 0.0;
    for (int k = 0; k < A; ++k) { /* for i in order :408 */
      const int i = order[e * A + k], a = actions[e * A + i];
      if (a == 0) continue; /* still :413 */
      const int x = p[2 * i], y = p[2 * i + 1], here = x * H + y, dir = gs[here] & 3;
      const int fx = x + DX[dir], fy = y + DY[dir];
      if (a == 1) gs[here] = (uint8_t)((dir + 3) & 3);        /* left :424-427 */
      else if (a == 2) gs[here] = (uint8_t)((dir + 1) & 3);   /* right :430-431 */
      else if (a == 3) {                                      /* forward :434-445 */
        if (fx < 0 || fy < 0 || fx >= W || fy >= H) { if (status) *status |= OC_ERR_OOB; continue; } /* reference: assert */
        const int f = fx * H + fy, ftype = gc[f] & 15;
        if (ftype != G_EMPTY) {
          if (ftype == G_GOAL) { /* terminated + _reward :436-438, :218-223 */
            term = 1;
            rewards[e * A + i] += 1 - 0.9 * ((double)step_count[e] / (double)max_steps);
          } /* switch: empty hook; anything else: nothing */
        } else { /* an agent only ever advances into an EMPTY cell (:441-444) */
          gc[f] = gc[here]; gs[f] = gs[here];
          gc[here] = G_EMPTY; gs[here] = 0;
          p[2 * i] = (uint8_t)fx; p[2 * i + 1] = (uint8_t)fy;
        }
      } else if (status) *status |= OC_ERR_BAD_ACTION; /* reference raises (multigrid.py:447 / :469) */
    }
    terminated[e] = term;
    truncated[e] = step_count[e] >= max_steps; /* :470-471 */
  }
  if (obs) oc_generic_encode(N, W, H, A, gcell, gstate, pos, obs);
  return 0;
}

/* MultiGridEnv.gen_obs for DefaultWorld (encode_dim 6), restated step by step as the reference performs it:
 * Agent.get_view_exts (agent.py:294-324) -> Grid.slice (grid.py:111-130, out of bounds = Wall) -> (dir + 1) x Grid.rotate_left
 * (grid.py:97-109) -> Grid.process_vis (grid.py:286-323; Wall and a closed / locked Door block sight, object.py:178-179, 223-224)
 * -> Grid.encode_for_agents with agent_pos = (V // 2, V - 1) (grid.py:254-284).  out: u8 [N][A][V][V][6].
 * `dirs` overrides the agents' directions (NULL = the dir stored with the agent cell). */
void oc_partial_view6(int64_t N, int W, int H, int A, int V, int see_through_walls, const uint8_t* gcell, const uint8_t* gstate,
                      const uint8_t* pos, const uint8_t* dirs, uint8_t* out) {
  enum { VMAX = 32, G_WALL = 2 };
  if (V < 1 || V > VMAX) return;
  const uint8_t WALL_GREY = (uint8_t)(G_WALL | (7 << 4)); /* Wall(world): colour "grey" = index 7 (constants.py:8-19) */
  for (int64_t e = 0; e < N; ++e)
    for (int k = 0; k < A; ++k) {
      const uint8_t* gc = gcell + e * W * H; const uint8_t* gs = gstate + e * W * H;
      const int px = pos[(e * A + k) * 2], py = pos[(e * A + k) * 2 + 1];
      const int dir = dirs ? dirs[e * A + k] : (gs[px * H + py] & 3);
      int topX, topY;
      if (dir == 0) { topX = px; topY = py - V / 2; }
      else if (dir == 1) { topX = px - V / 2; topY = py; }
      else if (dir == 2) { topX = px - V + 1; topY = py - V / 2; }
      else { topX = px - V / 2; topY = py - V + 1; }
      uint8_t c0[VMAX][VMAX], s0[VMAX][VMAX], c1[VMAX][VMAX], s1[VMAX][VMAX];
      for (int j = 0; j < V; ++j) /* slice */
        for (int i = 0; i < V; ++i) {
          const int x = topX + i, y = topY + j;
          if (x >= 0 && x < W && y >= 0 && y < H) { c0[i][j] = gc[x * H + y]; s0[i][j] = gs[x * H + y]; }
          else { c0[i][j] = WALL_GREY; s0[i][j] = 0; }
        }
      for (int r = 0; r < dir + 1; ++r) { /* rotate_left: new(j, V-1-i) = old(i, j) */
        for (int i = 0; i < V; ++i)
          for (int j = 0; j < V; ++j) { c1[j][V - 1 - i] = c0[i][j]; s1[j][V - 1 - i] = s0[i][j]; }
        memcpy(c0, c1, sizeof c0); memcpy(s0, s1, sizeof s0);
      }
      uint8_t mask[VMAX][VMAX];
      memset(mask, see_through_walls ? 1 : 0, sizeof mask);
      if (!see_through_walls) {
        mask[V / 2][V - 1] = 1;
        for (int j = V - 1; j >= 0; --j) {
          for (int i = 0; i < V - 1; ++i) {
            if (!mask[i][j]) continue;
            const int type = c0[i][j] & 15;
            if (type == G_WALL || (type == G_DOOR && s0[i][j] != 0)) continue; /* not see_behind() */
            mask[i + 1][j] = 1;
            if (j > 0) { mask[i + 1][j - 1] = 1; mask[i][j - 1] = 1; }
          }
          for (int i = V - 1; i >= 1; --i) {
            if (!mask[i][j]) continue;
            const int type = c0[i][j] & 15;
            if (type == G_WALL || (type == G_DOOR && s0[i][j] != 0)) continue;
            mask[i - 1][j] = 1;
            if (j > 0) { mask[i - 1][j - 1] = 1; mask[i][j - 1] = 1; }
          }
        }
      }
      uint8_t* o = out + ((e * A + k) * V * V) * 6;
      for (int i = 0; i < V; ++i)
        for (int j = 0; j < V; ++j) {
          uint8_t* q = o + (i * V + j) * 6;
          q[0] = q[1] = q[2] = q[3] = q[4] = q[5] = 0; /* unseen */
          if (!mask[i][j]) continue;
          const int type = c0[i][j] & 15;
          q[0] = (uint8_t)type; q[1] = c0[i][j] >> 4;
          if (type == G_DOOR) q[2] = s0[i][j];
          else if (type == G_AGENT) { q[4] = s0[i][j] & 3; q[5] = (i == V / 2 && j == V - 1); }
        }
    }
}
