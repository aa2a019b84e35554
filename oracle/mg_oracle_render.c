/* mg_oracle_render.c -- CPU restatement of the reference's rgb_array render for the Collect family.
 *
 * TEST INFRASTRUCTURE ONLY (see mg_oracle.h).  Pinned bit-for-bit against frames recorded from the unmodified reference
 * (oracle/gen_golden.py render -> tests/golden/render_*.npz, tests/test_oracle_golden.py).
 *
 * Follows, statement by statement:
 *   Grid.render        core/grid.py:183-221   tiles blitted at [j*ts, (j+1)*ts) x [i*ts, (i+1)*ts), float64 tile -> uint8 frame
 *   Grid.render_tile   core/grid.py:132-181   ts*3 supersampled image, object, the two grid lines, downsample(3)
 *   fill_coords / point_in_rect / point_in_circle / point_in_triangle / rotate_fn / downsample   utils/rendering.py:8-144
 *   Wall.render object.py:181-182, Ball.render object.py:320-321, Agent.render agent.py:105-117, COLORS constants.py:8-19
 * Highlighting (render(highlight=True), multigrid.py:560-595) is not restated: the default is off.
 *
 * Built with -ffp-contract=off: the reference evaluates every product and sum in its own rounding step. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mg_oracle.h"

static const uint8_t OC_COLORS[10][3] = {  /* constants.py:8-19, dict order = COLOR_TO_IDX */
    {228, 3, 3}, {255, 140, 0}, {255, 237, 0}, {0, 128, 38}, {0, 77, 255},
    {117, 7, 135}, {120, 79, 23}, {100, 100, 100}, {234, 153, 153}, {90, 170, 223}};

typedef struct { int kind; double cx, cy, ct, st; } shape_t;  /* 0 rect-all, 1 ball circle, 2 agent triangle, 3 / 4 grid lines */

static int in_triangle(double x, double y) {  /* point_in_triangle((0.12,0.19),(0.87,0.50),(0.12,0.81)), rendering.py:108-133 */
  const double ax = 0.12, ay = 0.19, bx = 0.87, by = 0.50, cx = 0.12, cy = 0.81;
  const double v0x = cx - ax, v0y = cy - ay, v1x = bx - ax, v1y = by - ay, v2x = x - ax, v2y = y - ay;
  const double dot00 = v0x * v0x + v0y * v0y, dot01 = v0x * v1x + v0y * v1y, dot02 = v0x * v2x + v0y * v2y;
  const double dot11 = v1x * v1x + v1y * v1y, dot12 = v1x * v2x + v1y * v2y;
  const double inv_denom = 1 / (dot00 * dot11 - dot01 * dot01);
  const double u = (dot11 * dot02 - dot01 * dot12) * inv_denom;
  const double v = (dot00 * dot12 - dot01 * dot02) * inv_denom;
  return (u >= 0) && (v >= 0) && (u + v) < 1;
}

static int shape_hit(const shape_t* s, double x, double y) {
  switch (s->kind) {
    case 0: return x >= 0 && x <= 1 && y >= 0 && y <= 1;                                   /* point_in_rect(0, 1, 0, 1) */
    case 1: return (x - 0.5) * (x - 0.5) + (y - 0.5) * (y - 0.5) <= 0.31 * 0.31;            /* point_in_circle(0.5, 0.5, 0.31) */
    case 2: {                                                                              /* rotate_fn(tri_fn, 0.5, 0.5, theta), rendering.py:46-57 */
      const double xx = x - s->cx, yy = y - s->cy;
      const double x2 = s->cx + xx * s->ct - yy * s->st;
      const double y2 = s->cy + yy * s->ct + xx * s->st;
      return in_triangle(x2, y2);
    }
    case 3: return x >= 0 && x <= 0.031 && y >= 0 && y <= 1;                               /* left grid line, grid.py:160 */
    default: return x >= 0 && x <= 1 && y >= 0 && y <= 0.031;                              /* top grid line, grid.py:161 */
  }
}

static void fill_coords(uint8_t* img, int S, const shape_t* s, const uint8_t* color, const uint8_t* bg_color) {  /* rendering.py:24-43 */
  for (int y = 0; y < S; ++y)
    for (int x = 0; x < S; ++x) {
      const double yf = (y + 0.5) / S, xf = (x + 0.5) / S;
      if (shape_hit(s, xf, yf)) memcpy(img + ((size_t)y * S + x) * 3, color, 3);
      else if (bg_color) memcpy(img + ((size_t)y * S + x) * 3, bg_color, 3);
    }
}

/* Grid.render_tile(world, obj, highlights=[], tile_size=ts, subdivs=3) -> out u8 [ts][ts][3] (the float64 tile cast the way
 * `img[ymin:ymax, xmin:xmax, :] = tile_img` casts it: truncation).  kind: -1 = None cell, 0 rect, 1 circle, 2 agent triangle
 * rotated by `dir`; bg = fill_coords' bg_color or NULL. */
static int render_tile(int kind, int dir, const uint8_t* fg, const uint8_t* bg, int ts, uint8_t* out) {
  const int S = ts * 3;
  uint8_t* img = (uint8_t*)calloc((size_t)S * S * 3, 1);
  if (!img) return -1;
  shape_t sh = {0, 0.5, 0.5, 0, 0};
  if (kind >= 0) {
    const double theta = 0.5 * 3.141592653589793 * dir;   /* agent.py:114: theta=0.5 * math.pi * self.dir */
    sh.kind = kind; sh.ct = cos(-theta); sh.st = sin(-theta);
    fill_coords(img, S, &sh, fg, bg);
  }
  const uint8_t line[3] = {100, 100, 100};
  sh.kind = 3; fill_coords(img, S, &sh, line, NULL);
  sh.kind = 4; fill_coords(img, S, &sh, line, NULL);
  /* downsample(img, 3): reshape [ts,3,ts,3,3], mean(axis=3) then mean(axis=1) in float64 (rendering.py:8-21) */
  for (int y = 0; y < ts; ++y)
    for (int x = 0; x < ts; ++x)
      for (int c = 0; c < 3; ++c) {
        double rows[3];
        for (int sy = 0; sy < 3; ++sy) {
          double acc = 0;
          for (int sx = 0; sx < 3; ++sx) acc += (double)img[((size_t)(3 * y + sy) * S + (3 * x + sx)) * 3 + c];
          rows[sy] = acc / 3;
        }
        const double m = ((rows[0] + rows[1]) + rows[2]) / 3;
        out[((size_t)y * ts + x) * 3 + c] = (uint8_t)m;
      }
  free(img);
  return 0;
}

/* CollectWorld tile from the cell's encode() triple: Wall.render object.py:181-182, Ball.render :320-321, Agent.render agent.py:105-117 */
int oc_render_tile(int type, int colour, int state, int ts, uint8_t* out) {
  if (ts <= 0 || type < 0 || type > 3 || colour < 0 || colour > 9) return -1;
  const int kind = type == OC_T_EMPTY ? -1 : (type == OC_T_WALL ? 0 : (type == OC_T_BALL ? 1 : 2));
  return render_tile(kind, type == OC_T_AGENT ? state : 0, OC_COLORS[colour], NULL, ts, out);
}

/* MazeSingleAgentEnv.render(): N envs on one map (field_map [S][S] index [x][y], MazeWorld codes 0 background / 2 flag /
 * 3 obstacle; maze.py:183-198), agent of env e at pos[e] = (x, y) facing dir[e] -> out u8 [N][S*ts][S*ts][3].
 * Floor white (object.py:147-148), Obstacle grey (:203-204), Flag red on white (:366-372), Agent blue on white
 * (maze.py:93-101, agent.py:105-117); MAZE_COLORS constants.py:37-49. */
int oc_render_maze(const uint8_t* field_map, int S, int64_t N, const int16_t* pos, const int8_t* dir, int ts, uint8_t* out) {
  static const uint8_t white[3] = {255, 250, 250}, red[3] = {228, 3, 3}, grey[3] = {100, 100, 100}, blue[3] = {0, 77, 255};
  const size_t tile_bytes = (size_t)ts * ts * 3, row_bytes = (size_t)S * ts * 3;
  uint8_t* tiles = (uint8_t*)calloc(8 * tile_bytes, 1);   /* 0 floor, 2 flag, 3 obstacle, 4 + dir agent */
  if (!tiles) return -1;
  int rc = render_tile(0, 0, white, NULL, ts, tiles) | render_tile(1, 0, red, white, ts, tiles + 2 * tile_bytes) |
           render_tile(0, 0, grey, NULL, ts, tiles + 3 * tile_bytes);
  for (int d = 0; d < 4; ++d) rc |= render_tile(2, d, blue, white, ts, tiles + (size_t)(4 + d) * tile_bytes);
  for (int64_t e = 0; e < N && !rc; ++e)
    for (int j = 0; j < S; ++j)
      for (int i = 0; i < S; ++i) {
        int code = field_map[i * S + j];
        if (pos[2 * e] == i && pos[2 * e + 1] == j) code = 4 + (dir[e] & 3);   /* the agent object occupies its cell (agent.py:195-196) */
        if (code == 1 || code > 7) { rc = -1; break; }
        const uint8_t* tile = tiles + (size_t)code * tile_bytes;
        for (int y = 0; y < ts; ++y)
          memcpy(out + ((size_t)e * S * ts + (size_t)j * ts + y) * row_bytes + (size_t)i * ts * 3, tile + (size_t)y * ts * 3, (size_t)ts * 3);
      }
  free(tiles);
  return rc ? -1 : 0;
}

/* Grid.render(tile_size) of N grids given as Grid.encode() observations u8 [N][W][H][3] -> out u8 [N][H*ts][W*ts][3]. */
int oc_render_grid(const uint8_t* obs, int64_t N, int W, int H, int ts, uint8_t* out) {
  uint8_t* cache = (uint8_t*)calloc((size_t)4 * 10 * 4 * ts * ts * 3, 1);
  uint8_t* have = (uint8_t*)calloc(4 * 10 * 4, 1);
  if (!cache || !have) { free(cache); free(have); return -1; }
  const size_t tile_bytes = (size_t)ts * ts * 3, row_bytes = (size_t)W * ts * 3;
  int rc = 0;
  for (int64_t e = 0; e < N && !rc; ++e)
    for (int j = 0; j < H && !rc; ++j)
      for (int i = 0; i < W; ++i) {
        const uint8_t* t = obs + ((e * W + i) * H + j) * 3;
        int type = t[0], colour = t[1], state = t[2];
        if (type == OC_T_EMPTY) { colour = 0; state = 0; }   /* None cell: no object, key = (tile_size,) */
        if (type > 3 || colour > 9 || state > 3) { rc = -1; break; }
        const int k = (type * 10 + colour) * 4 + state;
        uint8_t* tile = cache + (size_t)k * tile_bytes;
        if (!have[k]) { if (oc_render_tile(type, colour, state, ts, tile)) { rc = -1; break; } have[k] = 1; }
        for (int y = 0; y < ts; ++y)
          memcpy(out + ((size_t)e * H * ts + (size_t)j * ts + y) * row_bytes + (size_t)i * ts * 3, tile + (size_t)y * ts * 3, (size_t)ts * 3);
      }
  free(cache); free(have);
  return rc;
}

/* CtFMvNEnv.render() (Ctf1v1Env: variant_1v1): one map (field_map [S][S] index [x][y], CtfWorld codes 0 blue territory,
 * 1 red territory, 4 blue flag, 5 red flag, 6 obstacle; ctf.py:998-1031), n agents per env (the first num_blue are blue)
 * at pos [N][n][2] facing dir [N][n] with flags [N][n]: bit 0 terminated, bits 2-3 background colour (see oc_ctf_step).
 * Floor light_blue / light_red (object.py:147-148), Obstacle grey, Flag blue on light_blue / red on light_red (:366-372);
 * Agent triangle (agent.py:105-117) blue / red, blue_grey / red_grey once terminated (ctf.py:1316-1332, 1409-1418; the 1v1
 * env never recolours), on its sticky background colour; agent tiles are uncached (ctf.py:69).  CTF_COLORS constants.py:21-35. */
int oc_render_ctf(const uint8_t* field_map, int S, int64_t N, int n, int num_blue, int variant_1v1, const uint8_t* pos,
                  const uint8_t* dir, const uint8_t* flags, int ts, uint8_t* out) {
  static const uint8_t light_blue[3] = {240, 248, 255}, light_red[3] = {255, 228, 225}, grey[3] = {100, 100, 100},
                       blue[3] = {0, 77, 255}, red[3] = {228, 3, 3}, blue_grey[3] = {140, 146, 172}, red_grey[3] = {170, 152, 169};
  const size_t tile_bytes = (size_t)ts * ts * 3, row_bytes = (size_t)S * ts * 3;
  uint8_t* tiles = (uint8_t*)calloc((8 + 32) * tile_bytes, 1);   /* static codes 0..6, then agents 8 + ((team * 2 + grey) * 2 + bg_red) * 4 + dir */
  if (!tiles) return -1;
  int rc = render_tile(0, 0, light_blue, NULL, ts, tiles) | render_tile(0, 0, light_red, NULL, ts, tiles + tile_bytes) |
           render_tile(1, 0, blue, light_blue, ts, tiles + 4 * tile_bytes) | render_tile(1, 0, red, light_red, ts, tiles + 5 * tile_bytes) |
           render_tile(0, 0, grey, NULL, ts, tiles + 6 * tile_bytes);
  for (int team = 0; team < 2; ++team)
    for (int g = 0; g < 2; ++g)
      for (int bg = 0; bg < 2; ++bg)
        for (int d = 0; d < 4; ++d)
          rc |= render_tile(2, d, team ? (g ? red_grey : red) : (g ? blue_grey : blue), bg ? light_red : light_blue, ts,
                            tiles + (size_t)(8 + ((team * 2 + g) * 2 + bg) * 4 + d) * tile_bytes);
  for (int64_t e = 0; e < N && !rc; ++e)
    for (int j = 0; j < S; ++j)
      for (int i = 0; i < S; ++i) {
        int code = field_map[i * S + j];
        if (code == 2 || code == 3 || code > 6) { rc = -1; break; }
        for (int a = 0; a < n; ++a) {
          const uint8_t* p = pos + (e * n + a) * 2;
          if (p[0] != i || p[1] != j) continue;
          const int team = a >= num_blue, fl = flags[e * n + a], bgs = (fl >> 2) & 3;
          const int bg_red = bgs ? bgs == 2 : team;                 /* 0 = as constructed: the team's own light colour */
          const int g = (fl & 1) && !variant_1v1;
          code = 8 + ((team * 2 + g) * 2 + bg_red) * 4 + (dir[e * n + a] & 3);
        }
        const uint8_t* tile = tiles + (size_t)code * tile_bytes;
        for (int y = 0; y < ts; ++y)
          memcpy(out + ((size_t)e * S * ts + (size_t)j * ts + y) * row_bytes + (size_t)i * ts * 3, tile + (size_t)y * ts * 3, (size_t)ts * 3);
      }
  free(tiles);
  return rc ? -1 : 0;
}
