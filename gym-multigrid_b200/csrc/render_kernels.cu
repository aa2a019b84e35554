// render_kernels.cu -- rgb_array frames of selected envs:
//   MultiGridEnv.render(highlight=False)   multigrid.py:546-606
//   Grid.render / Grid.render_tile         core/grid.py:132-221
//   fill_coords, point_in_*, rotate_fn, downsample   utils/rendering.py:8-144
//
// The reference rasterises each distinct tile once (a class-level cache keyed by the object's encode() triple) and then
// blits tiles into the frame.  Same split here: the tile ATLAS - one ts x ts x 3 image per cell code - is rasterised on
// the host when a tile size is first asked for (a few hundred tiny images; double arithmetic in the reference's order of
// operations, libm cos / sin as CPython's math module calls them), and the per-frame work, the blit, is the kernel:
// one CTA per (frame, row of tiles) streams the atlas rows (L2-resident) into the frame with warp-contiguous 16-byte
// stores (directly, or through a shared-memory strip for tile sizes whose rows are not multiples of 16 bytes), so a
// frame costs exactly its own bytes of HBM writes.
#include <cmath>
#include <cstring>
#include <vector>

#include "mg_device.cuh"

namespace mg {

namespace {

struct TileSpec {      // what Grid.render_tile draws under the two grid lines
  int shape;           // 0 nothing (None cell), 1 full rect, 2 circle r = 0.31, 3 agent triangle
  int dir;             // triangle rotation: theta = 0.5 * pi * dir (agent.py:114)
  uint8_t fg[3];
  bool has_bg;         // fill_coords' bg_color (Flag / Agent with a bg_color): painted where the shape test fails
  uint8_t bg[3];
};

// constants.py:8-19 (COLORS, dict order = COLOR_TO_IDX) and :37-49 (MAZE_COLORS)
const uint8_t kColors[10][3] = {{228, 3, 3}, {255, 140, 0}, {255, 237, 0}, {0, 128, 38}, {0, 77, 255},
                                {117, 7, 135}, {120, 79, 23}, {100, 100, 100}, {234, 153, 153}, {90, 170, 223}};
const uint8_t kMazeWhite[3] = {255, 250, 250}, kMazeRed[3] = {228, 3, 3}, kMazeGrey[3] = {100, 100, 100}, kMazeBlue[3] = {0, 77, 255};
// CTF_COLORS constants.py:21-35
const uint8_t kCtfLightBlue[3] = {240, 248, 255}, kCtfLightRed[3] = {255, 228, 225}, kCtfBlueGrey[3] = {140, 146, 172}, kCtfRedGrey[3] = {170, 152, 169};

bool hit(const TileSpec& t, double x, double y, double ct, double st) {
  if (t.shape == 1) return x >= 0 && x <= 1 && y >= 0 && y <= 1;                          // point_in_rect(0, 1, 0, 1)
  if (t.shape == 2) return (x - 0.5) * (x - 0.5) + (y - 0.5) * (y - 0.5) <= 0.31 * 0.31;   // point_in_circle(0.5, 0.5, 0.31)
  // rotate_fn(point_in_triangle((0.12, 0.19), (0.87, 0.50), (0.12, 0.81)), cx=0.5, cy=0.5, theta)   rendering.py:46-57, 108-133
  const double rx = x - 0.5, ry = y - 0.5;
  const double px = 0.5 + rx * ct - ry * st, py = 0.5 + ry * ct + rx * st;
  const double a[2] = {0.12, 0.19}, b[2] = {0.87, 0.50}, c[2] = {0.12, 0.81};
  const double v0[2] = {c[0] - a[0], c[1] - a[1]}, v1[2] = {b[0] - a[0], b[1] - a[1]}, v2[2] = {px - a[0], py - a[1]};
  auto dot = [](const double* p, const double* q) { return p[0] * q[0] + p[1] * q[1]; };
  const double d00 = dot(v0, v0), d01 = dot(v0, v1), d02 = dot(v0, v2), d11 = dot(v1, v1), d12 = dot(v1, v2);
  const double inv = 1 / (d00 * d11 - d01 * d01);
  const double u = (d11 * d02 - d01 * d12) * inv, v = (d00 * d12 - d01 * d02) * inv;
  return (u >= 0) && (v >= 0) && (u + v) < 1;
}

// Grid.render_tile(..., tile_size=ts, subdivs=3) -> u8 [ts][ts][3], the float64 tile truncated as the frame assignment does
void raster_tile(const TileSpec& t, int ts, uint8_t* out) {
  const int S = ts * 3;
  std::vector<uint8_t> img((size_t)S * S * 3, 0);
  const double theta = 0.5 * 3.141592653589793 * t.dir;
  const double ct = std::cos(-theta), st = std::sin(-theta);
  const uint8_t line[3] = {100, 100, 100};
  for (int y = 0; y < S; ++y)
    for (int x = 0; x < S; ++x) {
      const double yf = (y + 0.5) / S, xf = (x + 0.5) / S;   // fill_coords: pixel centres in [0, 1]
      uint8_t* px = &img[((size_t)y * S + x) * 3];
      if (t.shape) {
        if (hit(t, xf, yf, ct, st)) std::memcpy(px, t.fg, 3);
        else if (t.has_bg) std::memcpy(px, t.bg, 3);
      }
      if ((xf >= 0 && xf <= 0.031 && yf >= 0 && yf <= 1) || (xf >= 0 && xf <= 1 && yf >= 0 && yf <= 0.031)) std::memcpy(px, line, 3);  // grid.py:160-161
    }
  for (int y = 0; y < ts; ++y)       // downsample(img, 3): mean over the 3 sub-columns, then over the 3 sub-rows, float64
    for (int x = 0; x < ts; ++x)
      for (int c = 0; c < 3; ++c) {
        double m[3];
        for (int sy = 0; sy < 3; ++sy) {
          const uint8_t* r = &img[((size_t)(3 * y + sy) * S + 3 * x) * 3 + c];
          m[sy] = (((double)r[0] + (double)r[3]) + (double)r[6]) / 3;
        }
        out[((size_t)y * ts + x) * 3 + c] = (uint8_t)(((m[0] + m[1]) + m[2]) / 3);
      }
}

}  // namespace

// Host: the atlas of a family for one tile size, u8 [256][ts][ts][3]; codes no object of the family's world maps to stay black.
//   Collect (CollectWorld): index = the packed grid byte  type | colour << 2 | dir << 6  (empty 0, wall 1, ball 2, agent 3)
//   Maze (MazeWorld, maze.py:93-101, 183-198): 0 background Floor (white), 2 Flag (red on white), 3 Obstacle (grey),
//         4 + dir = the agent (blue triangle on white)
//   CtF (CtfWorld): the map's codes 0 / 1 / 4 / 5 / 6, agents from 8 (team, grey, background colour, dir)
void build_render_atlas(int family, int ts, std::vector<uint8_t>& atlas) {
  const size_t tile = (size_t)ts * ts * 3;
  atlas.assign(256 * tile, 0);
  auto put = [&](int code, int shape, int dir, const uint8_t* fg, const uint8_t* bg) {
    TileSpec t{};
    t.shape = shape; t.dir = dir; t.has_bg = bg != nullptr;
    if (fg) std::memcpy(t.fg, fg, 3);
    if (bg) std::memcpy(t.bg, bg, 3);
    raster_tile(t, ts, &atlas[(size_t)code * tile]);
  };
  if (family == 0) {   // MG_FAMILY_COLLECT
    put(0, 0, 0, nullptr, nullptr);
    for (int colour = 0; colour < 10; ++colour) {
      put(1 | (colour << 2), 1, 0, kColors[colour], nullptr);                                       // Wall.render object.py:181-182
      put(2 | (colour << 2), 2, 0, kColors[colour], nullptr);                                       // Ball.render object.py:320-321
      for (int dir = 0; dir < 4; ++dir) put(3 | (colour << 2) | (dir << 6), 3, dir, kColors[colour], nullptr);  // Agent.render agent.py:105-117
    }
  } else if (family == 2) {   // MG_FAMILY_CTF (CtfWorld; ctf.py:998-1031, 765-815)
    put(0, 1, 0, kCtfLightBlue, nullptr);                                                           // Floor "blue_territory"
    put(1, 1, 0, kCtfLightRed, nullptr);                                                            // Floor "red_territory"
    put(4, 2, 0, kMazeBlue, kCtfLightBlue);                                                         // Flag blue on light_blue
    put(5, 2, 0, kMazeRed, kCtfLightRed);                                                           // Flag red on light_red
    put(6, 1, 0, kMazeGrey, nullptr);                                                               // Obstacle
    for (int team = 0; team < 2; ++team)        // agents: 8 + ((team * 2 + grey) * 2 + background is light_red) * 4 + dir
      for (int g = 0; g < 2; ++g)               // grey once terminated (ctf.py:1316-1332, 1409-1418)
        for (int bg = 0; bg < 2; ++bg)
          for (int dir = 0; dir < 4; ++dir)
            put(8 + ((team * 2 + g) * 2 + bg) * 4 + dir, 3, dir, team ? (g ? kCtfRedGrey : kMazeRed) : (g ? kCtfBlueGrey : kMazeBlue),
                bg ? kCtfLightRed : kCtfLightBlue);
  } else {             // MG_FAMILY_MAZE
    put(0, 1, 0, kMazeWhite, nullptr);                                                              // Floor.render object.py:147-148
    put(2, 2, 0, kMazeRed, kMazeWhite);                                                             // Flag.render object.py:366-372
    put(3, 1, 0, kMazeGrey, nullptr);                                                               // Obstacle.render object.py:203-204
    for (int dir = 0; dir < 4; ++dir) put(4 + dir, 3, dir, kMazeBlue, kMazeWhite);
  }
}

struct RenderParams {
  const uint8_t* cells;     // Collect: grid plane [N_pad][W*H] index x*H+y; Maze: the shared field_map [S*S] index x*S+y
  const uint8_t* agents;    // Maze / CtF: agent words (x, y, dir, flags), one row of `agent_stride` bytes per env
  int agent_stride, family; // 0 Collect, 1 Maze, 2 CtF
  int n_agents, num_blue, variant_1v1;
  long long N;
  const int32_t* env_ids;   // [n] envs to draw, or null = envs 0..n-1
  int n, W, H, ts, JB;      // JB = rows of tiles one CTA draws
  unsigned per_row_magic, cpt_magic, w_magic;   // floor(2^32 / d) + 1 for d = 16-byte chunks per frame row, per tile row, and W: q = umulhi(n, magic) (n * d < 2^32)
  const uint8_t* atlas;     // [256][ts][ts][3]
  uint8_t* out;             // [n][H*ts][W*ts][3]
  int32_t* status;
};

constexpr int kRenderThreads = 256, kRenderMaxTiles = 1024, kRenderStage = 32768;

// One CTA draws `JB` consecutive rows of tiles of one frame (blockIdx.x = frame * ceil(H / JB) + block of rows): JB * ts frame
// rows, ONE contiguous block of the output.
// MODE 0: 3 * ts is a multiple of 16 and `out` is 16-byte aligned: 16-byte copies atlas -> frame.
// MODE 1: any tile size / alignment: the block is assembled in shared memory, one thread per (frame row, tile) copying the
//         tile row G bytes at a time (G = 8, 4 or 1 divides 3 * ts), shifted so that shared and global addresses agree
//         mod 16, and leaves as 16-byte stores (+ ragged ends).  JB > 1 when tiles are small, so a CTA moves ~32 KB.
// MODE 2: frame rows too wide to stage: byte copies.
template <int MODE, int G>
__global__ void __launch_bounds__(kRenderThreads) render_kernel(const __grid_constant__ RenderParams p) {
  __shared__ uint32_t s_tile[kRenderMaxTiles];   // byte offset of each tile's image inside the atlas, [jj][i]
  __shared__ __align__(16) uint8_t s_buf[MODE == 1 ? kRenderStage + 16 : 16];
  const int JB = p.JB, nblk = (p.H + JB - 1) / JB;
  const int f = blockIdx.x / nblk, j0 = (blockIdx.x - f * nblk) * JB, jn = min(JB, p.H - j0), tid = threadIdx.x;
  const int ts = p.ts, trow = ts * 3, tile_bytes = ts * trow, W = p.W;
  long long e = p.env_ids ? p.env_ids[f] : f;
  if (e < 0 || e >= p.N) { if (tid == 0) atomicOr(p.status, 4 /* MG_ERR_OOB */); e = 0; }
  for (int t = tid; t < jn * W; t += kRenderThreads) {
    const int jj = t / W, i = t - jj * W, j = j0 + jj;
    int code;
    if (p.family == 0) {
      code = p.cells[e * (long long)(W * p.H) + i * p.H + j];
      if ((code & 3) == 2) code &= 0x3F;   // a ball's bit 6 is the step's internal "respawned" mark, not part of its appearance
    }
    else {
      code = p.cells[i * p.H + j];
      const uint8_t* a = p.agents + e * p.agent_stride;   // the agent object replaces the cell it stands on (agent.py:195-196)
      if (p.family == 1) { if (a[0] == i && a[1] == j) code = 4 + (a[2] & 3); }
      else
        for (int k = 0; k < p.n_agents; ++k, a += 4)
          if (a[0] == i && a[1] == j) {
            const int team = k >= p.num_blue, bgs = (a[3] >> 2) & 3;
            const int bg_red = bgs ? bgs == 2 : team;            // 0 = as constructed: the team's own light colour
            const int grey = (a[3] & 1) && !p.variant_1v1;       // the 1v1 env never recolours (ctf.py:551-654)
            code = 8 + ((team * 2 + grey) * 2 + bg_red) * 4 + (a[2] & 3);
          }
    }
    s_tile[t] = (uint32_t)code * (uint32_t)tile_bytes;
  }
  __syncthreads();
  const int frame_row = W * trow, nrows = jn * ts;
  uint8_t* dst0 = p.out + ((size_t)f * p.H + j0) * ts * (size_t)frame_row;
  if (MODE == 0) {
    const int cpt = trow / 16, per_row = W * cpt, total = ts * per_row;
    for (int jj = 0; jj < jn; ++jj) {
      const uint32_t* tile = s_tile + jj * W;
      uint8_t* d = dst0 + (size_t)jj * ts * frame_row;
      for (int c = tid; c < total; c += kRenderThreads) {
        const int y = (int)__umulhi((unsigned)c, p.per_row_magic), r = c - y * per_row, i = (int)__umulhi((unsigned)r, p.cpt_magic), k = r - i * cpt;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.atlas + tile[i] + y * trow) + k);
        reinterpret_cast<uint4*>(d + (size_t)y * frame_row)[r] = v;
      }
    }
  } else if (MODE == 1) {
    const int rows_pp = kRenderStage / frame_row, upt = trow / G;
    for (int y0 = 0; y0 < nrows; y0 += rows_pp) {
      const int rows = min(rows_pp, nrows - y0), bytes = rows * frame_row, total = rows * W;
      uint8_t* g0 = dst0 + (size_t)y0 * frame_row;
      const int shift = (int)(reinterpret_cast<uintptr_t>(g0) & 15);
      uint8_t* sb = s_buf + shift;
      const bool unit_ok = shift % G == 0;
      for (int c = tid; c < total; c += kRenderThreads) {
        const int yy = (int)__umulhi((unsigned)c, p.w_magic), i = c - yy * W, y = y0 + yy, jj = y / ts, ty = y - jj * ts;
        const uint8_t* src = p.atlas + s_tile[jj * W + i] + ty * trow;
        uint8_t* d = sb + yy * frame_row + i * trow;
        for (int k = 0; k < upt; ++k, src += G, d += G) {
          if (G == 8) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(src));
            if (unit_ok) *reinterpret_cast<uint2*>(d) = v;
            else {
#pragma unroll
              for (int b = 0; b < 4; ++b) { d[b] = (uint8_t)(v.x >> (8 * b)); d[4 + b] = (uint8_t)(v.y >> (8 * b)); }
            }
          } else if (G == 4) {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src));
            if (unit_ok) *reinterpret_cast<uint32_t*>(d) = v;
            else {
#pragma unroll
              for (int b = 0; b < 4; ++b) d[b] = (uint8_t)(v >> (8 * b));
            }
          } else {
            *d = __ldg(src);
          }
        }
      }
      __syncthreads();
      const int head = min(bytes, (16 - shift) & 15), nbody = (bytes - head) / 16, tail0 = head + nbody * 16;
      for (int c = tid; c < nbody; c += kRenderThreads)
        reinterpret_cast<uint4*>(g0 + head)[c] = *reinterpret_cast<const uint4*>(sb + head + 16 * c);
      if (tid < head) g0[tid] = sb[tid];
      if (tid < bytes - tail0) g0[tail0 + tid] = sb[tail0 + tid];
      __syncthreads();
    }
  } else {
    const int total = nrows * frame_row;
    for (int c = tid; c < total; c += kRenderThreads) {
      const int y = c / frame_row, r = c - y * frame_row, i = r / trow, k = r - i * trow, jj = y / ts, ty = y - jj * ts;
      dst0[c] = __ldg(p.atlas + s_tile[jj * W + i] + ty * trow + k);
    }
  }
}

cudaError_t launch_render(const uint8_t* cells, const uint8_t* agents, int agent_stride, int family, int n_agents, int num_blue, int variant_1v1,
                          long long N, const int32_t* env_ids, int n, int W, int H, int ts, const uint8_t* atlas, uint8_t* out, int32_t* status,
                          cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (W > kRenderMaxTiles) return cudaErrorInvalidValue;
  RenderParams p;
  p.cells = cells; p.agents = agents; p.agent_stride = agent_stride; p.family = family; p.N = N; p.env_ids = env_ids;
  p.n_agents = n_agents; p.num_blue = num_blue; p.variant_1v1 = variant_1v1;
  p.n = n; p.W = W; p.H = H; p.ts = ts; p.atlas = atlas; p.out = out; p.status = status;
  const int trow = ts * 3;
  const long long strip = (long long)ts * W * trow;            // bytes of one row of tiles
  int JB = (int)(kRenderStage / strip);                        // rows of tiles per CTA: ~32 KB of output, at least one
  if (JB < 1) JB = 1;
  if (JB > H) JB = H;
  if (JB * W > kRenderMaxTiles) JB = kRenderMaxTiles / W;
  p.JB = JB;
  auto magic = [](unsigned d) { return (unsigned)(4294967296ull / d) + 1u; };
  p.per_row_magic = trow % 16 == 0 ? magic((unsigned)(W * (trow / 16))) : 0u; p.cpt_magic = trow % 16 == 0 ? magic((unsigned)(trow / 16)) : 0u;
  p.w_magic = magic((unsigned)W);
  const unsigned blocks = (unsigned)n * (unsigned)((H + JB - 1) / JB);
  if (trow % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) render_kernel<0, 1><<<blocks, kRenderThreads, 0, st>>>(p);
  else if (W * trow > kRenderStage) render_kernel<2, 1><<<blocks, kRenderThreads, 0, st>>>(p);
  else if (trow % 8 == 0) render_kernel<1, 8><<<blocks, kRenderThreads, 0, st>>>(p);
  else if (trow % 4 == 0) render_kernel<1, 4><<<blocks, kRenderThreads, 0, st>>>(p);
  else render_kernel<1, 1><<<blocks, kRenderThreads, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace mg
