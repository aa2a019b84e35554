// view_kernels.cu -- partial (egocentric) observations: MultiGridEnv.gen_obs (multigrid.py:485-532)
//   = Agent.get_view_exts (agent.py:294-324) + Grid.slice (grid.py:111-130, out of bounds -> Wall)
//   + Grid.rotate_left x (dir + 1) (grid.py:97-109) + Grid.process_vis (grid.py:286-323)
//   + Grid.encode_for_agents (grid.py:254-284), encode_dim 3.
// The reference's gen_obs passes one positional argument too many to encode_for_agents
// (multigrid.py:526-528 vs grid.py:254) and raises; this follows the algorithm its pieces define.
//
// This path is instruction-issue bound, not HBM bound (147 output bytes per view need several hundred
// instructions), so the fast kernel (odd V <= 7; per-view body in view_device.cuh, shared with map_kernel's fused Maze
// mode) is written for instruction count: one thread per view with everything in registers (V is a template parameter,
// all loops unroll) -
//   * slice + rotations folded into one base index and two strides; cells are byte gathers from shared memory
//     (Collect: the tile's grid slab, staged by one TMA bulk load, with guard bands so that out-of-grid reads need
//     no predication; Maze: the static map pre-padded with the out-of-bounds filler, staged the same way);
//   * out-of-grid cells are selected by two V-bit range masks instead of per-cell bounds tests;
//   * process_vis as a bit-parallel (Kogge-Stone) flood over V-bit row masks: ~30 ALU ops per row;
//   * the masked codes are packed four per register in output order, expanded to (type, colour, state) bytes with
//     PRMT, re-aligned with one funnel shift per word and written to shared memory as 32-bit stores (the 147-byte
//     view pitch spreads a warp over all 32 banks); the tile's contiguous slab of views leaves as one TMA bulk store.
// Other view sizes (even V, V > 7) take the generic kernel below (runtime V, per-cell loops).  view6_kernel is the
// encode_dim-6 (DefaultWorld) variant for the generic family; toroid_fast_kernel the ToroidObservation wrapper.
#include <cstdlib>

#include "mg_device.cuh"
#include "smem_config.h"
#include "view_device.cuh"
#include "view_params.cuh"

namespace mg {

constexpr int kViewE = 64;        // envs per CTA (generic kernel; fast kernel, Collect)
constexpr int kViewThreads = 128;
constexpr int kViewMax = 15;      // largest view size
constexpr int kViewMazeE = 128;   // envs per CTA (fast kernel, Maze: one view per env)

template <int FAMILY, int V>
__global__ void __launch_bounds__(kViewThreads) view_fast_kernel(const __grid_constant__ ViewParams p) {
  static_assert(V % 2 == 1 && V <= 7, "fast path: odd view sizes up to 7");
  constexpr int VV = V * V;
  constexpr int E = FAMILY == MG_FAMILY_COLLECT ? kViewE : kViewMazeE;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, A = p.A;
  const long long e0 = (long long)blockIdx.x * E;
  const int n_here = (int)min((long long)E, p.N - e0);
  const int views = n_here * A;
  const uint32_t out_bytes_tile = (uint32_t)E * A * VV * 3;
  uint8_t* s_out = smem_raw;                                              // [E*A][VV*3]
  uint8_t* s_src = smem_raw + ((out_bytes_tile + 15u) & ~15u);            // Collect: guard | [E][cells] | guard; Maze: padded map
  const int guard = FAMILY == MG_FAMILY_COLLECT ? view_guard_bytes(V, p.H) : 0;

  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();
  if (tid == 0) {
    if (FAMILY == MG_FAMILY_COLLECT) {  // the grid plane is padded to whole tiles of >= 64 envs
      mbar_expect_tx(&bar, (uint32_t)E * p.cells);
      tma_load_1d(s_src + guard, p.grid + e0 * p.cells, (uint32_t)E * p.cells, &bar);
    } else {
      mbar_expect_tx(&bar, (uint32_t)p.map_padded_bytes);
      tma_load_1d(s_src, p.map_padded, (uint32_t)p.map_padded_bytes, &bar);
    }
  }
  // the first view's position and direction travel while the source tile loads (a tile is normally one view per thread)
  int x_n = 0, y_n = 0, dir_n = 3;
  auto fetch = [&](int v) {
    const long long gv = e0 * A + v;
    x_n = p.pos[gv * p.pos_stride]; y_n = p.pos[gv * p.pos_stride + 1];
    dir_n = p.dirs ? (p.dirs[gv * p.dir_stride] & 3) : 3;
  };
  if (tid < views) fetch(tid);
  mbar_wait(&bar, 0);

  const int pitch = FAMILY == MG_FAMILY_COLLECT ? p.H : p.pitch;
  for (int v = tid; v < views; v += kViewThreads) {
    const int x = x_n, y = y_n, dir = dir_n;
    if (v + kViewThreads < views) fetch(v + kViewThreads);
    int x0, y0, sa, sb;
    view_geometry<V>(x, y, dir, pitch, x0, y0, sa, sb);
    const bool a_is_x = dir & 1;  // dirs 1, 3: a walks along x, b along y
    const uint32_t mA = a_is_x ? range_mask(x0, dir == 3 ? 1 : -1, p.W, V) : range_mask(y0, dir == 0 ? 1 : -1, p.H, V);
    const uint32_t mB = a_is_x ? range_mask(y0, dir == 3 ? 1 : -1, p.H, V) : range_mask(x0, dir == 2 ? 1 : -1, p.W, V);
    const uint8_t* src;
    if (FAMILY == MG_FAMILY_COLLECT) src = s_src + guard + (v / A) * p.cells + x0 * pitch + y0;
    else src = s_src + (x0 + p.pad) * pitch + (y0 + p.pad);
    view_compute_store<FAMILY == MG_FAMILY_COLLECT, V>(src, sa, sb, mA, mB, p.oob_code, (uint32_t)p.agent_code | ((uint32_t)dir << 6),
                                                       p.see_through != 0, s_out, v);
  }
  fence_proxy_async_smem();
  __syncthreads();
  const uint32_t bytes = (uint32_t)views * VV * 3;
  const uint32_t bulk = p.out_bulk_ok ? (bytes & ~15u) : 0u;
  uint8_t* dst = p.out + e0 * A * VV * 3;
  if (tid == 0 && bulk) { tma_store_1d(dst, s_out, bulk); tma_commit(); }
  for (uint32_t i = bulk + tid; i < bytes; i += kViewThreads) dst[i] = s_out[i];
  if (tid == 0) tma_wait_read_all();
}

template <int FAMILY>
__device__ __forceinline__ uint8_t fetch_cell(const ViewParams& p, const uint8_t* g, int ax, int ay, int adir, int x, int y) {
  if (x < 0 || y < 0 || x >= p.W || y >= p.H) return p.oob_code;
  if (FAMILY == MG_FAMILY_COLLECT) return g[x * p.H + y];
  // Maze: the static map with the (single) agent drawn on top (maze.py:180-205)
  if (x == ax && y == ay) return (uint8_t)(p.agent_code | (adir << 6));
  return __ldg(p.map_codes + x * p.H + y);
}

template <int FAMILY>
__device__ __forceinline__ bool opaque(const ViewParams& p, uint8_t c) {
  // see_behind() is False only for Wall in CollectWorld (object.py:174-179); MazeWorld obstacles are built with
  // can_see_through=True (object.py:185-199), so inside a Maze only the out-of-bounds filler blocks sight
  return FAMILY == MG_FAMILY_COLLECT ? (c & 3) == T_WALL : c == p.oob_code;
}

template <int FAMILY>
__global__ void __launch_bounds__(kViewThreads) view_kernel(const __grid_constant__ ViewParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, A = p.A, V = p.V, VV = V * V, cells = p.cells;
  const long long e0 = (long long)blockIdx.x * kViewE;
  const int n_here = (int)min((long long)kViewE, p.N - e0);
  const int views = n_here * A;
  uint8_t* s_out = smem_raw;                                            // [kViewE*A][VV*3]
  uint8_t* s_code = s_out + (size_t)kViewE * A * VV * 3;                // [kViewThreads][VV] scratch
  uint8_t* s_grid = s_code + (size_t)kViewThreads * VV;                 // Collect: [kViewE][cells]
  s_grid = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_grid) + 15) & ~uintptr_t(15));

  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();
  if (FAMILY == MG_FAMILY_COLLECT) {
    if (tid == 0) {  // the grid plane is padded to whole tiles of >= 64 envs
      mbar_expect_tx(&bar, (uint32_t)kViewE * cells);
      tma_load_1d(s_grid, p.grid + e0 * cells, (uint32_t)kViewE * cells, &bar);
    }
    mbar_wait(&bar, 0);
  }

  uint8_t* code = s_code + (size_t)tid * VV;
  for (int v = tid; v < views; v += kViewThreads) {
    const int el = v / A, k = v - el * A;
    const long long e = e0 + el;
    const uint8_t* g = s_grid + (size_t)el * cells;
    const int x = p.pos[(e * A + k) * p.pos_stride], y = p.pos[(e * A + k) * p.pos_stride + 1];
    const int dir = p.dirs ? (p.dirs[(e * A + k) * p.dir_stride] & 3) : 3;
    const int hs = V / 2;
    uint32_t opq[kViewMax + 1], msk[kViewMax + 1];
    for (int b = 0; b < V; ++b) {
      uint32_t o = 0;
      for (int a = 0; a < V; ++a) {
        int wx, wy;  // slice + (dir + 1) rotate_left, as direct indices
        if (dir == 0)      { wx = x + V - 1 - b;      wy = y - hs + a; }          // facing right
        else if (dir == 1) { wx = x - hs + V - 1 - a; wy = y + V - 1 - b; }       // facing down
        else if (dir == 2) { wx = x - V + 1 + b;      wy = y - hs + V - 1 - a; }  // facing left
        else               { wx = x - hs + a;         wy = y - V + 1 + b; }       // facing up
        const uint8_t c = fetch_cell<FAMILY>(p, g, x, y, dir, wx, wy);
        code[a * V + b] = c;
        o |= (uint32_t)opaque<FAMILY>(p, c) << a;
      }
      opq[b] = o; msk[b] = 0;
    }
    if (p.see_through) {
      for (int b = 0; b < V; ++b) msk[b] = (1u << V) - 1;
    } else {  // process_vis (grid.py:286-323): rows bottom-up, each row left->right then right->left
      msk[V - 1] = 1u << hs;
      for (int j = V - 1; j >= 0; --j) {
        uint32_t m = msk[j], up = 0;
        const uint32_t clear = ~opq[j];
        for (int i = 0; i < V - 1; ++i)
          if ((m >> i) & (clear >> i) & 1u) { m |= 1u << (i + 1); up |= 3u << i; }
        for (int i = V - 1; i >= 1; --i)
          if ((m >> i) & (clear >> i) & 1u) { m |= 1u << (i - 1); up |= 3u << (i - 1); }
        msk[j] = m;
        if (j > 0) msk[j - 1] |= up;
      }
    }
    uint8_t* o = s_out + (size_t)v * VV * 3;
    for (int a = 0; a < V; ++a)
      for (int b = 0; b < V; ++b) {  // encode_for_agents: cells outside the mask stay (0, 0, 0)
        const uint8_t c = ((msk[b] >> a) & 1u) ? code[a * V + b] : 0;
        o[(a * V + b) * 3] = c & 3; o[(a * V + b) * 3 + 1] = (c >> 2) & 15; o[(a * V + b) * 3 + 2] = state_of(c);
      }
  }
  fence_proxy_async_smem();
  __syncthreads();
  const uint32_t bytes = (uint32_t)views * VV * 3;
  const uint32_t bulk = p.out_bulk_ok ? (bytes & ~15u) : 0u;
  uint8_t* dst = p.out + e0 * A * VV * 3;
  if (tid == 0 && bulk) { tma_store_1d(dst, s_out, bulk); tma_commit(); }
  for (uint32_t i = bulk + tid; i < bytes; i += kViewThreads) dst[i] = s_out[i];
  if (tid == 0) tma_wait_read_all();
}

// ToroidObservation.observation (wrappers/toroid.py:28-68): agent-centred, wrap-around one-hot planes,
// float32 [N][A][W][W][depth], depth = num_ball_types + num_agents; written at [y'][x'] like the reference.
// One thread per output cell (gathers its source cell), so every warp writes one contiguous run of floats.
__global__ void __launch_bounds__(256) toroid_kernel(const uint8_t* __restrict__ grid, const uint8_t* __restrict__ pos,
                                                     float* __restrict__ out, long long N, int W, int A, int nb) {
  pdl_launch_dependents();
  pdl_wait();
  const int cells = W * W, depth = nb + A;
  const long long total = N * A * cells;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long v = idx / cells;             // view index e*A + k
    const int c = (int)(idx - v * cells);
    const int ny = c / W, nx = c - ny * W;       // tor[new_coords[1], new_coords[0], ...]  toroid.py:58-66
    const long long e = v / A;
    const int px = pos[v * 2], py = pos[v * 2 + 1];
    int i = nx + px, j = ny + py;                // inverse of new = (i - px, j - py) wrapped into [0, W)
    if (i >= W) i -= W;
    if (j >= W) j -= W;
    const uint8_t code = grid[e * cells + i * W + j];
    const int type = code & 3;
    int ch = -1;
    if (type == T_WALL) ch = depth - 1;
    else if (type == T_BALL) ch = (code >> 2) & 15;
    else if (type == T_AGENT && !(i == px && j == py)) ch = depth - 2;  // another agent, not on this agent's cell
    float* o = out + idx * depth;
    for (int d = 0; d < depth; ++d) o[d] = (d == ch) ? 1.0f : 0.0f;
  }
}

// Fast toroid kernel: DEPTH is a template parameter, one CTA owns a tile of 32 envs whose grids and agent positions are
// staged in shared memory.  Phase 1: one thread per output cell gathers its (wrapped) source cell and records the cell's
// one-hot channel as a byte.  Phase 2: the tile's output - one contiguous run of floats - is written as fully coalesced
// 16-byte stores (a warp instruction covers 512 contiguous bytes; strided or scalar stores of the 20-byte cell records
// reach only ~60% of the write bandwidth).  Index arithmetic is 32-bit with host-computed multiply-shift reciprocals.
// Write-bound: the input is 1/40 of the output.
constexpr int kTorE = 32, kTorThreads = 256;

template <int DEPTH>
__global__ void __launch_bounds__(kTorThreads) toroid_fast_kernel(const uint8_t* __restrict__ grid, const uint8_t* __restrict__ pos,
                                                                  float* __restrict__ out, long long N, int W, int A,
                                                                  uint32_t cells_magic, uint32_t w_magic, uint32_t a_magic) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, cells = W * W;
  const long long e0 = (long long)blockIdx.x * kTorE;
  const int n_here = (int)min((long long)kTorE, N - e0);
  uint8_t* s_grid = smem_raw;                                   // [kTorE][cells]
  uint8_t* s_pos = s_grid + (size_t)kTorE * cells;              // [kTorE][A][2]
  int8_t* s_ch = reinterpret_cast<int8_t*>(s_pos + (size_t)kTorE * A * 2);   // [kTorE * A * cells] one-hot channel per output cell, -1 = none
  pdl_launch_dependents();
  pdl_wait();
  {  // e0 * cells is a multiple of 32, so the tile starts on a 16-byte boundary whenever the plane does
    const int bytes = n_here * cells;
    const uint8_t* src = grid + e0 * cells;
    for (int i = tid * 16; i + 16 <= bytes; i += kTorThreads * 16) *reinterpret_cast<uint4*>(s_grid + i) = *reinterpret_cast<const uint4*>(src + i);
    for (int i = (bytes & ~15) + tid; i < bytes; i += kTorThreads) s_grid[i] = src[i];
    for (int i = tid; i < n_here * A * 2; i += kTorThreads) s_pos[i] = pos[e0 * A * 2 + i];
  }
  __syncthreads();
  const int total = n_here * A * cells;                         // output cells of this tile
  for (int lc = tid; lc < total; lc += kTorThreads) {
    const int v = (int)__umulhi((uint32_t)lc, cells_magic);     // view of the tile = el * A + k
    const int c = lc - v * cells;
    const int ny = (int)__umulhi((uint32_t)c, w_magic), nx = c - ny * W;   // tor[new_coords[1], new_coords[0], ...]  toroid.py:58-66
    const int el = A == 1 ? v : (int)__umulhi((uint32_t)v, a_magic);
    const int px = s_pos[2 * v], py = s_pos[2 * v + 1];
    int i = nx + px, j = ny + py;                               // inverse of new = (i - px, j - py) wrapped into [0, W)
    if (i >= W) i -= W;
    if (j >= W) j -= W;
    const uint32_t code = s_grid[el * cells + i * W + j];
    const uint32_t type = code & 3u;
    int ch = -1;
    if (type == T_WALL) ch = DEPTH - 1;
    else if (type == T_BALL) ch = (int)((code >> 2) & 15u);
    else if (type == T_AGENT && !(i == px && j == py)) ch = DEPTH - 2;     // another agent, not on this agent's cell
    s_ch[lc] = (int8_t)ch;
  }
  __syncthreads();
  const int nfloat = total * DEPTH;
  float* o = out + e0 * A * cells * DEPTH;
  float4* o4 = reinterpret_cast<float4*>(o);
  for (int q = tid; 4 * q + 4 <= nfloat; q += kTorThreads) {
    float f[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int idx = 4 * q + t, cell = idx / DEPTH, d = idx - cell * DEPTH;
      f[t] = (int)s_ch[cell] == d ? 1.0f : 0.0f;
    }
    o4[q] = make_float4(f[0], f[1], f[2], f[3]);
  }
  for (int idx = (nfloat & ~3) + tid; idx < nfloat; idx += kTorThreads) {   // ragged end of the last tile
    const int cell = idx / DEPTH;
    o[idx] = (int)s_ch[cell] == idx - cell * DEPTH ? 1.0f : 0.0f;
  }
}

template <int DEPTH>
static cudaError_t launch_toroid_fast(const uint8_t* grid, const uint8_t* pos, float* out, long long N, int W, int A, cudaStream_t st) {
  const int cells = W * W;
  const size_t smem = (size_t)kTorE * cells * (1 + A) + (size_t)kTorE * A * 2 + 16;
  cudaError_t e = raise_smem_limit((const void*)toroid_fast_kernel<DEPTH>, smem);
  if (e != cudaSuccess) return e;
  auto magic = [](int d) { return d <= 1 ? 0u : (uint32_t)(4294967296ull / (unsigned)d) + 1u; };  // exact quotients for operands < 2^16
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((N + kTorE - 1) / kTorE)); cfg.blockDim = dim3(kTorThreads);
  cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = [] { const char* v = std::getenv("MG_PDL"); return !(v && v[0] == '0'); }();
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, toroid_fast_kernel<DEPTH>, grid, pos, out, N, W, A, magic(cells), magic(W), magic(A));
}

cudaError_t launch_toroid(const uint8_t* grid, const uint8_t* pos, float* out, long long N, int W, int A, int nb, cudaStream_t st) {
  static const bool generic = [] { const char* v = std::getenv("MG_TOROID_GENERIC"); return v && v[0] == '1'; }();
  const int depth = nb + A, cells = W * W;
  // fast path: tile indices stay below 2^16 (multiply-shift reciprocals), the output base is 16-byte aligned
  if (!generic && (long long)kTorE * A * cells < 65536 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0 &&
      (reinterpret_cast<uintptr_t>(grid) & 15u) == 0 && (size_t)kTorE * cells * (1 + A) + (size_t)kTorE * A * 2 + 16 <= 200 * 1024) {
    switch (depth) {
      case 2: return launch_toroid_fast<2>(grid, pos, out, N, W, A, st);
      case 3: return launch_toroid_fast<3>(grid, pos, out, N, W, A, st);
      case 4: return launch_toroid_fast<4>(grid, pos, out, N, W, A, st);
      case 5: return launch_toroid_fast<5>(grid, pos, out, N, W, A, st);
      case 6: return launch_toroid_fast<6>(grid, pos, out, N, W, A, st);
      case 7: return launch_toroid_fast<7>(grid, pos, out, N, W, A, st);
      case 8: return launch_toroid_fast<8>(grid, pos, out, N, W, A, st);
      default: break;
    }
  }
  const long long total = N * A * W * W;
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  toroid_kernel<<<blocks, 256, 0, st>>>(grid, pos, out, N, W, A, nb);
  return cudaGetLastError();
}

// MultiGridEnv.gen_obs with DefaultWorld (encode_dim 6; generic family): one thread per view, runtime V.  Wall and a
// closed / locked Door block sight (object.py:178-179, 223-224), out-of-grid cells are grey walls (grid.py:124-127), the
// agent's own cell (view cell (V/2, V-1)) carries the is_self byte (grid.py:279-281).  No shipped env reaches this path.
__global__ void __launch_bounds__(128) view6_kernel(const __grid_constant__ View6Params p) {
  constexpr int G_WALL = 2, G_DOOR = 4, G_AGENT = 10;   // DefaultWorld.OBJECT_TO_IDX (world.py:37-51)
  pdl_launch_dependents();
  pdl_wait();
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= p.N * p.A) return;
  const int V = p.V, hs = V / 2, W = p.W, H = p.H;
  const long long e = v / p.A;
  const uint8_t* gc = p.gcell + e * p.cells;
  const uint8_t* gs = p.gstate + e * p.cells;
  const int x = p.pos[v * 2], y = p.pos[v * 2 + 1];
  const int dir = p.dirs ? (p.dirs[v] & 3) : (gs[x * H + y] & 3);
  // world cell of view cell (a, b): slice + (dir + 1) x rotate_left folded into direct indices (same table as view_kernel)
  auto world = [&](int a, int b, int& wx, int& wy) {
    if (dir == 0)      { wx = x + V - 1 - b;      wy = y - hs + a; }
    else if (dir == 1) { wx = x - hs + V - 1 - a; wy = y + V - 1 - b; }
    else if (dir == 2) { wx = x - V + 1 + b;      wy = y - hs + V - 1 - a; }
    else               { wx = x - hs + a;         wy = y - V + 1 + b; }
  };
  uint32_t msk[kViewMax + 1];
  const uint32_t FULL = (1u << V) - 1u;
  if (p.see_through) {
    for (int b = 0; b < V; ++b) msk[b] = FULL;
  } else {  // process_vis (grid.py:286-323) with the bit-parallel row flood of view_fast_kernel
    for (int b = 0; b < V; ++b) msk[b] = 0;
    msk[V - 1] = 1u << hs;
    for (int j = V - 1; j >= 0; --j) {
      uint32_t opq = 0;
      for (int a = 0; a < V; ++a) {
        int wx, wy;
        world(a, j, wx, wy);
        bool o = true;   // outside the grid: Wall
        if (wx >= 0 && wy >= 0 && wx < W && wy < H) {
          const int type = gc[wx * H + wy] & 15;
          o = type == G_WALL || (type == G_DOOR && gs[wx * H + wy] != 0);
        }
        opq |= (uint32_t)o << a;
      }
      const uint32_t clear = ~opq & FULL;
      uint32_t m = msk[j];
      uint32_t F = m & clear, P = clear;
      F |= P & (F << 1); P &= P << 1;
      F |= P & (F << 2); P &= P << 2;
      F |= P & (F << 4); P &= P << 4;
      F |= P & (F << 8);
      F &= FULL >> 1;
      m |= F << 1;
      uint32_t up = F | (F << 1);
      uint32_t G = m & clear; P = clear;
      G |= P & (G >> 1); P &= P >> 1;
      G |= P & (G >> 2); P &= P >> 2;
      G |= P & (G >> 4); P &= P >> 4;
      G |= P & (G >> 8);
      G &= ~1u;
      m |= G >> 1;
      up |= G | (G >> 1);
      msk[j] = m;
      if (j > 0) msk[j - 1] |= up;
    }
  }
  uint16_t* o = reinterpret_cast<uint16_t*>(p.out + v * (long long)V * V * 6);
  for (int a = 0; a < V; ++a)
    for (int b = 0; b < V; ++b) {
      uint16_t w0 = 0, w1 = 0, w2 = 0;   // unseen
      if ((msk[b] >> a) & 1u) {
        int wx, wy;
        world(a, b, wx, wy);
        uint32_t c = G_WALL | (7u << 4), st = 0;
        if (wx >= 0 && wy >= 0 && wx < W && wy < H) { c = gc[wx * H + wy]; st = gs[wx * H + wy]; }
        const uint32_t type = c & 15u;
        w0 = (uint16_t)(type | ((c >> 4) << 8));
        if (type == G_DOOR) w1 = (uint16_t)st;
        else if (type == G_AGENT) w2 = (uint16_t)((st & 3u) | ((a == hs && b == V - 1) ? 0x100u : 0u));
      }
      o[(a * V + b) * 3] = w0; o[(a * V + b) * 3 + 1] = w1; o[(a * V + b) * 3 + 2] = w2;
    }
}

cudaError_t launch_view6(const View6Params& p, cudaStream_t st) {
  const long long views = p.N * p.A;
  view6_kernel<<<(unsigned)((views + 127) / 128), 128, 0, st>>>(p);
  return cudaGetLastError();
}

static bool view_is_fast(const ViewParams& p) {
  static const bool off = [] { const char* v = std::getenv("MG_VIEW_GENERIC"); return v && v[0] == '1'; }();
  if (off || !(p.V == 3 || p.V == 5 || p.V == 7)) return false;
  return p.family == MG_FAMILY_COLLECT || p.map_padded != nullptr;
}

static size_t view_fast_smem(const ViewParams& p) {
  const size_t E = p.family == MG_FAMILY_COLLECT ? kViewE : kViewMazeE;
  const size_t out = (E * p.A * p.V * p.V * 3 + 15) / 16 * 16;
  return out + (p.family == MG_FAMILY_COLLECT ? E * p.cells + 2 * (size_t)view_guard_bytes(p.V, p.H) : (size_t)p.map_padded_bytes);
}

size_t view_smem_bytes(const ViewParams& p) {
  if (view_is_fast(p)) return view_fast_smem(p);
  return (size_t)kViewE * p.A * p.V * p.V * 3 + (size_t)kViewThreads * p.V * p.V + 16 +
         (p.family == MG_FAMILY_COLLECT ? (size_t)kViewE * p.cells : 0);
}
int view_max() { return kViewMax; }
int view_tile_envs() { return kViewE; }

template <int FAMILY, int V>
static cudaError_t launch_fast(const ViewParams& p, size_t smem, cudaStream_t st) {
  cudaError_t e = raise_smem_limit((const void*)view_fast_kernel<FAMILY, V>, smem);   // raised by the first (eager) call
  if (e != cudaSuccess) return e;
  constexpr int E = FAMILY == MG_FAMILY_COLLECT ? kViewE : kViewMazeE;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((p.N + E - 1) / E)); cfg.blockDim = dim3(kViewThreads);
  cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = [] { const char* v = std::getenv("MG_PDL"); return !(v && v[0] == '0'); }();
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, view_fast_kernel<FAMILY, V>, p);
}

cudaError_t launch_view(const ViewParams& p, cudaStream_t st) {
  const size_t smem = view_smem_bytes(p);
  if (view_is_fast(p)) {
    const bool c = p.family == MG_FAMILY_COLLECT;
    switch (p.V) {
      case 3: return c ? launch_fast<MG_FAMILY_COLLECT, 3>(p, smem, st) : launch_fast<MG_FAMILY_MAZE, 3>(p, smem, st);
      case 5: return c ? launch_fast<MG_FAMILY_COLLECT, 5>(p, smem, st) : launch_fast<MG_FAMILY_MAZE, 5>(p, smem, st);
      default: return c ? launch_fast<MG_FAMILY_COLLECT, 7>(p, smem, st) : launch_fast<MG_FAMILY_MAZE, 7>(p, smem, st);
    }
  }
  cudaError_t e = raise_smem_limit((const void*)view_kernel<MG_FAMILY_COLLECT>, smem);
  if (e == cudaSuccess) e = raise_smem_limit((const void*)view_kernel<MG_FAMILY_MAZE>, smem);
  if (e != cudaSuccess) return e;
  const unsigned blocks = (unsigned)((p.N + kViewE - 1) / kViewE);
  if (p.family == MG_FAMILY_COLLECT) view_kernel<MG_FAMILY_COLLECT><<<blocks, kViewThreads, smem, st>>>(p);
  else view_kernel<MG_FAMILY_MAZE><<<blocks, kViewThreads, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace mg
