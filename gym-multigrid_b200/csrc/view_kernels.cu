// view_kernels.cu -- partial (egocentric) observations: MultiGridEnv.gen_obs (multigrid.py:485-532)
//   = Agent.get_view_exts (agent.py:294-324) + Grid.slice (grid.py:111-130, out of bounds -> Wall)
//   + Grid.rotate_left x (dir + 1) (grid.py:97-109) + Grid.process_vis (grid.py:286-323)
//   + Grid.encode_for_agents (grid.py:254-284), encode_dim 3.
// The reference's gen_obs passes one positional argument too many to encode_for_agents
// (multigrid.py:526-528 vs grid.py:254) and raises; this follows the algorithm its pieces define.
//
// One thread per (env, agent) view.  slice + rotations are folded into direct world indices; the
// visibility flood is a row sweep over V-bit masks; views are assembled in shared memory and the tile's
// contiguous slab of views leaves with one TMA bulk store.
#include "mg_device.cuh"
#include "view_params.cuh"

namespace mg {

constexpr int kViewE = 64;        // envs per CTA
constexpr int kViewThreads = 128;
constexpr int kViewMax = 15;      // largest view size


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
    const int x = p.pos[(e * A + k) * 2], y = p.pos[(e * A + k) * 2 + 1];
    const int dir = p.dirs ? p.dirs[e * A + k] : 3;
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
        o[(a * V + b) * 3] = c & 3; o[(a * V + b) * 3 + 1] = (c >> 2) & 15; o[(a * V + b) * 3 + 2] = c >> 6;
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

cudaError_t launch_toroid(const uint8_t* grid, const uint8_t* pos, float* out, long long N, int W, int A, int nb, cudaStream_t st) {
  const long long total = N * A * W * W;
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  toroid_kernel<<<blocks, 256, 0, st>>>(grid, pos, out, N, W, A, nb);
  return cudaGetLastError();
}

size_t view_smem_bytes(int family, int cells, int A, int V) {
  return (size_t)kViewE * A * V * V * 3 + (size_t)kViewThreads * V * V + 16 +
         (family == MG_FAMILY_COLLECT ? (size_t)kViewE * cells : 0);
}
int view_max() { return kViewMax; }
int view_tile_envs() { return kViewE; }

// opt the kernels in to `bytes` of dynamic shared memory (call outside stream capture, before the first launch)
cudaError_t configure_view_kernels(size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute((const void*)view_kernel<MG_FAMILY_COLLECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute((const void*)view_kernel<MG_FAMILY_MAZE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t launch_view(const ViewParams& p, cudaStream_t st) {
  const size_t smem = view_smem_bytes(p.family, p.cells, p.A, p.V);
  const unsigned blocks = (unsigned)((p.N + kViewE - 1) / kViewE);
  if (p.family == MG_FAMILY_COLLECT) view_kernel<MG_FAMILY_COLLECT><<<blocks, kViewThreads, smem, st>>>(p);
  else view_kernel<MG_FAMILY_MAZE><<<blocks, kViewThreads, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace mg
