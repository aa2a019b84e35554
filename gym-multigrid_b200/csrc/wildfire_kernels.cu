// wildfire_kernels.cu -- the Wildfire EXTENSION (no reference code exists; the rules are specified in
// include/multigrid_b200.h and restated by oracle/mg_oracle_wildfire.c).
//
// One CTA per env.  The env's terrain (one byte per cell) is staged in shared memory by a TMA bulk copy;
// warp 0 resolves the agents' moves in the step's random order with lane = agent: the acting lane broadcasts
// its target cell (__shfl_sync) and every other lane votes whether it stands there (__ballot_sync), which
// reproduces the sequential, order-dependent blocking without an occupancy grid; fire spread is a
// double-buffered 4-neighbour stencil over shared memory with one counter-based Philox block per group of
// four cells (skipped when no cell of the group can change); the observation is expanded in shared memory
// and leaves, like the new terrain, as one TMA bulk store.
#include <cstdlib>

#include "mg_device.cuh"
#include "wildfire_params.cuh"

namespace mg {

constexpr int kWfThreads = 256;
constexpr int WF_HEALTHY = 0, WF_BURNING = 1, WF_BURNT = 2;

__device__ __forceinline__ uint8_t wf_packed(int s) {  // healthy green, burning red, burnt grey (constants.py:8-19)
  return s == WF_HEALTHY ? cell(0, 3, 0) : (s == WF_BURNING ? cell(1, 0, 0) : cell(2, 7, 0));
}

struct WfSmem {
  uint8_t* told;   // [cells] terrain in / packed cells for the encode
  uint8_t* tnew;   // [cells]
  uint8_t* obs;    // [3*cells]
};

// all threads: terrain + agents -> packed cells (in `packed`) -> 3-byte encoding in s.obs
__device__ __forceinline__ void wf_encode(const WildfireParams& p, const uint8_t* terrain, uint8_t* packed, uint8_t* obs,
                                          const int* s_ax, const int* s_ay, const int* s_adir, int tid) {
  for (int i = tid; i < p.cells; i += kWfThreads) packed[i] = wf_packed(terrain[i]);
  __syncthreads();
  if (tid < p.A) packed[s_ax[tid] * p.H + s_ay[tid]] = (uint8_t)(cell(3, p.agent_colour[tid], 0) | (s_adir[tid] << 6));
  __syncthreads();
  const uint4* in = reinterpret_cast<const uint4*>(packed);
  uint4* out = reinterpret_cast<uint4*>(obs);
  for (int g = tid; g < p.cells / 16; g += kWfThreads) {
    uint4 a, b, c;
    expand16(in[g], a, b, c);
    out[3 * g] = a; out[3 * g + 1] = b; out[3 * g + 2] = c;
  }
}

// thread 0: reset of one env on shared memory (all healthy was written by all threads before)
__device__ __noinline__ void wf_reset_agents_fires(const WildfireParams& p, uint8_t* t, int* s_ax, int* s_ay, int* s_adir,
                                                   Rng<1>& r) {
  for (int f = 0; f < p.num_fires; ++f)
    for (;;) { const int i = (int)__umulhi(r.u32(), (uint32_t)p.cells); if (t[i] == WF_HEALTHY) { t[i] = WF_BURNING; break; } }
  for (int k = 0; k < p.A; ++k)
    for (;;) {
      const int i = (int)__umulhi(r.u32(), (uint32_t)p.cells);
      bool taken = false;
      for (int j = 0; j < k; ++j) taken |= (s_ax[j] * p.H + s_ay[j] == i);
      if (taken) continue;
      s_ax[k] = i / p.H; s_ay[k] = i % p.H; s_adir[k] = 3;
      break;
    }
}

__global__ void __launch_bounds__(kWfThreads) wildfire_kernel(const __grid_constant__ WildfireParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_ax[MG_MAX_WILDFIRE_AGENTS], s_ay[MG_MAX_WILDFIRE_AGENTS], s_adir[MG_MAX_WILDFIRE_AGENTS];
  __shared__ int s_order[MG_MAX_WILDFIRE_AGENTS], s_burning, s_flag;
  __shared__ int4 s_hdr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, A = p.A, W = p.W, H = p.H, cells = p.cells;
  const long long e = blockIdx.x;
  WfSmem s;
  s.told = smem_raw; s.tnew = smem_raw + cells; s.obs = smem_raw + 2 * (size_t)cells;
  uint8_t* g_terrain = p.terrain + e * cells;

  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); s_burning = 0; s_flag = 0; }
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();
  if (tid == 0) { mbar_expect_tx(&bar, (uint32_t)cells); tma_load_1d(s.told, g_terrain, (uint32_t)cells, &bar); }
  Rng<1> r;
  int4 h = make_int4(0, 0, 0, 0);
  if (tid < A) {
    const uchar4 a = reinterpret_cast<const uchar4*>(p.agents)[e * A + tid];
    s_ax[tid] = a.x; s_ay[tid] = a.y; s_adir[tid] = a.z;
  }
  if (tid == 0) {
    h = p.hdr[e];
    r.open_philox(p.seed, p.env_id_base + (unsigned long long)e, (uint32_t)h.z);
  }
  const bool do_reset = p.op == 0 && (!p.reset_mask || p.reset_mask[e]);
  mbar_wait(&bar, 0);
  __syncthreads();

  if (p.op == 1) {
    // ---- 1/2. ordered agent moves: warp 0, lane = agent
    if (tid == 0) {
      h.x += 1; h.y += 1;  // step_count, tick
      if (p.order) {
        for (int i = 0; i < A; ++i) s_order[i] = p.order[e * A + i];
      } else {  // Fisher-Yates over the env's Philox stream
        for (int i = 0; i < A; ++i) s_order[i] = i;
        for (int i = A - 1; i > 0; --i) { const int j = (int)__umulhi(r.u32(), (uint32_t)(i + 1)), t = s_order[i]; s_order[i] = s_order[j]; s_order[j] = t; }
      }
    }
    if (warp == 0) {
      __syncwarp();
      int x = lane < A ? s_ax[lane] : -1, y = lane < A ? s_ay[lane] : -1, dir = lane < A ? s_adir[lane] : 0;
      const int a = lane < A ? p.actions[e * A + lane] : 0;
      double rew = 0.0;
      for (int k = 0; k < A; ++k) {
        const int i = s_order[k];  // the acting agent; warp-uniform
        int dx = 0, dy = 0;
        if (lane == i && a >= 1 && a <= 4) { dx = (a == 4) - (a == 2); dy = (a == 3) - (a == 1); }
        const int nx = __shfl_sync(0xffffffffu, x + dx, i), ny = __shfl_sync(0xffffffffu, y + dy, i);
        const unsigned occupied = __ballot_sync(0xffffffffu, lane != i && lane < A && x == nx && y == ny);
        if (lane == i) {
          if ((dx | dy) && nx >= 0 && ny >= 0 && nx < W && ny < H && !occupied) {
            dir = dx == 1 ? 0 : (dy == 1 ? 1 : (dx == -1 ? 2 : 3));  // DIR_TO_VEC (constants.py:65-74)
            x = nx; y = ny;
          }
          if (s.told[x * H + y] == WF_BURNING) { s.told[x * H + y] = WF_BURNT; rew += 1.0; }  // extinguish
        }
        __syncwarp();
      }
      if (lane < A) { s_ax[lane] = x; s_ay[lane] = y; s_adir[lane] = dir; p.rewards[e * A + lane] = rew; }
    }
    __syncthreads();

    // ---- 3. fire dynamics: double-buffered 4-neighbour stencil, one Philox block per group of 4 cells
    const unsigned long long env_id = p.env_id_base + (unsigned long long)e;
    if (tid == 0) s_hdr = h;
    __syncthreads();
    const uint32_t tk = (uint32_t)s_hdr.y;
    int burning = 0;
    for (int g = tid; g < (cells + 3) / 4; g += kWfThreads) {
      int st[4], kk[4];
      bool active = false;
      for (int j = 0; j < 4; ++j) {
        const int i = 4 * g + j;
        st[j] = -1; kk[j] = 0;
        if (i >= cells) continue;
        const int x = i / H, y = i - x * H, sv = s.told[i];
        st[j] = sv;
        if (sv == WF_HEALTHY) {
          int k = 0;
          if (x > 0) k += s.told[i - H] == WF_BURNING;
          if (x < W - 1) k += s.told[i + H] == WF_BURNING;
          if (y > 0) k += s.told[i - 1] == WF_BURNING;
          if (y < H - 1) k += s.told[i + 1] == WF_BURNING;
          kk[j] = k;
          active |= k > 0;
        } else if (sv == WF_BURNING) {
          active = true;
        }
      }
      uint32_t u[4] = {0, 0, 0, 0};
      if (active) philox4x32_10((uint32_t)env_id, (uint32_t)(env_id >> 32), tk, 1u + (uint32_t)g, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), u);
      for (int j = 0; j < 4; ++j) {
        if (st[j] < 0) continue;
        int ns = st[j];
        if (st[j] == WF_BURNING) ns = u[j] < p.burnout_threshold ? WF_BURNT : WF_BURNING;
        else if (st[j] == WF_HEALTHY && kk[j] > 0) ns = u[j] < p.ignite_threshold[kk[j]] ? WF_BURNING : WF_HEALTHY;
        s.tnew[4 * g + j] = (uint8_t)ns;
        burning += ns == WF_BURNING;
      }
    }
    for (int o = 16; o > 0; o >>= 1) burning += __shfl_xor_sync(0xffffffffu, burning, o);
    if (lane == 0 && burning) atomicAdd(&s_burning, burning);
    __syncthreads();
    // ---- 4. termination, same-step autoreset
    const bool term = s_burning == 0, trunc = s_hdr.x >= p.max_steps;
    if (tid == 0) { p.terminated[e] = term; p.truncated[e] = trunc; }
    if (p.autoreset && (term || trunc)) {
      if (p.final_obs) {
        wf_encode(p, s.tnew, s.told, s.obs, s_ax, s_ay, s_adir, tid);
        __syncthreads();
        uint4* dst = reinterpret_cast<uint4*>(p.final_obs + e * 3 * cells);
        for (int i = tid; i < 3 * cells / 16; i += kWfThreads) dst[i] = reinterpret_cast<const uint4*>(s.obs)[i];
        __syncthreads();
      }
      for (int i = tid; i < cells; i += kWfThreads) s.tnew[i] = WF_HEALTHY;
      __syncthreads();
      if (tid == 0) { wf_reset_agents_fires(p, s.tnew, s_ax, s_ay, s_adir, r); h.x = 0; h.w += 1; }
      __syncthreads();
    }
  } else {
    // ---- reset(mask)
    for (int i = tid; i < cells; i += kWfThreads) s.tnew[i] = do_reset ? (uint8_t)WF_HEALTHY : s.told[i];
    __syncthreads();
    if (tid == 0 && do_reset) { wf_reset_agents_fires(p, s.tnew, s_ax, s_ay, s_adir, r); h.x = 0; h.w += 1; }
    __syncthreads();
  }

  // ---- 5. observation + write-back
  if (p.obs) wf_encode(p, s.tnew, s.told, s.obs, s_ax, s_ay, s_adir, tid);
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    tma_store_1d(g_terrain, s.tnew, (uint32_t)cells);
    if (p.obs) tma_store_1d(p.obs + e * 3 * cells, s.obs, (uint32_t)(3 * cells));
    tma_commit();
    h.z = (int)r.ctr;
    p.hdr[e] = h;
  }
  if (tid < A) reinterpret_cast<uchar4*>(p.agents)[e * A + tid] = make_uchar4((uint8_t)s_ax[tid], (uint8_t)s_ay[tid], (uint8_t)s_adir[tid], 0);
  if (tid == 0) tma_wait_read_all();
}

size_t wildfire_smem_bytes(int cells) { return (size_t)cells * 5 + 64; }

cudaError_t configure_wildfire_kernel(int cells) {
  return cudaFuncSetAttribute((const void*)wildfire_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wildfire_smem_bytes(cells));
}

cudaError_t launch_wildfire(const WildfireParams& p, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)p.N); cfg.blockDim = dim3(kWfThreads);
  cfg.dynamicSmemBytes = wildfire_smem_bytes(p.cells); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = [] { const char* v = std::getenv("MG_PDL"); return !(v && v[0] == '0'); }();
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, wildfire_kernel, p);
}

}  // namespace mg
