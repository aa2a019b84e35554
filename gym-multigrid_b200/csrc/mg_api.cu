// mg_api.cu -- implementation of the C ABI declared in include/multigrid_b200.h.
// Host-side only: validates configs, lays out the caller-owned state buffer, fills kernel
// parameter blocks and launches the sm_100a kernels.  No CPU compute path exists here: every
// entry point that produces results does so by launching a kernel on the handle's device.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "../../include/multigrid_b200.h"
#include "mg_device.cuh"

namespace mg {
cudaError_t launch_collect_step(int v, const CollectParams& p, cudaStream_t st);
cudaError_t launch_collect_reset(int v, const CollectParams& p, cudaStream_t st);
cudaError_t launch_encode3(int v, const uint8_t* grid, uint8_t* obs, long long N, int cells, int obs_bulk_ok, cudaStream_t st);
int num_tile_variants();
int tile_envs(int v);
cudaError_t configure_kernels(int v, int cells, int A);
size_t tile_smem(int v, int cells, int A);
}  // namespace mg

struct mg_env {
  mg_config cfg;
  int device;
  int tile;  // kernel tile variant (envs per CTA x threads), MG_TILE env var, default 0
  long long n_pad;
  size_t plane_off[MG_PLANE_COUNT], plane_bytes[MG_PLANE_COUNT], plane_row[MG_PLANE_COUNT], state_bytes;
  mg::CollectParams base;  // rules + constants; pointers filled per call
  mg_trace trace;
  bool has_trace;
  int32_t* d_status;
  uint8_t* d_wall_template;
  // staging for the *_host entry points (allocated on first use)
  int8_t* d_actions; uint8_t* d_obs; double* d_rewards; uint8_t* d_term; uint8_t* d_trunc; uint8_t* d_final;
  long long launches;
  unsigned long long* timeline;
  std::string err;
};

static thread_local std::string g_create_err;

static int fail(mg_env* env, const std::string& msg) {
  if (env) env->err = msg; else g_create_err = msg;
  return -1;
}
static int cuda_fail(mg_env* env, const char* what, cudaError_t e) {
  return fail(env, std::string(what) + ": " + cudaGetErrorString(e));
}
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int mg_abi_version(void) { return MG_ABI_VERSION; }

extern "C" const char* mg_last_error(const mg_env* env) { return env ? env->err.c_str() : g_create_err.c_str(); }

extern "C" int mg_create(const mg_config* cfg, int device, mg_env** out) {
  if (!cfg || !out) return fail(nullptr, "mg_create: null argument");
  *out = nullptr;
  if (cfg->struct_size != sizeof(mg_config)) return fail(nullptr, "mg_create: mg_config size mismatch (ABI)");
  if (cfg->family != MG_FAMILY_COLLECT) return fail(nullptr, "mg_create: unknown env family");
  const int W = cfg->width, H = cfg->height, A = cfg->num_agents, nb = cfg->num_ball_types;
  if (cfg->num_envs < 1) return fail(nullptr, "mg_create: num_envs must be >= 1");
  if (W < 3 || H < 3 || W > 255 || H > 255) return fail(nullptr, "mg_create: width/height must be in [3, 255] (Grid asserts >= 3, grid.py:19-20)");
  if (A < 1 || A > MG_MAX_AGENTS) return fail(nullptr, "mg_create: num_agents must be in [1, 8]");
  if (nb < 1 || nb > MG_MAX_BALL_TYPES) return fail(nullptr, "mg_create: num_ball_types must be in [1, 8]");
  for (int i = 0; i < A; ++i)
    if (cfg->agent_colour[i] < 0 || cfg->agent_colour[i] > 9) return fail(nullptr, "mg_create: agent colour index outside COLORS (constants.py:8-19)");
  for (int t = 0; t < nb; ++t)
    if (cfg->ball_colour[t] < 0 || cfg->ball_colour[t] > 9) return fail(nullptr, "mg_create: ball colour index outside COLORS (constants.py:8-19)");
  if (cfg->num_balls < 0) return fail(nullptr, "mg_create: num_balls < 0");
  switch (cfg->layout) {
    case MG_LAYOUT_EVEN_DIST: break;
    case MG_LAYOUT_QUADRANTS:
      if (nb > 4) return fail(nullptr, "mg_create: quadrants layout has 4 partitions (collect_game.py:275-280)");
      if (A + 1 >= W) return fail(nullptr, "mg_create: agents do not fit on row H-2");
      break;
    case MG_LAYOUT_ROOMS: {
      if (W != H) return fail(nullptr, "mg_create: rooms layout is square (collect_game.py:306-362 uses width on both axes)");
      const double q = (double)cfg->num_balls / nb;
      const int num_ball = (int)__builtin_nearbyint(q);
      if (num_ball <= 0 || (cfg->num_balls - 1) / num_ball >= nb || (cfg->num_balls - 1) / num_ball >= 4)
        return fail(nullptr, "mg_create: rooms layout: num_balls / len(balls_index) indexes past the partitions (reference: IndexError)");
      break;
    }
    case MG_LAYOUT_QUADRANTS_RESPAWN:
      if (cfg->num_balls < 3 || (cfg->num_balls - 1) / (cfg->num_balls / 3) >= 3)
        return fail(nullptr, "mg_create: quadrants_respawn layout needs num_balls divisible into 3 partitions (reference: IndexError)");
      if (A + 1 >= W) return fail(nullptr, "mg_create: agents do not fit on row H-2");
      break;
    default: return fail(nullptr, "mg_create: unknown layout");
  }
  // enough free cells for the rejection sampler to terminate
  if (cfg->num_balls + 3 + A > (W - 2) * (H - 2)) return fail(nullptr, "mg_create: more objects than free cells (place_obj would never return)");

  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, std::string("mg_create: no usable CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(ce));
  if (device < 0 || device >= ndev) return fail(nullptr, "mg_create: device index out of range");
  if ((ce = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, "cudaSetDevice", ce);
  cudaDeviceProp prop;
  if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, "cudaGetDeviceProperties", ce);
  if (prop.major != 10) return fail(nullptr, "mg_create: kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
  int tile = 0;
  if (const char* tv = std::getenv("MG_TILE")) tile = std::atoi(tv);
  if (tile < 0 || tile >= mg::num_tile_variants()) return fail(nullptr, "mg_create: MG_TILE out of range");
  // large grids: fall back to the smallest tile that fits the 227 KB of shared memory
  while (mg::tile_smem(tile, W * H, A) > (size_t)prop.sharedMemPerBlockOptin && tile != 6) tile = (tile == 0 ? 5 : 6);
  if (mg::tile_smem(tile, W * H, A) > (size_t)prop.sharedMemPerBlockOptin)
    return fail(nullptr, "mg_create: grid too large for the shared-memory tile (16 envs x 4*W*H bytes must fit 227 KB)");

  if ((ce = mg::configure_kernels(tile, W * H, A)) != cudaSuccess) return cuda_fail(nullptr, "cudaFuncSetAttribute(max dynamic shared memory)", ce);

  mg_env* env = new (std::nothrow) mg_env();
  if (!env) return fail(nullptr, "mg_create: out of host memory");
  env->cfg = *cfg;
  env->device = device;
  env->tile = tile;
  env->has_trace = false;
  env->launches = 0;
  env->timeline = nullptr;
  env->d_actions = nullptr; env->d_obs = nullptr; env->d_rewards = nullptr;
  env->d_term = nullptr; env->d_trunc = nullptr; env->d_final = nullptr;
  std::memset(&env->trace, 0, sizeof env->trace);
  const int E = mg::tile_envs(tile);
  env->n_pad = (cfg->num_envs + E - 1) / E * E;
  const size_t rows[MG_PLANE_COUNT] = {(size_t)W * H, (size_t)A * 2, 16, (size_t)A * nb * 4};
  size_t off = 0;
  for (int i = 0; i < MG_PLANE_COUNT; ++i) {
    env->plane_off[i] = off;
    env->plane_row[i] = rows[i];
    env->plane_bytes[i] = rows[i] * (size_t)env->n_pad;
    off = align_up(off + env->plane_bytes[i], 256);
  }
  env->state_bytes = off;

  mg::CollectParams& p = env->base;
  std::memset(&p, 0, sizeof p);
  p.W = W; p.H = H; p.cells = W * H; p.A = A; p.nb = nb;
  p.num_balls = cfg->num_balls; p.respawn = cfg->respawn != 0; p.layout = cfg->layout;
  p.fixed_horizon = cfg->fixed_horizon != 0; p.max_steps = cfg->max_steps; p.time_limit = cfg->time_limit;
  p.autoreset = cfg->autoreset != 0;
  for (int i = 0; i < A; ++i) p.agent_code[i] = mg::cell(mg::T_AGENT, cfg->agent_colour[i], 3);  // dir 3, multigrid.py:371-374
  for (int c = 0; c < 16; ++c) { p.type_of_colour[c] = -1; p.reward_of_colour[c] = 1.0; }  // Ball(world, index, 1) collect_game.py:391
  for (int t = nb - 1; t >= 0; --t) {
    p.ball_colour[t] = (uint8_t)cfg->ball_colour[t];
    p.type_of_colour[cfg->ball_colour[t]] = (int8_t)t;
    p.reward_of_colour[cfg->ball_colour[t]] = cfg->ball_reward[t];
  }
  p.N = cfg->num_envs; p.env_id_base = (unsigned long long)cfg->env_id_base; p.seed = cfg->seed;
  p.rng_mode = 1;

  if ((ce = cudaMalloc(&env->d_status, sizeof(int32_t))) != cudaSuccess) { delete env; return cuda_fail(nullptr, "cudaMalloc(status)", ce); }
  if ((ce = cudaMemset(env->d_status, 0, sizeof(int32_t))) != cudaSuccess) { cudaFree(env->d_status); delete env; return cuda_fail(nullptr, "cudaMemset(status)", ce); }
  p.status = env->d_status;
  {  // the layout's walls on an empty grid, index x*H + y (grid.py:66-89; collect_game.py:309-320 for Rooms)
    std::string t((size_t)W * H, '\0');
    auto set = [&](int x, int y) { t[(size_t)x * H + y] = (char)mg::WALL_GREY; };
    for (int i = 0; i < W; ++i) { set(i, 0); set(i, H - 1); }
    for (int j = 0; j < H; ++j) { set(0, j); set(W - 1, j); }
    if (cfg->layout == MG_LAYOUT_ROOMS) {
      const int ws = W / 2 - 1, m = W / 2;
      for (int i = 0; i < ws; ++i) { set(i, m); set(W - ws + i, m); set(m, i); set(m, W - ws + i); }
    }
    if ((ce = cudaMalloc(&env->d_wall_template, t.size())) != cudaSuccess ||
        (ce = cudaMemcpy(env->d_wall_template, t.data(), t.size(), cudaMemcpyHostToDevice)) != cudaSuccess) {
      cudaFree(env->d_status); delete env; return cuda_fail(nullptr, "wall template upload", ce);
    }
    p.wall_template = env->d_wall_template;
  }
  *out = env;
  return 0;
}

extern "C" int mg_destroy(mg_env* env) {
  if (!env) return 0;
  cudaSetDevice(env->device);
  cudaFree(env->d_status);
  cudaFree(env->d_wall_template);
  cudaFree(env->d_actions); cudaFree(env->d_obs); cudaFree(env->d_rewards);
  cudaFree(env->d_term); cudaFree(env->d_trunc); cudaFree(env->d_final);
  delete env;
  return 0;
}

extern "C" size_t mg_state_bytes(const mg_env* env) { return env ? env->state_bytes : 0; }
extern "C" size_t mg_obs_bytes(const mg_env* env) {
  return env ? (size_t)env->cfg.num_envs * env->cfg.width * env->cfg.height * 3 : 0;
}
extern "C" int mg_state_plane(const mg_env* env, int plane, size_t* offset, size_t* bytes, size_t* row_bytes) {
  if (!env || plane < 0 || plane >= MG_PLANE_COUNT) return -1;
  if (offset) *offset = env->plane_off[plane];
  if (bytes) *bytes = env->plane_bytes[plane];
  if (row_bytes) *row_bytes = env->plane_row[plane];
  return 0;
}

static void bind_state(mg_env* env, mg::CollectParams& p, void* state) {
  uint8_t* s = static_cast<uint8_t*>(state);
  p.grid = s + env->plane_off[MG_PLANE_GRID];
  p.agent_pos = s + env->plane_off[MG_PLANE_AGENT_POS];
  p.hdr = reinterpret_cast<int4*>(s + env->plane_off[MG_PLANE_HDR]);
  p.info = reinterpret_cast<int32_t*>(s + env->plane_off[MG_PLANE_INFO]);
}
static void bind_trace(mg_env* env, mg::CollectParams& p) {
  if (!env->has_trace) { p.rng_mode = 1; return; }
  const mg_trace& t = env->trace;
  p.rng_mode = 0;
  p.order = t.order; p.draws = t.draws; p.n_draws = t.n_draws; p.K = t.K;
  p.reset_draws = t.reset_draws; p.n_reset_draws = t.n_reset_draws; p.R = t.R;
  p.draws_used = t.draws_used; p.reset_draws_used = t.reset_draws_used;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int mg_set_trace(mg_env* env, const mg_trace* t) {
  if (!env) return -1;
  if (!t) { env->has_trace = false; return 0; }
  env->trace = *t;
  env->has_trace = true;
  return 0;
}

extern "C" int mg_reset(mg_env* env, void* state, const uint8_t* mask, uint8_t* obs, void* stream) {
  if (!env || !state) return fail(env, "mg_reset: null argument");
  if (!aligned16(state)) return fail(env, "mg_reset: state buffer must be 16-byte aligned");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  mg::CollectParams p = env->base;
  bind_state(env, p, state);
  bind_trace(env, p);
  p.reset_mask = mask; p.obs = obs; p.obs_bulk_ok = aligned16(obs);
  if ((ce = mg::launch_collect_reset(env->tile, p, static_cast<cudaStream_t>(stream))) != cudaSuccess) return cuda_fail(env, "collect_reset_kernel", ce);
  env->launches += 1;
  return 0;
}

static int step_device(mg_env* env, void* state, const mg_step_io* io, cudaStream_t st) {
  if (!io->actions || !io->rewards || !io->terminated || !io->truncated) return fail(env, "mg_step: actions/rewards/terminated/truncated must be non-null");
  if (io->final_obs && !io->obs) return fail(env, "mg_step: final_obs needs obs");
  mg::CollectParams p = env->base;
  bind_state(env, p, state);
  bind_trace(env, p);
  if (p.rng_mode == 0 && !p.order) return fail(env, "mg_step: trace mode needs the recorded agent order");
  p.actions = io->actions; p.obs = io->obs; p.rewards = io->rewards;
  p.terminated = io->terminated; p.truncated = io->truncated; p.final_obs = io->final_obs;
  p.obs_bulk_ok = aligned16(io->obs);
  p.io_bulk_ok = aligned16(io->actions) && aligned16(io->rewards) && aligned16(io->terminated) && aligned16(io->truncated);
  p.timeline = env->timeline;
  cudaError_t ce;
  if ((ce = mg::launch_collect_step(env->tile, p, st)) != cudaSuccess) return cuda_fail(env, "collect_step_kernel", ce);
  env->launches += 1;
  return 0;
}

extern "C" int mg_step(mg_env* env, void* state, const mg_step_io* io, void* stream) {
  if (!env || !state || !io) return fail(env, "mg_step: null argument");
  if (!aligned16(state)) return fail(env, "mg_step: state buffer must be 16-byte aligned");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  return step_device(env, state, io, static_cast<cudaStream_t>(stream));
}

extern "C" int mg_encode(mg_env* env, const void* state, uint8_t* obs, void* stream) {
  if (!env || !state || !obs) return fail(env, "mg_encode: null argument");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  const uint8_t* grid = static_cast<const uint8_t*>(state) + env->plane_off[MG_PLANE_GRID];
  if ((ce = mg::launch_encode3(env->tile, grid, obs, env->cfg.num_envs, env->cfg.width * env->cfg.height, aligned16(obs),
                               static_cast<cudaStream_t>(stream))) != cudaSuccess)
    return cuda_fail(env, "encode3_kernel", ce);
  env->launches += 1;
  return 0;
}

extern "C" int mg_step_host(mg_env* env, void* state, const mg_step_io* io, void* stream) {
  if (!env || !state || !io) return fail(env, "mg_step_host: null argument");
  if (!io->actions || !io->rewards || !io->terminated || !io->truncated) return fail(env, "mg_step_host: actions/rewards/terminated/truncated must be non-null");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t N = (size_t)env->cfg.num_envs, A = (size_t)env->cfg.num_agents, ob = mg_obs_bytes(env);
  if (!env->d_actions) {
    if ((ce = cudaMalloc(&env->d_actions, N * A)) != cudaSuccess) return cuda_fail(env, "cudaMalloc", ce);
    if ((ce = cudaMalloc(&env->d_obs, ob)) != cudaSuccess) return cuda_fail(env, "cudaMalloc", ce);
    if ((ce = cudaMalloc(&env->d_rewards, N * A * sizeof(double))) != cudaSuccess) return cuda_fail(env, "cudaMalloc", ce);
    if ((ce = cudaMalloc(&env->d_term, N)) != cudaSuccess) return cuda_fail(env, "cudaMalloc", ce);
    if ((ce = cudaMalloc(&env->d_trunc, N)) != cudaSuccess) return cuda_fail(env, "cudaMalloc", ce);
  }
  if (io->final_obs && !env->d_final) {
    if ((ce = cudaMalloc(&env->d_final, ob)) != cudaSuccess) return cuda_fail(env, "cudaMalloc", ce);
    if ((ce = cudaMemsetAsync(env->d_final, 0, ob, st)) != cudaSuccess) return cuda_fail(env, "cudaMemsetAsync", ce);
  }
  if ((ce = cudaMemcpyAsync(env->d_actions, io->actions, N * A, cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(env, "H2D actions", ce);
  mg_step_io dio;
  dio.actions = env->d_actions; dio.obs = io->obs ? env->d_obs : nullptr; dio.rewards = env->d_rewards;
  dio.terminated = env->d_term; dio.truncated = env->d_trunc; dio.final_obs = io->final_obs ? env->d_final : nullptr;
  if (step_device(env, state, &dio, st)) return -1;
  if (io->obs && (ce = cudaMemcpyAsync(io->obs, env->d_obs, ob, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H obs", ce);
  if ((ce = cudaMemcpyAsync(io->rewards, env->d_rewards, N * A * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H rewards", ce);
  if ((ce = cudaMemcpyAsync(io->terminated, env->d_term, N, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H terminated", ce);
  if ((ce = cudaMemcpyAsync(io->truncated, env->d_trunc, N, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H truncated", ce);
  if (io->final_obs && (ce = cudaMemcpyAsync(io->final_obs, env->d_final, ob, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H final_obs", ce);
  if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(env, "cudaStreamSynchronize", ce);
  return 0;
}

extern "C" int mg_status(mg_env* env, void* stream, int32_t* status_out) {
  if (!env || !status_out) return -1;
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((ce = cudaMemcpyAsync(status_out, env->d_status, sizeof(int32_t), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H status", ce);
  if ((ce = cudaMemsetAsync(env->d_status, 0, sizeof(int32_t), st)) != cudaSuccess) return cuda_fail(env, "memset status", ce);
  if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(env, "cudaStreamSynchronize", ce);
  return 0;
}

extern "C" int mg_debug_set_timeline(mg_env* env, uint64_t* timeline_dev) {
  if (!env) return -1;
  env->timeline = reinterpret_cast<unsigned long long*>(timeline_dev);
  return 0;
}

extern "C" int mg_tile_envs(const mg_env* env) { return env ? mg::tile_envs(env->tile) : -1; }

extern "C" int64_t mg_launch_count(const mg_env* env) { return env ? env->launches : 0; }
