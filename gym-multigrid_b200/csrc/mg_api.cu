// mg_api.cu -- implementation of the C ABI declared in include/multigrid_b200.h.
// Host-side only: validates configs, lays out the caller-owned state buffer, fills kernel
// parameter blocks and launches the sm_100a kernels.  No CPU compute path exists here: every
// entry point that produces results does so by launching a kernel on the handle's device.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/multigrid_b200.h"
#include "host_transport.h"
#include "mg_device.cuh"
#include "smem_config.h"

namespace mg {
cudaError_t launch_collect_step(int v, const CollectParams& p, cudaStream_t st);
cudaError_t launch_collect_reset(int v, const CollectParams& p, cudaStream_t st);
cudaError_t launch_encode3(int v, const uint8_t* grid, uint8_t* obs, long long N, int cells, int obs_bulk_ok, cudaStream_t st);
cudaError_t launch_collect_rollout(const CollectParams& p, int num_sms, cudaStream_t st);
cudaError_t configure_rollout_kernels(int cells, int A);
size_t rollout_smem_bytes(int cells, int A, int tile, bool with_delta);
int num_tile_variants();
int tile_envs(int v);
cudaError_t configure_kernels(int v, int cells, int A);
size_t tile_smem(int v, int cells, int A);
struct MapParams;
cudaError_t launch_map(const MapParams& p, cudaStream_t st);
cudaError_t configure_map_kernels(int L, int n, int cells, int obs_dtype, bool occ);
size_t map_smem_bytes(int L, int n, int cells, int obs_dtype, bool occ);
bool map_ctf_occ(int family, int nb, int nr, int cells);
bool map_can_fuse_policy(const MapParams& p);
bool map_obs_staged(int cells, int obs_dtype);
bool map_obs_tma(int L, int cells, int obs_dtype);
int map_tma_reps(int L, int cells, int obs_dtype);
size_t map_view_smem_bytes(int padded_bytes, int V);
cudaError_t configure_map_view_mode(size_t smem);
cudaError_t launch_map_info(const MapParams& p, double* out, cudaStream_t st);
cudaError_t launch_ctf_flat(const MapParams& p, const long long* tmpl, int L, void* out, int elem, cudaStream_t st);
int ctf_flat_tile_envs(int L, int elem);
int map_tile_envs();
}  // namespace mg
#include "map_params.cuh"
#include "view_params.cuh"
#include "wildfire_params.cuh"
#include "generic_params.cuh"
#include "policy_params.cuh"
namespace mg {
cudaError_t launch_ctf_policy(const PolicyParams& p, cudaStream_t st);
cudaError_t launch_view(const ViewParams& p, cudaStream_t st);
cudaError_t launch_view6(const View6Params& p, cudaStream_t st);
cudaError_t launch_toroid(const uint8_t* grid, const uint8_t* pos, float* out, long long N, int W, int A, int nb, cudaStream_t st);
void build_render_atlas(int family, int ts, std::vector<uint8_t>& atlas);
cudaError_t launch_render(const uint8_t* cells, const uint8_t* agents, int agent_stride, int family, int n_agents, int num_blue, int variant_1v1,
                          long long N, const int32_t* env_ids, int n, int W, int H, int ts, const uint8_t* atlas, uint8_t* out, int32_t* status,
                          cudaStream_t st);
size_t view_smem_bytes(const ViewParams& p);
int view_max();
int view_tile_envs();
cudaError_t launch_wildfire(const WildfireParams& p, cudaStream_t st);
cudaError_t configure_wildfire_kernel(int cells, int H);
size_t wildfire_smem_bytes(int cells, int H);
cudaError_t launch_generic(const GenericParams& p, cudaStream_t st);
cudaError_t configure_generic_kernel(int A, int cells);
size_t generic_smem_bytes(int A, int cells);
int generic_tile_envs();
}  // namespace mg

struct mg_env {
  int family;
  mg_config cfg;
  mg_map_config mcfg;
  mg::MapParams mbase;
  mg_wildfire_config wcfg;
  mg::WildfireParams wbase;
  mg_generic_config gcfg;
  mg::GenericParams gbase;
  mg_map_trace mtrace;
  std::vector<long long> flat_tmpl;        // CtF: static entries of the "flattened" observation (ctf.py:1084-1104), per-env slots 0
  long long* d_flat_tmpl;                  // ... uploaded on first use
  std::vector<std::pair<int, uint8_t*>> atlases;    // render: (tile_size, device atlas) built on first use, freed by mg_destroy
  uint8_t* d_policy_tables;        // CtF: tables of the scripted opponents (mg_set_red_policies), or null
  const uint16_t* pol_tr_patrol; const uint8_t* pol_tr_follow; const int8_t* pol_tr_action;   // mg_set_policy_trace (validation), or null
  uint4* d_view_table;             // Maze partial-observation mode: the memoised views of all S*S*4 agent states, or null
  mg::PolicyParams pbase;
  const int8_t* ext_red_actions;   // CtF: actions of an external enemy policy for the next steps (Philox mode), or null = RwPolicy
  int8_t* fused_red_out;           // CtF: mg_set_red_policy_fusion - every step decides the scripted opponents' actions itself, or null
  uint8_t* d_map_tables;  // field_map | obs_period | background / territory lists
  size_t obs_elem;        // bytes per obs element
  int act_cols, rew_cols;
  size_t map_codes_off;   // Maze: offset of the packed static map (partial views) inside d_map_tables
  size_t map_padded_off, map_padded_bytes;  // ... and of the copy padded with the out-of-map filler (fast view kernel)
  int map_pad;
  int device;
  int tile;  // kernel tile variant (envs per CTA x threads), MG_TILE env var, default 0
  long long n_pad;
  size_t plane_off[8], plane_bytes[8], plane_row[8], state_bytes;
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
  // compact host transports (mg_set_host_transport; Collect family)
  int num_sms; size_t smem_optin;
  int step_impl;             // 0 = tile kernel (collect_step_kernel), 1 = warp-tile kernel (collect_rollout_kernel, T = 1)
  bool rollout_ready;
  int transport;             // MG_TRANSPORT_*
  int host_threads;
  uint8_t* d_delta_blk;      // [16-byte header: int32 reset count][N x delta records]
  uint8_t* d_reset_rows;     // [N][reset_stride]: int32 env index + packed row
  uint8_t* h_delta_blk;      // page-locked mirrors of the two, plus the packed grid plane for full refreshes
  uint8_t* h_reset_rows;
  uint8_t* h_grid;
  size_t reset_stride;
  double reward_table[33];
  const uint8_t* mirror;     // the caller's obs buffer the delta transport patches in place
  bool mirror_valid;
  struct {
    bool active, refresh, plane; mg_step_io io;
    // asynchronous decode (mg_step_host_async with a compact transport): the completer thread waits for `ev`, then the host pool
    // decodes while the caller is free to enqueue other env batches; mg_step_host_wait helps and waits for `done`
    bool async; cudaStream_t st;
    std::atomic<int> done;        // raised after the last decode task (or on an error)
    std::atomic<int> pre;         // prerequisites of the reset-row task still open: the delta pass, the row copy
    int cuda_err; const char* err_what; bool corrupt;
    mg::HostTask t1, t2;
    mg::HostDeltaJob job;
  } pend;
  cudaEvent_t host_ev;           // recorded behind the step's last device-to-host copy
  // the steady-state delta step (H2D actions, header clear, kernel, D2H records) as ONE graph launch; re-captured when anything the
  // launches depend on differs from the capture (state / action pointers, the kernel's whole parameter block, the kernel variant)
  struct { cudaGraphExec_t exec; void* state; const void* h_actions; int step_impl, tile; mg::CollectParams p; int failures; } hgraph;
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
  // default tile: 32 envs x 96 threads (variant 9) - the fastest under fresh per-step actions, where most warps take the pickup /
  // respawn branch every step (profiles/r02_kbench_collect_variants.jsonl); MG_TILE selects another variant for experiments
  int tile = 9;
  if (const char* tv = std::getenv("MG_TILE")) tile = std::atoi(tv);
  if (tile < 0 || tile >= mg::num_tile_variants()) return fail(nullptr, "mg_create: MG_TILE out of range");
  // large grids: fall back to the smallest tile that fits the 227 KB of shared memory
  while (mg::tile_smem(tile, W * H, A) > (size_t)prop.sharedMemPerBlockOptin && tile != 6) tile = 6;
  if (mg::tile_smem(tile, W * H, A) > (size_t)prop.sharedMemPerBlockOptin)
    return fail(nullptr, "mg_create: grid too large for the shared-memory tile (16 envs x 4*W*H bytes must fit 227 KB)");

  if ((ce = mg::configure_kernels(tile, W * H, A)) != cudaSuccess) return cuda_fail(nullptr, "cudaFuncSetAttribute(max dynamic shared memory)", ce);

  mg_env* env = new (std::nothrow) mg_env();
  if (!env) return fail(nullptr, "mg_create: out of host memory");
  env->family = MG_FAMILY_COLLECT;
  env->num_sms = prop.multiProcessorCount; env->smem_optin = (size_t)prop.sharedMemPerBlockOptin;
  // mg_step's kernel: the warp-tile kernel (collect_rollout_kernels.cu, T = 1) unless MG_STEP_IMPL=tile asks for the CTA-tile kernel
  // (collect_kernels.cu), which also serves grids too large for a warp's shared-memory slice
  { const char* v = std::getenv("MG_STEP_IMPL"); env->step_impl = (v && v[0] == 't') ? 0 : 1; }
  env->map_codes_off = 0;
  env->d_map_tables = nullptr;
  env->obs_elem = 1; env->act_cols = cfg->num_agents; env->rew_cols = cfg->num_agents;
  env->cfg = *cfg;
  env->device = device;
  env->tile = tile;
  env->has_trace = false;
  env->launches = 0;
  env->timeline = nullptr;
  env->d_actions = nullptr; env->d_obs = nullptr; env->d_rewards = nullptr;
  env->d_term = nullptr; env->d_trunc = nullptr; env->d_final = nullptr;
  std::memset(&env->trace, 0, sizeof env->trace);
  const int E = mg::tile_envs(tile) > mg::view_tile_envs() ? mg::tile_envs(tile) : mg::view_tile_envs();  // both tile sizes are powers of two
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
  // Ball.reward is an attribute of the ball OBJECT with two possible values in the reference:
  //   placed by _gen_grid: balls_reward[type] (collect_game.py:98-101, :252, :287, :354, :359), QuadrantsRespawn: the literal 1 (:393);
  //   placed by _respawn:  balls_reward[COLOUR index] (:130, :409).  Where that raises in the reference (colour >= len(balls_reward):
  //   IndexError) the initial value is kept.  If the two differ for a colour this env can hold, respawned balls carry bit 6.
  // The info counter of a pickup is indexed by the ball's COLOUR index, not its type: info[keys[num_ball_types * i + ball_idx]] with
  // ball_idx = COLOR_TO_IDX[fwd_cell.color] (collect_game.py:139,147).  type_of_colour holds that column, -1 where the reference
  // would run past agent i's num_ball_types columns (it raises or bumps a neighbour's counter there; no counter is touched here).
  for (int c = 0; c < 16; ++c) { p.type_of_colour[c] = (int8_t)(c < nb ? c : -1); p.reward_initial[c] = 1.0; }
  for (int t = nb - 1; t >= 0; --t) {
    p.ball_colour[t] = (uint8_t)cfg->ball_colour[t];
    if (cfg->layout != MG_LAYOUT_QUADRANTS_RESPAWN) p.reward_initial[cfg->ball_colour[t]] = cfg->ball_reward[t];
  }
  for (int c = 0; c < 16; ++c) p.reward_respawned[c] = c < nb ? cfg->ball_reward[c] : p.reward_initial[c];
  p.mark_respawned = 0;
  if (cfg->respawn) {
    if (cfg->layout == MG_LAYOUT_QUADRANTS_RESPAWN) {
      for (int c = 0; c < 3; ++c) if (p.reward_initial[c] != p.reward_respawned[c]) p.mark_respawned = 1;
    } else {
      for (int t = 0; t < nb; ++t) if (p.reward_initial[cfg->ball_colour[t]] != p.reward_respawned[cfg->ball_colour[t]]) p.mark_respawned = 1;
    }
  }
  env->reward_table[0] = 0.0;
  for (int c = 0; c < 16; ++c) { env->reward_table[1 + c] = p.reward_initial[c]; env->reward_table[17 + c] = p.reward_respawned[c]; }
  p.N = cfg->num_envs; p.env_id_base = (unsigned long long)cfg->env_id_base; p.seed = cfg->seed;
  p.rng_mode = 1;
  // early observation store: OFF by default.  It pays only while agents stand still (round 1's frozen-action benchmark): with fresh
  // actions every env patches ~4 cells x 3 bytes of the slab in global memory, and those 8 x 10^5 partial-sector writes per launch cost
  // as much L2 bandwidth as storing the slab again (13.9 vs 10.8 us per 65 536-env launch on one stream).  MG_EARLY_OBS=1 enables it.
  { const char* v = std::getenv("MG_EARLY_OBS"); p.early_obs = (v && v[0] == '1'); }

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

static void drain_host_step(mg_env* env);

extern "C" int mg_destroy(mg_env* env) {
  if (!env) return 0;
  cudaSetDevice(env->device);
  drain_host_step(env);
  if (env->host_ev) cudaEventDestroy(env->host_ev);
  if (env->hgraph.exec) cudaGraphExecDestroy(env->hgraph.exec);
  cudaFree(env->d_status);
  cudaFree(env->d_wall_template);
  cudaFree(env->d_map_tables);
  for (auto& a : env->atlases) cudaFree(a.second);
  cudaFree(env->d_flat_tmpl);
  cudaFree(env->d_policy_tables);
  cudaFree(env->d_view_table);
  cudaFree(env->d_actions); cudaFree(env->d_obs); cudaFree(env->d_final);   // rewards / term / trunc live inside the d_obs block
  cudaFree(env->d_delta_blk); cudaFree(env->d_reset_rows);
  cudaFreeHost(env->h_delta_blk); cudaFreeHost(env->h_reset_rows); cudaFreeHost(env->h_grid);
  delete env;
  return 0;
}

extern "C" size_t mg_state_bytes(const mg_env* env) { return env ? env->state_bytes : 0; }
extern "C" size_t mg_obs_bytes(const mg_env* env) {
  if (!env) return 0;
  if (env->family == MG_FAMILY_WILDFIRE) return (size_t)env->wcfg.num_envs * env->wcfg.width * env->wcfg.height * 3;
  if (env->family == MG_FAMILY_GENERIC) return (size_t)env->gcfg.num_envs * env->gcfg.num_agents * env->gcfg.width * env->gcfg.height * 6;
  if (env->family == MG_FAMILY_MAZE && env->mbase.view_V) return (size_t)env->mcfg.num_envs * env->mbase.view_V * env->mbase.view_V * 3;
  if (env->family != MG_FAMILY_COLLECT) return (size_t)env->mcfg.num_envs * env->mcfg.size * env->mcfg.size * env->obs_elem;
  return (size_t)env->cfg.num_envs * env->cfg.width * env->cfg.height * 3;
}
extern "C" int mg_state_plane(const mg_env* env, int plane, size_t* offset, size_t* bytes, size_t* row_bytes) {
  if (!env || plane < 0 || plane >= (env->family == MG_FAMILY_WILDFIRE ? (int)MG_WF_PLANE_COUNT : env->family == MG_FAMILY_GENERIC ? (int)MG_GEN_PLANE_COUNT : (int)MG_PLANE_COUNT)) return -1;
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

// ------------------------------------------------------------------------------ Maze / CtF
static size_t gcd_sz(size_t a, size_t b) { while (b) { size_t t = a % b; a = b; b = t; } return a; }

extern "C" int mg_create_map(const mg_map_config* cfg, int device, mg_env** out) {
  if (!cfg || !out) return fail(nullptr, "mg_create_map: null argument");
  *out = nullptr;
  if (cfg->struct_size != sizeof(mg_map_config)) return fail(nullptr, "mg_create_map: mg_map_config size mismatch (ABI)");
  const bool maze = cfg->family == MG_FAMILY_MAZE;
  if (!maze && cfg->family != MG_FAMILY_CTF) return fail(nullptr, "mg_create_map: family must be MG_FAMILY_MAZE or MG_FAMILY_CTF");
  const int S = cfg->size, cells = S * S;
  if (cfg->num_envs < 1) return fail(nullptr, "mg_create_map: num_envs must be >= 1");
  if (S < 3 || S > 255 || !cfg->field_map) return fail(nullptr, "mg_create_map: need a square field_map with 3 <= size <= 255");
  const int nb = maze ? 1 : cfg->num_blue, nr = maze ? 0 : cfg->num_red, n = nb + nr;
  if (nb < 1 || nr < 0 || n > MG_MAX_MAP_AGENTS || (!maze && nr < 1)) return fail(nullptr, "mg_create_map: agent counts out of range (1..16 agents in total)");
  if (cfg->max_steps < 1) return fail(nullptr, "mg_create_map: max_steps must be >= 1");
  if (cfg->variant_1v1 && (maze || nb != 1 || nr != 1)) return fail(nullptr, "mg_create_map: variant_1v1 needs the CtF family with one blue and one red agent");
  if (cfg->variant_1v1 && cfg->obstacle_penalty != 0.0)
    return fail(nullptr, "mg_create_map: Ctf1v1Env is only defined for obstacle_penalty == 0 (the reference's own step raises otherwise, ctf.py:639)");
  // cell lists in np.where order (row-major over field_map[x][y])
  std::string bg, bt, rt;  // uint16 lists packed in strings
  // entries are packed cells x | y << 8 (cell index i = x * S + y), so the kernels never divide by S
  auto push = [S](std::string& v, int c) { uint16_t u = (uint16_t)((c / S) | ((c % S) << 8)); v.append(reinterpret_cast<const char*>(&u), 2); };
  int blue_flag = -1, red_flag = -1;
  for (int i = 0; i < cells; ++i) {
    const int c = cfg->field_map[i];
    if (maze) {
      if (c > 3) return fail(nullptr, "mg_create_map: Maze map codes must be 0 background, 2 flag, 3 obstacle (world.py:81-91)");
      if (c == 0) push(bg, i);
    } else {
      if (c > 6 || c == 2 || c == 3) return fail(nullptr, "mg_create_map: CtF map codes must be 0/1 territory, 4/5 flags, 6 obstacle (world.py:66-79)");
      if (c == 0) push(bt, i);
      if (c == 1) push(rt, i);
      if (c == 4 && blue_flag < 0) blue_flag = i;   // list(zip(*np.where(...)))[0]  ctf.py:757-763
      if (c == 5 && red_flag < 0) red_flag = i;
    }
  }
  if (maze && bg.empty()) return fail(nullptr, "mg_create_map: Maze map has no background cell to start on");
  if (!maze) {
    if (blue_flag < 0 || red_flag < 0) return fail(nullptr, "mg_create_map: CtF map needs a blue flag (4) and a red flag (5)");
    push(bt, blue_flag); push(rt, red_flag);  // territory lists end with the flag cell (ctf.py:765-773)
    if ((int)bt.size() / 2 < nb || (int)rt.size() / 2 < nr) return fail(nullptr, "mg_create_map: more agents than territory cells");
  }
  const size_t L = (size_t)cells / gcd_sz((size_t)cells, 16) * 16;  // lcm(cells, 16)

  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, std::string("mg_create_map: no usable CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(ce));
  if (device < 0 || device >= ndev) return fail(nullptr, "mg_create_map: device index out of range");
  if ((ce = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, "cudaSetDevice", ce);
  cudaDeviceProp prop;
  if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, "cudaGetDeviceProperties", ce);
  if (prop.major != 10) return fail(nullptr, "mg_create_map: kernels are built for sm_100a only");
  const bool occ = mg::map_ctf_occ(cfg->family, nb, nr, cells);
  if (mg::map_smem_bytes((int)L, n, cells, cfg->obs_dtype, occ) > (size_t)prop.sharedMemPerBlockOptin)
    return fail(nullptr, "mg_create_map: map too large: lcm(size*size, 16) bytes must fit in shared memory");
  if ((ce = mg::configure_map_kernels((int)L, n, cells, cfg->obs_dtype, occ)) != cudaSuccess) return cuda_fail(nullptr, "cudaFuncSetAttribute", ce);

  mg_env* env = new (std::nothrow) mg_env();
  if (!env) return fail(nullptr, "mg_create_map: out of host memory");
  env->family = cfg->family;
  env->map_codes_off = 0;
  env->mcfg = *cfg; env->mcfg.field_map = nullptr;
  if (!maze) {   // blue agent pairs | red agent pairs | blue flag | red flag | blue_territory | red_territory | obstacle | terminated
    std::vector<long long>& t = env->flat_tmpl;
    t.assign((size_t)2 * n, 0);
    auto xy = [&](int c) { t.push_back(c / S); t.push_back(c % S); };
    xy(blue_flag); xy(red_flag);
    for (const std::string* v : {&bt, &rt})
      for (size_t k = 0; k + 1 < v->size(); k += 2) { t.push_back((unsigned char)(*v)[k]); t.push_back((unsigned char)(*v)[k + 1]); }
    for (int i = 0; i < cells; ++i) if (cfg->field_map[i] == 6) xy(i);
    t.resize(t.size() + (size_t)(cfg->variant_1v1 ? 1 : n), 0);   // terminated per agent; Ctf1v1Env: is_red_agent_defeated alone (ctf.py:359-371)
  }
  env->device = device; env->tile = 0; env->has_trace = false; env->launches = 0; env->timeline = nullptr;
  env->d_actions = nullptr; env->d_obs = nullptr; env->d_rewards = nullptr; env->d_term = nullptr; env->d_trunc = nullptr;
  env->d_final = nullptr; env->d_wall_template = nullptr; env->d_status = nullptr; env->d_map_tables = nullptr;
  std::memset(&env->mtrace, 0, sizeof env->mtrace);
  env->ext_red_actions = nullptr;
  env->obs_elem = cfg->obs_dtype == MG_OBS_U8 ? 1 : 8;
  env->act_cols = nb; env->rew_cols = 1;
  const int E = mg::map_tile_envs();
  env->n_pad = (cfg->num_envs + E - 1) / E * E;
  int slots = 1;
  while (slots < n) slots *= 2;
  const size_t rows[MG_MAP_PLANE_COUNT] = {(size_t)slots * 4, 16};
  size_t off = 0;
  for (int i = 0; i < MG_MAP_PLANE_COUNT; ++i) {
    env->plane_off[i] = off; env->plane_row[i] = rows[i]; env->plane_bytes[i] = rows[i] * (size_t)env->n_pad;
    off = align_up(off + env->plane_bytes[i], 256);
  }
  env->state_bytes = off;

  // device tables: field_map | obs_period | cell lists
  std::string period(L, '\0');
  for (size_t k = 0; k < L; ++k) {
    const int i = (int)(k % cells);
    // Maze obs is field_map as is (maze.py:245-260); CtF returns the transpose (ctf.py:1163): obs[y][x] = map[x][y]
    period[k] = (char)(maze ? cfg->field_map[i] : cfg->field_map[(i % S) * S + (i / S)]);
  }
  const size_t o_map = 0, o_per = align_up((size_t)cells, 256), o_bg = align_up(o_per + L, 256),
               o_bt = align_up(o_bg + bg.size(), 256), o_rt = align_up(o_bt + bt.size(), 256), o_pk = align_up(o_rt + rt.size(), 256),
               o_pad = align_up(o_pk + cells, 256), o_d2 = 0;  /* o_d2 set below */
  const int pad = 14, pitch = S + 2 * pad;   // view_size <= 15: a view reaches at most 14 cells beyond the map
  const size_t padded_bytes = align_up((size_t)pitch * pitch, 16);
  const size_t o_d2b = align_up(o_pad + padded_bytes, 256), o_tile = align_up(o_d2b + (size_t)3 * cells * sizeof(int32_t), 256);
  const bool staged = mg::map_obs_staged(cells, cfg->obs_dtype);
  const size_t tile_bytes = staged ? (size_t)mg::map_tile_envs() * cells : 0;   // image of one tile's static observation slab
  const size_t total = align_up(o_tile + tile_bytes, 256) + 256;
  (void)o_d2;
  std::string blob(total, '\0');
  std::memcpy(&blob[o_map], cfg->field_map, cells);
  std::memcpy(&blob[o_per], period.data(), L);
  for (size_t k = 0; k < tile_bytes; ++k) blob[o_tile + k] = period[k % L];   // tile envs * cells is a multiple of L
  if (!bg.empty()) std::memcpy(&blob[o_bg], bg.data(), bg.size());
  if (!bt.empty()) std::memcpy(&blob[o_bt], bt.data(), bt.size());
  if (!rt.empty()) std::memcpy(&blob[o_rt], rt.data(), rt.size());
  if (maze)  // packed (type | colour << 2) static map for partial views: Floor "background" white, Flag red, Obstacle grey (maze.py:183-198)
    for (int i = 0; i < cells; ++i) {
      const int c = cfg->field_map[i];
      blob[o_pk + i] = (char)(c == 0 ? mg::cell(0, 10, 0) : (c == 2 ? mg::cell(2, 0, 0) : mg::cell(3, 7, 0)));
    }
  if (maze) {
    std::memset(&blob[o_pad], (int)mg::cell(3, 7, 1), padded_bytes);   // the out-of-map filler (see mg_gen_obs)
    for (int x = 0; x < S; ++x) std::memcpy(&blob[o_pad + (size_t)(x + pad) * pitch + pad], &blob[o_pk + (size_t)x * S], S);
  }
  env->map_padded_off = o_pad; env->map_padded_bytes = padded_bytes; env->map_pad = pad;
  {  // _get_info tables: min squared distance from every cell to the cell lists of maze.py:262-269 / ctf.py:1165-1182
    int32_t* d2 = reinterpret_cast<int32_t*>(&blob[o_d2b]);
    std::vector<int32_t> col((size_t)cells);
    for (int t = 0; t < 3; ++t) {
      // Maze: flag list, obstacle list; CtF: blue territory + blue flag, red territory + red flag (ctf.py:765-773), obstacle list
      const int ca = maze ? (t == 0 ? 2 : (t == 1 ? 3 : -1)) : (t == 0 ? 0 : (t == 1 ? 1 : 6));
      const int cb = maze ? ca : (t == 0 ? 4 : (t == 1 ? 5 : 6));
      // exact squared Euclidean distance to the nearest listed cell in two separable passes, O(S^3) instead of O(S^4):
      // col[x][y] = min over x' of (x - x')^2 with (x', y) listed; d2[x][y] = min over y' of col[x][y'] + (y - y')^2
      for (int y = 0; y < S; ++y)
        for (int x = 0; x < S; ++x) {
          int best = -1;
          for (int xx = 0; xx < S; ++xx) {
            const int c = cfg->field_map[xx * S + y];
            if (c != ca && c != cb) continue;
            const int v = (x - xx) * (x - xx);
            if (best < 0 || v < best) best = v;
          }
          col[(size_t)x * S + y] = best;
        }
      for (int x = 0; x < S; ++x)
        for (int y = 0; y < S; ++y) {
          int best = -1;
          for (int yy = 0; yy < S; ++yy) {
            const int g = col[(size_t)x * S + yy];
            if (g < 0) continue;
            const int v = g + (y - yy) * (y - yy);
            if (best < 0 || v < best) best = v;
          }
          d2[t * cells + x * S + y] = best;
        }
    }
  }
  if ((ce = cudaMalloc(&env->d_map_tables, total)) != cudaSuccess ||
      (ce = cudaMemcpy(env->d_map_tables, blob.data(), total, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (ce = cudaMalloc(&env->d_status, sizeof(int32_t))) != cudaSuccess ||
      (ce = cudaMemset(env->d_status, 0, sizeof(int32_t))) != cudaSuccess) {
    cudaFree(env->d_map_tables); cudaFree(env->d_status); delete env;
    return cuda_fail(nullptr, "mg_create_map: device tables", ce);
  }
  mg::MapParams& p = env->mbase;
  std::memset(&p, 0, sizeof p);
  p.S = S; p.cells = cells; p.nb = nb; p.nr = nr; p.n = n; p.family = cfg->family; p.max_steps = cfg->max_steps;
  p.autoreset = cfg->autoreset != 0; p.obs_dtype = cfg->obs_dtype; p.variant_1v1 = cfg->variant_1v1 != 0;
  p.flag_reward = cfg->flag_reward; p.obstacle_penalty = cfg->obstacle_penalty; p.step_penalty = cfg->step_penalty;
  p.battle_reward = cfg->battle_reward; p.battle_range = cfg->battle_range; p.randomness = cfg->randomness;
  p.n_background = (int)bg.size() / 2; p.len_blue = (int)bt.size() / 2; p.len_red = (int)rt.size() / 2;
  p.blue_flag = maze ? 0 : (blue_flag / S) | ((blue_flag % S) << 8); p.red_flag = maze ? 0 : (red_flag / S) | ((red_flag % S) << 8);
  p.L16 = (int)(L / 16); p.L16_magic = (unsigned)(4294967296ull / (L / 16)) + 1u; p.tile_mod_L16 = mg::map_tile_envs() % (int)(L / 16);
  // integer restatements of the battle tests (ctf.py:1368, 1392-1407), formed with the reference's double arithmetic
  p.d2_max = -1;
  for (int d2 = 0; d2 <= 2 * 255 * 255 && std::sqrt((double)d2) <= cfg->battle_range; ++d2) p.d2_max = d2;
  auto win_threshold = [](double pb) -> unsigned long long {
    const double t = std::ceil(pb * 4294967296.0);   // (double)u / 2^32 < pb  <=>  u < ceil(pb * 2^32): scaling by 2^32 is exact
    return t <= 0.0 ? 0ull : (t >= 4294967296.0 ? 4294967296ull : (unsigned long long)t);
  };
  p.thr_blue_home = win_threshold(cfg->randomness); p.thr_red_home = win_threshold(1.0 - cfg->randomness); p.thr_even = win_threshold(0.5);
  p.obs_staged = mg::map_obs_staged(cells, cfg->obs_dtype) ? 1 : 0;
  p.obs_tma = mg::map_obs_tma((int)L, cells, cfg->obs_dtype) ? 1 : 0;
  p.tma_reps = mg::map_tma_reps((int)L, cells, cfg->obs_dtype);
  p.N = cfg->num_envs; p.env_id_base = (unsigned long long)cfg->env_id_base; p.seed = cfg->seed;
  p.field_map = env->d_map_tables + o_map; p.obs_period = env->d_map_tables + o_per; p.L = (int)L;
  p.obs_tile = staged ? env->d_map_tables + o_tile : nullptr;
  p.background = reinterpret_cast<const uint16_t*>(env->d_map_tables + o_bg);
  p.blue_terr = reinterpret_cast<const uint16_t*>(env->d_map_tables + o_bt);
  p.red_terr = reinterpret_cast<const uint16_t*>(env->d_map_tables + o_rt);
  p.status = env->d_status; p.rng_mode = 1;
  p.d2_tables = reinterpret_cast<const int32_t*>(env->d_map_tables + o_d2b);
  env->map_codes_off = o_pk;
  *out = env;
  return 0;
}

extern "C" int mg_set_map_trace(mg_env* env, const mg_map_trace* t) {
  if (!env || env->family == MG_FAMILY_COLLECT) return -1;
  if (!t) { env->has_trace = false; return 0; }
  env->mtrace = *t;
  env->has_trace = true;
  return 0;
}

static int map_launch(mg_env* env, void* state, int op, const mg_step_io* io, const uint8_t* mask, void* obs, cudaStream_t st) {
  mg::MapParams p = env->mbase;
  uint8_t* s = static_cast<uint8_t*>(state);
  p.agents = s + env->plane_off[MG_MAP_PLANE_AGENTS]; p.row_bytes = (int)env->plane_row[MG_MAP_PLANE_AGENTS];
  p.hdr = reinterpret_cast<int4*>(s + env->plane_off[MG_MAP_PLANE_HDR]);
  p.op = op; p.reset_mask = mask;
  p.red_actions = env->ext_red_actions;
  if (env->has_trace) {
    const mg_map_trace& t = env->mtrace;
    p.rng_mode = 0;
    p.start_index = t.start_index; p.blue_place = t.blue_place; p.red_place = t.red_place; p.red_actions = t.red_actions;
    p.order = t.order; p.blue_win = t.blue_win; p.KB = t.KB; p.battles_used = t.battles_used;
    const bool maze = env->family == MG_FAMILY_MAZE;
    const bool need_reset = op == 0 || p.autoreset;
    if (need_reset && maze && !p.start_index) return fail(env, "trace mode: Maze reset needs start_index");
    if (need_reset && !maze && (!p.blue_place || !p.red_place)) return fail(env, "trace mode: CtF reset needs blue_place / red_place");
    if (op == 1 && !maze && (!p.red_actions || (!p.order && !p.variant_1v1))) return fail(env, "trace mode: CtF step needs red_actions and order");
  }
  if (op == 1) {
    p.actions = io->actions; p.obs = io->obs; p.rewards = io->rewards; p.terminated = io->terminated;
    p.truncated = io->truncated; p.final_obs = io->final_obs;
  } else {
    p.obs = obs;
  }
  if (p.obs && !aligned16(p.obs)) return fail(env, "obs buffer must be 16-byte aligned");
  if (p.view_V && p.final_obs) return fail(env, "final_obs is not available in partial-observation mode");
  cudaError_t ce;
  if (op == 1 && env->fused_red_out && !env->has_trace) {   // the scripted opponents decide as part of this step (mg_set_red_policy_fusion)
    if (!env->d_policy_tables) return fail(env, "mg_step: red-policy fusion is on but no policies are set (mg_set_red_policies)");
    p.red_actions = env->fused_red_out;
    const mg::PolicyParams& q = env->pbase;
    if (!env->pol_tr_follow && mg::map_can_fuse_policy(p)) {   // one launch: the 2v2 lean kernel with the policy prologue
      p.pol_on = 1; p.pol_n_along = q.n_along;
      p.pol_first_move = q.first_move; p.pol_goal = q.patrol_goal; p.pol_border = q.on_border; p.pol_along = q.along;
      for (int k = 0; k < 2; ++k) { p.pol_kind[k] = q.kind[k]; p.pol_thr[k] = q.thr[k]; }
      p.pol_out = env->fused_red_out;
    } else {                                                   // any other configuration: the policy kernel first, same stream
      mg::PolicyParams pp = q;
      pp.agents = p.agents; pp.row_bytes = p.row_bytes; pp.hdr = p.hdr; pp.seed = p.seed; pp.out = env->fused_red_out;
      pp.tr_patrol = env->pol_tr_patrol; pp.tr_follow = env->pol_tr_follow; pp.tr_action = env->pol_tr_action;
      if ((ce = mg::launch_ctf_policy(pp, st)) != cudaSuccess) return cuda_fail(env, "ctf_policy_kernel", ce);
      env->launches += 1;
    }
  }
  if ((ce = mg::launch_map(p, st)) != cudaSuccess) return cuda_fail(env, "map_kernel", ce);
  env->launches += 1;
  return 0;
}

extern "C" int mg_set_red_actions(mg_env* env, const int8_t* red_actions_dev) {
  if (!env) return -1;
  if (env->family != MG_FAMILY_CTF) return fail(env, "mg_set_red_actions: CtF family only");
  env->ext_red_actions = red_actions_dev;
  return 0;
}

extern "C" int mg_set_red_policies(mg_env* env, const mg_red_policies* t) {
  if (!env) return -1;
  if (env->family != MG_FAMILY_CTF) return fail(env, "mg_set_red_policies: CtF family only");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  if ((ce = cudaDeviceSynchronize()) != cudaSuccess) return cuda_fail(env, "cudaDeviceSynchronize", ce);   // a launch may still read the old tables
  cudaFree(env->d_policy_tables);
  env->d_policy_tables = nullptr;
  env->fused_red_out = nullptr;   // fusion is re-armed per table set
  if (!t) return 0;
  if (t->struct_size != sizeof(mg_red_policies)) return fail(env, "mg_set_red_policies: mg_red_policies size mismatch (ABI)");
  const mg::MapParams& m = env->mbase;
  if (t->num_red != m.nr) return fail(env, "mg_set_red_policies: num_red differs from the handle's num_red_agents");
  if (!t->first_move || !t->patrol_goal || !t->on_border) return fail(env, "mg_set_red_policies: null table");
  const size_t cells = (size_t)m.cells;
  if (cells > 65535) return fail(env, "mg_set_red_policies: map too large for 16-bit cell indices");
  bool patrols = false;
  mg::PolicyParams p{};
  for (int k = 0; k < m.nr; ++k) {
    if (t->kind[k] < MG_POLICY_RW || t->kind[k] > MG_POLICY_PATROL_FIGHT) return fail(env, "mg_set_red_policies: unknown policy kind");
    if (!(t->randomness[k] >= 0.0 && t->randomness[k] <= 1.0)) return fail(env, "mg_set_red_policies: randomness must be in [0, 1]");
    patrols |= t->kind[k] == MG_POLICY_PATROL || t->kind[k] == MG_POLICY_PATROL_FIGHT;
    p.kind[k] = t->kind[k];
    const double th = std::ceil(t->randomness[k] * 4294967296.0);   // (double)u / 2^32 < r  <=>  u < ceil(r * 2^32)
    p.thr[k] = th <= 0.0 ? 0ull : (th >= 4294967296.0 ? 4294967296ull : (unsigned long long)th);
  }
  if (t->n_along < 0 || (t->n_along > 0 && !t->along_border)) return fail(env, "mg_set_red_policies: bad along_border");
  for (size_t i = 0; i < cells * cells; ++i)
    if (t->first_move[i] > 4) return fail(env, "mg_set_red_policies: first_move holds a value outside CtfActions");
  for (size_t i = 0; i < cells; ++i) {
    if (patrols && t->patrol_goal[i] >= cells) return fail(env, "mg_set_red_policies: patrol_goal outside the map (empty border?)");
    if (patrols && t->on_border[i] && t->n_along == 0) return fail(env, "mg_set_red_policies: a border without patrol candidates (the reference raises there)");
  }
  for (int i = 0; i < t->n_along; ++i)
    if (t->along_border[i] >= cells) return fail(env, "mg_set_red_policies: along_border outside the map");
  const size_t o_fm = 0, o_goal = align_up(cells * cells, 16), o_border = o_goal + align_up(cells * 2, 16),
               o_along = o_border + align_up(cells, 16), total = o_along + align_up((size_t)t->n_along * 2 + 2, 16);
  std::vector<uint8_t> host(total, 0);
  std::memcpy(host.data() + o_fm, t->first_move, cells * cells);
  std::memcpy(host.data() + o_goal, t->patrol_goal, cells * 2);
  std::memcpy(host.data() + o_border, t->on_border, cells);
  if (t->n_along) std::memcpy(host.data() + o_along, t->along_border, (size_t)t->n_along * 2);
  if ((ce = cudaMalloc(&env->d_policy_tables, total)) != cudaSuccess) return cuda_fail(env, "cudaMalloc", ce);
  if ((ce = cudaMemcpy(env->d_policy_tables, host.data(), total, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(env, "policy table upload", ce);
  p.S = m.S; p.cells = m.cells; p.nb = m.nb; p.nr = m.nr; p.n_along = t->n_along;
  p.blue_flag_cell = (m.blue_flag & 255) * m.S + (m.blue_flag >> 8);
  p.N = m.N; p.seed = m.seed; p.env_id_base = m.env_id_base;
  p.field_map = m.field_map;
  p.first_move = env->d_policy_tables + o_fm;
  p.patrol_goal = reinterpret_cast<const uint16_t*>(env->d_policy_tables + o_goal);
  p.on_border = env->d_policy_tables + o_border;
  p.along = reinterpret_cast<const uint16_t*>(env->d_policy_tables + o_along);
  env->pbase = p;
  return 0;
}

extern "C" int mg_set_carry_agent_flags(mg_env* env, int on) {
  if (!env) return -1;
  if (env->family != MG_FAMILY_CTF || env->mbase.variant_1v1) return fail(env, "mg_set_carry_agent_flags: CtFMvN handles only");
  env->mbase.carry_flags = on ? 1 : 0;
  return 0;
}

extern "C" int mg_set_red_policy_fusion(mg_env* env, int8_t* red_actions_dev) {
  if (!env) return -1;
  if (env->family != MG_FAMILY_CTF) return fail(env, "mg_set_red_policy_fusion: CtF family only");
  if (red_actions_dev && !env->d_policy_tables) return fail(env, "mg_set_red_policy_fusion: no policies set (mg_set_red_policies)");
  env->fused_red_out = red_actions_dev;
  return 0;
}

extern "C" int mg_set_policy_trace(mg_env* env, const uint16_t* patrol_target_dev, const uint8_t* follow_dev, const int8_t* action_dev) {
  if (!env) return -1;
  if (env->family != MG_FAMILY_CTF) return fail(env, "mg_set_policy_trace: CtF family only");
  if ((follow_dev == nullptr) != (action_dev == nullptr) || (follow_dev == nullptr) != (patrol_target_dev == nullptr))
    return fail(env, "mg_set_policy_trace: pass all three arrays, or three NULLs to return to Philox draws");
  env->pol_tr_patrol = patrol_target_dev; env->pol_tr_follow = follow_dev; env->pol_tr_action = action_dev;
  return 0;
}

extern "C" int mg_red_policy_actions(mg_env* env, const void* state, int8_t* red_actions_dev, void* stream) {
  if (!env || !state || !red_actions_dev) return fail(env, "mg_red_policy_actions: null argument");
  if (env->family != MG_FAMILY_CTF) return fail(env, "mg_red_policy_actions: CtF family only");
  if (!env->d_policy_tables) return fail(env, "mg_red_policy_actions: no policies set (mg_set_red_policies)");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  mg::PolicyParams p = env->pbase;
  const uint8_t* base = static_cast<const uint8_t*>(state);
  p.agents = base + env->plane_off[MG_MAP_PLANE_AGENTS];
  p.row_bytes = (int)env->plane_row[MG_MAP_PLANE_AGENTS];
  p.hdr = reinterpret_cast<const int4*>(base + env->plane_off[MG_MAP_PLANE_HDR]);
  p.seed = env->mbase.seed;      // mg_set_seed may have re-keyed the handle since the tables were set
  p.out = red_actions_dev;
  p.tr_patrol = env->pol_tr_patrol; p.tr_follow = env->pol_tr_follow; p.tr_action = env->pol_tr_action;
  if ((ce = mg::launch_ctf_policy(p, static_cast<cudaStream_t>(stream))) != cudaSuccess) return cuda_fail(env, "ctf_policy_kernel", ce);
  env->launches += 1;
  return 0;
}

extern "C" int mg_set_partial_obs(mg_env* env, int view_size, int see_through_walls) {
  if (!env) return -1;
  if (env->family != MG_FAMILY_MAZE) return fail(env, "mg_set_partial_obs: Maze family only (Collect / generic handles use mg_gen_obs)");
  if (view_size != 0 && view_size != 3 && view_size != 5 && view_size != 7) return fail(env, "mg_set_partial_obs: view_size must be 0 (off), 3, 5 or 7");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  mg::MapParams& p = env->mbase;
  if ((ce = cudaDeviceSynchronize()) != cudaSuccess) return cuda_fail(env, "cudaDeviceSynchronize", ce);   // a launch may still read the old table
  cudaFree(env->d_view_table);
  env->d_view_table = nullptr; p.view_table = nullptr; p.view_row16 = 0;
  if (view_size) {
    p.map_padded = env->d_map_tables + env->map_padded_off; p.pad = env->map_pad; p.pitch = p.S + 2 * env->map_pad;
    p.map_padded_bytes = (int)env->map_padded_bytes;
    p.view_oob = mg::cell(3, 7, 1); p.view_agent = mg::cell(1, 4, 0);   // as mg_gen_obs: out-of-map filler, Agent(color="blue")
    // Memoised views: the map never changes and the agent is the only moving object, so a view is a pure function of (x, y, dir).  The
    // view kernel - slice, rotations, process_vis, encode - runs ONCE over all S*S*4 agent states; a step then copies its env's 3*V*V
    // bytes out of the table.  Measured (1 M envs, 64x64, V = 7): 57.0 us against 55.6 us for computing every view in the step - the
    // uncoalesced 16-byte row gathers cost the load/store unit what the view arithmetic costs the ALUs - so the table is used only
    // when asked for (MG_VIEW_TABLE=1) or when the padded map is too large to be staged in shared memory (maps beyond ~190x190).
    const char* tv = std::getenv("MG_VIEW_TABLE");
    const bool map_fits = mg::map_view_smem_bytes((int)env->map_padded_bytes, view_size) <= 227 * 1024;
    if ((tv && tv[0] == '1') || (!map_fits && !(tv && tv[0] == '0'))) {
      const int S = p.S, VV3 = 3 * view_size * view_size, row16 = (VV3 + 15) / 16;
      const size_t M = (size_t)S * S * 4;
      std::vector<uint8_t> states(M * 4 + 1024, 0);
      for (int x = 0; x < S; ++x)
        for (int y = 0; y < S; ++y)
          for (int d = 0; d < 4; ++d) {
            uint8_t* q = &states[((((size_t)x * S + y) << 2) | d) * 4];
            q[0] = (uint8_t)x; q[1] = (uint8_t)y; q[2] = (uint8_t)d;
          }
      uint8_t *d_states = nullptr, *d_tmp = nullptr;
      uint4* table = nullptr;
      mg::ViewParams q;
      std::memset(&q, 0, sizeof q);
      q.V = view_size; q.see_through = see_through_walls != 0; q.family = MG_FAMILY_MAZE;
      q.W = q.H = S; q.cells = S * S; q.A = 1; q.N = (long long)M;
      q.map_codes = env->d_map_tables + env->map_codes_off;
      q.map_padded = p.map_padded; q.pad = p.pad; q.pitch = p.pitch; q.map_padded_bytes = p.map_padded_bytes;
      q.oob_code = p.view_oob; q.agent_code = p.view_agent;
      if (mg::view_smem_bytes(q) > 227 * 1024) q.map_padded = nullptr;   // large padded maps: the generic view kernel reads the map through L1
      if ((ce = cudaMalloc(&d_states, states.size())) != cudaSuccess || (ce = cudaMalloc(&d_tmp, M * VV3 + 16)) != cudaSuccess ||
          (ce = cudaMalloc(&table, M * row16 * 16)) != cudaSuccess ||
          (ce = cudaMemcpy(d_states, states.data(), states.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
          (ce = cudaMemset(table, 0, M * row16 * 16)) != cudaSuccess) {
        cudaFree(d_states); cudaFree(d_tmp); cudaFree(table);
        return cuda_fail(env, "mg_set_partial_obs: view table", ce);
      }
      q.pos = d_states; q.pos_stride = 4; q.dirs = d_states + 2; q.dir_stride = 4;
      q.out = d_tmp; q.out_bulk_ok = 1;
      if ((ce = mg::launch_view(q, nullptr)) != cudaSuccess ||
          (ce = cudaMemcpy2D(table, (size_t)row16 * 16, d_tmp, (size_t)VV3, (size_t)VV3, M, cudaMemcpyDeviceToDevice)) != cudaSuccess ||
          (ce = cudaDeviceSynchronize()) != cudaSuccess) {
        cudaFree(d_states); cudaFree(d_tmp); cudaFree(table);
        return cuda_fail(env, "mg_set_partial_obs: view table", ce);
      }
      cudaFree(d_states); cudaFree(d_tmp);
      env->d_view_table = table; p.view_table = table; p.view_row16 = row16;
      env->launches += 1;
    }
    const size_t smem = mg::map_view_smem_bytes(p.view_table ? 0 : (int)env->map_padded_bytes, view_size);
    if (smem > 227 * 1024) return fail(env, "mg_set_partial_obs: padded map does not fit in shared memory");
    if ((ce = mg::configure_map_view_mode(smem)) != cudaSuccess) return cuda_fail(env, "cudaFuncSetAttribute", ce);
  }
  p.view_V = view_size; p.view_see_through = see_through_walls != 0;
  // the observation size changed: mg_step_host re-creates its device staging block on the next call
  cudaFree(env->d_actions); cudaFree(env->d_obs); cudaFree(env->d_final);
  env->d_actions = nullptr; env->d_obs = nullptr; env->d_rewards = nullptr; env->d_term = nullptr; env->d_trunc = nullptr; env->d_final = nullptr;
  return 0;
}

extern "C" int mg_map_info(mg_env* env, const void* state, double* out, void* stream) {
  if (!env || !state || !out) return fail(env, "mg_map_info: null argument");
  if (env->family != MG_FAMILY_MAZE && env->family != MG_FAMILY_CTF) return fail(env, "mg_map_info: Maze and CtF families only");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  mg::MapParams p = env->mbase;
  p.agents = const_cast<uint8_t*>(static_cast<const uint8_t*>(state)) + env->plane_off[MG_MAP_PLANE_AGENTS];
  p.row_bytes = (int)env->plane_row[MG_MAP_PLANE_AGENTS];
  if ((ce = mg::launch_map_info(p, out, static_cast<cudaStream_t>(stream))) != cudaSuccess) return cuda_fail(env, "map_info_kernel", ce);
  env->launches += 1;
  return 0;
}

// length of the "flattened" observation: 3 n + 4 + 2 (|blue_territory| + |red_territory| + |obstacle|), territories include their flag cell
extern "C" int mg_ctf_flat_len(const mg_env* env) {
  if (!env || env->family != MG_FAMILY_CTF) return -1;
  return (int)env->flat_tmpl.size();
}

static int ctf_flat(mg_env* env, const void* state, void* out, int elem, void* stream) {
  if (!env || !state || !out) return fail(env, "mg_ctf_flat_obs: null argument");
  if (env->family != MG_FAMILY_CTF) return fail(env, "mg_ctf_flat_obs: CtF family only");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  const int L = (int)env->flat_tmpl.size();
  if (mg::ctf_flat_tile_envs(L, elem) < 2) return fail(env, "mg_ctf_flat_obs: the map's cell lists are too long for the kernel's shared-memory tile");
  if (!env->d_flat_tmpl) {
    if ((ce = cudaMalloc(&env->d_flat_tmpl, (size_t)L * sizeof(long long))) != cudaSuccess) return cuda_fail(env, "cudaMalloc", ce);
    if ((ce = cudaMemcpy(env->d_flat_tmpl, env->flat_tmpl.data(), (size_t)L * sizeof(long long), cudaMemcpyHostToDevice)) != cudaSuccess)
      return cuda_fail(env, "template upload", ce);
  }
  mg::MapParams p = env->mbase;
  p.agents = const_cast<uint8_t*>(static_cast<const uint8_t*>(state)) + env->plane_off[MG_MAP_PLANE_AGENTS];
  p.row_bytes = (int)env->plane_row[MG_MAP_PLANE_AGENTS];
  if ((ce = mg::launch_ctf_flat(p, env->d_flat_tmpl, L, out, elem, static_cast<cudaStream_t>(stream))) != cudaSuccess)
    return cuda_fail(env, "ctf_flat_kernel", ce);
  env->launches += 1;
  return 0;
}

extern "C" int mg_ctf_flat_obs(mg_env* env, const void* state, int64_t* out, void* stream) { return ctf_flat(env, state, out, 8, stream); }
extern "C" int mg_ctf_flat_obs_u8(mg_env* env, const void* state, uint8_t* out, void* stream) { return ctf_flat(env, state, out, 1, stream); }

// -------------------------------------------------------------------------------- Wildfire
extern "C" int mg_create_wildfire(const mg_wildfire_config* cfg, int device, mg_env** out) {
  if (!cfg || !out) return fail(nullptr, "mg_create_wildfire: null argument");
  *out = nullptr;
  if (cfg->struct_size != sizeof(mg_wildfire_config)) return fail(nullptr, "mg_create_wildfire: mg_wildfire_config size mismatch (ABI)");
  if (cfg->family != MG_FAMILY_WILDFIRE) return fail(nullptr, "mg_create_wildfire: family must be MG_FAMILY_WILDFIRE");
  const int W = cfg->width, H = cfg->height, cells = W * H, A = cfg->num_agents;
  if (cfg->num_envs < 1) return fail(nullptr, "mg_create_wildfire: num_envs must be >= 1");
  if (W < 2 || H < 2 || W > 255 || H > 255 || cells % 16) return fail(nullptr, "mg_create_wildfire: need 2 <= W, H <= 255 and W*H a multiple of 16 (TMA tiles)");
  if (A < 1 || A > MG_MAX_WILDFIRE_AGENTS) return fail(nullptr, "mg_create_wildfire: num_agents must be in [1, 32] (one warp resolves the moves)");
  if (cfg->num_fires < 0 || cfg->num_fires + A > cells) return fail(nullptr, "mg_create_wildfire: num_fires + num_agents exceed the grid");
  if (cfg->max_steps < 1) return fail(nullptr, "mg_create_wildfire: max_steps must be >= 1");
  for (int i = 0; i < A; ++i)
    if (cfg->agent_colour[i] < 0 || cfg->agent_colour[i] > 9) return fail(nullptr, "mg_create_wildfire: agent colour index outside COLORS");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, std::string("mg_create_wildfire: no usable CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(ce));
  if (device < 0 || device >= ndev) return fail(nullptr, "mg_create_wildfire: device index out of range");
  if ((ce = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, "cudaSetDevice", ce);
  cudaDeviceProp prop;
  if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, "cudaGetDeviceProperties", ce);
  if (prop.major != 10) return fail(nullptr, "mg_create_wildfire: kernels are built for sm_100a only");
  if (mg::wildfire_smem_bytes(cells, H) > (size_t)prop.sharedMemPerBlockOptin) return fail(nullptr, "mg_create_wildfire: grid too large for one CTA's shared memory");
  if ((ce = mg::configure_wildfire_kernel(cells, H)) != cudaSuccess) return cuda_fail(nullptr, "cudaFuncSetAttribute", ce);
  mg_env* env = new (std::nothrow) mg_env();
  if (!env) return fail(nullptr, "mg_create_wildfire: out of host memory");
  env->family = MG_FAMILY_WILDFIRE;
  env->wcfg = *cfg;
  env->device = device; env->tile = 0; env->has_trace = false; env->launches = 0; env->timeline = nullptr;
  env->d_actions = nullptr; env->d_obs = nullptr; env->d_rewards = nullptr; env->d_term = nullptr; env->d_trunc = nullptr;
  env->d_final = nullptr; env->d_wall_template = nullptr; env->d_status = nullptr; env->d_map_tables = nullptr;
  env->map_codes_off = 0;
  std::memset(&env->trace, 0, sizeof env->trace);
  env->obs_elem = 1; env->act_cols = A; env->rew_cols = A;
  env->n_pad = cfg->num_envs;
  const size_t rows[MG_WF_PLANE_COUNT] = {(size_t)cells, (size_t)A * 4, 16};
  size_t off = 0;
  for (int i = 0; i < MG_WF_PLANE_COUNT; ++i) {
    env->plane_off[i] = off; env->plane_row[i] = rows[i]; env->plane_bytes[i] = rows[i] * (size_t)env->n_pad;
    off = align_up(off + env->plane_bytes[i], 256);
  }
  env->state_bytes = off;
  if ((ce = cudaMalloc(&env->d_status, sizeof(int32_t))) != cudaSuccess || (ce = cudaMemset(env->d_status, 0, sizeof(int32_t))) != cudaSuccess) {
    cudaFree(env->d_status); delete env; return cuda_fail(nullptr, "cudaMalloc(status)", ce);
  }
  mg::WildfireParams& p = env->wbase;
  std::memset(&p, 0, sizeof p);
  p.W = W; p.H = H; p.cells = cells; p.A = A; p.num_fires = cfg->num_fires; p.max_steps = cfg->max_steps; p.autoreset = cfg->autoreset != 0;
  for (int k = 0; k < 5; ++k) p.ignite_threshold[k] = cfg->ignite_threshold[k];
  p.burnout_threshold = cfg->burnout_threshold;
  p.rw_magic = (H >= 4) ? (uint32_t)(4294967296ull / (unsigned)(H / 4)) + 1u : 0u;
  p.rv_magic = (H % 16 == 0) ? (uint32_t)(4294967296ull / (unsigned)(H / 16)) + 1u : p.rw_magic;
  for (int i = 0; i < A; ++i) p.agent_colour[i] = (uint8_t)cfg->agent_colour[i];
  p.N = cfg->num_envs; p.env_id_base = (unsigned long long)cfg->env_id_base; p.seed = cfg->seed;
  *out = env;
  return 0;
}

static int wildfire_launch(mg_env* env, void* state, int op, const mg_step_io* io, const uint8_t* mask, uint8_t* obs, cudaStream_t st) {
  mg::WildfireParams p = env->wbase;
  uint8_t* s = static_cast<uint8_t*>(state);
  p.terrain = s + env->plane_off[MG_WF_PLANE_TERRAIN]; p.agents = s + env->plane_off[MG_WF_PLANE_AGENTS];
  p.hdr = reinterpret_cast<int4*>(s + env->plane_off[MG_WF_PLANE_HDR]);
  p.op = op; p.reset_mask = mask;
  p.order = env->has_trace ? env->trace.order : nullptr;
  if (op == 1) {
    p.actions = io->actions; p.obs = io->obs; p.rewards = io->rewards; p.terminated = io->terminated;
    p.truncated = io->truncated; p.final_obs = io->final_obs;
  } else {
    p.obs = obs;
  }
  if ((p.obs && !aligned16(p.obs)) || (p.final_obs && !aligned16(p.final_obs))) return fail(env, "obs buffers must be 16-byte aligned");
  cudaError_t ce;
  if ((ce = mg::launch_wildfire(p, st)) != cudaSuccess) return cuda_fail(env, "wildfire_kernel", ce);
  env->launches += 1;
  return 0;
}

// --------------------------------------------------------------------------------- generic
extern "C" int mg_create_generic(const mg_generic_config* cfg, int device, mg_env** out) {
  if (!cfg || !out) return fail(nullptr, "mg_create_generic: null argument");
  *out = nullptr;
  if (cfg->struct_size != sizeof(mg_generic_config)) return fail(nullptr, "mg_create_generic: mg_generic_config size mismatch (ABI)");
  if (cfg->family != MG_FAMILY_GENERIC) return fail(nullptr, "mg_create_generic: family must be MG_FAMILY_GENERIC");
  const int W = cfg->width, H = cfg->height, A = cfg->num_agents, cells = W * H;
  if (cfg->num_envs < 1 || W < 3 || H < 3 || W > 255 || H > 255 || A < 1 || A > 8 || cfg->max_steps < 1)
    return fail(nullptr, "mg_create_generic: need num_envs >= 1, 3 <= W, H <= 255 (grid.py:19-20), 1 <= num_agents <= 8, max_steps >= 1");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, std::string("mg_create_generic: no usable CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(ce));
  if (device < 0 || device >= ndev) return fail(nullptr, "mg_create_generic: device index out of range");
  if ((ce = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, "cudaSetDevice", ce);
  cudaDeviceProp prop;
  if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, "cudaGetDeviceProperties", ce);
  if (prop.major != 10) return fail(nullptr, "mg_create_generic: kernels are built for sm_100a only");
  if ((size_t)mg::generic_tile_envs() * A * cells >= 65536 || mg::generic_smem_bytes(A, cells) > (size_t)prop.sharedMemPerBlockOptin)
    return fail(nullptr, "mg_create_generic: num_agents * width * height too large for one tile (8 envs' observations must fit in shared memory)");
  if ((ce = mg::configure_generic_kernel(A, cells)) != cudaSuccess) return cuda_fail(nullptr, "cudaFuncSetAttribute", ce);
  mg_env* env = new (std::nothrow) mg_env();
  if (!env) return fail(nullptr, "mg_create_generic: out of host memory");
  env->family = MG_FAMILY_GENERIC;
  env->gcfg = *cfg;
  env->device = device; env->tile = 0; env->has_trace = false; env->launches = 0; env->timeline = nullptr;
  env->d_actions = nullptr; env->d_obs = nullptr; env->d_rewards = nullptr; env->d_term = nullptr; env->d_trunc = nullptr;
  env->d_final = nullptr; env->d_wall_template = nullptr; env->d_status = nullptr; env->d_map_tables = nullptr;
  env->map_codes_off = 0;
  std::memset(&env->trace, 0, sizeof env->trace);
  env->obs_elem = 1; env->act_cols = A; env->rew_cols = A;
  const int E = mg::generic_tile_envs();
  env->n_pad = (cfg->num_envs + E - 1) / E * E;
  const size_t rows[MG_GEN_PLANE_COUNT] = {(size_t)cells, (size_t)cells, (size_t)A * 2, 16, (size_t)cells, (size_t)cells, (size_t)A * 2};
  size_t off = 0;
  for (int i = 0; i < MG_GEN_PLANE_COUNT; ++i) {
    env->plane_off[i] = off; env->plane_row[i] = rows[i]; env->plane_bytes[i] = rows[i] * (size_t)env->n_pad;
    off = align_up(off + env->plane_bytes[i], 256);
  }
  env->state_bytes = off;
  if ((ce = cudaMalloc(&env->d_status, sizeof(int32_t))) != cudaSuccess || (ce = cudaMemset(env->d_status, 0, sizeof(int32_t))) != cudaSuccess) {
    cudaFree(env->d_status); delete env; return cuda_fail(nullptr, "cudaMalloc(status)", ce);
  }
  mg::GenericParams& p = env->gbase;
  std::memset(&p, 0, sizeof p);
  p.W = W; p.H = H; p.cells = cells; p.A = A; p.max_steps = cfg->max_steps; p.autoreset = cfg->autoreset != 0;
  p.cells_magic = (uint32_t)(4294967296ull / (unsigned)cells) + 1u; p.per_env_magic = (uint32_t)(4294967296ull / (unsigned)(A * cells)) + 1u;
  p.half_magic = cells >= 2 ? (uint32_t)(4294967296ull / (unsigned)(cells / 2)) + 1u : 0u;
  p.N = cfg->num_envs; p.env_id_base = (unsigned long long)cfg->env_id_base; p.seed = cfg->seed; p.status = env->d_status;
  *out = env;
  return 0;
}

static int generic_launch(mg_env* env, void* state, int op, const mg_step_io* io, const uint8_t* mask, uint8_t* obs, cudaStream_t st) {
  mg::GenericParams p = env->gbase;
  uint8_t* s = static_cast<uint8_t*>(state);
  p.gcell = s + env->plane_off[MG_GEN_PLANE_CELL]; p.gstate = s + env->plane_off[MG_GEN_PLANE_STATE];
  p.pos = s + env->plane_off[MG_GEN_PLANE_POS]; p.hdr = reinterpret_cast<int4*>(s + env->plane_off[MG_GEN_PLANE_HDR]);
  p.icell = s + env->plane_off[MG_GEN_PLANE_INIT_CELL]; p.istate = s + env->plane_off[MG_GEN_PLANE_INIT_STATE];
  p.ipos = s + env->plane_off[MG_GEN_PLANE_INIT_POS];
  p.op = op; p.reset_mask = mask;
  p.order = env->has_trace ? env->trace.order : nullptr;
  if (op == 1) {
    p.actions = io->actions; p.obs = io->obs; p.rewards = io->rewards; p.terminated = io->terminated;
    p.truncated = io->truncated; p.final_obs = io->final_obs;
  } else {
    p.obs = obs;
  }
  if ((reinterpret_cast<uintptr_t>(p.obs) & 1) || (reinterpret_cast<uintptr_t>(p.final_obs) & 1)) return fail(env, "obs buffers must be 2-byte aligned");
  cudaError_t ce;
  if ((ce = mg::launch_generic(p, st)) != cudaSuccess) return cuda_fail(env, "generic_kernel", ce);
  env->launches += 1;
  return 0;
}

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
  if (env->family == MG_FAMILY_WILDFIRE) return wildfire_launch(env, state, 0, nullptr, mask, obs, static_cast<cudaStream_t>(stream));
  if (env->family == MG_FAMILY_GENERIC) return generic_launch(env, state, 0, nullptr, mask, obs, static_cast<cudaStream_t>(stream));
  if (env->family != MG_FAMILY_COLLECT) return map_launch(env, state, 0, nullptr, mask, obs, static_cast<cudaStream_t>(stream));
  mg::CollectParams p = env->base;
  bind_state(env, p, state);
  bind_trace(env, p);
  env->mirror_valid = false;   // the delta transport's host mirror no longer matches the state
  p.reset_mask = mask; p.obs = obs; p.obs_bulk_ok = aligned16(obs);
  if ((ce = mg::launch_collect_reset(env->tile, p, static_cast<cudaStream_t>(stream))) != cudaSuccess) return cuda_fail(env, "collect_reset_kernel", ce);
  env->launches += 1;
  return 0;
}

// the warp-tile kernel needs 2 warps' worth of shared memory per CTA: opt in once per handle; != 0 if the grid is too large for it
static int rollout_prepare(mg_env* env) {
  if (env->rollout_ready) return 0;
  if (mg::rollout_smem_bytes(env->base.cells, env->base.A, 32, true) > env->smem_optin) return -1;
  if (mg::configure_rollout_kernels(env->base.cells, env->base.A) != cudaSuccess) return -1;
  env->rollout_ready = true;
  return 0;
}

// the Collect step's kernel parameter block for one call (everything the launch depends on besides env->step_impl / env->tile)
static int collect_step_params(mg_env* env, void* state, const mg_step_io* io, bool with_delta, mg::CollectParams& p) {
  p = env->base;
  bind_state(env, p, state);
  bind_trace(env, p);
  if (p.rng_mode == 0 && !p.order) return fail(env, "mg_step: trace mode needs the recorded agent order");
  p.actions = io->actions; p.obs = io->obs; p.rewards = io->rewards;
  p.terminated = io->terminated; p.truncated = io->truncated; p.final_obs = io->final_obs;
  p.obs_bulk_ok = aligned16(io->obs);
  p.io_bulk_ok = aligned16(io->actions) && aligned16(io->rewards) && aligned16(io->terminated) && aligned16(io->truncated);
  p.timeline = env->timeline;
  if (with_delta) {
    p.delta = env->d_delta_blk + 16; p.delta_stride = mg::delta_record_bytes(p.cells, p.A); p.delta_wide = p.cells > 256;
    p.reset_count = reinterpret_cast<int32_t*>(env->d_delta_blk);
    p.reset_rows = env->d_reset_rows; p.reset_stride = (int)env->reset_stride;
  }
  return 0;
}

static int step_device(mg_env* env, void* state, const mg_step_io* io, cudaStream_t st, bool with_delta = false) {
  if (!io->actions || !io->rewards || !io->terminated || !io->truncated) return fail(env, "mg_step: actions/rewards/terminated/truncated must be non-null");
  if (io->final_obs && !io->obs) return fail(env, "mg_step: final_obs needs obs");
  if (env->family == MG_FAMILY_WILDFIRE) return wildfire_launch(env, state, 1, io, nullptr, nullptr, st);
  if (env->family == MG_FAMILY_GENERIC) return generic_launch(env, state, 1, io, nullptr, nullptr, st);
  if (env->family != MG_FAMILY_COLLECT) return map_launch(env, state, 1, io, nullptr, nullptr, st);
  mg::CollectParams p;
  if (collect_step_params(env, state, io, with_delta, p)) return -1;
  cudaError_t ce;
  if (env->step_impl == 1 && rollout_prepare(env) == 0) {
    p.T = 1;
    if ((ce = mg::launch_collect_rollout(p, env->num_sms, st)) != cudaSuccess) return cuda_fail(env, "collect_rollout_kernel", ce);
  } else {
    if ((ce = mg::launch_collect_step(env->tile, p, st)) != cudaSuccess) return cuda_fail(env, "collect_step_kernel", ce);
  }
  env->launches += 1;
  return 0;
}

extern "C" int mg_rollout(mg_env* env, void* state, const mg_rollout_io* io, void* stream) {
  if (!env || !state || !io) return fail(env, "mg_rollout: null argument");
  if (env->family != MG_FAMILY_COLLECT) return fail(env, "mg_rollout: Collect family only");
  if (!aligned16(state)) return fail(env, "mg_rollout: state buffer must be 16-byte aligned");
  if (io->steps < 1) return fail(env, "mg_rollout: steps must be >= 1");
  if (!io->rewards || !io->terminated || !io->truncated) return fail(env, "mg_rollout: rewards/terminated/truncated must be non-null");
  if (io->final_obs && !io->obs) return fail(env, "mg_rollout: final_obs needs obs");
  if (env->has_trace && io->steps != 1) return fail(env, "mg_rollout: trace replay (mg_set_trace) carries one step of recorded draws: steps must be 1");
  if (env->has_trace && !io->actions) return fail(env, "mg_rollout: trace replay needs actions");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  if (rollout_prepare(env)) return fail(env, "mg_rollout: the grid is too large for the warp-tile kernel's shared memory (32 envs x 4*W*H bytes per warp)");
  env->mirror_valid = false;
  mg::CollectParams p = env->base;
  bind_state(env, p, state);
  bind_trace(env, p);
  if (p.rng_mode == 0 && !p.order) return fail(env, "mg_rollout: trace mode needs the recorded agent order");
  p.T = io->steps;
  p.actions = io->actions; p.actions_out = io->actions_out; p.obs = io->obs; p.rewards = io->rewards;
  p.terminated = io->terminated; p.truncated = io->truncated; p.final_obs = io->final_obs;
  const bool strides_ok = io->steps == 1 || (p.N % 16 == 0);   // step-major arrays: every step's slab must start 16-byte aligned for TMA
  p.obs_bulk_ok = aligned16(io->obs) && strides_ok;
  p.io_bulk_ok = (!io->actions || aligned16(io->actions)) && aligned16(io->rewards) && aligned16(io->terminated) && aligned16(io->truncated) && strides_ok;
  p.timeline = env->timeline;
  if ((ce = mg::launch_collect_rollout(p, env->num_sms, static_cast<cudaStream_t>(stream))) != cudaSuccess) return cuda_fail(env, "collect_rollout_kernel", ce);
  env->launches += 1;
  return 0;
}

extern "C" int mg_step(mg_env* env, void* state, const mg_step_io* io, void* stream) {
  if (!env || !state || !io) return fail(env, "mg_step: null argument");
  if (!aligned16(state)) return fail(env, "mg_step: state buffer must be 16-byte aligned");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  env->mirror_valid = false;
  return step_device(env, state, io, static_cast<cudaStream_t>(stream));
}

extern "C" int mg_encode(mg_env* env, const void* state, uint8_t* obs, void* stream) {
  if (!env || !state || !obs) return fail(env, "mg_encode: null argument");
  if (env->family != MG_FAMILY_COLLECT) return fail(env, "mg_encode: Collect family only (Grid.encode); Maze/CtF observations come from mg_reset/mg_step");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  const uint8_t* grid = static_cast<const uint8_t*>(state) + env->plane_off[MG_PLANE_GRID];
  if ((ce = mg::launch_encode3(env->tile, grid, obs, env->cfg.num_envs, env->cfg.width * env->cfg.height, aligned16(obs),
                               static_cast<cudaStream_t>(stream))) != cudaSuccess)
    return cuda_fail(env, "encode3_kernel", ce);
  env->launches += 1;
  return 0;
}

extern "C" int mg_gen_obs(mg_env* env, const void* state, const uint8_t* dirs, int view_size, int see_through_walls,
                          uint8_t* out, void* stream) {
  if (!env || !state || !out) return fail(env, "mg_gen_obs: null argument");
  if (env->family == MG_FAMILY_CTF || env->family == MG_FAMILY_WILDFIRE) return fail(env, "mg_gen_obs: Collect, Maze and generic families only");
  if (view_size < 1 || view_size > mg::view_max()) return fail(env, "mg_gen_obs: view_size must be in [1, 15]");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  if (env->family == MG_FAMILY_GENERIC) {   // DefaultWorld, encode_dim 6: out u8 [N][A][V][V][6]
    if (reinterpret_cast<uintptr_t>(out) & 1) return fail(env, "mg_gen_obs: out must be 2-byte aligned");
    mg::View6Params q;
    std::memset(&q, 0, sizeof q);
    const uint8_t* gsb = static_cast<const uint8_t*>(state);
    q.W = env->gcfg.width; q.H = env->gcfg.height; q.cells = q.W * q.H; q.A = env->gcfg.num_agents; q.N = env->gcfg.num_envs;
    q.V = view_size; q.see_through = see_through_walls != 0;
    q.gcell = gsb + env->plane_off[MG_GEN_PLANE_CELL]; q.gstate = gsb + env->plane_off[MG_GEN_PLANE_STATE];
    q.pos = gsb + env->plane_off[MG_GEN_PLANE_POS]; q.dirs = dirs; q.out = out;
    if ((ce = mg::launch_view6(q, static_cast<cudaStream_t>(stream))) != cudaSuccess) return cuda_fail(env, "view6_kernel", ce);
    env->launches += 1;
    return 0;
  }
  mg::ViewParams p;
  std::memset(&p, 0, sizeof p);
  const uint8_t* s = static_cast<const uint8_t*>(state);
  p.V = view_size; p.see_through = see_through_walls != 0; p.family = env->family;
  if (env->family == MG_FAMILY_COLLECT) {
    p.W = env->cfg.width; p.H = env->cfg.height; p.A = env->cfg.num_agents; p.N = env->cfg.num_envs;
    p.grid = s + env->plane_off[MG_PLANE_GRID]; p.pos = s + env->plane_off[MG_PLANE_AGENT_POS]; p.pos_stride = 2;
    p.dirs = dirs; p.dir_stride = 1;                // NULL = 3 for every agent
    p.oob_code = mg::WALL_GREY;                     // Grid.slice: Wall(self.world) outside the grid (grid.py:124-127)
  } else {
    p.W = p.H = env->mcfg.size; p.A = 1; p.N = env->mcfg.num_envs;
    p.pos = s + env->plane_off[MG_MAP_PLANE_AGENTS]; p.pos_stride = (int)env->plane_row[MG_MAP_PLANE_AGENTS];
    p.dirs = dirs ? dirs : p.pos + 2; p.dir_stride = dirs ? 1 : p.pos_stride;
    p.map_codes = env->d_map_tables + env->map_codes_off;
    p.map_padded = env->d_map_tables + env->map_padded_off; p.pad = env->map_pad; p.pitch = p.W + 2 * env->map_pad;
    p.map_padded_bytes = (int)env->map_padded_bytes;
    p.oob_code = mg::cell(3, 7, 1);                 // extension: MazeWorld has no wall (world.py:81-91); an opaque obstacle-coloured filler, state 1 marks it
    p.agent_code = mg::cell(1, 4, 0);               // Agent(color="blue", type="agent") maze.py:93-101
  }
  p.cells = p.W * p.H;
  p.out = out; p.out_bulk_ok = aligned16(out);
  if (mg::view_smem_bytes(p) > 227 * 1024) { p.map_padded = nullptr; }  // large padded maps: the generic kernel reads the map through L1
  if (mg::view_smem_bytes(p) > 227 * 1024) return fail(env, "mg_gen_obs: view tile does not fit in shared memory (reduce view_size)");
  if ((ce = mg::launch_view(p, static_cast<cudaStream_t>(stream))) != cudaSuccess) return cuda_fail(env, "view_kernel", ce);
  env->launches += 1;
  return 0;
}

extern "C" int mg_render(mg_env* env, const void* state, const int32_t* env_ids, int n, int tile_size, uint8_t* out, void* stream) {
  if (!env || !state || !out) return fail(env, "mg_render: null argument");
  if (env->family != MG_FAMILY_COLLECT && env->family != MG_FAMILY_MAZE && env->family != MG_FAMILY_CTF)
    return fail(env, "mg_render: Collect, Maze and CtF families only (the generic family draws DefaultWorld objects; Wildfire has no reference renderer)");
  if (n < 0 || tile_size < 1 || tile_size > 64) return fail(env, "mg_render: n >= 0 and 1 <= tile_size <= 64");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  const uint8_t* atlas = nullptr;
  for (auto& a : env->atlases) if (a.first == tile_size) atlas = a.second;
  if (!atlas) {   // first frame at this tile size: rasterise the tiles on the host (Grid.render_tile's cache, grid.py:147-150) and upload
    std::vector<uint8_t> host;
    mg::build_render_atlas(env->family == MG_FAMILY_COLLECT ? 0 : (env->family == MG_FAMILY_MAZE ? 1 : 2), tile_size, host);
    uint8_t* d = nullptr;
    if ((ce = cudaMalloc(&d, host.size())) != cudaSuccess) return cuda_fail(env, "cudaMalloc atlas", ce);
    if ((ce = cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice)) != cudaSuccess) { cudaFree(d); return cuda_fail(env, "atlas upload", ce); }
    env->atlases.emplace_back(tile_size, d);
    atlas = d;
  }
  const uint8_t* s = static_cast<const uint8_t*>(state);
  if (env->family == MG_FAMILY_COLLECT)
    ce = mg::launch_render(s + env->plane_off[MG_PLANE_GRID], nullptr, 0, 0, 0, 0, 0, env->cfg.num_envs, env_ids, n, env->cfg.width,
                           env->cfg.height, tile_size, atlas, out, env->d_status, static_cast<cudaStream_t>(stream));
  else
    ce = mg::launch_render(env->mbase.field_map, s + env->plane_off[MG_MAP_PLANE_AGENTS], (int)env->plane_row[MG_MAP_PLANE_AGENTS],
                           env->family == MG_FAMILY_MAZE ? 1 : 2, env->mbase.n, env->mbase.nb, env->mbase.variant_1v1, env->mcfg.num_envs,
                           env_ids, n, env->mcfg.size, env->mcfg.size, tile_size, atlas, out, env->d_status, static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) return cuda_fail(env, "render_kernel", ce);
  if (n > 0) env->launches += 1;
  return 0;
}

extern "C" int mg_toroid_obs(mg_env* env, const void* state, float* out, void* stream) {
  if (!env || !state || !out) return fail(env, "mg_toroid_obs: null argument");
  if (env->family != MG_FAMILY_COLLECT) return fail(env, "mg_toroid_obs: Collect family only (wrappers/toroid.py wraps CollectGameEnv)");
  if (env->cfg.width != env->cfg.height) return fail(env, "mg_toroid_obs: square grids only (the reference indexes [y][x] into a (W, H) array)");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  const uint8_t* s = static_cast<const uint8_t*>(state);
  if ((ce = mg::launch_toroid(s + env->plane_off[MG_PLANE_GRID], s + env->plane_off[MG_PLANE_AGENT_POS], out, env->cfg.num_envs,
                              env->cfg.width, env->cfg.num_agents, env->cfg.num_ball_types, static_cast<cudaStream_t>(stream))) != cudaSuccess)
    return cuda_fail(env, "toroid_kernel", ce);
  env->launches += 1;
  return 0;
}

static size_t env_count(const mg_env* env) {
  return (size_t)(env->family == MG_FAMILY_COLLECT ? env->cfg.num_envs : env->family == MG_FAMILY_WILDFIRE ? env->wcfg.num_envs
                  : env->family == MG_FAMILY_GENERIC ? env->gcfg.num_envs : env->mcfg.num_envs);
}
static void host_layout(const mg_env* env, size_t* off_rewards, size_t* off_term, size_t* off_trunc, size_t* total) {
  const size_t N = env_count(env), R = (size_t)env->rew_cols;
  *off_rewards = align_up(mg_obs_bytes(env), 256);
  *off_term = align_up(*off_rewards + N * R * sizeof(double), 256);
  *off_trunc = align_up(*off_term + N, 256);
  *total = align_up(*off_trunc + N, 256);
}
extern "C" int mg_host_layout(const mg_env* env, size_t* off_rewards, size_t* off_terminated, size_t* off_truncated, size_t* total_bytes) {
  if (!env || !off_rewards || !off_terminated || !off_truncated || !total_bytes) return -1;
  host_layout(env, off_rewards, off_terminated, off_truncated, total_bytes);
  return 0;
}

// ---- compact host transports (Collect family) ----------------------------------------------------------------------
static int transport_alloc(mg_env* env) {
  if (env->d_delta_blk) return 0;
  const size_t N = (size_t)env->cfg.num_envs, cells = (size_t)env->base.cells;
  const size_t R = (size_t)mg::delta_record_bytes(env->base.cells, env->base.A);
  env->reset_stride = align_up(cells, 4) + 4;
  cudaError_t ce;
  if ((ce = cudaMalloc(&env->d_delta_blk, 16 + N * R)) != cudaSuccess ||
      (ce = cudaMalloc(&env->d_reset_rows, N * env->reset_stride)) != cudaSuccess ||
      (ce = cudaHostAlloc(&env->h_delta_blk, 16 + N * R, cudaHostAllocDefault)) != cudaSuccess ||
      (ce = cudaHostAlloc(&env->h_reset_rows, N * env->reset_stride, cudaHostAllocDefault)) != cudaSuccess ||
      (ce = cudaHostAlloc(&env->h_grid, N * cells, cudaHostAllocDefault)) != cudaSuccess) {
    cudaFree(env->d_delta_blk); cudaFree(env->d_reset_rows);
    cudaFreeHost(env->h_delta_blk); cudaFreeHost(env->h_reset_rows); cudaFreeHost(env->h_grid);
    env->d_delta_blk = env->d_reset_rows = env->h_delta_blk = env->h_reset_rows = env->h_grid = nullptr;
    return cuda_fail(env, "mg_set_host_transport: staging buffers", ce);
  }
  return 0;
}

extern "C" int mg_set_host_transport(mg_env* env, int mode, int host_threads) {
  if (!env) return -1;
  if (mode != MG_TRANSPORT_FULL && mode != MG_TRANSPORT_PACKED && mode != MG_TRANSPORT_DELTA) return fail(env, "mg_set_host_transport: unknown mode");
  if (mode != MG_TRANSPORT_FULL && env->family != MG_FAMILY_COLLECT) return fail(env, "mg_set_host_transport: the compact transports exist for the Collect family only");
  if (env->pend.active) return fail(env, "mg_set_host_transport: a host step is still in flight (mg_step_host_wait first)");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  // (the staging buffers - 220 bytes per env, half of it page-locked - are allocated by the first host step, not here: most handles
  //  are only ever stepped with device tensors)
  env->host_threads = host_threads > 0 ? host_threads : mg::host_default_threads();
  if (mode != MG_TRANSPORT_FULL) mg::host_pool_ensure(env->host_threads);
  env->transport = mode;
  env->mirror = nullptr; env->mirror_valid = false;
  return 0;
}

extern "C" int mg_host_invalidate(mg_env* env) {
  if (!env) return -1;
  env->mirror_valid = false;
  return 0;
}

extern "C" int mg_host_expand_plane(const uint8_t* grid, uint8_t* obs, size_t n_cells, int host_threads) {
  if (!grid || !obs) return -1;
  const int t = host_threads > 0 ? host_threads : mg::host_default_threads();
  mg::host_pool_ensure(t);
  mg::host_expand_plane(grid, obs, n_cells, t);
  return 0;
}

extern "C" int mg_host_apply_delta(const uint8_t* records, size_t n, int cells, int num_agents, const double* reward_table, uint8_t* obs,
                                   double* rewards, uint8_t* terminated, uint8_t* truncated, uint8_t* final_obs, int host_threads) {
  if (!records || !reward_table || cells < 1 || num_agents < 1 || num_agents > MG_MAX_AGENTS) return -1;
  const int t = host_threads > 0 ? host_threads : mg::host_default_threads();
  mg::host_pool_ensure(t);
  mg::HostDeltaJob j;
  j.records = records; j.n = n; j.stride = mg::delta_record_bytes(cells, num_agents); j.wide = cells > 256; j.cells = cells; j.A = num_agents;
  j.reward_table = reward_table; j.obs = obs; j.skip_patches = 0; j.rewards = rewards; j.terminated = terminated; j.truncated = truncated;
  j.final_obs = final_obs;
  mg::host_apply_delta(j, t);
  return 0;
}

extern "C" int mg_delta_record_bytes(int cells, int num_agents) { return mg::delta_record_bytes(cells, num_agents); }

static bool host_graph_enabled() {
  static const bool on = [] { const char* v = std::getenv("MG_HOST_GRAPH"); return !(v && v[0] == '0'); }();
  return on;
}
// H2D actions -> step -> D2H of the step's compact results into the handle's page-locked staging; decoded by finish_host_step
static int step_host_compact(mg_env* env, void* state, const mg_step_io* io, cudaStream_t st) {
  if (!aligned16(state)) return fail(env, "mg_step_host: state buffer must be 16-byte aligned");
  if (transport_alloc(env)) return -1;
  const size_t N = (size_t)env->cfg.num_envs, A = (size_t)env->act_cols, cells = (size_t)env->base.cells;
  const size_t R = (size_t)mg::delta_record_bytes(env->base.cells, env->base.A);
  cudaError_t ce;
  if (!env->d_actions) {
    size_t off_r, off_t, off_u, total;
    host_layout(env, &off_r, &off_t, &off_u, &total);
    int8_t* da = nullptr; uint8_t* dobs = nullptr;
    if ((ce = cudaMalloc(&da, N * A)) != cudaSuccess) return cuda_fail(env, "cudaMalloc", ce);
    if ((ce = cudaMalloc(&dobs, total)) != cudaSuccess) { cudaFree(da); return cuda_fail(env, "cudaMalloc", ce); }
    if ((ce = cudaMemsetAsync(dobs, 0, total, st)) != cudaSuccess) { cudaFree(da); cudaFree(dobs); return cuda_fail(env, "cudaMemsetAsync", ce); }
    env->d_actions = da; env->d_obs = dobs;
    env->d_rewards = reinterpret_cast<double*>(dobs + off_r); env->d_term = dobs + off_t; env->d_trunc = dobs + off_u;
  }
  const bool delta = env->transport == MG_TRANSPORT_DELTA;
  // the delta transport patches the caller's obs buffer in place: it must be the buffer the previous host step filled
  const bool refresh = delta && io->obs && (!env->mirror_valid || env->mirror != io->obs);
  const bool want_plane = io->obs && (!delta || refresh);
  const bool dev_final = io->final_obs && (!delta || refresh);   // no valid mirror to take the terminal rows from
  const size_t ob = mg_obs_bytes(env);
  if (dev_final && !env->d_final) {
    if ((ce = cudaMalloc(&env->d_final, ob)) != cudaSuccess) return cuda_fail(env, "cudaMalloc", ce);
    if ((ce = cudaMemsetAsync(env->d_final, 0, ob, st)) != cudaSuccess) return cuda_fail(env, "cudaMemsetAsync", ce);
  }
  mg_step_io dio;
  dio.actions = env->d_actions; dio.rewards = env->d_rewards; dio.terminated = env->d_term; dio.truncated = env->d_trunc;
  dio.obs = dev_final ? static_cast<uint8_t*>(env->d_obs) : nullptr;   // the expanded observation stays off the wire
  dio.final_obs = dev_final ? env->d_final : nullptr;
  // steady state of the delta transport: the four stream operations never change from step to step - one graph launch
  bool capturing = false;
  if (delta && !want_plane && !dev_final && host_graph_enabled() && env->hgraph.failures < 3) {
    mg::CollectParams p;
    if (collect_step_params(env, state, &dio, true, p)) return -1;
    auto& g = env->hgraph;
    if (g.exec && g.state == state && g.h_actions == io->actions && g.step_impl == env->step_impl && g.tile == env->tile &&
        std::memcmp(&g.p, &p, sizeof p) == 0) {
      if ((ce = cudaGraphLaunch(g.exec, st)) != cudaSuccess) return cuda_fail(env, "cudaGraphLaunch", ce);
      env->launches += 1;
      env->pend.active = true; env->pend.refresh = false; env->pend.plane = false; env->pend.io = *io;
      return 0;
    }
    cudaPointerAttributes pa;   // pageable actions cannot be captured (and are not asynchronous either)
    const bool pinned = cudaPointerGetAttributes(&pa, io->actions) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();
    if (pinned && env->step_impl == 1 && rollout_prepare(env) == 0) {
      if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        capturing = true;
        g.state = state; g.h_actions = io->actions; g.step_impl = env->step_impl; g.tile = env->tile; g.p = p;
      } else {
        (void)cudaGetLastError();
        g.failures += 1;
      }
    }
  }
  auto end_capture = [&](bool ok) -> int {   // closes the capture; on success instantiates and launches the graph
    if (!capturing) return 0;
    capturing = false;
    cudaGraph_t graph = nullptr;
    cudaError_t e2 = cudaStreamEndCapture(st, &graph);
    if (ok && e2 == cudaSuccess && graph) {
      e2 = cudaGraphInstantiate(&env->hgraph.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e2 == cudaSuccess && (e2 = cudaGraphLaunch(env->hgraph.exec, st)) == cudaSuccess) return 0;
      if (env->hgraph.exec) { cudaGraphExecDestroy(env->hgraph.exec); env->hgraph.exec = nullptr; }
    } else if (graph) {
      cudaGraphDestroy(graph);
    }
    (void)cudaGetLastError();
    env->hgraph.exec = nullptr;
    env->hgraph.failures += 1;
    if (std::getenv("MG_DEBUG")) std::fprintf(stderr, "multigrid_b200: host-step graph capture failed (%s), plain launches instead\n", cudaGetErrorString(e2));
    return -1;
  };
  for (int attempt = 0; attempt < 2; ++attempt) {   // a failed capture is followed by one plain pass
    const bool was_capturing = capturing;
    bool ok = true;
    if ((ce = cudaMemcpyAsync(env->d_actions, io->actions, N * A, cudaMemcpyHostToDevice, st)) != cudaSuccess) ok = false;
    if (ok && delta && (ce = cudaMemsetAsync(env->d_delta_blk, 0, 16, st)) != cudaSuccess) ok = false;
    if (!ok && !was_capturing) return cuda_fail(env, "H2D actions / header clear", ce);
    if (ok && step_device(env, state, &dio, st, delta)) { if (!was_capturing) return -1; ok = false; }
    if (ok && delta && (ce = cudaMemcpyAsync(env->h_delta_blk, env->d_delta_blk, 16 + N * R, cudaMemcpyDeviceToHost, st)) != cudaSuccess) {
      if (!was_capturing) return cuda_fail(env, "D2H delta records", ce);
      ok = false;
    }
    if (!was_capturing) break;
    if (end_capture(ok) == 0) break;   // captured, instantiated and launched
  }
  if (delta) {
  } else {
    size_t off_r, off_t, off_u, total;
    host_layout(env, &off_r, &off_t, &off_u, &total);
    const uint8_t* hr = reinterpret_cast<const uint8_t*>(io->rewards);
    if (io->terminated == hr + (off_t - off_r) && io->truncated == hr + (off_u - off_r)) {   // one block on the host as well: one copy
      if ((ce = cudaMemcpyAsync(io->rewards, env->d_rewards, total - off_r, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H rewards / flags", ce);
    } else {
      if ((ce = cudaMemcpyAsync(io->rewards, env->d_rewards, N * (size_t)env->rew_cols * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H rewards", ce);
      if ((ce = cudaMemcpyAsync(io->terminated, env->d_term, N, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H terminated", ce);
      if ((ce = cudaMemcpyAsync(io->truncated, env->d_trunc, N, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H truncated", ce);
    }
  }
  if (want_plane) {
    const uint8_t* grid = static_cast<const uint8_t*>(state) + env->plane_off[MG_PLANE_GRID];
    if ((ce = cudaMemcpyAsync(env->h_grid, grid, N * cells, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H grid plane", ce);
  }
  if (dev_final && (ce = cudaMemcpyAsync(io->final_obs, env->d_final, ob, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H final_obs", ce);
  env->pend.active = true; env->pend.refresh = refresh; env->pend.plane = want_plane; env->pend.io = *io;
  return 0;
}

// ---- asynchronous decode ---------------------------------------------------------------------------------------------
// A step enqueued by mg_step_host_async is watched by the host pool's idle threads (cudaEventQuery on its event); the thread that
// sees it land turns the decode into pool tasks, which the workers run while the caller's thread enqueues its other env batches:
// the host half of step k overlaps the device half of step k + 1 AND the caller's own work.  Rows of envs that autoreset are
// fetched by the watching thread (a count-sized copy) while the delta pass runs; the row task starts when both are done.
static bool host_async_enabled() {
  static const bool on = [] { const char* v = std::getenv("MG_HOST_ASYNC"); return !(v && v[0] == '0'); }();
  return on;
}

static void pend_finish(mg_env* env) { env->pend.done.store(1, std::memory_order_release); }
static void pend_rows_ready(mg_env* env) {   // called twice per step with reset rows: after the delta pass and after the row copy
  if (env->pend.pre.fetch_sub(1, std::memory_order_acq_rel) == 1) mg::host_submit(&env->pend.t2);
}

// on the thread that saw the step's event complete: build and submit the decode tasks
static void start_async_decode(mg_env* env) {
  auto& pd = env->pend;
  const mg_step_io& io = pd.io;
  const size_t N = (size_t)env->cfg.num_envs, cells = (size_t)env->base.cells;
  const int T = env->host_threads;
  if (env->transport == MG_TRANSPORT_PACKED) {
    if (!io.obs) { pend_finish(env); return; }
    mg::host_expand_task(&pd.t1, env->h_grid, io.obs, N * cells, T);
    pd.t1.then = [env] { pend_finish(env); };
    mg::host_submit(&pd.t1);
    return;
  }
  mg::HostDeltaJob& j = pd.job;
  j.records = env->h_delta_blk + 16; j.n = N; j.stride = mg::delta_record_bytes(env->base.cells, env->base.A);
  j.wide = env->base.cells > 256; j.cells = env->base.cells; j.A = env->base.A; j.reward_table = env->reward_table;
  j.obs = io.obs; j.skip_patches = pd.refresh; j.rewards = io.rewards; j.terminated = io.terminated; j.truncated = io.truncated;
  j.final_obs = pd.refresh ? nullptr : io.final_obs;
  mg::host_delta_task(&pd.t1, &j, T);
  if (!io.obs) {
    pd.t1.then = [env] { pend_finish(env); };
    mg::host_submit(&pd.t1);
    return;
  }
  if (pd.refresh) {   // the whole mirror again from the packed plane, after rewards / flags
    mg::host_expand_task(&pd.t2, env->h_grid, io.obs, N * cells, T);
    pd.t2.then = [env] { pend_finish(env); };
    pd.t1.then = [env] { mg::host_submit(&env->pend.t2); };
    mg::host_submit(&pd.t1);
    return;
  }
  int32_t count;
  std::memcpy(&count, env->h_delta_blk, 4);
  if (count < 0 || (size_t)count > N) { pd.corrupt = true; pend_finish(env); return; }
  if (count == 0) {
    pd.t1.then = [env] { pend_finish(env); };
    mg::host_submit(&pd.t1);
    return;
  }
  mg::host_rows_task(&pd.t2, env->h_reset_rows, env->reset_stride, (size_t)count, env->base.cells, io.obs, N, T);
  pd.t2.then = [env] { pend_finish(env); };
  pd.pre.store(2, std::memory_order_release);
  pd.t1.then = [env] { pend_rows_ready(env); };
  mg::host_submit(&pd.t1);
  cudaError_t ce = cudaSetDevice(env->device);   // a pool thread: its current device is whatever it touched last
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(env->h_reset_rows, env->d_reset_rows, (size_t)count * env->reset_stride, cudaMemcpyDeviceToHost, pd.st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(pd.st);
  if (ce != cudaSuccess) { pd.cuda_err = (int)ce; pd.err_what = "D2H reset rows"; }   // the rows are still applied (stale): the wait reports the error
  pend_rows_ready(env);
}

// steps in flight, watched by the pool's idle threads (host_set_poller): probe = cudaEventQuery over them
static std::mutex g_inflight_mu;
static std::vector<mg_env*> g_inflight;
static void* probe_inflight() {
  std::lock_guard<std::mutex> lk(g_inflight_mu);
  for (size_t i = 0; i < g_inflight.size(); ++i) {
    mg_env* env = g_inflight[i];
    const cudaError_t ce = cudaEventQuery(env->host_ev);
    if (ce == cudaErrorNotReady) continue;
    if (ce != cudaSuccess) { env->pend.cuda_err = (int)ce; env->pend.err_what = "cudaEventQuery"; }
    g_inflight.erase(g_inflight.begin() + (long)i);
    return env;
  }
  return nullptr;
}
static void start_inflight(void* v) {
  mg_env* env = static_cast<mg_env*>(v);
  mg::host_poll_add(-1);
  if (env->pend.cuda_err) { pend_finish(env); return; }
  start_async_decode(env);
}

// behind step_host_compact on the async path: mark the step and put it on the watch list
static int launch_async_decode(mg_env* env, cudaStream_t st) {
  cudaError_t ce;
  if (!env->host_ev) {
    if ((ce = cudaEventCreateWithFlags(&env->host_ev, cudaEventDisableTiming)) != cudaSuccess) { env->pend.active = false; return cuda_fail(env, "cudaEventCreate", ce); }
    mg::host_set_poller(probe_inflight, start_inflight);
  }
  if ((ce = cudaEventRecord(env->host_ev, st)) != cudaSuccess) { env->pend.active = false; return cuda_fail(env, "cudaEventRecord", ce); }
  env->pend.async = true; env->pend.st = st; env->pend.cuda_err = 0; env->pend.err_what = nullptr; env->pend.corrupt = false;
  env->pend.done.store(0, std::memory_order_release);
  {
    std::lock_guard<std::mutex> lk(g_inflight_mu);
    g_inflight.push_back(env);
  }
  mg::host_poll_add(1);
  return 0;
}

// mg_step_host_wait on the async path: work on queued decode chunks until this step's are done
static int wait_async_decode(mg_env* env) {
  mg::host_help_until(env->pend.done);
  auto& pd = env->pend;
  pd.active = false; pd.async = false;
  if (pd.cuda_err) { env->mirror_valid = false; return cuda_fail(env, pd.err_what ? pd.err_what : "host step", (cudaError_t)pd.cuda_err); }
  if (pd.corrupt) { env->mirror_valid = false; return fail(env, "mg_step_host: corrupt reset count"); }
  if (env->transport == MG_TRANSPORT_DELTA) {
    if (!pd.io.obs) env->mirror_valid = false;   // nothing was patched: a mirror kept by the caller is now stale
    else { env->mirror = pd.io.obs; env->mirror_valid = true; }
  }
  return 0;
}

static void drain_host_step(mg_env* env) {
  if (env->pend.active && env->pend.async) wait_async_decode(env);
}

// after the stream has drained: decode the staged results into the caller's buffers
static int finish_host_step(mg_env* env, cudaStream_t st) {
  if (!env->pend.active) return 0;
  env->pend.active = false;
  const mg_step_io& io = env->pend.io;
  const size_t N = (size_t)env->cfg.num_envs, cells = (size_t)env->base.cells;
  const int T = env->host_threads;
  if (env->transport == MG_TRANSPORT_PACKED) {
    if (io.obs) mg::host_expand_plane(env->h_grid, io.obs, N * cells, T);
    return 0;
  }
  mg::HostDeltaJob j;
  j.records = env->h_delta_blk + 16; j.n = N; j.stride = mg::delta_record_bytes(env->base.cells, env->base.A);
  j.wide = env->base.cells > 256; j.cells = env->base.cells; j.A = env->base.A; j.reward_table = env->reward_table;
  j.obs = io.obs; j.skip_patches = env->pend.refresh; j.rewards = io.rewards; j.terminated = io.terminated; j.truncated = io.truncated;
  j.final_obs = env->pend.refresh ? nullptr : io.final_obs;
  mg::host_apply_delta(j, T);
  if (!io.obs) { env->mirror_valid = false; return 0; }   // nothing was patched: a mirror kept by the caller is now stale
  if (env->pend.refresh) {
    mg::host_expand_plane(env->h_grid, io.obs, N * cells, T);
  } else {
    int32_t count;
    std::memcpy(&count, env->h_delta_blk, 4);
    if (count < 0 || (size_t)count > N) return fail(env, "mg_step_host: corrupt reset count");
    if (count > 0) {   // rows of the envs that autoreset this step: a second, count-sized copy
      cudaError_t ce;
      if ((ce = cudaMemcpyAsync(env->h_reset_rows, env->d_reset_rows, (size_t)count * env->reset_stride, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
        return cuda_fail(env, "D2H reset rows", ce);
      if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(env, "cudaStreamSynchronize", ce);
      mg::host_apply_rows(env->h_reset_rows, env->reset_stride, (size_t)count, env->base.cells, io.obs, N, T);
    }
  }
  env->mirror = io.obs; env->mirror_valid = true;
  return 0;
}

static int step_host_enqueue(mg_env* env, void* state, const mg_step_io* io, void* stream, bool wait) {
  if (!env || !state || !io) return fail(env, "mg_step_host: null argument");
  if (!io->actions || !io->rewards || !io->terminated || !io->truncated) return fail(env, "mg_step_host: actions/rewards/terminated/truncated must be non-null");
  if (io->final_obs && !io->obs) return fail(env, "mg_step_host: final_obs needs obs");
  if (env->pend.active) return fail(env, "mg_step_host: the previous host step has not been waited for (mg_step_host_wait)");
  cudaError_t ce;
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (env->transport != MG_TRANSPORT_FULL) {
    if (step_host_compact(env, state, io, st)) return -1;
    if (!wait) return host_async_enabled() ? launch_async_decode(env, st) : 0;
    if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) { env->pend.active = false; return cuda_fail(env, "cudaStreamSynchronize", ce); }
    return finish_host_step(env, st);
  }
  const size_t N = env_count(env);
  const size_t A = (size_t)env->act_cols, R = (size_t)env->rew_cols, ob = mg_obs_bytes(env);
  // device staging: ONE block laid out obs | rewards | terminated | truncated (256-byte aligned parts, see mg_host_layout):
  // a caller whose host buffers use the same layout gets a single device-to-host copy per step
  size_t off_r, off_t, off_u, total;
  host_layout(env, &off_r, &off_t, &off_u, &total);
  if (!env->d_actions) {   // committed to the handle only when both allocations succeeded
    int8_t* da = nullptr; uint8_t* dobs = nullptr;
    if ((ce = cudaMalloc(&da, N * A)) != cudaSuccess) return cuda_fail(env, "cudaMalloc", ce);
    if ((ce = cudaMalloc(&dobs, total)) != cudaSuccess) { cudaFree(da); return cuda_fail(env, "cudaMalloc", ce); }
    if ((ce = cudaMemsetAsync(dobs, 0, total, st)) != cudaSuccess) { cudaFree(da); cudaFree(dobs); return cuda_fail(env, "cudaMemsetAsync", ce); }
    env->d_actions = da; env->d_obs = dobs;
    env->d_rewards = reinterpret_cast<double*>(dobs + off_r);
    env->d_term = dobs + off_t; env->d_trunc = dobs + off_u;
  }
  if (io->final_obs && !env->d_final) {
    if ((ce = cudaMalloc(&env->d_final, ob)) != cudaSuccess) return cuda_fail(env, "cudaMalloc", ce);
    if ((ce = cudaMemsetAsync(env->d_final, 0, ob, st)) != cudaSuccess) return cuda_fail(env, "cudaMemsetAsync", ce);
  }
  if ((ce = cudaMemcpyAsync(env->d_actions, io->actions, N * A, cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(env, "H2D actions", ce);
  mg_step_io dio;
  dio.actions = env->d_actions; dio.obs = io->obs ? static_cast<uint8_t*>(static_cast<void*>(env->d_obs)) : nullptr; dio.rewards = env->d_rewards;
  dio.terminated = env->d_term; dio.truncated = env->d_trunc; dio.final_obs = io->final_obs ? env->d_final : nullptr;
  env->mirror_valid = false;
  if (step_device(env, state, &dio, st)) return -1;
  const uint8_t* hb = static_cast<const uint8_t*>(static_cast<const void*>(io->obs));
  const bool one_copy = io->obs && reinterpret_cast<const uint8_t*>(io->rewards) == hb + off_r && io->terminated == hb + off_t && io->truncated == hb + off_u;
  if (one_copy) {
    if ((ce = cudaMemcpyAsync(io->obs, env->d_obs, total, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H results", ce);
  } else {
    if (io->obs && (ce = cudaMemcpyAsync(io->obs, env->d_obs, ob, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H obs", ce);
    if ((ce = cudaMemcpyAsync(io->rewards, env->d_rewards, N * R * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H rewards", ce);
    if ((ce = cudaMemcpyAsync(io->terminated, env->d_term, N, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H terminated", ce);
    if ((ce = cudaMemcpyAsync(io->truncated, env->d_trunc, N, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H truncated", ce);
  }
  if (io->final_obs && (ce = cudaMemcpyAsync(io->final_obs, env->d_final, ob, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(env, "D2H final_obs", ce);
  if (wait && (ce = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(env, "cudaStreamSynchronize", ce);
  return 0;
}

extern "C" int mg_step_host(mg_env* env, void* state, const mg_step_io* io, void* stream) {
  return step_host_enqueue(env, state, io, stream, true);
}

extern "C" int mg_step_host_async(mg_env* env, void* state, const mg_step_io* io, void* stream) {
  return step_host_enqueue(env, state, io, stream, false);
}

extern "C" int mg_step_host_wait(mg_env* env, void* stream) {
  if (!env) return -1;
  cudaError_t ce;
  if (env->pend.active && env->pend.async) return wait_async_decode(env);
  if ((ce = cudaSetDevice(env->device)) != cudaSuccess) return cuda_fail(env, "cudaSetDevice", ce);
  if ((ce = cudaStreamSynchronize(static_cast<cudaStream_t>(stream))) != cudaSuccess) { env->pend.active = false; return cuda_fail(env, "cudaStreamSynchronize", ce); }
  return finish_host_step(env, static_cast<cudaStream_t>(stream));
}

extern "C" int mg_set_seed(mg_env* env, uint64_t seed) {
  if (!env) return -1;
  env->base.seed = seed; env->mbase.seed = seed; env->wbase.seed = seed; env->gbase.seed = seed;   // only the handle's family reads its block
  env->cfg.seed = seed; env->mcfg.seed = seed;
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

extern "C" int mg_tile_envs(const mg_env* env) {
  if (!env) return -1;
  if (env->family == MG_FAMILY_WILDFIRE) return 1;
  if (env->family == MG_FAMILY_GENERIC) return mg::generic_tile_envs();
  return env->family == MG_FAMILY_COLLECT ? mg::tile_envs(env->tile) : mg::map_tile_envs();
}

extern "C" int64_t mg_launch_count(const mg_env* env) { return env ? env->launches : 0; }

extern "C" int mg_stream_idle(void* stream) {
  const cudaError_t ce = cudaStreamQuery(static_cast<cudaStream_t>(stream));
  if (ce == cudaSuccess) return 1;
  if (ce == cudaErrorNotReady) return 0;
  (void)cudaGetLastError();
  return -1;
}
