/* multigrid_b200.h -- C ABI of the B200-native batched gridworld simulator.
 *
 * Drop-in boundary.  The reference (Tran-Research-Group/gym-multigrid) is pure Python and has
 * no FFI of its own: its hot path sits behind the gymnasium Env API.  This header is the C
 * boundary a reference-side binding would call (ctypes stub in INTEGRATION.md); each entry
 * point names the reference interface it replaces (file:line under the reference checkout).
 *
 * Conventions
 *   - plain C, no C++/torch types; every call returns an int status (0 = ok, < 0 = error,
 *     text via mg_last_error); nothing throws or aborts.
 *   - the CALLER owns every buffer (PyTorch tensors on the handle's device).  The handle owns
 *     only the config, tiny device constant tables and, for the *_host calls, staging buffers.
 *   - device calls are asynchronous: work is enqueued on `stream` (a cudaStream_t passed as
 *     void*; NULL = legacy default stream), no implicit synchronisation.
 *   - one handle per device; calls on one handle are serialised by the caller.
 *   - there is NO CPU fallback: mg_create fails when no CUDA device is usable.
 */
#ifndef MULTIGRID_B200_H
#define MULTIGRID_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_ABI_VERSION 1
#define MG_MAX_AGENTS 8
#define MG_MAX_BALL_TYPES 8

/* env families (reference: gym_multigrid/envs/{collect_game,maze,ctf}.py) */
enum { MG_FAMILY_COLLECT = 0 };

/* Collect layouts = the reference's _gen_grid variants */
enum { MG_LAYOUT_EVEN_DIST = 0,         /* CollectGameEvenDist          collect_game.py:227-259 */
       MG_LAYOUT_QUADRANTS = 1,         /* CollectGameQuadrants         collect_game.py:261-300 */
       MG_LAYOUT_ROOMS = 2,             /* CollectGameRooms             collect_game.py:302-362 */
       MG_LAYOUT_QUADRANTS_RESPAWN = 3  /* CollectGameQuadrantsRespawn  collect_game.py:372-409 */ };

/* Constructor arguments: CollectGameEnv.__init__ kwargs (collect_game.py:17-72) as registered
 * in gym_multigrid/__init__.py:6-147, plus the vector-env additions (num_envs, seed, ...). */
typedef struct mg_config {
  uint32_t struct_size;                     /* = sizeof(mg_config), ABI check */
  int32_t family;                           /* MG_FAMILY_* */
  int64_t num_envs;                         /* envs on THIS device */
  int64_t env_id_base;                      /* global id of env 0 (multi-GPU shards; keys the RNG) */
  int32_t width, height;                    /* size / grid_size */
  int32_t num_agents;                       /* len(agents_index) */
  int32_t num_ball_types;                   /* len(balls_index) */
  int32_t agent_colour[MG_MAX_AGENTS];      /* agents_index */
  int32_t ball_colour[MG_MAX_BALL_TYPES];   /* balls_index */
  double ball_reward[MG_MAX_BALL_TYPES];    /* balls_reward */
  int32_t num_balls;                        /* np.sum(num_balls) */
  int32_t respawn;
  int32_t layout;                           /* MG_LAYOUT_* */
  int32_t fixed_horizon;                    /* CollectGameRoomsFixedHorizon.step :368-370 */
  int32_t max_steps;                        /* env-internal truncation (100, collect_game.py:65) */
  int32_t time_limit;                       /* max_episode_steps of the registration; 0 = none */
  int32_t autoreset;                        /* 0 = off; 1 = gymnasium 0.29.1 same-step autoreset */
  uint64_t seed;                            /* Philox key (production RNG mode) */
} mg_config;

/* Planes inside the caller-owned state buffer (struct of arrays, one row per env; rows are padded
 * to a whole number of kernel tiles, N_pad >= N; every plane starts 256-byte aligned). */
enum { MG_PLANE_GRID = 0,      /* u8  [N_pad][W*H]  packed cell = type | colour<<2 | state<<6, index x*H+y */
       MG_PLANE_AGENT_POS = 1, /* u8  [N_pad][A][2] (x, y)                              Agent.pos */
       MG_PLANE_HDR = 2,       /* i32 [N_pad][4]    step_count, collected_balls, Philox block counter, episodes */
       MG_PLANE_INFO = 3,      /* i32 [N_pad][A*nb] env.info counters, index nb*agent + ball_type */
       MG_PLANE_COUNT = 4 };

/* step inputs/outputs; device pointers for mg_step, host pointers for mg_step_host.
 * Replaces CollectGameEnv.step's return tuple (collect_game.py:183-214). */
typedef struct mg_step_io {
  const int8_t* actions;  /* [N][A]  CollectActions 0..3 north/east/south/west; other values = no-op */
  uint8_t* obs;           /* [N][W][H][3] Grid.encode() (grid.py:223-252); NULL = skip */
  double* rewards;        /* [N][A] float64, as np.zeros(len(actions)) */
  uint8_t* terminated;    /* [N] */
  uint8_t* truncated;     /* [N] */
  uint8_t* final_obs;     /* [N][W][H][3] terminal obs of envs that autoreset this step; NULL = skip */
} mg_step_io;

/* Validation mode: replay RNG outputs recorded from the reference's generators
 * (np.random.permutation collect_game.py:186; random.randint multigrid.py:225-230).
 * All device pointers.  Pass NULL to mg_set_trace to return to Philox mode. */
typedef struct mg_trace {
  const uint8_t* order;         /* [N][A]   per-step agent update order */
  const uint8_t* draws;         /* [N][K]   per-step randint outputs (x, y, x, y, ...) */
  const int32_t* n_draws;       /* [N]      valid entries of draws per env */
  int32_t K;
  const uint8_t* reset_draws;   /* [N][R]   randint outputs consumed by reset / autoreset */
  const int32_t* n_reset_draws; /* [N] */
  int32_t R;
  int32_t* draws_used;          /* [N] out: entries consumed by the step (NULL = skip) */
  int32_t* reset_draws_used;    /* [N] out: entries consumed by the reset (NULL = skip) */
} mg_trace;

/* bits of the device status word (mg_status) */
#define MG_ERR_TRACE_OVERFLOW 1 /* replay ran out of recorded draws */
#define MG_ERR_TRACE_RANGE 2    /* a recorded draw lies outside the requested [lo, hi] */
#define MG_ERR_OOB 4            /* an agent tried to leave the grid (reference asserts, grid.py:62-63) */

typedef struct mg_env mg_env;

int mg_abi_version(void);

/* MultiGridEnv.__init__ / gymnasium.make (multigrid.py:28-89; __init__.py:6-147). */
int mg_create(const mg_config* cfg, int device, mg_env** out);
int mg_destroy(mg_env* env);
const char* mg_last_error(const mg_env* env); /* env may be NULL: error of the last failed mg_create */

/* sizes and layout of the caller-owned buffers */
size_t mg_state_bytes(const mg_env* env);
size_t mg_obs_bytes(const mg_env* env);                       /* N*W*H*3 */
int mg_state_plane(const mg_env* env, int plane, size_t* offset, size_t* bytes, size_t* row_bytes);

/* MultiGridEnv.reset + _gen_grid (multigrid.py:114-153; collect_game.py:107-119 + layout).
 * mask_dev: u8[N], NULL = all envs.  obs_dev may be NULL. */
int mg_reset(mg_env* env, void* state_dev, const uint8_t* mask_dev, uint8_t* obs_dev, void* stream);

/* CollectGameEnv.step + Grid.encode, fused, all envs in lockstep (collect_game.py:183-214). */
int mg_step(mg_env* env, void* state_dev, const mg_step_io* io_dev, void* stream);

/* Grid.encode alone: state grid plane -> obs (grid.py:223-252). */
int mg_encode(mg_env* env, const void* state_dev, uint8_t* obs_dev, void* stream);

/* Same as mg_step with HOST buffers: copies actions host->device, steps, copies obs / rewards /
 * flags device->host and waits.  This is the call a gymnasium-style user makes with numpy
 * arrays; buffers should be page-locked for full PCIe bandwidth. */
int mg_step_host(mg_env* env, void* state_dev, const mg_step_io* io_host, void* stream);

int mg_set_trace(mg_env* env, const mg_trace* trace_dev);

/* device status word: read (synchronises the stream) and clear */
int mg_status(mg_env* env, void* stream, int32_t* status_out);

/* Profiling hooks (no reference counterpart).  mg_debug_set_timeline: device buffer u64[tiles][8]; each
 * CTA of the step kernel records globaltimer ns at {start, inputs landed, stepped, autoreset done,
 * encoded, stores drained}; NULL switches it off.  mg_tile_envs: envs per CTA tile (N_pad granule). */
int mg_debug_set_timeline(mg_env* env, uint64_t* timeline_dev);
int mg_tile_envs(const mg_env* env);

/* number of kernel launches issued through this handle since creation */
int64_t mg_launch_count(const mg_env* env);

#ifdef __cplusplus
}
#endif
#endif
