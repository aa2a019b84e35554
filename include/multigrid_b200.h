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
enum { MG_FAMILY_COLLECT = 0, MG_FAMILY_MAZE = 1, MG_FAMILY_CTF = 2, MG_FAMILY_WILDFIRE = 3, MG_FAMILY_GENERIC = 4 };

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
enum { MG_PLANE_GRID = 0,      /* u8  [N_pad][W*H]  packed cell = type | colour<<2 | state<<6, index x*H+y; state = agent dir.  A ball
                                  placed by _respawn carries bit 6 when the config gives it a reward different from a ball
                                  placed by _gen_grid (collect_game.py:130/:409 vs :101/:393); its observed STATE stays 0 */
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

/* T steps in ONE launch (Collect handles): the loop `for t in range(T): env.step(actions[t])` (collect_game.py:183-214 called T times)
 * with the env state held in shared memory across the steps - state traffic and launch latency are paid once per call, which is
 * what small batches need (4 096 envs: one launch per step is latency-bound).  Bit-identical to T calls of mg_step with the same
 * actions, autoresets included.  All device pointers; arrays are step-major. */
typedef struct mg_rollout_io {
  int32_t steps;           /* T >= 1 */
  const int8_t* actions;   /* [T][N][A], or NULL = uniform random policy drawn on the device: one Philox4x32-10 block per env and step,
                              counter (env id lo, env id hi, step_count, 2^30 | episode), key = seed; agent i takes bits 2i..2i+1 of word 0 */
  uint8_t* obs;            /* [T][N][W][H][3] or NULL (rewards-only rollouts, e.g. planners scoring action sequences) */
  double* rewards;         /* [T][N][A] */
  uint8_t* terminated;     /* [T][N] */
  uint8_t* truncated;      /* [T][N] */
  uint8_t* final_obs;      /* [T][N][W][H][3] or NULL */
  int8_t* actions_out;     /* [T][N][A] or NULL: with the on-device policy, the actions that were taken */
} mg_rollout_io;
int mg_rollout(mg_env* env, void* state_dev, const mg_rollout_io* io_dev, void* stream);

/* Grid.encode alone: state grid plane -> obs (grid.py:223-252). */
int mg_encode(mg_env* env, const void* state_dev, uint8_t* obs_dev, void* stream);

/* MultiGridEnv.gen_obs (multigrid.py:485-532): per agent the egocentric view_size x view_size window in front of
 * the agent (slice + rotate_left x (dir+1) + process_vis + encode_for_agents), out u8 [N][A][V][V][3].
 * dirs_dev: u8 [N][A] agent directions or NULL (Collect: 3, the only direction its agents ever have; Maze: the
 * state's dir plane).  Collect: cells outside the grid are grey walls (grid.py:124-127).  Maze (a composition the
 * reference does not ship: MazeWorld has no wall, world.py:81-91): outside cells are an opaque obstacle-coloured
 * filler (3, 7, 1) -- an extension, parity unpinned for that one code.  Generic family (DefaultWorld, encode_dim 6):
 * out u8 [N][A][V][V][6]; walls and closed / locked doors block sight (object.py:178-179, 223-224); dirs NULL = the
 * dir stored with each agent's cell. */
int mg_gen_obs(mg_env* env, const void* state_dev, const uint8_t* dirs_dev, int view_size, int see_through_walls,
               uint8_t* out_dev, void* stream);

/* ToroidObservation.observation (wrappers/toroid.py:28-68) for every env and agent: agent-centred wrap-around
 * one-hot planes, out f32 [N][A][W][H][num_ball_types + num_agents].  Collect family, square grids. */
int mg_toroid_obs(mg_env* env, const void* state_dev, float* out_dev, void* stream);

/* MultiGridEnv.render() frames (render_mode "rgb_array", highlight off: multigrid.py:546-606, Grid.render grid.py:183-221,
 * Grid.render_tile :132-181, utils/rendering.py) of `n` envs: env_ids_dev int32 [n] on the device, or NULL = envs 0..n-1;
 * out u8 [n][H*tile_size][W*tile_size][3].  The reference's default tile size is 32 (constants.py:5).  Collect, Maze and
 * CtF handles (CtF agents: grey once terminated, on the sticky background colour kept in flags bits 2-3 of the agent word).
 * The per-code tile images are rasterised once per tile size on the host when first asked for (the reference's tile cache); the call itself is one kernel that blits them.  An env id outside [0, N) draws env 0 and sets MG_ERR_OOB. */
int mg_render(mg_env* env, const void* state_dev, const int32_t* env_ids_dev, int n, int tile_size, uint8_t* out_dev, void* stream);

/* Re-key the Philox streams of the following launches (`reset(seed=...)`, multigrid.py:114-119 -> gymnasium's np_random(seed)).
 * The per-env block counters live in the caller's state buffer (header word 2): zero them as well to make what follows a
 * function of the seed alone. */
int mg_set_seed(mg_env* env, uint64_t seed);

/* Same as mg_step with HOST buffers: copies actions host->device, steps, copies obs / rewards /
 * flags device->host and waits.  This is the call a gymnasium-style user makes with numpy
 * arrays; buffers should be page-locked for full PCIe bandwidth. */
int mg_step_host(mg_env* env, void* state_dev, const mg_step_io* io_host, void* stream);

/* The two halves of mg_step_host - gymnasium's VectorEnv.step_async / step_wait (the pattern AsyncVectorEnv is driven with):
 * _async enqueues H2D actions -> step -> D2H results on `stream` and returns at once; _wait blocks until that stream has
 * drained.  With two handles on two streams the device-to-host copy of one env batch overlaps the host work and the step
 * of the other, which keeps the PCIe link busy back to back.  The host buffers must stay untouched until _wait returns. */
int mg_step_host_async(mg_env* env, void* state_dev, const mg_step_io* io_host, void* stream);
int mg_step_host_wait(mg_env* env, void* stream);

/* Layout that lets mg_step_host return everything with ONE device-to-host copy: if the caller's host buffers are parts of one
 * (page-locked) block with io->rewards = io->obs + *off_rewards, io->terminated = io->obs + *off_terminated, io->truncated =
 * io->obs + *off_truncated (block size *total_bytes), the results travel in a single cudaMemcpyAsync; any other placement is
 * served by one copy per array.  Offsets depend on the observation mode (mg_set_partial_obs). */
int mg_host_layout(const mg_env* env, size_t* off_rewards, size_t* off_terminated, size_t* off_truncated, size_t* total_bytes);

/* What mg_step_host[_async] moves over PCIe for the observation (Collect handles; the step always runs on the device):
 *   MG_TRANSPORT_FULL    the expanded observation, 3 bytes per cell (default);
 *   MG_TRANSPORT_PACKED  the packed grid plane, 1 byte per cell (type | colour << 2 | state << 6), expanded to the caller's
 *                        (W, H, 3) uint8 array on `host_threads` host threads inside mg_step_host / mg_step_host_wait;
 *   MG_TRANSPORT_DELTA   one record per env with the <= 3 * num_agents cells the step wrote, the pickups (rewards are a table
 *                        lookup on the host) and the flags - mg_delta_record_bytes() bytes, 16 for 2 agents on a 10x10 grid -
 *                        plus the packed rows of the envs that autoreset in this step.  io->obs is then a PERSISTENT mirror that
 *                        the library patches in place: pass the same host buffer on every call and do not write to it.  A new
 *                        buffer, or a mirror invalidated by mg_reset / mg_step / mg_host_invalidate (call it after writing the
 *                        state buffer directly), is refreshed in full from the packed plane on the next host step.
 * Same results as MG_TRANSPORT_FULL, byte for byte.  host_threads <= 0: the cores of the process's affinity mask (at most 32). */
enum { MG_TRANSPORT_FULL = 0, MG_TRANSPORT_PACKED = 1, MG_TRANSPORT_DELTA = 2 };
int mg_set_host_transport(mg_env* env, int mode, int host_threads);
int mg_host_invalidate(mg_env* env);
/* The host-side decoders on their own (no device involved): packed cells -> Grid.encode() triples (grid.py:223-252), and
 * the delta records of n envs applied to an observation mirror.  reward_table: double [33], entry 0 = 0.0, entry
 * 1 + (colour | respawned << 4) = the reward of such a ball.  Record layout: byte 0 = n_changes | terminated << 5 |
 * truncated << 6 | autoreset << 7; bytes 1..A = per agent 0 or 1 + (colour | respawned << 4) of the ball it picked up; then
 * 3A entries (cell index x*H+y as u8 if W*H <= 256 else u16 little endian, packed cell code u8); padded to a multiple of 4. */
int mg_host_expand_plane(const uint8_t* grid_host, uint8_t* obs_host, size_t n_cells, int host_threads);
int mg_delta_record_bytes(int cells, int num_agents);
int mg_host_apply_delta(const uint8_t* records, size_t n, int cells, int num_agents, const double* reward_table, uint8_t* obs,
                        double* rewards, uint8_t* terminated, uint8_t* truncated, uint8_t* final_obs, int host_threads);

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

/* 1 = every operation enqueued on `stream` so far has completed, 0 = work is pending, -1 = error (cudaStreamQuery).  For callers
 * of mg_step_host_async that keep a private stream per env batch: the private stream only has to wait for the caller's compute
 * stream when that one is busy. */
int mg_stream_idle(void* stream);

/* ===================================================================================== Maze and CtF
 * Static text-map families: MazeSingleAgentEnv (envs/maze.py:26-377) and CtFMvNEnv (envs/ctf.py:657-1433).
 * The map (utils/map.py:22-39, field_map = np.loadtxt(path).T, indexed [x][y]) is shared by all envs and
 * lives in handle-owned device tables; per-env state is only the agents.  Square maps only: the
 * reference mixes width and height (maze.py:68-70 vs :184-186, ctf.py:745-747) and raises an AssertionError
 * from Grid.set on any non-square map, so there is no non-square behaviour to reproduce. */
#define MG_MAX_MAP_AGENTS 16

enum { MG_OBS_U8 = 0,        /* "map" codes as uint8 (compact default) */
       MG_OBS_REFERENCE = 1  /* the reference's dtype: float64 for Maze (maze.py:246), int64 for CtF (ctf.py:1138) */ };

typedef struct mg_map_config {
  uint32_t struct_size;       /* = sizeof(mg_map_config) */
  int32_t family;             /* MG_FAMILY_MAZE | MG_FAMILY_CTF */
  int64_t num_envs;
  int64_t env_id_base;
  int32_t size;               /* W == H */
  const uint8_t* field_map;   /* HOST pointer, [size*size], index x*size + y; MazeWorld / CtfWorld codes (world.py:66-91) */
  int32_t num_blue, num_red;  /* CtF: num_blue_agents / num_red_agents; Maze: 1 / 0 */
  double flag_reward;         /* flag_reward */
  double battle_reward;       /* battle_reward_ratio * flag_reward      (ctf.py:725), formed by the caller in double */
  double obstacle_penalty;    /* obstacle_penalty_ratio * flag_reward   (ctf.py:726, maze.py:349) */
  double step_penalty;        /* step_penalty_ratio * flag_reward       (ctf.py:727, maze.py:350) */
  double battle_range;        /* ctf.py:667 */
  double randomness;          /* ctf.py:668 */
  int32_t max_steps;
  int32_t autoreset;
  int32_t obs_dtype;          /* MG_OBS_* */
  int32_t variant_1v1;        /* CtF only: 1 = Ctf1v1Env rules (ctf.py:50-654): fixed order blue then red, no shuffle draw, losing a
                                 battle ends the episode instead of defeating blue, no collision penalty; needs num_blue = num_red = 1 */
  uint64_t seed;
} mg_map_config;

/* planes of the map families' state buffer */
enum { MG_MAP_PLANE_AGENTS = 0, /* u8  [N_pad][row]: agent i (blue first, n = num_blue + num_red) at bytes 4i .. 4i+3 =
                                   x, y (Agent.pos), dir (Agent.dir), flags (bit0 terminated / defeated, bit1 collided,
                                   agent.py:97-100; CtF bits 2-3: Agent.bg_color, 0 as constructed, 1 light_blue, 2 light_red,
                                   ctf.py:1214-1230); row = 4 * (n rounded up to a power of two) bytes */
       MG_MAP_PLANE_HDR = 1,    /* i32 [N_pad][4]  step_count, CtF game_stats bits (ctf.py:1068-1073: bit0 blue_flag_captured,
                                   bit1 red_flag_captured, bit 8+i agent i defeated in a battle; cleared by reset), Philox block
                                   counter, episodes */
       MG_MAP_PLANE_COUNT = 2 };

/* Validation mode for the map families: recorded outputs of the reference's RNG call sites. */
typedef struct mg_map_trace {
  const int32_t* start_index;  /* [N]           Maze reset: np.random.randint(0, len(background))  maze.py:204 */
  const int32_t* blue_place;   /* [N][num_blue] CtF reset: np_random.choice(len(blue_territory), k, replace=False) ctf.py:1034 */
  const int32_t* red_place;    /* [N][num_red]  ctf.py:1041 */
  const int8_t* red_actions;   /* [N][num_red]  RwPolicy: np_random.integers(0, 5)  policy/ctf/heuristic.py:72 */
  const uint8_t* order;        /* [N][n]        np_random.shuffle(agent_indices)    ctf.py:1245 */
  const uint8_t* blue_win;     /* [N][KB]       battle outcomes np_random.choice    ctf.py:1393-1403 */
  int32_t KB;
  int32_t* battles_used;       /* [N] out (NULL = skip) */
} mg_map_trace;

#define MG_ERR_BAD_ACTION 8     /* action outside the action set (reference: ValueError, maze.py:286, ctf.py:1200) */

/* MazeSingleAgentEnv.__init__ / CtFMvNEnv.__init__.  mg_reset / mg_step / mg_step_host / mg_status /
 * mg_destroy work on the returned handle; mg_step_io then means: actions int8 [N][1] (Maze, MazeActions) or
 * [N][num_blue] (CtF, CtfActions: 0 stay 1 left 2 down 3 right 4 up); rewards float64 [N] (scalar reward);
 * obs [N][size][size] of obs_dtype (Maze: [x][y] as _encode_map maze.py:245-260; CtF: [y][x] as
 * _encode_map().T ctf.py:1137-1163). */
int mg_create_map(const mg_map_config* cfg, int device, mg_env** out);
int mg_set_map_trace(mg_env* env, const mg_map_trace* trace_dev);

/* CtF handles: actions of the red agents for the following steps, int8 [N][num_red] on the device, read by every mg_step until
 * changed (the caller overwrites the buffer between steps) - the reference's `enemy_policies` argument (ctf.py:666) for any
 * policy that is not the built-in RwPolicy: a learned opponent, self-play, a host-side A* policy.  NULL = RwPolicy drawn on the
 * device (default).  Shuffle order and battle outcomes stay with the env's Philox stream. */
int mg_set_red_actions(mg_env* env, const int8_t* red_actions_dev);

/* CtF handles: the reference's scripted opponents (policy/ctf/heuristic.py:40-463) decided ON THE DEVICE for every env.
 * mg_set_red_policies copies the tables (host pointers, read during the call) to the device; mg_red_policy_actions is one
 * launch that writes the red team's actions int8 [N][num_red] from the current state - bind the same buffer with
 * mg_set_red_actions and call it before each mg_step.  Targets are the reference's (closest blue agent, blue flag, border
 * patrol, fight while a blue agent stands on red ground); the move towards a target is the first step of the reference's
 * A* route (policy/ctf/utils.py:17-120), which the caller tabulates per (cell, target) by running that A* - cells are
 * indexed x * size + y; "follow the route with probability randomness, else a uniform action" (heuristic.py:150-175) draws
 * from the env's Philox generator (blocks disjoint from the step's), the device stand-in for numpy's Generator as for
 * RwPolicy.  NULL tables = forget them. */
enum { MG_POLICY_RW = 0, MG_POLICY_FIGHT = 1, MG_POLICY_CAPTURE = 2, MG_POLICY_PATROL = 3, MG_POLICY_PATROL_FIGHT = 4 };
typedef struct mg_red_policies {
  uint32_t struct_size;         /* sizeof(mg_red_policies) */
  int32_t num_red;              /* must equal the handle's num_red_agents */
  int32_t kind[16];             /* MG_POLICY_* per red agent */
  double randomness[16];        /* probability of following the route (heuristic.py:89, default 0.75) */
  const uint8_t* first_move;    /* [cells][cells] start-major: CtfActions value of the first move from `start` towards `target` */
  const uint16_t* patrol_goal;  /* [cells] closest border cell of every cell (PatrolPolicy off the border, heuristic.py:336) */
  const uint8_t* on_border;     /* [cells] 1 = the cell is in PatrolPolicy.border */
  const uint16_t* along_border; /* [n_along] border cells next to a border cell, duplicates kept (heuristic.py:323-333) */
  int32_t n_along;              /* may be 0 only if no agent patrols */
} mg_red_policies;
int mg_set_red_policies(mg_env* env, const mg_red_policies* tables);
/* Host-only helper that fills mg_red_policies.first_move: for every (start, target) pair of a rows x cols map - cell index
 * x * cols + y - the CtfActions value of the first move of the route the reference's A* returns (policy/ctf/utils.py:17-120, its
 * frontier order and neighbour order restated; blocked[cell] != 0 marks the cells the reference treats as blocking, map value
 * 8, utils.py:73); `stay` when start == target; a pair without a route gets the move towards the target if adjacent, else
 * `stay` (the reference raises there, heuristic.py:172).  first_move: [cells][cells] start-major.  No device involved. */
int mg_astar_first_moves(const uint8_t* blocked, int32_t rows, int32_t cols, uint8_t* first_move);
int mg_red_policy_actions(mg_env* env, const void* state, int8_t* red_actions_dev, void* stream);
/* Fold that launch into the step: from now on every mg_step / mg_step_host* of the handle first decides the red team's actions
 * from the state it is given - exactly what mg_red_policy_actions(env, state, red_actions_dev, stream) on the step's stream would
 * write, and red_actions_dev [N][num_red] int8 receives them as before - and then steps with them, as the reference's step calls
 * its enemy policies before moving anybody (ctf.py:1297-1301).  One launch where the configuration has a fused kernel (2v2,
 * Philox draws, staged u8 map observation, no final_obs: the policy prologue runs on the agent words the step has loaded anyway);
 * any other configuration, and the validation trace, run ctf_policy_kernel ahead of the step kernel on the same stream.
 * NULL = off.  mg_set_red_policies switches it off (new tables, new call). */
int mg_set_red_policy_fusion(mg_env* env, int8_t* red_actions_dev);
/* CtFMvN handles: on != 0 makes every following reset (mg_reset, masked or not, and the same-step autoreset) keep each agent's
 * flag byte - terminated, collided, background colour - instead of clearing it: ONE reference env instance stepped through several
 * episodes.  The reference assigns Agent.terminated / collided only in Agent.__init__ and in step (core/agent.py:97-100;
 * ctf.py:1231-1236, 1316-1332, 1409-1418) and no reset() touches them (multigrid.py:114-153, ctf.py:1050-1075), so an agent
 * defeated in one episode starts the next one defeated (pinned by tests/golden/ctf_*_carry.npz, recorded from one
 * reference instance per session).  Default off = every reset starts a fresh instance, which is what the reference's own tests do. */
int mg_set_carry_agent_flags(mg_env* env, int on);
/* Validation mode of mg_red_policy_actions: replay the outputs the reference's generator produced instead of drawing from Philox -
 * per (env, red agent), all [N][num_red] on the device: the cell PatrolPolicy drew on the border (np_random.choice over the border
 * cells with a border neighbour, heuristic.py:323-334; cell index x * size + y), whether the route is followed
 * (np_random.choice([True, False], p=[randomness, 1 - randomness]), :150-151) and the uniform action (np_random.integers(0, 5), :72,
 * :175).  Entries a decision does not draw are ignored.  Three NULLs = back to Philox. */
int mg_set_policy_trace(mg_env* env, const uint16_t* patrol_target_dev, const uint8_t* follow_dev, const int8_t* action_dev);

/* CtF handles: `_get_obs()` with observation_option="flattened" (ctf.py:1084-1104; what the reference's RL script trains on,
 * scripts/main_mvn_ctf_rl.py:15-21) for every env: out int64 [N][L] on the device, L = mg_ctf_flat_len() =
 * 3 n + 4 + 2 (|blue_territory| + |red_territory| + |obstacle|): blue agent (x, y) pairs, red agent pairs, blue flag, red flag,
 * the three cell lists in the reference's order, int(agent.terminated) per agent.  The "positional" dict (ctf.py:1112-1135) is
 * the same vector cut at the key boundaries. */
int mg_ctf_flat_len(const mg_env* env);
int mg_ctf_flat_obs(mg_env* env, const void* state_dev, int64_t* out_dev, void* stream);
/* the same vector as u8 [N][L] (every entry is a coordinate < 256 or a flag): 1/8 of the bytes for consumers on the device */
int mg_ctf_flat_obs_u8(mg_env* env, const void* state_dev, uint8_t* out_dev, void* stream);

/* Maze handles: observation mode of mg_reset / mg_step / mg_step_host.  view_size 0 (default) = the "map" observation;
 * 3 / 5 / 7 = MultiGridEnv.gen_obs partial views u8 [N][1][V][V][3] computed by the SAME launch that steps the envs
 * (BASELINE config 4; same cells as mg_gen_obs would return after the step).  final_obs is not available in this mode. */
int mg_set_partial_obs(mg_env* env, int view_size, int see_through_walls);

/* `_get_info()` of every env (maze.py:262-269; ctf.py:1165-1182, 434-452) -> out float64, device:
 *   Maze [N][2]  = d_a_f, d_a_ob
 *   CtF  [N][11] = d_ba_ra, d_ba_bf, d_ba_rf, d_ra_bf, d_ra_rf, d_bf_rf, d_ba_bb, d_ba_rb, d_ra_bb, d_ra_rb, d_ba_ob
 * ("ba" = agents[0], "ra" = agents[1] - the second agent of the list, a red one only with a single blue agent, as in the
 * reference).  A distance to an empty cell list is +inf (the reference raises there). */
int mg_map_info(mg_env* env, const void* state_dev, double* out_dev, void* stream);

/* ========================================================================================= Wildfire
 * EXTENSION: the reference has no Wildfire code (only a README heading, README.md:43); BASELINE.json config 5
 * asks for it, so the rules below are OUR specification (docs in DESIGN.md section 10), checked against an
 * in-repo CPU oracle (oracle/mg_oracle_wildfire.c) - parity vs the reference is unpinned by construction.
 *
 * World: every cell of a W x H grid is a tree: healthy (0), burning (1) or burnt (2); up to 32 agents walk on
 * top (5-way CtfActions: 0 stay, 1 left (0,-1), 2 down (-1,0), 3 right (0,+1), 4 up (+1,0)).  One step:
 *   1. step_count += 1, tick += 1
 *   2. agents act in a per-step random order (Philox Fisher-Yates, or replayed): the target cell is entered
 *      unless it is off the grid or holds another agent (order-dependent blocking, as in CtF); dir follows
 *      DIR_TO_VEC; then, moved or not, a burning cell under the agent is extinguished (-> burnt), reward[i] += 1
 *   3. fire dynamics from the post-agent terrain (double-buffered stencil): a burning cell burns out (-> burnt)
 *      iff u < burnout_threshold; a healthy cell with k >= 1 burning 4-neighbours ignites iff u < ignite_threshold[k];
 *      u = word (cell & 3) of Philox4x32-10(key = seed, counter = (env id lo, env id hi, tick, 1 + cell / 4))
 *   4. terminated = no burning cell left; truncated = step_count >= max_steps
 *   5. obs = (OBJECT_IDX, COLOR_IDX, STATE) per cell, u8 [W][H][3]: healthy (0, 3 green, 0), burning (1, 0 red, 0),
 *      burnt (2, 7 grey, 0), agent (3, agent_colour[i], dir)
 * reset: all healthy, num_fires distinct burning cells, then agents on distinct cells (Philox rejection sampling). */
#define MG_MAX_WILDFIRE_AGENTS 32

typedef struct mg_wildfire_config {
  uint32_t struct_size;
  int32_t family;                 /* MG_FAMILY_WILDFIRE */
  int64_t num_envs, env_id_base;
  int32_t width, height;
  int32_t num_agents;
  int32_t agent_colour[MG_MAX_WILDFIRE_AGENTS];
  int32_t num_fires;              /* burning cells after reset */
  uint32_t ignite_threshold[5];   /* floor(2^32 * (1 - (1 - alpha)^k)), k = 0..4 (entry 0 unused) */
  uint32_t burnout_threshold;     /* floor(2^32 * beta) */
  int32_t max_steps;
  int32_t autoreset;
  uint64_t seed;
} mg_wildfire_config;

enum { MG_WF_PLANE_TERRAIN = 0, /* u8  [N_pad][W*H] 0 healthy 1 burning 2 burnt, index x*H + y */
       MG_WF_PLANE_AGENTS = 1,  /* u8  [N_pad][A][4] x, y, dir, 0 */
       MG_WF_PLANE_HDR = 2,     /* i32 [N_pad][4] step_count, tick, Philox block counter, episodes */
       MG_WF_PLANE_COUNT = 3 };

/* mg_reset / mg_step / mg_step_host / mg_status / mg_destroy work on the handle; mg_step_io: actions int8 [N][A],
 * rewards f64 [N][A], obs u8 [N][W][H][3].  mg_set_trace (order only) replays agent orders. */
int mg_create_wildfire(const mg_wildfire_config* cfg, int device, mg_env** out);

/* ================================================================ generic MultiGridEnv.step (DefaultWorld)
 * The base-class step (multigrid.py:397-483) with DefaultWorld (world.py:33-52, encode_dim 6): still / left /
 * right / forward, goal termination with `_reward` = 1 - 0.9 * step_count / max_steps (float64, multigrid.py:218-223),
 * per-agent full-grid observations `encode_for_agents` (grid.py:254-284).  Other actions raise in the reference
 * (`self.actions.available`, multigrid.py:447) and set MG_ERR_BAD_ACTION here.  `_gen_grid` is env-specific in the
 * reference, so the layout is injected: the caller fills the INIT_* planes (episode-start snapshot) and mg_reset /
 * autoreset restore them. */
typedef struct mg_generic_config {
  uint32_t struct_size;
  int32_t family;          /* MG_FAMILY_GENERIC */
  int64_t num_envs, env_id_base;
  int32_t width, height, num_agents;   /* num_agents <= 8 */
  int32_t max_steps, autoreset;
  uint64_t seed;
} mg_generic_config;

enum { MG_GEN_PLANE_CELL = 0,       /* u8 [N_pad][W*H] type | colour << 4 (type 1 = empty), index x*H + y */
       MG_GEN_PLANE_STATE = 1,      /* u8 [N_pad][W*H] door state 0 open 1 closed 2 locked / agent dir */
       MG_GEN_PLANE_POS = 2,        /* u8 [N_pad][A][2] */
       MG_GEN_PLANE_HDR = 3,        /* i32 [N_pad][4] step_count, 0, Philox block counter, episodes */
       MG_GEN_PLANE_INIT_CELL = 4, MG_GEN_PLANE_INIT_STATE = 5, MG_GEN_PLANE_INIT_POS = 6,
       MG_GEN_PLANE_COUNT = 7 };

/* mg_step_io: actions int8 [N][A] (DefaultActions 0 still 1 left 2 right 3 forward), rewards f64 [N][A],
 * obs u8 [N][A][W][H][6].  mg_set_trace (order only) replays np.random.permutation outputs (multigrid.py:402). */
int mg_create_generic(const mg_generic_config* cfg, int device, mg_env** out);

#ifdef __cplusplus
}
#endif
#endif
