/* mg_oracle.h -- CPU restatement ("oracle") of the gym-multigrid hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (gym-multigrid_b200/ + include/multigrid_b200.h) never links, imports or calls it.
 *
 * Parity pin: every function here is checked bit-for-bit against traces recorded from
 * the unmodified reference (oracle/gen_golden.py -> tests/golden/ (npz files), replayed by
 * tests/test_oracle_golden.py).  The reference ships no golden vectors of its own
 * (SURVEY.md section 4), so those recorded traces are the pin.
 *
 * Citations are file:line under the reference checkout (/root/reference).
 */
#ifndef MG_ORACLE_H
#define MG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OC_MAX_AGENTS 8
#define OC_MAX_BALL_TYPES 8

/* Packed Collect cell: type | colour << 2 | state << 6 (CollectWorld, core/world.py:54-64;
 * colours core/constants.py:8-19; agent state = dir, core/agent.py:119-126). */
#define OC_T_EMPTY 0
#define OC_T_WALL 1
#define OC_T_BALL 2
#define OC_T_AGENT 3
#define OC_CELL(type, colour, state) ((uint8_t)((type) | ((colour) << 2) | ((state) << 6)))
#define OC_WALL_GREY OC_CELL(OC_T_WALL, 7, 0)

enum { OC_LAYOUT_EVEN_DIST = 0,        /* CollectGameEvenDist._gen_grid        collect_game.py:236-259 */
       OC_LAYOUT_QUADRANTS = 1,        /* CollectGameQuadrants._gen_grid       collect_game.py:266-300 */
       OC_LAYOUT_ROOMS = 2,            /* CollectGameRooms._gen_grid           collect_game.py:306-362 */
       OC_LAYOUT_QUADRANTS_RESPAWN = 3 /* CollectGameQuadrantsRespawn          collect_game.py:376-409 */ };

typedef struct {
  int32_t width, height;
  int32_t num_agents;                       /* len(agents_index) */
  int32_t num_ball_types;                   /* len(balls_index) */
  int32_t agent_colour[OC_MAX_AGENTS];      /* agents_index */
  int32_t ball_colour[OC_MAX_BALL_TYPES];   /* balls_index */
  double ball_reward[OC_MAX_BALL_TYPES];    /* balls_reward */
  int32_t num_balls;                        /* np.sum(num_balls), collect_game.py:37 */
  int32_t respawn;
  int32_t layout;
  int32_t fixed_horizon;                    /* CollectGameRoomsFixedHorizon.step :368-370 */
  int32_t max_steps;                        /* env-internal, 100 (collect_game.py:65) */
  int32_t time_limit;                       /* gymnasium TimeLimit from registration; 0 = none */
} oc_collect_cfg;

/* Batched struct-of-arrays env state (same planes the CUDA library keeps in HBM). */
typedef struct {
  uint8_t* grid;       /* [N][W*H] packed cells, index x*H + y (the obs order, grid.py:234-250) */
  uint8_t* agent_pos;  /* [N][A][2] (x, y) */
  int32_t* step_count; /* [N] */
  int32_t* collected;  /* [N] */
  int32_t* info;       /* [N][A*num_ball_types]  info[keys[nb*i + colour]] collect_game.py:147 */
  uint32_t* rng_ctr;   /* [N] Philox block counter (production RNG mode only) */
} oc_collect_state;

/* Where random numbers come from.
 * mode 0 (trace): outputs recorded from the reference's own generators are replayed:
 *    order = np.random.permutation(A) (collect_game.py:186), draws = random.randint
 *    outputs in call order (multigrid.py:225-230 via place_obj :316-321).
 * mode 1 (philox): counter-based Philox4x32-10 keyed by (seed), counter (env id, ctr). */
typedef struct {
  int32_t mode;
  const uint8_t* order;   /* [N][A]      trace */
  const uint8_t* draws;   /* [N][K]      trace */
  const int32_t* n_draws; /* [N]         trace: valid entries per env */
  int32_t K;
  int32_t* draws_used;    /* [N] out, may be NULL */
  uint64_t seed;          /* philox */
  uint64_t env_id_base;   /* philox: global id of env 0 of this shard */
} oc_rng_src;

/* status bits accumulated into *status (may be NULL) */
#define OC_ERR_TRACE_OVERFLOW 1 /* trace ran out of recorded draws */
#define OC_ERR_TRACE_RANGE 2    /* a recorded draw is outside the requested [lo, hi] */
#define OC_ERR_OOB 4            /* an agent tried to leave the grid (reference would assert, grid.py:62-63) */

void oc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Grid.encode for encode_dim 3 (grid.py:223-252): n_cells packed cells -> 3*n_cells bytes. */
void oc_encode3(const uint8_t* cells, int64_t n_cells, uint8_t* obs);

/* 1 = this config makes a respawned ball's reward differ from an initial ball's for some colour, so balls placed by
 * _respawn carry bit 6 of their cell in the state (see mg_oracle.c: reward_initial / reward_respawned). */
int oc_collect_marks_respawned(const oc_collect_cfg* cfg);

/* reset: CollectGameEnv.reset + _gen_grid (collect_game.py:107-119 + layout).  mask may be NULL
 * (= all).  obs may be NULL. */
int oc_collect_reset(const oc_collect_cfg* cfg, int64_t N, oc_collect_state* st, const uint8_t* mask,
                     const oc_rng_src* rng, uint8_t* obs, int32_t* status, int nthreads);

/* step: CollectGameEnv.step (collect_game.py:183-214).  autoreset: 0 = none; 1 = gymnasium 0.29.1
 * same-step autoreset (obs returned for finished envs is the reset obs; the terminal obs goes to
 * final_obs when non-NULL).  reset_rng is used for the autoreset in trace mode (philox mode keeps
 * drawing from the env's own stream). */
int oc_collect_step(const oc_collect_cfg* cfg, int64_t N, oc_collect_state* st, const int8_t* actions,
                    const oc_rng_src* rng, uint8_t* obs, double* rewards, uint8_t* terminated,
                    uint8_t* truncated, int autoreset, const oc_rng_src* reset_rng, uint8_t* final_obs,
                    int32_t* status, int nthreads);

#ifdef __cplusplus
}
#endif
#endif

/* ===================================================================== Maze and Capture-the-Flag
 * Both run on a static text map (utils/map.py:22-39: field_map = np.loadtxt(path).T, indexed
 * [x][y]); per-env state is only the agents.  Square maps only (the reference mixes width/height,
 * maze.py:68-70 vs :184-186). */
#ifndef MG_ORACLE_MAP_H
#define MG_ORACLE_MAP_H
#ifdef __cplusplus
extern "C" {
#endif

#define OC_MAX_CTF_AGENTS 16
#define OC_MAX_BATTLES 64

typedef struct {
  int32_t size;                /* W == H */
  const uint8_t* field_map;    /* [W*H] index x*H + y */
  /* maze (maze.py:31-40): products formed by the caller in double exactly as the reference does */
  double flag_reward, obstacle_penalty, step_penalty;
  int32_t max_steps;
  /* ctf (ctf.py:662-679) */
  int32_t num_blue, num_red;
  double battle_range, randomness, battle_reward;
  int32_t variant_1v1;         /* 1 = Ctf1v1Env (ctf.py:50-654): fixed order blue->red, losing a battle ends the episode */
  int32_t carry_agent_flags;   /* 1 = one env INSTANCE stepped through several episodes: reset keeps Agent.terminated / collided /
                                  bg_color as the reference's does (agent.py:97-100 are the only assignments outside step);
                                  0 = every reset starts a fresh instance */
} oc_map_cfg;

typedef struct {
  uint8_t* pos;        /* [N][n][2] (x, y) */
  uint8_t* dir;        /* [N][n] */
  uint8_t* flags;      /* [N][n] bit0 = terminated (defeated), bit1 = collided  (ctf only) */
  int32_t* step_count; /* [N] */
  uint32_t* rng_ctr;   /* [N] */
  int32_t* stats;      /* [N] or NULL: CtF game_stats bits (ctf.py:1068-1073): 0 blue_flag_captured, 1 red_flag_captured, 8+i agent i defeated in a battle */
} oc_map_state;

typedef struct {
  int32_t mode;               /* 0 trace, 1 philox */
  /* maze reset: np.random.randint(0, len(background)) output (maze.py:204) */
  const int32_t* start_index; /* [N] */
  /* ctf reset: np_random.choice(len(territory), k, replace=False) outputs (ctf.py:1034, 1041) */
  const int32_t* blue_place;  /* [N][num_blue] */
  const int32_t* red_place;   /* [N][num_red] */
  /* ctf step: RwPolicy integers(0,5) (heuristic.py:72), np_random.shuffle (ctf.py:1245), battle choice (:1393-1403) */
  const int8_t* red_actions;  /* [N][num_red]; also honoured in Philox mode: actions of an external enemy policy (no RwPolicy draw) */
  const uint8_t* order;       /* [N][n] */
  const uint8_t* blue_win;    /* [N][KB] */
  int32_t KB;
  int32_t* battles_used;      /* [N] out, may be NULL */
  uint64_t seed, env_id_base;
} oc_map_rng;

/* MazeSingleAgentEnv (maze.py).  obs: u8 [N][W][H] "map" codes (the reference returns the same values as float64). */
int oc_maze_reset(const oc_map_cfg* c, int64_t N, oc_map_state* st, const uint8_t* mask, const oc_map_rng* rng,
                  uint8_t* obs, int32_t* status);
int oc_maze_step(const oc_map_cfg* c, int64_t N, oc_map_state* st, const int8_t* actions, const oc_map_rng* rng,
                 uint8_t* obs, double* reward, uint8_t* terminated, uint8_t* truncated, int autoreset,
                 uint8_t* final_obs, int32_t* status);

/* CtFMvNEnv (ctf.py:657-1433).  obs: u8 [N][H][W] "map" codes = _encode_map().T (the reference returns int64). */
int oc_ctf_reset(const oc_map_cfg* c, int64_t N, oc_map_state* st, const uint8_t* mask, const oc_map_rng* rng,
                 uint8_t* obs, int32_t* status);
int oc_ctf_step(const oc_map_cfg* c, int64_t N, oc_map_state* st, const int8_t* blue_actions, const oc_map_rng* rng,
                uint8_t* obs, double* reward, uint8_t* terminated, uint8_t* truncated, int autoreset,
                uint8_t* final_obs, int32_t* status);

/* _get_info of the map families: Maze [N][2] (d_a_f, d_a_ob), CtF [N][11] (ctf.py:1165-1182 key order) */
int oc_map_info(const oc_map_cfg* c, int is_maze, int64_t N, const oc_map_state* st, double* out);
/* observation_option="flattened" (ctf.py:1084-1104): int64 [N][L]; out NULL = only return L */
int oc_ctf_flattened(const oc_map_cfg* c, int64_t N, const oc_map_state* st, int64_t* out);

/* Scripted CtF opponents decided for every env (the rule of csrc/policy_kernels.cu; targets as policy/ctf/heuristic.py): kind 0 rw,
 * 1 fight, 2 capture, 3 patrol, 4 patrol_fight; tables indexed by cell = x * size + y; out int8 [N][num_red]. */
int oc_ctf_policy_actions(const oc_map_cfg* c, int64_t N, const oc_map_state* st, const int32_t* episode, const int32_t* kind,
                          const double* randomness, const uint8_t* first_move, const uint16_t* patrol_goal,
                          const uint8_t* on_border, const uint16_t* along, int32_t n_along, uint64_t seed, uint64_t env_id_base,
                          int8_t* out);

#define OC_ERR_BAD_ACTION 8 /* action outside the env's action set (reference: ValueError, maze.py:286, ctf.py:1200) */

#ifdef __cplusplus
}
#endif
#endif

/* ============================================================================ partial views
 * MultiGridEnv.gen_obs (multigrid.py:485-532): per agent an egocentric V x V window in front of the
 * agent = Grid.slice (grid.py:111-130, out-of-bounds -> Wall) + rotate_left x (dir+1) (grid.py:97-109) +
 * process_vis (grid.py:286-323) + encode_for_agents (grid.py:254-284), encode_dim 3.
 * grid: packed Collect cells [N][W*H]; pos [N][A][2]; dirs [N][A] (NULL = 3, the only direction a Collect
 * agent ever has, multigrid.py:371-374).  out: u8 [N][A][V][V][3].
 * oob_code: packed cell shown outside the grid (reference: grey Wall = OC_WALL_GREY); opaque_rule 0: walls block
 * sight (reference); 1: only oob_code does (the Maze composition, an extension). */
#ifdef __cplusplus
extern "C"
#endif
void oc_partial_view3(const uint8_t* grid, const uint8_t* pos, const uint8_t* dirs, int64_t N, int W, int H, int A,
                      int V, int see_through_walls, int oob_code, int opaque_rule, uint8_t* out);

/* ToroidObservation.observation (wrappers/toroid.py:28-68): per agent an agent-centred, wrap-around one-hot
 * float32 tensor [W][H][depth], depth = num_ball_types + num_agents, written at [y'][x'] (the reference's axis
 * swap), channels: ball colour index | depth-2 = another agent (at a different cell) | depth-1 = wall.
 * grid: packed Collect cells [N][W*H] (square); out: f32 [N][A][W][H][depth]. */
#ifdef __cplusplus
extern "C"
#endif
void oc_toroid(const uint8_t* grid, const uint8_t* pos, int64_t N, int W, int A, int num_ball_types, float* out);

/* ==================================================================================== Wildfire (extension)
 * No reference code exists (README.md:43 is a heading only): this restates OUR specification
 * (include/multigrid_b200.h "Wildfire"), so the CUDA kernels have an independent
 * checker.  Parity versus the reference is unpinned by construction. */
#define OC_MAX_WF_AGENTS 32
typedef struct {
  int32_t width, height, num_agents;
  int32_t agent_colour[OC_MAX_WF_AGENTS];
  int32_t num_fires;
  uint32_t ignite_threshold[5];
  uint32_t burnout_threshold;
  int32_t max_steps;
} oc_wf_cfg;

typedef struct {
  uint8_t* terrain;  /* [N][W*H] */
  uint8_t* agents;   /* [N][A][4] x, y, dir, 0 */
  int32_t* hdr;      /* [N][4] step_count, tick, rng ctr, episodes */
} oc_wf_state;

#ifdef __cplusplus
extern "C" {
#endif
int oc_wf_reset(const oc_wf_cfg* c, int64_t N, oc_wf_state* st, const uint8_t* mask, uint64_t seed, uint64_t env_id_base,
                uint8_t* obs);
int oc_wf_step(const oc_wf_cfg* c, int64_t N, oc_wf_state* st, const int8_t* actions, const uint8_t* order /* NULL = Philox */,
               uint64_t seed, uint64_t env_id_base, uint8_t* obs, double* rewards, uint8_t* terminated, uint8_t* truncated,
               int autoreset, uint8_t* final_obs);
#ifdef __cplusplus
}
#endif

/* ============================================================ generic MultiGridEnv.step (DefaultWorld)
 * The base-class step (multigrid.py:397-483) with DefaultWorld (world.py:33-52, encode_dim 6) and per-agent
 * full-grid observations `encode_for_agents` (grid.py:254-284).  Only still/left/right/forward are defined:
 * for any other action the reference evaluates `self.actions.available` (multigrid.py:447), which no action
 * enum has, and raises.  Cells: gcell = type | colour << 4 (type 1 = empty), gstate = door state / agent dir. */
#ifdef __cplusplus
extern "C" {
#endif
int oc_generic_step(int64_t N, int W, int H, int A, int max_steps, uint8_t* gcell, uint8_t* gstate, uint8_t* pos /*[N][A][2]*/,
                    int32_t* step_count, const int8_t* actions, const uint8_t* order, uint8_t* obs /*[N][A][W][H][6]*/,
                    double* rewards, uint8_t* terminated, uint8_t* truncated, int32_t* status);
void oc_generic_encode(int64_t N, int W, int H, int A, const uint8_t* gcell, const uint8_t* gstate, const uint8_t* pos, uint8_t* obs);
/* MultiGridEnv.gen_obs with DefaultWorld (encode_dim 6): u8 [N][A][V][V][6]; dirs NULL = the dir stored with the agent cell */
void oc_partial_view6(int64_t N, int W, int H, int A, int V, int see_through_walls, const uint8_t* gcell, const uint8_t* gstate,
                      const uint8_t* pos, const uint8_t* dirs, uint8_t* out);
#ifdef __cplusplus
}
#endif

/* ============================================================ rgb_array render (Collect family)
 * Grid.render / Grid.render_tile / utils/rendering.py, highlight off; see mg_oracle_render.c. */
#ifdef __cplusplus
extern "C" {
#endif
int oc_render_tile(int type, int colour, int state, int tile_size, uint8_t* out /*[ts][ts][3]*/);
int oc_render_grid(const uint8_t* obs /*[N][W][H][3] Grid.encode()*/, int64_t N, int W, int H, int tile_size,
                   uint8_t* out /*[N][H*ts][W*ts][3]*/);
int oc_render_maze(const uint8_t* field_map /*[S][S]*/, int S, int64_t N, const int16_t* pos /*[N][2]*/, const int8_t* dir /*[N]*/,
                   int tile_size, uint8_t* out /*[N][S*ts][S*ts][3]*/);
int oc_render_ctf(const uint8_t* field_map /*[S][S]*/, int S, int64_t N, int n, int num_blue, int variant_1v1, const uint8_t* pos /*[N][n][2]*/,
                  const uint8_t* dir /*[N][n]*/, const uint8_t* flags /*[N][n]*/, int tile_size, uint8_t* out /*[N][S*ts][S*ts][3]*/);
#ifdef __cplusplus
}
#endif
