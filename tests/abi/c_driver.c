/* c_driver.c -- the C ABI used from plain C: no Python, no torch, only include/multigrid_b200.h and the CUDA runtime for
 * the caller-owned buffers.  Creates `multigrid-collect-respawn-clustered-v0` (registered kwargs, gym_multigrid/__init__.py:122-134),
 * resets, runs `steps` steps with a deterministic action pattern and writes the final observations, the per-step reward
 * sums, the flag counts and rendered frames of three envs as raw bytes to `out_path`; tests/test_c_abi_gpu.py recomputes them
 * with the oracle.
 *
 *   cc c_driver.c -I include -I $CUDA/include -L pkg -lmultigrid_b200 -L $CUDA/lib64 -lcudart -o c_driver
 *   ./c_driver <num_envs> <steps> <seed> <out_path> */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "multigrid_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s num_envs steps seed out_path\n", argv[0]); return 1; }
  const int64_t N = atoll(argv[1]);
  const int steps = atoi(argv[2]);
  const int A = 2, W = 10, H = 10;
  mg_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.struct_size = sizeof cfg; cfg.family = MG_FAMILY_COLLECT; cfg.num_envs = N; cfg.env_id_base = 0;
  cfg.width = W; cfg.height = H; cfg.num_agents = A; cfg.num_ball_types = 3;
  cfg.agent_colour[0] = 3; cfg.agent_colour[1] = 5;
  for (int i = 0; i < 3; ++i) { cfg.ball_colour[i] = i; cfg.ball_reward[i] = 1.0; }
  cfg.num_balls = 15; cfg.respawn = 1; cfg.layout = MG_LAYOUT_QUADRANTS_RESPAWN; cfg.max_steps = 100; cfg.time_limit = 50;
  cfg.autoreset = 1; cfg.seed = (uint64_t)atoll(argv[3]);
  mg_env* env = NULL;
  if (mg_create(&cfg, 0, &env) != 0) { fprintf(stderr, "mg_create: %s\n", mg_last_error(NULL)); return 3; }

  void* state; uint8_t *obs, *term, *trunc; double* rew; int8_t* act;
  const size_t sb = mg_state_bytes(env), ob = mg_obs_bytes(env);
  CK(cudaMalloc(&state, sb)); CK(cudaMemset(state, 0, sb));
  CK(cudaMalloc((void**)&obs, ob)); CK(cudaMalloc((void**)&rew, N * A * sizeof(double)));
  CK(cudaMalloc((void**)&term, N)); CK(cudaMalloc((void**)&trunc, N)); CK(cudaMalloc((void**)&act, N * A));
  if (mg_reset(env, state, NULL, obs, NULL) != 0) { fprintf(stderr, "mg_reset: %s\n", mg_last_error(env)); return 4; }

  int8_t* h_act = (int8_t*)malloc(N * A);
  double* h_rew = (double*)malloc(N * A * sizeof(double));
  uint8_t* h_flag = (uint8_t*)malloc(N);
  FILE* f = fopen(argv[4], "wb");
  if (!f) return 5;
  mg_step_io io = {act, obs, rew, term, trunc, NULL};
  for (int t = 0; t < steps; ++t) {
    for (int64_t e = 0; e < N; ++e)
      for (int i = 0; i < A; ++i) h_act[e * A + i] = (int8_t)((e * 7 + i * 3 + (int64_t)t * 5 + (e >> 3)) % 4);
    CK(cudaMemcpy(act, h_act, N * A, cudaMemcpyHostToDevice));
    if (mg_step(env, state, &io, NULL) != 0) { fprintf(stderr, "mg_step: %s\n", mg_last_error(env)); return 6; }
    CK(cudaMemcpy(h_rew, rew, N * A * sizeof(double), cudaMemcpyDeviceToHost));
    double sum = 0.0;
    for (int64_t k = 0; k < N * A; ++k) sum += h_rew[k];
    int64_t flags[2] = {0, 0};
    CK(cudaMemcpy(h_flag, term, N, cudaMemcpyDeviceToHost));
    for (int64_t e = 0; e < N; ++e) flags[0] += h_flag[e];
    CK(cudaMemcpy(h_flag, trunc, N, cudaMemcpyDeviceToHost));
    for (int64_t e = 0; e < N; ++e) flags[1] += h_flag[e];
    fwrite(&sum, sizeof sum, 1, f); fwrite(flags, sizeof flags, 1, f);
  }
  uint8_t* h_obs = (uint8_t*)malloc(ob);
  CK(cudaMemcpy(h_obs, obs, ob, cudaMemcpyDeviceToHost));
  fwrite(h_obs, 1, ob, f);
  /* MultiGridEnv.render() frames of three envs through mg_render (env ids on the device, tile size 32) */
  const int32_t h_ids[3] = {0, (int32_t)(N / 2), (int32_t)(N - 1)};
  int32_t* ids; uint8_t* frames;
  const size_t fb = (size_t)3 * (H * 32) * (W * 32) * 3;
  CK(cudaMalloc((void**)&ids, sizeof h_ids)); CK(cudaMalloc((void**)&frames, fb));
  CK(cudaMemcpy(ids, h_ids, sizeof h_ids, cudaMemcpyHostToDevice));
  if (mg_render(env, state, ids, 3, 32, frames, NULL) != 0) { fprintf(stderr, "mg_render: %s\n", mg_last_error(env)); return 8; }
  uint8_t* h_frames = (uint8_t*)malloc(fb);
  CK(cudaMemcpy(h_frames, frames, fb, cudaMemcpyDeviceToHost));
  fwrite(h_frames, 1, fb, f);
  int32_t status = -1;
  if (mg_status(env, NULL, &status) != 0 || status != 0) { fprintf(stderr, "status word %d\n", status); return 7; }
  fclose(f);
  printf("c_driver ok: %lld envs x %d steps, %lld launches\n", (long long)N, steps, (long long)mg_launch_count(env));
  mg_destroy(env);
  return 0;
}
