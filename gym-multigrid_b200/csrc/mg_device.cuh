// mg_device.cuh -- device-side building blocks shared by the sm_100a kernels:
// packed cell codes, Philox4x32-10, the per-env random source (trace replay or Philox),
// and the TMA (cp.async.bulk) / mbarrier PTX wrappers used to move env tiles.
//
// Reference citations are file:line under the gym-multigrid checkout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/multigrid_b200.h"

namespace mg {

// ---------------------------------------------------------------------------- cell codes
// CollectWorld (core/world.py:54-64): empty 0, wall 1, ball 2, agent 3; colours
// core/constants.py:8-19 (< 10); agent state = dir (< 4, core/agent.py:119-126).
// One byte per cell: type | colour << 2 | state << 6; the empty cell is 0 = (0, 0, 0).
constexpr int T_EMPTY = 0, T_WALL = 1, T_BALL = 2, T_AGENT = 3;
__host__ __device__ constexpr uint8_t cell(int type, int colour, int state) {
  return (uint8_t)(type | (colour << 2) | (state << 6));
}
constexpr uint8_t WALL_GREY = cell(T_WALL, 7, 0);  // Wall(world) colour "grey" (object.py:174-176)

// ------------------------------------------------------------------------------- params
struct CollectParams {
  // geometry / rules
  int W, H, cells, A, nb;
  int num_balls, respawn, layout, fixed_horizon, max_steps, time_limit, autoreset;
  uint8_t agent_code[MG_MAX_AGENTS];   // cell(T_AGENT, agents_index[i], dir = 3)
  uint8_t ball_colour[MG_MAX_BALL_TYPES];
  int8_t type_of_colour[16];
  double reward_initial[16];      // fwd_cell.reward of a ball placed by _gen_grid, by colour (QuadrantsRespawn: the literal 1, collect_game.py:393)
  double reward_respawned[16];    // ... of a ball placed by _respawn: balls_reward[colour] (collect_game.py:130, :409)
  int mark_respawned;             // 1 = the two tables differ somewhere: respawned balls carry bit 6 of their cell (never shown: a ball's STATE is 0)
  long long N;            // envs on this device
  unsigned long long env_id_base, seed;
  // state planes (caller-owned buffer)
  uint8_t* grid;          // [N_pad][cells]
  uint8_t* agent_pos;     // [N_pad][A][2]
  int4* hdr;              // [N_pad] {step_count, collected, rng_ctr, episodes}
  int32_t* info;          // [N_pad][A*nb]
  const uint8_t* wall_template;  // [cells] handle-owned: the layout's walls on an empty grid (reset starts from it)
  // io
  const int8_t* actions;
  uint8_t* obs;
  double* rewards;
  uint8_t* terminated;
  uint8_t* truncated;
  uint8_t* final_obs;
  const uint8_t* reset_mask;
  // trace replay (rng_mode 0) -- pointers may be null in Philox mode (rng_mode 1)
  int rng_mode;
  const uint8_t* order;
  const uint8_t* draws;
  const int32_t* n_draws;
  int K;
  const uint8_t* reset_draws;
  const int32_t* n_reset_draws;
  int R;
  int32_t* draws_used;
  int32_t* reset_draws_used;
  int32_t* status;
  int obs_bulk_ok;        // obs base pointer is 16-byte aligned
  int io_bulk_ok;         // actions / rewards / terminated / truncated pointers are 16-byte aligned
  unsigned long long* timeline;  // optional [tiles][8] per-CTA phase timestamps (globaltimer ns), profiling only
  int early_obs;          // 1 = full tiles store the pre-step observation slab while the agents are stepped and patch the <= 3A changed cells in place
  // rollout (collect_rollout_kernels.cu): T steps per launch; actions / obs / rewards / flags are [T][N][...] arrays
  int T;
  int roll_tile;          // envs per warp (4 / 8 / 16 / 32), chosen per launch from the batch size
  int8_t* actions_out;    // uniform on-device policy (actions == NULL): the actions taken, [T][N][A], or NULL
  // compact host transport (mg_set_host_transport, MG_TRANSPORT_DELTA): one record per env with the cells the step wrote, plus the
  // packed rows of the envs that autoreset, compacted behind a device counter.  Record (delta_stride bytes, see delta_record_bytes):
  //   byte 0      n_changes (bits 0-4) | terminated << 5 | truncated << 6 | autoreset << 7
  //   bytes 1..A  per agent: 0 = no pickup, else 1 + (ball colour | respawned << 4)   (the reward is a table lookup on the host)
  //   then 3A entries (cell index u8 when W*H <= 256 else u16 little endian, packed cell code u8), n_changes of them valid
  uint8_t* delta;
  int delta_stride, delta_wide;
  int32_t* reset_count;   // zeroed by the host before the launch
  uint8_t* reset_rows;    // [N][reset_stride]: int32 env index (on this device) followed by the env's fresh packed row
  int reset_stride;       // 4 + W*H rounded up to a multiple of 4
};

__host__ __device__ inline int delta_record_bytes(int cells, int A) {
  const int entry = cells <= 256 ? 2 : 3;
  return (1 + A + 3 * A * entry + 3) & ~3;
}

// ------------------------------------------------------------------------ Philox4x32-10
// Salmon et al. SC'11.  Production-mode generator (the reference's python `random` and legacy
// numpy MT19937 streams cannot be reproduced on a device; they are replayed in trace mode).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Per-env cursor over the random source.  MODE 0 = trace replay, MODE 1 = Philox.
template <int MODE>
struct Rng {
  const uint8_t* draws; int n, k;                       // trace
  uint32_t k0, k1, id0, id1, ctr, b0, b1, b2, b3; int have;  // philox
  uint32_t h16; int nh16;                                  // ... half-word buffer of u16()
  int err;

  __device__ __forceinline__ void open_trace(const uint8_t* d, int n_) { draws = d; n = n_; k = 0; err = 0; have = 0; nh16 = 0; }
  __device__ __forceinline__ void open_philox(unsigned long long seed, unsigned long long env_id, uint32_t ctr_) {
    k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); id0 = (uint32_t)env_id; id1 = (uint32_t)(env_id >> 32);
    ctr = ctr_; have = 0; nh16 = 0; err = 0; k = 0; n = 0; draws = nullptr;
  }
  // 16 random bits: small draws (an action out of 5, a Fisher-Yates index) take half a Philox word each, so the ~5 draws of a 2v2
  // CtF step fit one Philox block instead of two (the block is ~100 of the step's ~1000 instructions)
  __device__ __forceinline__ uint32_t u16() {
    if (nh16 == 0) { h16 = u32(); nh16 = 2; }
    const uint32_t v = h16 & 0xFFFFu;
    h16 >>= 16; --nh16;
    return v;
  }
  __device__ __forceinline__ uint32_t u32() {
    if (have == 0) {
      uint32_t o[4];
      philox4x32_10(id0, id1, ctr, 0u, k0, k1, o);
      b0 = o[0]; b1 = o[1]; b2 = o[2]; b3 = o[3];
      ++ctr; have = 4;
    }
    const uint32_t v = b0;
    b0 = b1; b1 = b2; b2 = b3; --have;
    return v;
  }
  // MultiGridEnv._rand_int = random.randint(low, high): INCLUSIVE bounds (multigrid.py:225-230)
  __device__ __forceinline__ int rand_int(int lo, int hi) {
    if (MODE == 0) {
      if (k >= n) { err |= MG_ERR_TRACE_OVERFLOW; return lo; }
      const int v = draws[k++];
      if (v < lo || v > hi) err |= MG_ERR_TRACE_RANGE;
      return v;
    } else {
      return lo + (int)__umulhi(u32(), (uint32_t)(hi - lo + 1));
    }
  }
  // One placement candidate (x, y) of place_obj (multigrid.py:316-321: x is drawn before y).
  // Trace mode replays the two recorded randint outputs; Philox mode spends ONE 32-bit word per
  // candidate (x from the low half, y from the high half), i.e. four candidates per Philox block.
  __device__ __forceinline__ void rand_pair(int lox, int hix, int loy, int hiy, int& x, int& y) {
    if (MODE == 0) {
      x = rand_int(lox, hix);
      y = rand_int(loy, hiy);
    } else {
      const uint32_t w = u32();
      x = lox + (int)(((w & 0xFFFFu) * (uint32_t)(hix - lox + 1)) >> 16);
      y = loy + (int)(((w >> 16) * (uint32_t)(hiy - loy + 1)) >> 16);
    }
  }
};

// ------------------------------------------------------------------- TMA / mbarrier PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(phase)
      : "memory");
}
// barrier over a subset of the CTA's warps (id 1..15; `count` threads, a multiple of 32)
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// global -> shared bulk copy (TMA, 1-D).  16-byte aligned addresses, size % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy (TMA, 1-D).
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... until the bulk stores of this thread have been PERFORMED (their global writes are complete), not just read
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Programmatic dependent launch (PDL): the next kernel in the stream may be scheduled while this one
// drains; it blocks in pdl_wait() until every prerequisite grid has completed and flushed its writes.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------- encode
// Grid.encode, encode_dim 3 (grid.py:223-252 + object.py:58-74 + agent.py:119-126):
// 16 packed cells (one uint4) -> 48 obs bytes (three uint4), byte order (type, colour, state).
// four cells given as separate type / colour / state byte planes -> 12 interleaved obs bytes
__device__ __forceinline__ void interleave3(uint32_t t, uint32_t c, uint32_t s, uint32_t& o0, uint32_t& o1, uint32_t& o2);
// STATE is the agent's dir (type 3); a Collect ball (type 2) may carry the internal "respawned" mark in bit 6, which is not
// part of the observation: WorldObj.encode gives (type, colour, 0) for it (object.py:58-74).
// MARK = false: the caller knows no ball of this handle carries the mark (CollectParams::mark_respawned == 0, the case of every
// registered config), which saves the four masking operations per word.
template <bool MARK = true>
__device__ __forceinline__ void expand4(uint32_t w, uint32_t& o0, uint32_t& o1, uint32_t& o2) {
  uint32_t s = (w >> 6) & 0x03030303u;
  if (MARK) {
    const uint32_t ball = (w >> 1) & ~w & 0x01010101u;
    s &= ~(ball * 3u);
  }
  interleave3(w & 0x03030303u, (w >> 2) & 0x0F0F0F0Fu, s, o0, o1, o2);
}
__device__ __forceinline__ uint8_t state_of(uint8_t c) { return (c & 3) == T_BALL ? 0 : (uint8_t)(c >> 6); }
__device__ __forceinline__ void interleave3(uint32_t t, uint32_t c, uint32_t s, uint32_t& o0, uint32_t& o1, uint32_t& o2) {
  // out bytes: t0 c0 s0 t1 | c1 s1 t2 c2 | s2 t3 c3 s3
  const uint32_t tc = __byte_perm(t, c, 0x5140);   // t0 c0 t1 c1  (bytes: [t0, c0, t1, c1])
  o0 = __byte_perm(tc, s, 0x2410);                 // t0 c0 s0 t1
  const uint32_t cs = __byte_perm(c, s, 0x6251);   // c1 s1 c2 s2
  o1 = __byte_perm(cs, t, 0x2610);                 // c1 s1 t2 c2
  const uint32_t st = __byte_perm(s, t, 0x3072);   // s2 t3 .  s3
  o2 = __byte_perm(st, c, 0x3710);                 // s2 t3 c3 s3
}

// the same with an all-zero state plane, three PRMTs instead of six: a selector nibble with bit 3 set replicates the SIGN of the
// byte it names, and every type / colour byte is < 128, so such a nibble yields 0x00 (PTX prmt, default mode; the __byte_perm
// intrinsic masks that bit away, hence the asm)
__device__ __forceinline__ uint32_t prmt_raw(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ void interleave3_zero_state(uint32_t t, uint32_t c, uint32_t& o0, uint32_t& o1, uint32_t& o2) {
  o0 = prmt_raw(t, c, 0x1840u);   // t0 c0 0  t1
  o1 = prmt_raw(t, c, 0x6285u);   // c1 0  t2 c2
  o2 = prmt_raw(t, c, 0x8738u);   // 0  t3 c3 0
}

template <bool MARK = true>
__device__ __forceinline__ void expand16(const uint4 in, uint4& a, uint4& b, uint4& c) {
  uint32_t o[12];
  expand4<MARK>(in.x, o[0], o[1], o[2]);
  expand4<MARK>(in.y, o[3], o[4], o[5]);
  expand4<MARK>(in.z, o[6], o[7], o[8]);
  expand4<MARK>(in.w, o[9], o[10], o[11]);
  a = make_uint4(o[0], o[1], o[2], o[3]);
  b = make_uint4(o[4], o[5], o[6], o[7]);
  c = make_uint4(o[8], o[9], o[10], o[11]);
}

}  // namespace mg
