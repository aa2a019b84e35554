// policy_kernels.cu -- the reference's scripted CtF opponents (policy/ctf/heuristic.py) decided on the device for every env:
// one thread per env writes the red team's actions [N][num_red] that the following mg_step reads (mg_set_red_actions).
//
// What is reference behaviour and what is stand-in:
//   * targets are the reference's: FightPolicy -> the closest blue agent, first of equals in index order, defeated ones
//     included (heuristic.py:216-226, utils/map.py:56-61; squared integer distances order like the float norms);
//     CapturePolicy -> the blue flag (:265-272); PatrolPolicy -> on the border a uniformly drawn border cell that has a
//     border neighbour, else the closest border cell (:321-338); PatrolFightPolicy -> fight while any blue agent stands on
//     red territory or the red flag, else patrol (:434-463);
//   * the move towards a target is the first step of the reference's A* route (policy/ctf/utils.py:17-120, tie-breaking
//     included), looked up in a [cell][target] table the host fills by running that A* (the route depends on nothing else);
//   * "follow the route with probability `randomness`, else a uniform action" (heuristic.py:150-175) and the patrol draw use
//     the env's Philox generator - the device stand-in for numpy's Generator, like RwPolicy's draw in the step kernel - on
//     blocks of their own: counter (env id, 16 * step_count + block, 2^31 | episode), so they never meet the step's blocks
//     (4th counter word 0).  Draw order per red agent as in the reference: patrol target, follow-or-not, random action.
#include "mg_device.cuh"
#include "policy_params.cuh"
#include "policy_device.cuh"
#include "../../include/multigrid_b200.h"

namespace mg {

namespace {

__global__ void __launch_bounds__(128) ctf_policy_kernel(const PolicyParams p) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.N) return;
  // one 32-bit load per agent: x | y << 8 | dir << 16 | flags << 24 (rows are 4 * 2^k bytes, 4-byte aligned)
  const uint32_t* row = reinterpret_cast<const uint32_t*>(p.agents + e * p.row_bytes);
  const int4 h = p.hdr[e];
  const unsigned long long id = p.env_id_base + (unsigned long long)e;
  PolicyRng r;
  r.open(p.seed, id, h.x, h.w);
  const int S = p.S;
  // "a blue agent stands on red ground", shared by every red agent of the env
  bool intruder = false;
  for (int i = 0; i < p.nb; ++i) {
    const uint32_t w = row[i];
    const int code = __ldg(p.field_map + (w & 255u) * S + ((w >> 8) & 255u));
    intruder |= code == 1 || code == 5;   // observation["red_territory"] = red territory cells + the red flag (ctf.py:765-769)
  }
  const bool trace = p.tr_follow != nullptr;
  for (int k = 0; k < p.nr; ++k) {
    const int kind = p.kind[k];
    const long long ek = e * p.nr + k;
    int a;
    if (kind == MG_POLICY_RW) {
      a = trace ? p.tr_action[ek] : r.below(5);
    } else {
      const uint32_t me = row[p.nb + k];
      const int x = me & 255u, y = (me >> 8) & 255u, cell = x * S + y;
      int target;
      if (kind == MG_POLICY_CAPTURE) {
        target = p.blue_flag_cell;
      } else if (kind == MG_POLICY_FIGHT || (kind == MG_POLICY_PATROL_FIGHT && intruder)) {
        int best = 0x7fffffff;
        target = cell;
        for (int i = 0; i < p.nb; ++i) {
          const uint32_t w = row[i];
          const int bx = w & 255u, by = (w >> 8) & 255u, dx = bx - x, dy = by - y, d2 = dx * dx + dy * dy;
          if (d2 < best) { best = d2; target = bx * S + by; }
        }
      } else if (__ldg(p.on_border + cell)) {
        target = trace ? (int)p.tr_patrol[ek] : (int)__ldg(p.along + r.below(p.n_along));
      } else {
        target = __ldg(p.patrol_goal + cell);
      }
      const int mv = __ldg(p.first_move + (size_t)cell * p.cells + target);
      if (trace) {
        a = p.tr_follow[ek] ? mv : p.tr_action[ek];
      } else {
        const bool follow = (unsigned long long)r.u32() < p.thr[k];
        a = follow ? mv : r.below(5);
      }
    }
    p.out[ek] = (int8_t)a;
  }
}

}  // namespace

cudaError_t launch_ctf_policy(const PolicyParams& p, cudaStream_t st) {
  const int threads = 128;
  const long long blocks = (p.N + threads - 1) / threads;
  ctf_policy_kernel<<<(unsigned)blocks, threads, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace mg
