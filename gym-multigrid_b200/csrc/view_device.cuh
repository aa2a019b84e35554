// view_device.cuh -- the per-view body of the fast partial-view path (MultiGridEnv.gen_obs, encode_dim 3), shared by
// view_fast_kernel (view_kernels.cu) and the fused Maze step + partial-view mode of map_kernel (map_kernels.cu).
#pragma once
#include "mg_device.cuh"

namespace mg {

// V-bit mask of the t in [0, V) with 0 <= u0 + s*t < L  (s = +1 / -1)
__device__ __forceinline__ uint32_t range_mask(int u0, int s, int L, int V) {
  int lo = s > 0 ? -u0 : u0 - (L - 1), hi = s > 0 ? L - 1 - u0 : u0;
  lo = max(lo, 0); hi = min(hi, V - 1);
  return hi < lo ? 0u : (((2u << hi) - 1u) & ~((1u << lo) - 1u));
}

__host__ __device__ constexpr int view_guard_bytes(int V, int H) { return ((V - 1) * (H + 1) + 15) / 16 * 16; }

// Geometry of one view: world cell of view cell (a, b) = (x0 + a*ax + b*bx, y0 + a*ay + b*by); in a row-major source of row
// pitch `pitch` its linear index is i0 + a*sa + b*sb.  slice + (dir + 1) x rotate_left folded (agent.py:294-324, grid.py:97-130).
template <int V>
__device__ __forceinline__ void view_geometry(int x, int y, int dir, int pitch, int& x0, int& y0, int& sa, int& sb) {
  constexpr int HS = V / 2;
  if (dir == 3)      { x0 = x - HS;      y0 = y - (V - 1); sa = pitch;  sb = 1; }       // facing up
  else if (dir == 1) { x0 = x + HS;      y0 = y + (V - 1); sa = -pitch; sb = -1; }      // facing down
  else if (dir == 0) { x0 = x + (V - 1); y0 = y - HS;      sa = 1;      sb = -pitch; }  // facing right
  else               { x0 = x - (V - 1); y0 = y + HS;      sa = -1;     sb = pitch; }   // facing left
}

// One view: gather V x V cells from `src` (= address of view cell (0, 0); reads may leave the env's grid, the caller
// provides guard bands / a padded map), visibility sweep, encode, and store the 3*V*V bytes at s_out + v * 3*V*V.
//   COLLECT: cells outside the range masks mA (over a) / mB (over b) become `oob_code`; walls block sight.
//   Maze (COLLECT = false): the source is pre-padded; view cell (V/2, V-1) shows `agent_cell`; only `oob_code` blocks sight.
//   Both: mA / mB = the view cells (over a / over b) that lie inside the grid / map.
template <bool COLLECT, int V>
__device__ __forceinline__ void view_compute_store(const uint8_t* src, int sa, int sb, uint32_t mA, uint32_t mB, uint32_t oob_code,
                                                   uint32_t agent_cell, bool see_through, uint8_t* s_out, int v) {
  static_assert(V % 2 == 1 && V <= 7, "fast path: odd view sizes up to 7");
  constexpr int VV = V * V, HS = V / 2, NPK = (VV + 3) / 4, NW = (3 * VV + 1) / 4, NFULL = (3 * VV - 3) / 4;
  constexpr uint32_t FULL = (1u << V) - 1u;
  uint32_t pk[NPK], opq[V], msk[V];
#pragma unroll
  for (int k = 0; k < NPK; ++k) pk[k] = 0;
#pragma unroll
  for (int b = 0; b < V; ++b) {
    const uint32_t rowv = ((mB >> b) & 1u) ? mA : 0u;
    uint32_t o = 0;
    if (COLLECT) {
#pragma unroll
      for (int a = 0; a < V; ++a) {
        uint32_t c = src[a * sa + b * sb];
        c = ((rowv >> a) & 1u) ? c : oob_code;
        o |= (uint32_t)((c & 3u) == (uint32_t)T_WALL) << a;     // see_behind() is False only for Wall (object.py:174-179)
        pk[(a * V + b) / 4] |= c << (8 * ((a * V + b) % 4));
      }
    }
    // Maze: only the out-of-map filler blocks sight, and which view cells lie outside the map is geometry - the complement of the
    // range masks - not something to find by comparing 49 cell values with the filler code; the cells are gathered AFTER the
    // visibility sweep below, and only the visible ones
    opq[b] = COLLECT ? o : (~rowv & FULL);
    msk[b] = 0;
  }
  if (see_through) {
#pragma unroll
    for (int b = 0; b < V; ++b) msk[b] = FULL;
  } else {  // process_vis (grid.py:286-323): rows bottom-up; inside a row left->right, then right->left
    msk[V - 1] = 1u << HS;
#pragma unroll
    for (int j = V - 1; j >= 0; --j) {
      const uint32_t clear = ~opq[j] & FULL;
      uint32_t m = msk[j];
      // left -> right: F = cells that are visible AND transparent once the sweep has passed them
      uint32_t F = m & clear, P = clear;
      F |= P & (F << 1); P &= P << 1;
      F |= P & (F << 2);
      if (V > 4) { P &= P << 2; F |= P & (F << 4); }
      F &= FULL >> 1;                       // the loop runs i = 0 .. V-2
      m |= F << 1;
      uint32_t up = F | (F << 1);
      // right -> left
      uint32_t G = m & clear; P = clear;
      G |= P & (G >> 1); P &= P >> 1;
      G |= P & (G >> 2);
      if (V > 4) { P &= P >> 2; G |= P & (G >> 4); }
      G &= ~1u;                             // the loop runs i = V-1 .. 1
      m |= G >> 1;
      up |= G | (G >> 1);
      msk[j] = m;
      if (j > 0) msk[j - 1] |= up;
    }
  }
  // encode_for_agents: cells outside the mask stay (0, 0, 0)
  if (COLLECT) {
#pragma unroll
    for (int a = 0; a < V; ++a)
#pragma unroll
      for (int b = 0; b < V; ++b)
        if (!((msk[b] >> a) & 1u)) pk[(a * V + b) / 4] &= ~(0xFFu << (8 * ((a * V + b) % 4)));
  } else {   // Maze: the mask is known before a single cell has been read - load the visible cells, leave the others zero
#pragma unroll
    for (int b = 0; b < V; ++b)
#pragma unroll
      for (int a = 0; a < V; ++a) {
        uint32_t c = 0;
        if ((msk[b] >> a) & 1u) c = src[a * sa + b * sb];
        if (a == HS && b == V - 1) c = agent_cell;  // the agent stands at view cell (V/2, V-1), always visible (the sweep starts there)
        pk[(a * V + b) / 4] |= c << (8 * ((a * V + b) % 4));
      }
  }
  uint32_t w[NW + 2];
#pragma unroll
  for (int k = 0; k < NPK; ++k) {
    uint32_t o0, o1, o2;
    expand4<COLLECT>(pk[k], o0, o1, o2);   // only Collect grids can hold a marked ball (bit 6 of a type-2 cell); Maze type 2 is the flag, state 0
    w[3 * k] = o0;
    if (3 * k + 1 < NW + 2) w[3 * k + 1] = o1;
    if (3 * k + 2 < NW + 2) w[3 * k + 2] = o2;
  }
  // 3*VV bytes at byte offset v*3*VV: h head bytes up to the next word boundary, NFULL aligned words, 3-h tail bytes
  uint8_t* dst = s_out + (size_t)v * (3 * VV);
  const int h = (int)((4u - ((uint32_t)v * (3u * VV) & 3u)) & 3u);
  uint32_t* q = reinterpret_cast<uint32_t*>(dst + h);
#pragma unroll
  for (int j = 0; j < NFULL; ++j) q[j] = __funnelshift_r(w[j], w[j + 1], 8 * h);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const bool head = i < h;
    dst[head ? i : 4 * NFULL + i] = (uint8_t)((head ? w[0] : w[NFULL]) >> (8 * i));
  }
}

// A memoised view (MapParams::view_table): copy the 3*V*V bytes of `row` (16-byte aligned, zero padded) to s_out + v * 3*V*V - the same
// head / aligned words / tail placement as view_compute_store, since consecutive views of a tile are 3*V*V (odd) bytes apart.
template <int V>
__device__ __forceinline__ void view_copy_store(const uint4* row, uint8_t* s_out, int v) {
  constexpr int VV = V * V, NFULL = (3 * VV - 3) / 4, NQ = (NFULL + 1 + 3) / 4;
  uint32_t w[4 * NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    const uint4 q = __ldg(row + k);
    w[4 * k] = q.x; w[4 * k + 1] = q.y; w[4 * k + 2] = q.z; w[4 * k + 3] = q.w;
  }
  uint8_t* dst = s_out + (size_t)v * (3 * VV);
  const int h = (int)((4u - ((uint32_t)v * (3u * VV) & 3u)) & 3u);
  uint32_t* q = reinterpret_cast<uint32_t*>(dst + h);
#pragma unroll
  for (int j = 0; j < NFULL; ++j) q[j] = __funnelshift_r(w[j], w[j + 1], 8 * h);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const bool head = i < h;
    dst[head ? i : 4 * NFULL + i] = (uint8_t)((head ? w[0] : w[NFULL]) >> (8 * i));
  }
}

}  // namespace mg
