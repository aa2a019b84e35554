// astar_host.cu -- host-only: the first-move table of the scripted CtF opponents (mg_astar_first_moves).
//
// For every (start, target) pair of a map: the action of the first move of the route the reference's A* returns
// (policy/ctf/utils.py:17-120), which is what DestinationPolicy.act turns into an action (heuristic.py:140-172).  Which
// of the equally short routes the reference returns is decided by its frontier order, so that is what is restated here:
//   * the frontier is a heap of whole records (f, g, h, parent, loc) compared as nested tuples (utils.py:9-14): equal f
//     falls through to g, h, then the PARENT CHAIN (record by record, recursively), then the cell;
//   * neighbours in the order (0,+1), (0,-1), (+1,0), (-1,0) (utils.py:64); a cell already on the frontier or already
//     expanded is replaced only by a strictly smaller f (utils.py:96-116);
//   * a cell blocks iff the caller marked it (the reference: map value 8, utils.py:73).
// Replaced frontier records are dropped lazily when they surface; the minimum of a total order does not depend on the
// container, so the pop sequence is the reference's.  Checked cell for cell against the Python restatement
// (policy/ctf/utils.py of this package, itself pinned to the reference's routes) in tests/test_policies.py.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <deque>
#include <queue>
#include <thread>
#include <vector>

#include "../../include/multigrid_b200.h"

namespace {

struct Node {
  int f, g, h;
  const Node* parent;
  int x, y;
};

int compare(const Node* a, const Node* b) {  // tuple order of (f, g, h, parent, (x, y))
  while (true) {
    if (a == b) return 0;
    if (a->f != b->f) return a->f < b->f ? -1 : 1;
    if (a->g != b->g) return a->g < b->g ? -1 : 1;
    if (a->h != b->h) return a->h < b->h ? -1 : 1;
    if (a->parent != b->parent) {            // equal g = equal depth: both chains end together
      const int c = compare(a->parent, b->parent);
      if (c) return c;
    }
    if (a->x != b->x) return a->x < b->x ? -1 : 1;
    if (a->y != b->y) return a->y < b->y ? -1 : 1;
    return 0;
  }
}

struct Later {
  bool operator()(const Node* a, const Node* b) const { return compare(a, b) > 0; }
};

inline int manhattan(int x0, int y0, int x1, int y1) { return std::abs(x0 - x1) + std::abs(y0 - y1); }

// CtfActions value of a unit step (heuristic.py:160-170), 0 = stay for anything else
inline uint8_t action_of(int dx, int dy) {
  if (dx == 0 && dy == -1) return 1;
  if (dx == -1 && dy == 0) return 2;
  if (dx == 0 && dy == 1) return 3;
  if (dx == 1 && dy == 0) return 4;
  return 0;
}

// rows [s0, s1) of the table: every target for each of these start cells
void fill_rows(const uint8_t* blocked, int rows, int cols, int s0, int s1, uint8_t* first_move) {
  const int cells = rows * cols;
  static const int DX[4] = {0, 0, 1, -1}, DY[4] = {1, -1, 0, 0};
  std::vector<const Node*> live(cells), expanded(cells);
  for (int s = s0; s < s1; ++s) {
    const int sx = s / cols, sy = s % cols;
    for (int t = 0; t < cells; ++t) {
      const int tx = t / cols, ty = t % cols;
      std::deque<Node> pool;   // stable addresses
      std::priority_queue<const Node*, std::vector<const Node*>, Later> frontier;
      std::fill(live.begin(), live.end(), nullptr);
      std::fill(expanded.begin(), expanded.end(), nullptr);
      const int h0 = manhattan(sx, sy, tx, ty);
      pool.push_back(Node{h0, 0, h0, nullptr, sx, sy});
      live[s] = &pool.back();
      frontier.push(&pool.back());
      const Node* goal = nullptr;
      while (!frontier.empty()) {
        const Node* n = frontier.top();
        frontier.pop();
        const int c = n->x * cols + n->y;
        if (live[c] != n) continue;          // a replaced record surfacing late
        live[c] = nullptr;
        expanded[c] = n;
        if (c == t) { goal = n; break; }
        for (int d = 0; d < 4; ++d) {
          const int nx = n->x + DX[d], ny = n->y + DY[d];
          if (nx < 0 || ny < 0 || nx >= rows || ny >= cols || blocked[nx * cols + ny]) continue;
          const int nc = nx * cols + ny, g = n->g + 1, h = manhattan(nx, ny, tx, ty), f = g + h;
          if (expanded[nc]) {
            if (f >= expanded[nc]->f) continue;
            expanded[nc] = nullptr;          // utils.py:96-103: back onto the frontier
          } else if (live[nc] && f >= live[nc]->f) {
            continue;
          }
          pool.push_back(Node{f, g, h, n, nx, ny});
          live[nc] = &pool.back();
          frontier.push(&pool.back());
        }
      }
      int nx = tx, ny = ty;                  // no second cell on the route (start == target, or no route): the target itself
      if (goal && goal->parent) {
        const Node* n = goal;
        while (n->parent->parent) n = n->parent;
        nx = n->x; ny = n->y;
      }
      first_move[(size_t)s * cells + t] = action_of(nx - sx, ny - sy);
    }
  }
}

}  // namespace

extern "C" int mg_astar_first_moves(const uint8_t* blocked, int32_t rows, int32_t cols, uint8_t* first_move) {
  if (!blocked || !first_move || rows < 1 || cols < 1 || (long long)rows * cols > 65535) return -1;
  const int cells = rows * cols;
  int nt = (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, std::min(cells / 16, 64)));   // the start cells are independent: one slice of rows per host thread
  if (nt == 1) {
    fill_rows(blocked, rows, cols, 0, cells, first_move);
    return 0;
  }
  std::vector<std::thread> pool;
  for (int i = 0; i < nt; ++i)
    pool.emplace_back(fill_rows, blocked, (int)rows, (int)cols, (int)((long long)cells * i / nt), (int)((long long)cells * (i + 1) / nt), first_move);
  for (auto& th : pool) th.join();
  return 0;
}
