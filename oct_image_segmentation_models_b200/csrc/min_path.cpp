// Boundary extraction after the network: exact-tie-break Dijkstra on the
// (W+2) x H grid graph, a native restatement of what the reference does in pure Python
// (reference min_path_processing/graph_search.py: run_dijkstras :5-105, create_graph_structure
// :108-225 with max_grad=1, append_firstlast_cols :337-357, delineate_boundary :360-428,
// segment_maps :519-572).  Host code: the reference keeps this stage on the CPU and so do we; it
// lives in liboctseg.so so that predict(graph_search=True) needs nothing but the library.
//
// Behaviour that must be reproduced bit for bit (tests/golden/minpath_golden.npz):
//   * heap order = (dist, prio, insertion counter): prio 0 for the same-column "down" neighbour,
//     otherwise 1 + position in the reference's neighbour list;
//   * edge cost = 2 - (p_u + p_v) in float64, p = uint8 / 255 (no clamp: `np.max(x, 0)` is an axis);
//   * search stops when the bottom-right node is finalised; back-trace writes the row of every
//     interior column; result is stored as uint16.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/octseg.h"

namespace octseg { void set_error(const std::string &msg); }

namespace {

struct Entry {
  double d;
  int32_t prio;
  int64_t cnt;
  int32_t n, v;
};
inline bool less_than(const Entry &a, const Entry &b) {
  if (a.d != b.d) return a.d < b.d;
  if (a.prio != b.prio) return a.prio < b.prio;
  return a.cnt < b.cnt;
}
struct MinHeap {
  std::vector<Entry> h;
  void push(const Entry &e) {
    h.push_back(e);
    size_t i = h.size() - 1;
    while (i > 0) {
      size_t p = (i - 1) / 2;
      if (!less_than(h[i], h[p])) break;
      std::swap(h[i], h[p]);
      i = p;
    }
  }
  Entry pop() {
    Entry top = h[0];
    h[0] = h.back();
    h.pop_back();
    size_t i = 0, n = h.size();
    for (;;) {
      size_t l = 2 * i + 1, r = l + 1, m = i;
      if (l < n && less_than(h[l], h[m])) m = l;
      if (r < n && less_than(h[r], h[m])) m = r;
      if (m == i) break;
      std::swap(h[i], h[m]);
      i = m;
    }
    return top;
  }
  bool empty() const { return h.empty(); }
};

// neighbours of node (row i, col j) in reference order; returns count
inline int neighbours(int i, int j, int gw, int gh, int32_t out[4]) {
  const int right = (j + 1) + i * gw, down = j + (i + 1) * gw;
  const int dup = (j + 1) + (i - 1) * gw, ddown = (j + 1) + (i + 1) * gw;
  const bool last_col = j == gw - 1, first_col = j == 0;
  int n = 0;
  if (i == gh - 1) {
    if (last_col) return 0;
    out[n++] = right;
    if (i - 1 >= 0) out[n++] = dup;
    return n;
  }
  if (i == 0) {
    if (last_col) { out[n++] = down; return n; }
    out[n++] = right;
    if (first_col) out[n++] = down;
    out[n++] = ddown;
    return n;
  }
  if (last_col) { out[n++] = down; return n; }
  out[n++] = right;
  if (first_col) out[n++] = down;
  out[n++] = dup;
  out[n++] = ddown;
  return n;
}

// Per-thread scratch, reused over the maps a thread processes.  The search touches nodes in an irregular order, so
// the node state is kept as small as it can be -- ONE byte per node: 0 = open, 1..4 = finalised, reached by the move
// right / down / diagonal-up / diagonal-down (the predecessor is implied), 5 = the start node -- i.e. 263 KB for a
// 512 x 512 map instead of 3.4 MB of double / int32 / flag arrays that had to be allocated, zero-filled and faulted in
// per map (that, not the search, was most of the time, and it did not scale over threads).  Probabilities come
// straight from the uint8 map through a 256-entry table of the reference's float64 uint8 / 255.
struct Scratch {
  std::vector<uint8_t> state;
  std::vector<double> delin;
  MinHeap q;
  double lut[256];
  Scratch() { for (int i = 0; i < 256; ++i) lut[i] = (double)i / 255.0; }
  void fit(size_t n_nodes, size_t width) {
    if (state.size() < n_nodes) state.resize(n_nodes);
    std::memset(state.data(), 0, n_nodes);
    if (delin.size() < width) delin.resize(width);
  }
};

// one boundary map [W][H] uint8 -> rows[W] uint16
void delineate(const uint8_t *map_t, int W, int H, uint16_t *rows, Scratch &S) {
  const int gw = W + 2, gh = H;
  const int n_nodes = gw * gh, max_ind = n_nodes - 1;
  S.fit((size_t)n_nodes, (size_t)W);
  uint8_t *state = S.state.data();
  const double *lut = S.lut;
  // probability of node (row r, col c): appended first / last columns are 1.0, the others uint8 / 255 of the
  // TRANSPOSED map (callers hand maps as [W][H], reference prediction/prediction.py:134-135)
  auto prob = [&](int r, int c) -> double {
    return (c == 0 || c == gw - 1) ? 1.0 : lut[map_t[(size_t)(c - 1) * H + r]];
  };
  MinHeap &q = S.q;
  q.h.clear();
  if (q.h.capacity() < 4 * (size_t)gw + 1024) q.h.reserve(4 * (size_t)gw + 1024);
  q.push({0.0, 0, 0, 0, 0});
  int64_t add_count = 1;
  while (!q.empty()) {
    const Entry e = q.pop();
    if (state[e.n]) continue;
    const int vr = e.n / gw, vc = e.n % gw;
    {
      // how was this node reached from e.v?  (same node for the start entry)
      const int pr = e.v / gw, pc = e.v % gw;
      uint8_t code = 5;
      if (e.v != e.n) code = (pr == vr) ? 1 : (pc == vc) ? 2 : (pr == vr + 1) ? 3 : 4;
      state[e.n] = code;
    }
    if (e.n == max_ind) break;
    const double pv = prob(vr, vc);
    int32_t nb[4];
    const int cnt = neighbours(vr, vc, gw, gh, nb);
    for (int i = 0; i < cnt; ++i) {
      const int n = nb[i];
      if (state[n]) continue;
      const int nr = n / gw, nc = n % gw;
      const double edge = 2.0 - (pv + prob(nr, nc));
      const int prio = (nc == vc && nr == vr + 1) ? 0 : i + 1;
      q.push({e.d + edge, prio, add_count, n, e.n});
      ++add_count;
    }
  }
  double *delin = S.delin.data();
  std::fill(delin, delin + W, 0.0);
  // walk back to (0,0); the reference appends coords then writes delin[col-1] = row in list order,
  // i.e. the node CLOSEST to the start wins for a column visited more than once
  int cr = max_ind / gw, cc = max_ind % gw;
  while (!(cc == 0 && cr == 0)) {
    if (cc != 0 && cc != gw - 1) delin[cc - 1] = cr;
    switch (state[(size_t)cr * gw + cc]) {
      case 1: cc -= 1; break;               // came by "right"
      case 2: cr -= 1; break;               // "down"
      case 3: cr += 1; cc -= 1; break;      // "diagonal up"
      case 4: cr -= 1; cc -= 1; break;      // "diagonal down"
      default: cr = 0; cc = 0; break;       // unreachable: would mean a broken chain
    }
  }
  for (int c = 0; c < W; ++c) rows[c] = (uint16_t)delin[c];
}

}  // namespace

extern "C" int32_t octseg_min_path_segment(const uint8_t *maps_t, int32_t n_maps, int32_t width, int32_t height,
                                           uint16_t *rows_out, int32_t n_threads) {
  if (!maps_t || !rows_out) { octseg::set_error("null argument"); return 1; }
  if (n_maps < 0 || width <= 0 || height <= 0) { octseg::set_error("bad map shape"); return 1; }
  const int nt = std::max(1, std::min(n_threads <= 0 ? 1 : n_threads, n_maps));
  auto work = [&](int t) {
    Scratch S;      // one per worker thread, reused over its maps
    for (int m = t; m < n_maps; m += nt)
      delineate(maps_t + (size_t)m * width * height, width, height, rows_out + (size_t)m * width, S);
  };
  if (nt == 1) { work(0); return 0; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
  for (auto &x : th) x.join();
  return 0;
}
