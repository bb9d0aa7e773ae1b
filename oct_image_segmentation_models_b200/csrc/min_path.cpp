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

// one boundary map [W][H] uint8 -> rows[W] uint16
void delineate(const uint8_t *map_t, int W, int H, uint16_t *rows) {
  const int gw = W + 2, gh = H;
  const int n_nodes = gw * gh, max_ind = n_nodes - 1;
  // probability of node (col c, row r): appended first/last columns are 1.0
  std::vector<double> prob((size_t)n_nodes);
  for (int r = 0; r < gh; ++r) {
    prob[(size_t)r * gw] = 1.0;
    prob[(size_t)r * gw + gw - 1] = 1.0;
    for (int c = 1; c < gw - 1; ++c) prob[(size_t)r * gw + c] = (double)map_t[(size_t)(c - 1) * H + r] / 255.0;
  }
  std::vector<uint8_t> done((size_t)n_nodes, 0);
  std::vector<int32_t> prev((size_t)n_nodes, -1);
  MinHeap q;
  q.h.reserve(4 * (size_t)gw + 1024);
  q.push({0.0, 0, 0, 0, 0});
  int64_t add_count = 1;
  while (!q.empty()) {
    const Entry e = q.pop();
    if (done[e.n]) continue;
    done[e.n] = 1;
    prev[e.n] = e.v;
    if (e.n == max_ind) break;
    const int vr = e.n / gw, vc = e.n % gw;
    const double pv = prob[e.n];
    int32_t nb[4];
    const int cnt = neighbours(vr, vc, gw, gh, nb);
    for (int i = 0; i < cnt; ++i) {
      const int n = nb[i];
      if (done[n]) continue;
      const double edge = 2.0 - (pv + prob[n]);
      const int nr = n / gw, nc = n % gw;
      const int prio = (nc == vc && nr == vr + 1) ? 0 : i + 1;
      q.push({e.d + edge, prio, add_count, n, e.n});
      ++add_count;
    }
  }
  std::vector<double> delin((size_t)W, 0.0);
  int node = max_ind;
  int cc = node % gw, cr = node / gw;
  int p = prev[node];
  // walk back to (0,0); the reference appends coords then writes delin[col-1] = row in list order,
  // i.e. the node CLOSEST to the start wins for a column visited more than once
  while (!(cc == 0 && cr == 0)) {
    if (cc != 0 && cc != gw - 1) delin[cc - 1] = cr;
    cc = p % gw; cr = p / gw;
    p = prev[p];
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
    for (int m = t; m < n_maps; m += nt)
      delineate(maps_t + (size_t)m * width * height, width, height, rows_out + (size_t)m * width);
  };
  if (nt == 1) { work(0); return 0; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
  for (auto &x : th) x.join();
  return 0;
}
