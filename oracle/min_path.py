"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (comparator for boundary parity).

Restatement of the reference boundary extraction
(reference: oct_image_segmentation_models/min_path_processing/graph_search.py):
  create_graph_structure :108-225  -> neighbours() below (generated on the fly,
                                      same neighbour ORDER, which feeds the tie-break)
  run_dijkstras          :5-105    -> _dijkstra()
  append_firstlast_cols  :337-357, delineate_boundary :360-428 -> delineate_boundary()
  segment_maps           :519-572  -> segment_maps()

PINNED: tests/test_oracle_minpath.py runs the unmodified reference file (loaded
by path, in the build container only) against this restatement on clean, flipped
and pure-noise maps, and tests/golden/minpath_*.npz hold reference outputs for
the GPU box where /root/reference does not exist.

Behaviour that must be kept:
  * heap entries are (dist, prio, insertion_counter, node, prev); prio = 0 for the
    same-column "down" neighbour, else 1 + position in the neighbour list;
  * edge cost = 2 - (p_u + p_v) in float64 -- `np.max(x, 0)` in the reference is an
    axis argument, not a clamp;
  * early exit when the bottom-right node is finalised;
  * result cast to uint16 by assignment into a uint16 array.
"""
from heapq import heappop, heappush

import numpy as np


def neighbours(node: int, gw: int, gh: int):
    """Neighbour list of `node` in the (W+2) x H grid, reference order
    (graph_search.py:139-223, max_grad=1)."""
    i, j = divmod(node, gw)
    right = (j + 1) + i * gw
    down = j + (i + 1) * gw
    dup = (j + 1) + (i - 1) * gw
    ddown = (j + 1) + (i + 1) * gw
    last_col, first_col = j == gw - 1, j == 0
    if i == gh - 1:                       # last row
        if last_col:
            return []
        return [right, dup] if i - 1 >= 0 else [right]
    if i == 0:                            # first row (gh > 1 here)
        if last_col:
            return [down]
        if first_col:
            return [right, down, ddown]
        return [right, ddown]
    if last_col:                          # middle rows
        return [down]
    if first_col:
        return [right, down, dup, ddown]
    return [right, dup, ddown]


def _dijkstra(prob_map: np.ndarray):
    """prob_map: float64 [W+2, H].  Returns prev[] for finalised nodes (-1 otherwise)."""
    gw, gh = prob_map.shape
    n_nodes = gw * gh
    max_ind = n_nodes - 1
    done = np.zeros(n_nodes, dtype=bool)
    prev = np.full(n_nodes, -1, dtype=np.int64)
    pm = prob_map  # indexed [col][row]
    q = [(0, 0, 0, 0, 0)]
    add_count = 1
    while q:
        path_len, _, _, v, a = heappop(q)
        if done[v]:
            continue
        done[v] = True
        prev[v] = a
        if v == max_ind:
            break
        vr, vc = divmod(v, gw)
        pv = pm[vc][vr]
        for i, n in enumerate(neighbours(v, gw, gh)):
            nr, nc = divmod(n, gw)
            edge_len = 2 - (pv + pm[nc][nr])
            if not done[n]:
                prio = 0 if (nc == vc and nr == vr + 1) else i + 1
                heappush(q, (path_len + edge_len, prio, add_count, n, v))
                add_count += 1
    return prev


def delineate_boundary(prob_map: np.ndarray) -> np.ndarray:
    """prob_map float64 [W, H] in [0,1] -> float64 [W] row per column."""
    h = prob_map.shape[1]
    pm = np.concatenate((np.ones((1, h)), prob_map, np.ones((1, h))), axis=0)
    prev = _dijkstra(pm)
    gw, gh = pm.shape
    node = gw * gh - 1
    delin = np.zeros(gw - 2)
    coords = []
    coord = (node % gw, node // gw)
    p = prev[node]
    while coord != (0, 0):
        coords.append(coord)
        coord = (p % gw, p // gw)
        p = prev[p]
    for c, r in coords:
        if c != 0 and c != gw - 1:
            delin[c - 1] = r
    return delin


def segment_maps(boundary_maps_t: np.ndarray) -> np.ndarray:
    """boundary_maps_t: uint8 [K-1, W, H] (already transposed as the reference
    callers do, prediction.py:134-135) -> uint16 [K-1, W] boundary rows."""
    pm = boundary_maps_t / 255
    out = np.zeros((pm.shape[0], pm.shape[1]), dtype="uint16")
    for m in range(pm.shape[0]):
        out[m, :] = delineate_boundary(pm[m])
    return out
