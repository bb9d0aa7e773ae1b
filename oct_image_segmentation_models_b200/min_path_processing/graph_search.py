"""Boundary search with the reference's function names (min_path_processing/graph_search.py).

`segment_maps(prob_maps, truths, graph_structure)` keeps the reference's signature and return
triple; the Dijkstra itself runs in liboctseg.so (csrc/min_path.cpp, exact tie-breaking).  The
adjacency-list graph the reference builds per image (0.26 s at 512x512) is not needed natively:
`create_graph_structure` returns a small descriptor so existing call sites keep working.
"""
import ctypes as C
import os

import numpy as np

from .. import _native as nat


def create_graph_structure(shape, max_grad=1):
    if max_grad != 1:
        raise NotImplementedError("only max_grad=1 (the reference's default and only use) is supported")
    return {"width": int(shape[0]), "height": int(shape[1]), "max_grad": 1}


def calc_errors(prediction, truth):
    width = prediction.shape[0]
    error = np.zeros((width,), dtype="float64")
    for i in range(width):
        if np.isnan(truth[i]) or truth[i] <= 0:
            error[i] = np.nan
        else:
            error[i] = prediction[i].astype("float64") - truth[i]
    return error


def segment_maps(prob_maps, truths, graph_structure=None, n_threads=None, return_prob_maps=True):
    """prob_maps: uint8 [n_maps, width, height]; returns (predictions uint16 [n_maps, width],
    errors float64 [n_maps, width], prob_maps / 255) as the reference does (graph_search.py:519-572).
    return_prob_maps=False skips the third value (None): the float64 copy of every map costs ~8x the search itself
    (2 MB written per 512x512 map against 0.24 ms of native Dijkstra) and none of the pipeline's callers reads it."""
    maps = np.ascontiguousarray(prob_maps, dtype=np.uint8)
    n_maps, width, height = maps.shape
    if graph_structure is not None and isinstance(graph_structure, dict):
        if graph_structure["width"] != width or graph_structure["height"] != height:
            raise ValueError("graph structure does not match the map shape")
    predictions = np.zeros((n_maps, width), dtype="uint16")
    if n_threads is None:
        n_threads = min(n_maps, os.cpu_count() or 1)
    nat.check(nat.load().octseg_min_path_segment(maps.ctypes.data_as(C.c_void_p), n_maps, width, height,
                                                 predictions.ctypes.data_as(C.c_void_p), int(n_threads)))
    errors = np.zeros((n_maps, width), dtype="float64")
    if truths is not None:
        for m in range(n_maps):
            errors[m:, ] = calc_errors(predictions[m], truths[m, :])   # reference quirk kept (:568-570)
    return predictions, errors, (maps / 255 if return_prob_maps else None)


def calculate_overall_errors(errors):
    nb = errors.shape[0]
    out = [np.zeros((nb,), dtype="float64") for _ in range(4)]
    for b in range(nb):
        out[0][b] = np.nanmean(np.abs(errors[b]))
        out[1][b] = np.nanmean(errors[b])
        out[2][b] = np.nanstd(np.abs(errors[b]))
        out[3][b] = np.nanstd(errors[b])
    return out
