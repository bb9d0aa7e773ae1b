"""Pins the post-processing / min-path oracle restatements against outputs produced by the
reference's own code (tests/golden/*.npz + generating scripts), and -- in the build container,
where /root/reference exists -- against the live reference."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

from oracle import min_path, postproc

GOLD = Path(__file__).parent / "golden"
REF_GS = Path("/root/reference/oct_image_segmentation_models/min_path_processing/graph_search.py")


def test_minpath_matches_reference_golden():
    g = np.load(GOLD / "minpath_golden.npz")
    names = sorted(k[:-7] for k in g.files if k.endswith("_maps_t"))
    assert len(names) >= 6
    for nm in names:
        got = min_path.segment_maps(g[nm + "_maps_t"])
        assert got.dtype == np.uint16
        assert np.array_equal(got, g[nm + "_pred"]), nm


def test_postproc_matches_reference_golden():
    g = np.load(GOLD / "postproc_golden.npz")
    for nm in ("a", "b", "ties", "edges"):
        am, cat = postproc.perform_argmax(g[nm + "_probs"], bin=True)
        assert np.array_equal(am, g[nm + "_argmax"]), nm
        assert np.array_equal(cat, g[nm + "_cat"]), nm
        maps = postproc.convert_predictions_to_maps_semantic(cat, bg_ilm=True, bg_csi=False)
        assert maps.dtype == np.uint8
        assert np.array_equal(maps, g[nm + "_maps"]), nm
    for nm in ("a", "b"):
        _, cat = postproc.perform_argmax(g[nm + "_probs"], bin=True)
        maps = postproc.convert_predictions_to_maps_semantic(cat, bg_ilm=False, bg_csi=True)
        assert np.array_equal(maps, g[nm + "_maps_csi"]), nm
    # the documented edge quirks survive: 254 at the np.roll wrap row
    assert 254 in np.unique(g["edges_maps"])


def test_trained_fixture_is_self_consistent():
    g = np.load(GOLD / "trained_small_unet.npz")
    segs = np.stack([postproc.boundaries_from_probs(g["probs"][i:i + 1]) for i in range(2)])
    assert np.array_equal(segs, g["segs"][:2])


@pytest.mark.skipif(not REF_GS.exists(), reason="reference tree only exists in the build container")
def test_minpath_matches_live_reference():
    spec = importlib.util.spec_from_file_location("ref_graph_search", str(REF_GS))
    gs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gs)
    rng = np.random.default_rng(5)
    for shape in ((2, 24, 20), (1, 33, 17)):
        maps = rng.integers(0, 256, size=shape, dtype=np.uint8)
        G = gs.create_graph_structure((shape[1], shape[2], 1))
        assert np.array_equal(gs.segment_maps(maps, None, G)[0], min_path.segment_maps(maps))
        # neighbour generator == reference adjacency lists (order matters for tie-breaks)
        gw, gh = shape[1] + 2, shape[2]
        for node in range(gw * gh):
            assert min_path.neighbours(node, gw, gh) == G[node]


def test_native_min_path_matches_reference_golden_and_oracle():
    """csrc/min_path.cpp (host code inside liboctseg.so, runs without a GPU) reproduces the
    reference's boundaries bit for bit, including noise / exact-tie maps."""
    from oct_image_segmentation_models_b200.min_path_processing import graph_search
    g = np.load(GOLD / "minpath_golden.npz")
    for nm in sorted(k[:-7] for k in g.files if k.endswith("_maps_t")):
        maps = g[nm + "_maps_t"]
        G = graph_search.create_graph_structure((maps.shape[1], maps.shape[2], 1))
        pred, err, pm = graph_search.segment_maps(maps, None, G)
        assert pred.dtype == np.uint16 and np.array_equal(pred, g[nm + "_pred"]), nm
        assert pm.dtype == np.float64
    rng = np.random.default_rng(12)
    for shape in ((3, 37, 29), (2, 64, 48), (1, 8, 1), (1, 1, 5)):
        maps = rng.integers(0, 256, size=shape, dtype=np.uint8)
        maps[rng.random(shape) < 0.3] = 255        # many exact ties
        got = graph_search.segment_maps(maps, None, None, n_threads=2)[0]
        assert np.array_equal(got, min_path.segment_maps(maps)), shape
