"""Generates tests/golden/minpath_golden.npz by running the UNMODIFIED reference
graph search (loaded by path from /root/reference, build container only) on seeded maps.
Run: python tests/golden/make_minpath_golden.py"""
import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oct_image_segmentation_models_b200.common.synthetic import synthetic_bscan  # noqa: E402
from oracle import postproc  # noqa: E402

REF = "/root/reference/oct_image_segmentation_models/min_path_processing/graph_search.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_graph_search", REF)
    gs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gs)
    return gs


def cases():
    rng = np.random.default_rng(2024)
    out = {}
    # clean layered maps, several sizes
    for name, (i, h, w) in {"clean_a": (3, 96, 128), "clean_b": (5, 64, 48), "clean_c": (9, 128, 64)}.items():
        _, lab, _ = synthetic_bscan(i, h, w)
        out[name] = lab[..., 0]
    # label maps with random flips (argmax noise)
    _, lab, _ = synthetic_bscan(11, 96, 96)
    lab = lab[..., 0].copy()
    flip = rng.random(lab.shape) < 0.01
    lab[flip] = rng.integers(0, 4, size=int(flip.sum()))
    out["flipped"] = lab
    return out, rng


def main():
    gs = load_reference()
    labs, rng = cases()
    store = {}
    for name, lab in labs.items():
        probs = postproc.to_categorical(lab, 4)[None]
        _, cat = postproc.perform_argmax(probs)
        maps = postproc.convert_predictions_to_maps_semantic(cat)
        mt = postproc.maps_for_graph_search(maps[0])
        G = gs.create_graph_structure((mt.shape[1], mt.shape[2], 1))
        pred = gs.segment_maps(mt, None, G)[0]
        store[name + "_maps_t"] = mt
        store[name + "_pred"] = pred
    # raw noise maps: worst case for tie-breaks
    noise = rng.integers(0, 256, size=(3, 40, 32), dtype=np.uint8)
    G = gs.create_graph_structure((40, 32, 1))
    store["noise_maps_t"] = noise
    store["noise_pred"] = gs.segment_maps(noise, None, G)[0]
    # binary-ish maps with many exact ties
    ties = (rng.random((2, 36, 28)) < 0.15).astype(np.uint8) * 255
    G = gs.create_graph_structure((36, 28, 1))
    store["ties_maps_t"] = ties
    store["ties_pred"] = gs.segment_maps(ties, None, G)[0]
    np.savez_compressed(Path(__file__).parent / "minpath_golden.npz", **store)
    print("wrote", len(store), "arrays")


if __name__ == "__main__":
    main()
