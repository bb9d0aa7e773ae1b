"""Generates tests/golden/postproc_golden.npz by EXECUTING the reference's own numpy
post-processing functions (source text read from /root/reference at generation time and
exec'd with stubs for the two Keras symbols they touch).  Nothing is copied into the repo.
Run: python tests/golden/make_postproc_golden.py"""
import ast
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oct_image_segmentation_models_b200.common.synthetic import synthetic_bscan  # noqa: E402

REF = "/root/reference/oct_image_segmentation_models/common/utils.py"
WANTED = {"convert_maps_uint8", "perform_argmax", "convert_predictions_to_maps_semantic"}


def load_reference_functions():
    src = Path(REF).read_text()
    tree = ast.parse(src)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    for n in body:
        n.decorator_list = []
    mod = ast.Module(body=body, type_ignores=[])

    class _K:
        @staticmethod
        def image_data_format():
            return "channels_last"

    def to_categorical(y, num_classes):
        y = np.asarray(y, dtype="int64")
        out = np.zeros(y.shape + (num_classes,), dtype="float32")
        np.put_along_axis(out, y[..., None], 1.0, axis=-1)
        return out

    ns = {"np": np, "K": _K, "to_categorical": to_categorical}
    exec(compile(mod, REF, "exec"), ns)
    return ns


def main():
    ns = load_reference_functions()
    rng = np.random.default_rng(77)
    store = {}
    for name, (i, h, w) in {"a": (2, 48, 40), "b": (8, 64, 32)}.items():
        _, lab, _ = synthetic_bscan(i, h, w)
        probs = rng.random((1, h, w, 4)).astype(np.float32) * 0.2
        np.put_along_axis(probs, lab[None].astype(np.int64), 1.0, axis=-1)
        probs /= probs.sum(-1, keepdims=True)
        am, cat = ns["perform_argmax"](probs, bin=True)
        maps = ns["convert_predictions_to_maps_semantic"](np.array(cat), bg_ilm=True, bg_csi=False)
        maps2 = ns["convert_predictions_to_maps_semantic"](np.array(cat), bg_ilm=False, bg_csi=True)
        store[f"{name}_probs"], store[f"{name}_argmax"], store[f"{name}_cat"] = probs, am, cat
        store[f"{name}_maps"], store[f"{name}_maps_csi"] = maps, maps2
    # edge quirks: boundary at row 1 and at the last row, plus exact ties in the probabilities
    lab = np.zeros((1, 12, 6), dtype=np.int64)
    lab[:, 1:, :] = 1
    lab[:, 6:, :] = 2
    lab[:, 11:, :] = 3
    probs = np.full((1, 12, 6, 4), 0.25, dtype=np.float32)   # all ties -> argmax 0 everywhere
    probs2 = np.zeros((1, 12, 6, 4), dtype=np.float32)
    np.put_along_axis(probs2, lab[..., None], 1.0, axis=-1)
    for nm, p in (("ties", probs), ("edges", probs2)):
        am, cat = ns["perform_argmax"](p, bin=True)
        store[f"{nm}_probs"], store[f"{nm}_argmax"], store[f"{nm}_cat"] = p, am, cat
        store[f"{nm}_maps"] = ns["convert_predictions_to_maps_semantic"](np.array(cat), bg_ilm=True, bg_csi=False)
    np.savez_compressed(Path(__file__).parent / "postproc_golden.npz", **store)
    print("wrote", len(store), "arrays; edge map uniques:", np.unique(store["edges_maps"]))


if __name__ == "__main__":
    main()
