"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

Restatement of the post-processing between the network and the boundary search
(reference: oct_image_segmentation_models/common/utils.py):
  convert_maps_uint8                      :73-77
  perform_argmax(bin=True)                :80-112   (np.argmax = first max on ties,
                                                     to_categorical -> float32 one-hot,
                                                     transpose to [N,K,H,W])
  convert_predictions_to_maps_semantic    :115-168  (np.gradient / np.roll edge quirks
                                                     kept: 254 at the wrap row, rows 0+1)
and of how the callers hand maps to the graph search
(reference prediction/prediction.py:134-135, evaluation/evaluation.py:291-292).
"""
import numpy as np


def to_categorical(labels: np.ndarray, num_classes: int) -> np.ndarray:
    out = np.zeros(labels.shape + (num_classes,), dtype=np.float32)
    np.put_along_axis(out, labels[..., None].astype(np.int64), 1.0, axis=-1)
    return out


def perform_argmax(predictions: np.ndarray, bin: bool = True):
    """predictions [N,H,W,K] -> [argmax [N,H,W], categorical [N,K,H,W]]."""
    num_maps = predictions.shape[3]
    argmax_pred = np.argmax(predictions, axis=3)
    if bin:
        categorical_pred = np.transpose(to_categorical(argmax_pred, num_maps), (0, 3, 1, 2))
    else:
        categorical_pred = np.transpose(predictions, (0, 3, 1, 2))
    return [argmax_pred, categorical_pred]


def _to_u8(g: np.ndarray) -> np.ndarray:
    g = g * 255
    return g.astype("uint8")


def convert_predictions_to_maps_semantic(categorical_pred: np.ndarray, bg_ilm: bool = True,
                                         bg_csi: bool = False) -> np.ndarray:
    n, k, h, w = categorical_pred.shape
    out = np.zeros((n, k - 1, h, w), dtype="uint8")
    for s in range(n):
        for m in range(1, k):
            if (m == 1 and bg_ilm) or (m == k - 1 and bg_csi):
                g = -np.gradient(categorical_pred[s, m - 1], axis=0)
            else:
                g = np.gradient(categorical_pred[s, m], axis=0)
            g = np.where(g < 0, 0, g) * 2
            g = g - np.roll(g, -1, axis=0)
            g = np.where(g < 0, 0, g)
            out[s, m - 1] = _to_u8(g)
    return out


def maps_for_graph_search(boundary_maps: np.ndarray) -> np.ndarray:
    """[K-1,H,W] -> [K-1,W,H], as the callers transpose before segment_maps."""
    return np.transpose(boundary_maps, (0, 2, 1))


def boundaries_from_probs(probs: np.ndarray):
    """Full chain for one image: probs [1,H,W,K] -> uint16 [K-1,W] via the oracle
    min-path restatement."""
    from . import min_path
    _, cat = perform_argmax(probs, bin=True)
    maps = convert_predictions_to_maps_semantic(cat, bg_ilm=True, bg_csi=False)
    return min_path.segment_maps(maps_for_graph_search(maps[0]))
