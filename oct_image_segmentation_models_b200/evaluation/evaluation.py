"""`evaluate_model(EvaluationParameters) -> List[EvaluationOutput]` (reference
evaluation/evaluation.py:73-448).

Same contract per image (labels, one-hot prediction, boundary maps, optional graph-search boundaries
and their errors against the ground-truth boundaries, Dice metrics); the forward pass runs batched on
the GPU, argmax + boundary maps come from the device kernel, the boundary search from the native
min-path.  HDF5/CSV/PNG result writers and the dataset-level aggregation are out of scope
(SURVEY section 2 #5): per-image arrays are returned and, when a save flag asks, stored as .npz.
"""
import logging as log
from pathlib import Path
from typing import List, Optional

import numpy as np

from ..common import (EVALUATION_METRIC_DICE_CLASSES, EVALUATION_METRIC_DICE_MACRO, EVALUATION_METRIC_DICE_MICRO,
                      custom_metrics, dataset_loader as dl, utils)
from ..min_path_processing import graph_search
from ..min_path_processing.utils import generate_boundary
from ..models import get_model_class
from .evaluation_parameters import EvaluationParameters


class EvaluationOutput:
    def __init__(self, image: np.ndarray, image_name: Path, image_segments: np.ndarray, image_output_dir: Path,
                 predicted_labels: np.ndarray, categorical_pred: np.ndarray, boundary_maps: np.ndarray,
                 gs_pred_segs: Optional[np.ndarray], errors: Optional[np.ndarray], mean_abs_err: Optional[np.ndarray],
                 mean_err: Optional[np.ndarray], abs_err_sd: Optional[np.ndarray], err_sd: Optional[np.ndarray],
                 metrics: Optional[dict] = None) -> None:
        self.image = image
        self.image_name = image_name
        self.image_segments = image_segments
        self.image_output_dir = image_output_dir
        self.predicted_labels = predicted_labels
        self.categorical_pred = categorical_pred
        self.boundary_maps = boundary_maps
        self.gs_pred_segs = gs_pred_segs
        self.errors = errors
        self.mean_abs_err = mean_abs_err
        self.mean_err = mean_err
        self.abs_err_sd = abs_err_sd
        self.err_sd = err_sd
        self.metrics = metrics or {}


def evaluate_model(eval_params: EvaluationParameters, batch_size: int = 64) -> List[EvaluationOutput]:
    ds = dl.open_dataset(eval_params.test_dataset_path)
    eval_images, eval_labels, eval_image_names = dl.load_testing_data(ds)
    out_dirs = [eval_params.save_foldername / Path(f"image_{i}") for i in range(eval_images.shape[0])]
    # ground-truth boundaries [n_images, K-1, W]  (evaluation.py:86-88)
    eval_segments = np.swapaxes(generate_boundary(np.squeeze(eval_labels, axis=3), axis=1), 0, 1)
    K = eval_params.num_classes
    try:
        get_model_class(eval_params.loaded_model.name)
    except ValueError as e:
        log.error(e)
        exit(1)
    unsupported = set(eval_params.metrics) - {EVALUATION_METRIC_DICE_CLASSES, EVALUATION_METRIC_DICE_MACRO,
                                              EVALUATION_METRIC_DICE_MICRO}
    if unsupported:
        raise NotImplementedError(f"metrics {sorted(unsupported)} need the surface-distance package (out of scope)")
    engine = eval_params.loaded_model.engine
    dice_macro = custom_metrics.dice_coef_macro(False, K)
    dice_micro = custom_metrics.dice_coef_micro(False, K)
    outputs: List[EvaluationOutput] = []
    for i0 in range(0, len(eval_images), batch_size):
        chunk = np.ascontiguousarray(eval_images[i0:i0 + batch_size])
        labels, maps = engine.predict_maps(chunk, bg_ilm=eval_params.bg_ilm, bg_csi=eval_params.bg_csi)
        segs = errs = None
        if eval_params.graph_search:
            maps_t = np.ascontiguousarray(np.transpose(maps, (0, 1, 3, 2)))
            n, km1, W, H = maps_t.shape
            segs = graph_search.segment_maps(maps_t.reshape(-1, W, H), None, None, return_prob_maps=False)[0].reshape(n, km1, W)
        for k in range(len(chunk)):
            i = i0 + k
            cat = np.transpose(utils.to_categorical(labels[k], K), (2, 0, 1))          # [K,H,W], binarised
            truth_cat = utils.to_categorical(np.squeeze(eval_labels[i], axis=2), K)    # [H,W,K]
            m = {}
            if EVALUATION_METRIC_DICE_CLASSES in eval_params.metrics:
                m[EVALUATION_METRIC_DICE_CLASSES] = custom_metrics.soft_dice_class(
                    np.transpose(truth_cat, (2, 0, 1))[None], cat[None])[0]
            if EVALUATION_METRIC_DICE_MACRO in eval_params.metrics:
                m[EVALUATION_METRIC_DICE_MACRO] = dice_macro(truth_cat[None], np.transpose(cat, (1, 2, 0))[None])
            if EVALUATION_METRIC_DICE_MICRO in eval_params.metrics:
                m[EVALUATION_METRIC_DICE_MICRO] = float(dice_micro(truth_cat[None], np.transpose(cat, (1, 2, 0))[None]))
            gs = err = stats = None
            if segs is not None:
                gs = segs[k]
                err = np.stack([graph_search.calc_errors(gs[b], eval_segments[i][b].astype("float64"))
                                for b in range(gs.shape[0])])
                stats = graph_search.calculate_overall_errors(err)
            _save(eval_params, out_dirs[i], labels[k], cat, maps[k], gs)
            outputs.append(EvaluationOutput(
                image=eval_images[i], image_name=eval_image_names[i], image_segments=eval_segments[i],
                image_output_dir=out_dirs[i], predicted_labels=labels[k], categorical_pred=cat, boundary_maps=maps[k],
                gs_pred_segs=gs, errors=err, mean_abs_err=None if stats is None else stats[0],
                mean_err=None if stats is None else stats[1], abs_err_sd=None if stats is None else stats[2],
                err_sd=None if stats is None else stats[3], metrics=m))
    return outputs


def _save(ep, out_dir, labels, cat, maps, gs):
    sp = ep.save_params
    if not (sp.predicted_labels or sp.categorical_pred or sp.boundary_maps):
        return
    Path(out_dir).mkdir(parents=True, exist_ok=True)
    arrays = {}
    if sp.predicted_labels:
        arrays["predicted_labels"] = labels
    if sp.categorical_pred:
        arrays["categorical_pred"] = cat
    if sp.boundary_maps:
        arrays["boundary_maps"] = maps
    if gs is not None:
        arrays["gs_pred_segs"] = gs
    np.savez(Path(out_dir) / "evaluations.npz", **arrays)
