"""`predict(PredictionParams) -> List[PredictionOutput]` (reference prediction/prediction.py:48-186).

Same per-image contract and outputs as the reference; the network forward runs batched on the
GPU (the reference calls Keras once per image, :70-81), post-processing is the same numpy code
path, file outputs are .npz (no h5py / matplotlib offline) and only written when a save flag asks.
"""
import logging as log
import time
from pathlib import Path
from typing import List, Union

import numpy as np

from ..common import utils
from ..models import get_model_class
from .prediction_parameters import PredictionParams


class PredictionOutput:
    def __init__(self, image: np.ndarray, image_name: Path, image_output_dir: Path,
                 predicted_labels: np.ndarray, categorical_pred: np.ndarray, boundary_maps: np.ndarray,
                 gs_pred_segs: Union[np.ndarray, None]) -> None:
        self.image = image
        self.image_name = image_name
        self.image_output_dir = image_output_dir
        self.predicted_labels = predicted_labels
        self.categorical_pred = categorical_pred
        self.boundary_maps = boundary_maps
        self.gs_pred_segs = gs_pred_segs


def predict(predict_params: PredictionParams, batch_size: int = 64) -> List[PredictionOutput]:
    dataset = predict_params.dataset
    try:
        model_class = get_model_class(predict_params.loaded_model.name)
    except ValueError as e:
        log.error(e)
        exit(1)
    model_container = model_class(**predict_params.model_config)
    preprocess = model_container.get_preprocess_input_fn()
    model = predict_params.loaded_model
    images = dataset.images
    outputs: List[PredictionOutput] = []
    for i0 in range(0, len(images), batch_size):
        chunk = np.asarray(images[i0:i0 + batch_size])
        t0 = time.time()
        engine = getattr(model, "engine", None)
        dev_labels = dev_maps = probs = None
        if chunk.dtype == np.uint8 and engine is not None:
            # What the reference keeps of a forward pass is the argmax and the boundary maps built from it
            # (reference prediction.py:97-104); both are produced on the GPU and 1 + (K-1) bytes per pixel come
            # back instead of 4*K bytes of probabilities (bit-identical to the numpy chain on the same labels:
            # tests/test_gpu_parity.py::test_device_boundary_maps_match_reference_semantics).
            dev_labels, dev_maps = engine.predict_maps(np.ascontiguousarray(chunk), bg_ilm=True, bg_csi=False)
        elif chunk.dtype == np.uint8:
            probs = model.predict(chunk)                       # fused on-device x/255
        else:
            probs = model.predict(preprocess(chunk), verbose=2, batch_size=1)
        predict_time = (time.time() - t0) / len(chunk)
        for k in range(len(chunk)):
            i = i0 + k
            if dev_labels is not None:
                num_maps = predict_params.num_classes
                predicted_labels = dev_labels[k:k + 1].astype(np.int64)          # np.argmax's dtype
                categorical_pred = np.transpose(utils.to_categorical(predicted_labels, num_maps), axes=(0, 3, 1, 2))
                boundary_maps = dev_maps[k:k + 1]
            else:
                predicted_labels, categorical_pred = utils.perform_argmax(probs[k:k + 1], bin=True)
                boundary_maps = utils.convert_predictions_to_maps_semantic(np.array(categorical_pred),
                                                                           bg_ilm=True, bg_csi=False)
            predicted_labels = np.squeeze(predicted_labels)
            categorical_pred = np.squeeze(categorical_pred)
            boundary_maps = np.squeeze(boundary_maps)
            gs_pred_segs = None
            if predict_params.graph_search:
                from ..min_path_processing import graph_search
                boundary_maps_t = np.transpose(boundary_maps, axes=[0, 2, 1])
                gs_pred_segs, _, _ = graph_search.segment_maps(boundary_maps_t, None, None, return_prob_maps=False)
            _save_image_prediction_results(predict_params, dataset.image_output_dirs[i], predicted_labels,
                                           categorical_pred, boundary_maps, gs_pred_segs, predict_time)
            outputs.append(PredictionOutput(image=images[i], image_name=dataset.image_names[i],
                                            image_output_dir=dataset.image_output_dirs[i],
                                            predicted_labels=predicted_labels, categorical_pred=categorical_pred,
                                            boundary_maps=boundary_maps, gs_pred_segs=gs_pred_segs))
            log.info(f"DONE processing image number {i}: {dataset.image_names[i]}")
    return outputs


def _save_image_prediction_results(pp, out_dir, predicted_labels, categorical_pred, boundary_maps,
                                   gs_pred_segs, predict_time):
    sp = pp.save_params
    if not (sp.predicted_labels or sp.categorical_pred or sp.boundary_maps):
        return
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    arrays = {"predict_time": np.float64(predict_time)}
    if sp.predicted_labels:
        arrays["predicted_labels"] = predicted_labels
    if sp.categorical_pred:
        arrays["categorical_pred"] = categorical_pred
    if sp.boundary_maps:
        arrays["boundary_maps"] = boundary_maps
    if gs_pred_segs is not None:
        arrays["gs_pred_segs"] = gs_pred_segs
    np.savez(out_dir / "prediction_info.npz", **arrays)
