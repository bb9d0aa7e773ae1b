"""Host utilities with the reference's names and semantics (common/utils.py).

`load_model_and_config` is the drop-in boundary (reference :27-70): it returns
(model, model_config) where `model` quacks like the Keras object the callers use.
`perform_argmax` / `convert_predictions_to_maps_semantic` are the CPU numpy post-processing
between the network and the boundary search (reference :80-168) -- vectorised here, checked
against reference outputs in tests/test_postproc_product.py.
"""
import datetime
import hashlib
import json
import logging as log
from pathlib import Path, PurePosixPath
from typing import Tuple, Union

import numpy as np

from ..models import get_model_class
from ..models.keras_like import B200Model


def get_timestamp():
    return datetime.datetime.now().strftime("%Y-%m-%d_%H_%M_%S")


def load_model_and_config(model_path: Union[Path, PurePosixPath], **kwargs) -> Tuple[B200Model, dict]:
    if not isinstance(model_path, (Path, PurePosixPath)):
        raise TypeError("model_path must be a pathlib path")
    mlflow_tracking_uri = kwargs.pop("mlflow_tracking_uri", {})
    kwargs.pop("mlflow_run_uuid", {})
    if mlflow_tracking_uri:
        log.error("MLflow model loading is outside the B200 hot path; load from a local model file")
        exit(1)
    with open(Path(model_path).parent / Path("model_config.json"), "r") as config_file:
        model_config = json.load(config_file)
    from ..models.keras_like import read_weight_file
    name, weights = read_weight_file(Path(model_path))
    model_class = get_model_class(name)
    loaded_model = model_class(**model_config).build_model(precision=kwargs.pop("precision", None),
                                                           device=kwargs.pop("device", 0))
    loaded_model.set_weights(weights)
    return loaded_model, model_config


def to_categorical(y, num_classes):
    y = np.asarray(y, dtype=np.int64)
    out = np.zeros(y.shape + (num_classes,), dtype=np.float32)
    np.put_along_axis(out, y[..., None], 1.0, axis=-1)
    return out


def convert_maps_uint8(prob_maps):
    prob_maps *= 255
    return prob_maps.astype("uint8")


def perform_argmax(predictions, bin=True):
    num_maps = predictions.shape[3]
    argmax_pred = np.argmax(predictions, axis=3)
    if bin:
        categorical_pred = np.transpose(to_categorical(argmax_pred, num_maps), axes=(0, 3, 1, 2))
    else:
        categorical_pred = np.transpose(predictions, axes=(0, 3, 1, 2))
    return [argmax_pred, categorical_pred]


def convert_predictions_to_maps_semantic(categorical_pred, bg_ilm=True, bg_csi=False):
    num_samples, num_maps, img_height, img_width = categorical_pred.shape
    boundary_maps = np.zeros((num_samples, num_maps - 1, img_height, img_width), dtype="uint8")
    for map_ind in range(1, num_maps):
        if (map_ind == 1 and bg_ilm is True) or (map_ind == num_maps - 1 and bg_csi is True):
            grad_map = -np.gradient(categorical_pred[:, map_ind - 1], axis=1)
        else:
            grad_map = np.gradient(categorical_pred[:, map_ind], axis=1)
        grad_map = np.where(grad_map < 0, 0, grad_map) * 2
        grad_map = grad_map - np.roll(grad_map, -1, axis=1)
        grad_map = np.where(grad_map < 0, 0, grad_map)
        boundary_maps[:, map_ind - 1] = convert_maps_uint8(grad_map)
    return boundary_maps


def md5(file_path: Path) -> str:
    log.info(f"Calculating md5 of file: {file_path}")
    with open(file_path, "rb") as file_to_check:
        return hashlib.md5(file_to_check.read()).hexdigest()
