"""Evaluation metrics with the reference's names and math (common/custom_metrics.py:19-100), in
numpy: thresholded Dice micro / macro and the per-class soft Dice.  (Average surface distance and
Hausdorff need the `surface-distance` package: outside the hot path.)"""
import numpy as np

from . import TRAINING_MONITOR_METRIC_DICE_MACRO, TRAINING_MONITOR_METRIC_DICE_MICRO


def _one_hot(y, num_classes):
    y = np.asarray(y).astype(np.int64)
    out = np.zeros(y.shape + (num_classes,), np.float32)
    np.put_along_axis(out, y[..., None], 1.0, axis=-1)
    return out


def dice_coef_micro(is_y_true_sparse: bool, num_classes: int):
    def _dice_coef_micro(y_true, y_pred):
        if is_y_true_sparse:
            y_true = _one_hot(np.squeeze(y_true), num_classes)
        y_true_f = np.asarray(y_true, np.float32).ravel()
        y_pred_f = (np.asarray(y_pred, np.float32).ravel() > 0.5).astype(np.float32)
        return 2.0 * np.sum(y_true_f * y_pred_f) / (np.sum(y_true_f) + np.sum(y_pred_f))

    _dice_coef_micro.__name__ = "dice_coef_micro"
    return _dice_coef_micro


def dice_coef_macro(is_y_true_sparse: bool, num_classes: int):
    def _dice_coef_macro(y_true, y_pred, eps=1e-05):
        if is_y_true_sparse:
            y_true = _one_hot(np.squeeze(y_true), num_classes)
        y_true = np.asarray(y_true, np.float32)
        y_pred = (np.asarray(y_pred) > 0.5).astype(np.float32)
        axes = tuple(range(1, y_pred.ndim - 1))
        inter = np.sum(y_true * y_pred, axis=axes)
        denom = np.sum(y_true, axis=axes) + np.sum(y_pred, axis=axes)
        return float(np.mean((2.0 * inter + eps) / (denom + eps)))

    _dice_coef_macro.__name__ = "dice_coef_macro"
    return _dice_coef_macro


training_monitor_metric_objects = {
    TRAINING_MONITOR_METRIC_DICE_MACRO: dice_coef_macro,
    TRAINING_MONITOR_METRIC_DICE_MICRO: dice_coef_micro,
}


def soft_dice_class(y_true, y_pred, eps=1e-5):
    axes = tuple(range(2, len(y_pred.shape)))
    intersect = np.sum(y_pred * y_true, axis=axes)
    denom = np.sum(y_pred + y_true, axis=axes)
    return ((2.0 * intersect) + eps) / (denom + eps)
