"""Loss registry with the reference's shape (common/custom_losses.py:230-255): `custom_loss_objects[name]` is
{"function": factory, "takes_sparse": bool} and `factory(num_classes=, is_y_true_sparse=, **loss_fn_kwargs)`
returns the loss object handed to `model.compile(loss=...)` (training/training.py:195-217).

The accelerated path implements ONE loss family in CUDA -- the reference's `weighted_categorical_crossentropy`
(common/custom_losses.py:11-37: renormalise, clip to [K.epsilon(), 1 - K.epsilon()], -sum y * log p * w; Keras
SUM_OVER_BATCH_SIZE reduction) -- fused with the 1x1 head and its gradient (csrc/train_kernels.cu head_loss_kernel).
It is registered here under its reference function name, and `categorical_crossentropy` is the same kernel with unit
weights.  The reference's other registered names (focal / dice / bce variants) are kept in the table so that a
reference config fails with a precise message instead of a KeyError."""
from typing import Optional, Sequence

import numpy as np

K_EPSILON = 1e-7


class WeightedCategoricalCrossentropy:
    """Loss object for `compile(loss=...)`: carries the class weights the CUDA step consumes; calling it evaluates the
    same formula on host arrays (used by tests and for small validation sets)."""
    name = "weighted_categorical_crossentropy"

    def __init__(self, weights: Optional[Sequence[float]], num_classes: int, is_y_true_sparse: bool = True):
        self.num_classes = int(num_classes)
        self.is_y_true_sparse = bool(is_y_true_sparse)
        w = np.ones(self.num_classes, np.float32) if weights is None else np.asarray(weights, np.float32)
        if w.shape != (self.num_classes,):
            raise ValueError(f"class weights must have {self.num_classes} entries, got shape {w.shape}")
        self.weights = w

    def __call__(self, y_true, y_pred):
        p = np.asarray(y_pred, np.float32)
        p = p / p.sum(-1, keepdims=True)
        p = np.clip(p, K_EPSILON, 1 - K_EPSILON)
        y = np.asarray(y_true)
        if self.is_y_true_sparse or y.shape[-1] != self.num_classes or y.ndim == p.ndim - 1:
            lab = y.reshape(p.shape[:-1]).astype(np.int64)
            pt = np.take_along_axis(p, lab[..., None], -1)[..., 0]
            return -np.log(pt) * self.weights[lab]
        return -np.sum(y * np.log(p) * self.weights, -1)

    def get_config(self):
        return {"name": self.name, "weights": self.weights.tolist()}


def weighted_categorical_crossentropy(num_classes: int, is_y_true_sparse: bool = True, weights=None, **_):
    return WeightedCategoricalCrossentropy(weights, num_classes, is_y_true_sparse)


def categorical_crossentropy(num_classes: int, is_y_true_sparse: bool = True, **_):
    loss = WeightedCategoricalCrossentropy(None, num_classes, is_y_true_sparse)
    loss.name = "categorical_crossentropy"
    return loss


def _not_accelerated(name):
    def factory(**_):
        raise NotImplementedError(
            f"loss '{name}' is registered by the reference (common/custom_losses.py:230-255) but has no CUDA kernel "
            "here; the accelerated train step implements 'weighted_categorical_crossentropy' / 'categorical_crossentropy'")
    factory.__name__ = name
    return factory


custom_loss_objects = {
    "weighted_categorical_crossentropy": {"function": weighted_categorical_crossentropy, "takes_sparse": True},
    "categorical_crossentropy": {"function": categorical_crossentropy, "takes_sparse": True},
    # reference names without a kernel (fail loudly, with the reason)
    "bce_dice_loss": {"function": _not_accelerated("bce_dice_loss"), "takes_sparse": False},
    "dice_loss_micro": {"function": _not_accelerated("dice_loss_micro"), "takes_sparse": False},
    "dice_loss_macro": {"function": _not_accelerated("dice_loss_macro"), "takes_sparse": False},
    "focal_loss": {"function": _not_accelerated("focal_loss"), "takes_sparse": True},
    "bce_focal_loss": {"function": _not_accelerated("bce_focal_loss"), "takes_sparse": False},
    "focal_dice_loss": {"function": _not_accelerated("focal_dice_loss"), "takes_sparse": True},
}
ACCELERATED_LOSSES = ("weighted_categorical_crossentropy", "categorical_crossentropy")
