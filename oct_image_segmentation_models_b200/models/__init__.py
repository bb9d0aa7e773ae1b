"""Model registry (reference models/__init__.py:9-22)."""
from typing import Type

from . import base_model
from . import unet

DEEPLABV3PLUS_MODEL_NAME = "deeplabv3plus"


class _DeeplabV3PlusUnavailable(base_model.BaseModel):
    """Registry placeholder: DeepLabV3+ needs ImageNet ResNet50 weights (a download) and is not
    on the hot path this package accelerates (component 3)."""

    def __init__(self, **kwargs):
        raise NotImplementedError("deeplabv3plus is outside the B200 hot path (U-Net only)")

    def build_model(self):  # pragma: no cover
        raise NotImplementedError

    def get_preprocess_input_fn(self):  # pragma: no cover
        raise NotImplementedError


model_name_map = {
    DEEPLABV3PLUS_MODEL_NAME: _DeeplabV3PlusUnavailable,
    unet.UNET_MODEL_NAME: unet.UNet,
}


def get_model_class(model_name: str) -> Type[base_model.BaseModel]:
    if not isinstance(model_name, str):
        raise TypeError("model_name must be str")
    model_class = model_name_map.get(model_name)

    if model_class is None:
        raise ValueError(f"Model name: '{model_name}' could not be found.")

    return model_class
