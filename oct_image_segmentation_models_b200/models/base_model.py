"""Abstract model container.

Mirrors the reference container class (reference:
oct_image_segmentation_models/models/base_model.py:8-36): four keyword-only
geometry attributes, `get_config()` returning exactly those four keys, and two
abstract hooks (`build_model`, `get_preprocess_input_fn`).
"""
import abc
from typing import Callable


class BaseModel(abc.ABC):
    def __init__(
        self,
        *,
        input_channels: int,
        num_classes: int,
        image_height: int,
        image_width: int,
    ):
        for name, val in (
            ("input_channels", input_channels),
            ("num_classes", num_classes),
            ("image_height", image_height),
            ("image_width", image_width),
        ):
            # the reference is @typechecked; keep the same failure mode (TypeError)
            if not isinstance(val, int) or isinstance(val, bool):
                raise TypeError(f"{name} must be int, got {type(val).__name__}")
        self.input_channels = input_channels
        self.num_classes = num_classes
        self.image_height = image_height
        self.image_width = image_width

    @abc.abstractmethod
    def build_model(self):
        raise NotImplementedError("Must be implemented in subclasses.")

    def get_config(self) -> dict:
        return {
            "input_channels": self.input_channels,
            "num_classes": self.num_classes,
            "image_height": self.image_height,
            "image_width": self.image_width,
        }

    @abc.abstractmethod
    def get_preprocess_input_fn(self) -> Callable:
        raise NotImplementedError("Must be implemented in subclasses.")
