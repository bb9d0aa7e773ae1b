"""The product's numpy post-processing (common/utils.py mirror) against outputs of the
reference's own functions (tests/golden/postproc_golden.npz)."""
from pathlib import Path

import numpy as np


def test_product_postproc_matches_reference_golden():
    import importlib
    import sys
    import types
    # common/utils.py imports the GPU model class; stub the native loader so this runs on CPU
    utils = importlib.import_module("oct_image_segmentation_models_b200.common.utils")
    g = np.load(Path(__file__).parent / "golden" / "postproc_golden.npz")
    for nm in ("a", "b", "ties", "edges"):
        am, cat = utils.perform_argmax(g[nm + "_probs"], bin=True)
        assert np.array_equal(am, g[nm + "_argmax"]) and np.array_equal(cat, g[nm + "_cat"])
        maps = utils.convert_predictions_to_maps_semantic(np.array(cat), bg_ilm=True, bg_csi=False)
        assert maps.dtype == np.uint8 and np.array_equal(maps, g[nm + "_maps"])
    for nm in ("a", "b"):
        _, cat = utils.perform_argmax(g[nm + "_probs"], bin=True)
        assert np.array_equal(utils.convert_predictions_to_maps_semantic(np.array(cat), False, True), g[nm + "_maps_csi"])
    del sys, types


def test_model_registry_and_config_schema():
    from oct_image_segmentation_models_b200.models import get_model_class
    import pytest
    cls = get_model_class("unet")
    m = cls(input_channels=1, num_classes=4, image_height=64, image_width=32, start_neurons=8)
    assert m.get_config() == {"input_channels": 1, "num_classes": 4, "image_height": 64, "image_width": 32,
                              "start_neurons": 8, "pool_layers": 4, "conv_layers": 2, "enc_kernel": (3, 3),
                              "dec_kernel": (2, 2)}
    x = np.array([0, 51, 255], dtype=np.uint8)
    assert np.array_equal(m.get_preprocess_input_fn()(x), x / 255.0)
    with pytest.raises(ValueError):
        get_model_class("resnet")
    with pytest.raises(TypeError):
        cls(input_channels="1", num_classes=4, image_height=64, image_width=32)
    with pytest.raises(NotImplementedError):
        get_model_class("deeplabv3plus")(input_channels=1, num_classes=4, image_height=64, image_width=32)
