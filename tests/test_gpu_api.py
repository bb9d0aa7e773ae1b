"""Drop-in surface on the GPU: load_model_and_config / PredictionParams / predict / train_model
behave like the reference's API (same call shapes, same outputs as the oracle chain)."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import postproc
from oracle.unet_oracle import OracleUNet

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "trained_small_unet.npz"
CFG = dict(input_channels=1, num_classes=4, image_height=64, image_width=64, start_neurons=8, pool_layers=2,
           conv_layers=2, enc_kernel=[3, 3], dec_kernel=[2, 2])


def _saved_model(tmp_path, precision):
    from oct_image_segmentation_models_b200.models import get_model_class
    g = np.load(GOLD)
    weights = [g[f"w{i:03d}"] for i in range(len([k for k in g.files if k.startswith("w")]))]
    model = get_model_class("unet")(**CFG).build_model(precision=precision)
    model.set_weights(weights)
    model.save(tmp_path / "model_epoch01.hdf5")
    model.close()
    (tmp_path / "model_config.json").write_text(json.dumps(CFG))
    return g, weights


def test_predict_api_matches_oracle_chain(tmp_path):
    from oct_image_segmentation_models_b200.common.dataset import Dataset
    from oct_image_segmentation_models_b200.prediction import prediction, prediction_parameters as pp
    import os
    os.environ["OCTSEG_PRECISION"] = "fp32"
    try:
        g, weights = _saved_model(tmp_path, "fp32")
        imgs = g["images"]
        ds = Dataset(imgs, None, [Path(f"img{i}") for i in range(len(imgs))], [tmp_path / f"o{i}" for i in range(len(imgs))])
        params = pp.PredictionParams(tmp_path / "model_epoch01.hdf5", None, None, ds, tmp_path,
                                     pp.PredictionSaveParams(png_images=False))
        assert params.num_classes == 4 and params.loaded_model.name == "unet"
        outs = prediction.predict(params)
    finally:
        os.environ.pop("OCTSEG_PRECISION", None)
    ref = OracleUNet(weights, **{k: (tuple(v) if isinstance(v, list) else v) for k, v in CFG.items()
                                 if k not in ("image_height", "image_width")}).predict(imgs)
    for i, o in enumerate(outs):
        am, cat = postproc.perform_argmax(ref[i:i + 1])
        maps = postproc.convert_predictions_to_maps_semantic(cat)
        assert np.array_equal(o.predicted_labels, am[0])
        assert np.array_equal(o.categorical_pred, cat[0])
        assert np.array_equal(o.boundary_maps, maps[0])
        assert (tmp_path / f"o{i}" / "prediction_info.npz").exists()
    # Keras-style call: model.predict(preprocess(x)) == fused uint8 path, bit for bit
    model = params.loaded_model
    a = model.predict(imgs[:2] / 255.0, verbose=2, batch_size=1)
    b = model.predict(imgs[:2])
    assert np.array_equal(a, b)
    model.close()


def test_train_model_api_learns(tmp_path):
    from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch
    from oct_image_segmentation_models_b200.training.training import train_model
    from oct_image_segmentation_models_b200.training.training_parameters import TrainingParams
    tr_i, tr_l = synthetic_batch(0, 32, 64, 64)
    va_i, va_l = synthetic_batch(500, 4, 64, 64)
    np.savez(tmp_path / "ds.npz", train_images=tr_i, train_labels=tr_l, val_images=va_i, val_labels=va_l)
    tp = TrainingParams("unet", tmp_path / "ds.npz", None, tmp_path / "run", "adam", "weighted_categorical_crossentropy",
                        "dice_coef_macro", epochs=6, batch_size=8,
                        model_hyperparameters=dict(start_neurons=8, pool_layers=2, conv_layers=2),
                        opt_params=dict(learning_rate=3e-3), class_weight=[0.5, 1.0, 2.0, 1.0],
                        model_save_monitor=("val_loss", "min"))
    import os
    os.environ["OCTSEG_INIT_SEED"] = "7"
    try:
        model, hist = train_model(tp)
    finally:
        os.environ.pop("OCTSEG_INIT_SEED", None)
    assert hist[-1]["loss"] < hist[0]["loss"] * 0.7
    # (validation runs with the BN *moving* statistics, which at momentum 0.99 lag far behind after 24 steps --
    #  Keras behaves the same -- so only finiteness is asserted here; the loss curve is checked against the
    #  oracle in test_gpu_train.py)
    assert np.isfinite(hist[-1]["val_loss"]) and 0.0 <= hist[-1]["val_acc"] <= 1.0
    cfg = json.loads((tmp_path / "run" / "model_config.json").read_text())
    assert cfg["num_classes"] == 4 and cfg["image_height"] == 64 and cfg["pool_layers"] == 2
    assert list((tmp_path / "run").glob("model_epoch*.hdf5"))
    # the checkpoint round-trips through load_model_and_config
    from oct_image_segmentation_models_b200.common.utils import load_model_and_config
    ck = sorted((tmp_path / "run").glob("model_epoch*.hdf5"))[-1]
    m2, cfg2 = load_model_and_config(ck)
    assert cfg2 == cfg and m2.output.shape[-1] == 4
    m2.close()
    model.close()


def test_evaluate_model_api_with_hdf5_dataset_and_graph_search(tmp_path):
    """EvaluationParameters / evaluate_model on an HDF5 test set (written by the built-in HDF5 writer, read
    back through dataset_loader) with graph search: boundaries equal the oracle chain, errors are measured
    against generate_boundary(ground truth), Dice metrics are reported."""
    from oct_image_segmentation_models_b200.common import hdf5_min
    from oct_image_segmentation_models_b200.evaluation import evaluation, evaluation_parameters as ep
    import os
    os.environ["OCTSEG_PRECISION"] = "fp32"
    try:
        g, weights = _saved_model(tmp_path, "fp32")          # model_epoch01.hdf5 is a real HDF5 (Keras layout)
        with open(tmp_path / "model_epoch01.hdf5", "rb") as fh:
            assert fh.read(4) == b"\x89HDF"
        with hdf5_min.H5Writer(tmp_path / "test.hdf5") as f:
            f.create_dataset("test_images", g["images"])
            f.create_dataset("test_labels", g["labels"])
            f.create_dataset("test_images_source", np.array([f"img_{i}.png".encode() for i in range(len(g["images"]))]))
        params = ep.EvaluationParameters(tmp_path / "model_epoch01.hdf5", None, None, tmp_path / "test.hdf5",
                                         tmp_path / "eval", ep.EvaluationSaveParams(png_images=False), True,
                                         ["dice_coef_classes", "dice_coef_macro", "dice_coef_micro"])
        outs = evaluation.evaluate_model(params)
    finally:
        os.environ.pop("OCTSEG_PRECISION", None)
    assert len(outs) == len(g["images"]) and str(outs[2].image_name) == "img_2.png"
    for i, o in enumerate(outs):
        assert np.array_equal(o.gs_pred_segs, g["segs"][i])             # oracle min-path boundaries
        assert o.errors.shape == o.gs_pred_segs.shape and np.nanmax(np.abs(o.errors)) <= 6
        assert o.metrics["dice_coef_macro"] > 0.9 and o.metrics["dice_coef_classes"].shape == (4,)
        assert (tmp_path / "eval" / f"image_{i}" / "evaluations.npz").exists()
    params.loaded_model.close()
    with pytest.raises(SystemExit):
        ep.EvaluationParameters(tmp_path / "model_epoch01.hdf5", None, None, tmp_path / "test.hdf5", tmp_path / "e2",
                                ep.EvaluationSaveParams(), False, ["not_a_metric"])
