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
    """train_model(TrainingParams) end to end as the reference runs it (training/training.py:135-408): opt_con(**opt_params),
    loss from the registry, compile + fit with ModelCheckpoint / SaveEpochInfo / EarlyStopping, DataGenerator feed."""
    from oct_image_segmentation_models_b200.common import hdf5_min
    from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch
    from oct_image_segmentation_models_b200.training.optimizers import Adam
    from oct_image_segmentation_models_b200.training.training import train_model
    from oct_image_segmentation_models_b200.training.training_parameters import TrainingParams
    tr_i, tr_l = synthetic_batch(0, 32, 64, 64)
    va_i, va_l = synthetic_batch(500, 8, 64, 64)
    np.savez(tmp_path / "ds.npz", train_images=tr_i, train_labels=tr_l, val_images=va_i, val_labels=va_l)
    tp = TrainingParams("unet", tmp_path / "ds.npz", None, tmp_path / "run", Adam, "weighted_categorical_crossentropy",
                        "dice_coef_macro", epochs=6, batch_size=8,
                        model_hyperparameters=dict(start_neurons=8, pool_layers=2, conv_layers=2),
                        opt_params=dict(learning_rate=3e-3), class_weight=[0.5, 1.0, 2.0, 1.0],
                        model_save_monitor=("val_dice_coef_macro", "max"), patience=50)
    import os
    os.environ["OCTSEG_INIT_SEED"] = "7"
    try:
        model, hist = train_model(tp)
    finally:
        os.environ.pop("OCTSEG_INIT_SEED", None)
    h = hist.history
    assert set(h) >= {"loss", "val_loss", "dice_coef_macro", "val_dice_coef_macro"} and len(h["loss"]) == 6
    assert h["loss"][-1] < h["loss"][0] * 0.7
    # (validation runs with the BN *moving* statistics, which at momentum 0.99 lag far behind after 24 steps --
    #  Keras behaves the same -- so only finiteness / range is asserted; the loss curve is checked in test_gpu_train.py)
    assert np.isfinite(h["val_loss"][-1]) and 0.0 <= h["val_dice_coef_macro"][-1] <= 1.0
    run = model.results_folder
    assert run.parent == tmp_path / "run" and run.name.endswith("_unet")       # <results>/<timestamp>_<architecture>
    cfg = json.loads((run / "model_config.json").read_text())
    assert cfg["num_classes"] == 4 and cfg["image_height"] == 64 and cfg["pool_layers"] == 2
    cks = sorted(run.glob("model_epoch*.hdf5"))
    assert cks and (run / "stats_epoch06.hdf5").exists() and not (run / "stats_epoch05.hdf5").exists()
    np.testing.assert_allclose(hdf5_min.H5File(run / "stats_epoch06.hdf5")["train_loss"].read(), h["loss"])
    # the checkpoint carries the optimizer state (as ModelCheckpoint's model.save does) and round-trips
    ow, tc = hdf5_min.load_keras_optimizer_weights(cks[-1])
    assert ow[0][0] == "Adam/iter:0" and int(ow[0][1]) > 0 and json.loads(tc)["optimizer_config"]["config"]["learning_rate"] == 3e-3
    from oct_image_segmentation_models_b200.common.utils import load_model_and_config
    m2, cfg2 = load_model_and_config(cks[-1])
    assert cfg2 == cfg and m2.output.shape[-1] == 4
    m2.close()
    model.close()
    # a reference loss name without a kernel ends like the reference's unknown-loss path: log.error + exit(1)
    tp_bad = TrainingParams("unet", tmp_path / "ds.npz", None, tmp_path / "run2", Adam, "focal_loss", "dice_coef_macro",
                            epochs=1, batch_size=8, model_hyperparameters=dict(start_neurons=8, pool_layers=2))
    with pytest.raises(SystemExit):
        train_model(tp_bad)


def test_fit_callback_protocol_and_resume_from_checkpoint(tmp_path, monkeypatch):
    """Keras protocol of B200Model.fit (reference training.py:401-407, training_callbacks.py:35-64) and
    `initial_model` resume: 2 + 2 epochs through a checkpoint (weights + Adam m, v, iterations) == 4 epochs.
    (Trained on the fp32 kernels so that the two trajectories can be compared tightly; the default trains in bf16.)"""
    monkeypatch.setenv("OCTSEG_TRAIN_PRECISION", "fp32")
    from oct_image_segmentation_models_b200.common.custom_losses import custom_loss_objects
    from oct_image_segmentation_models_b200.common.data_generator import DataGenerator
    from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch
    from oct_image_segmentation_models_b200.models import get_model_class
    from oct_image_segmentation_models_b200.models.keras_like import load_model
    from oct_image_segmentation_models_b200.training import training_callbacks as tcb
    from oct_image_segmentation_models_b200.training.optimizers import Adam
    tr_i, tr_l = synthetic_batch(0, 16, 64, 64)
    va_i, va_l = synthetic_batch(500, 8, 64, 64)
    container = get_model_class("unet")(**CFG)

    def build():
        m = B200(container)
        loss = custom_loss_objects["weighted_categorical_crossentropy"]["function"](num_classes=4, is_y_true_sparse=True,
                                                                                    weights=[0.5, 1.0, 2.0, 1.0])
        m.compile(optimizer=Adam(learning_rate=2e-3), loss=loss, metrics=["dice_coef_micro"])
        return m

    def B200(c):
        from oct_image_segmentation_models_b200.models.keras_like import B200Model
        return B200Model("unet", c.spec_kwargs(), precision="fp32", init_seed=3)

    def gens():
        pre = container.get_preprocess_input_fn()
        return (DataGenerator(tr_i, tr_l, 8, [], "none", (), False, pre, shuffle=False),
                DataGenerator(va_i, va_l, 8, [], "none", (), False, pre, shuffle=False))

    events = []

    class Rec(tcb.Callback):
        def on_train_begin(self, logs=None): events.append("train_begin")
        def on_epoch_begin(self, epoch, logs=None): events.append(f"begin{epoch}")
        def on_epoch_end(self, epoch, logs=None): events.append((epoch, sorted(logs)))
        def on_train_end(self, logs=None): events.append("train_end")

    full = build()
    tg, vg = gens()
    hist = full.fit(x=tg, validation_data=vg, epochs=4, callbacks=[Rec()], verbose=0)
    assert events[0] == "train_begin" and events[1] == "begin0" and events[-1] == "train_end"
    assert events[2] == (0, ["dice_coef_micro", "epoch_time", "loss", "val_dice_coef_micro", "val_loss"])
    assert hist.history["loss"][-1] < hist.history["loss"][0]
    w_full = full.get_weights()
    full.close()

    # library-generated dropout masks depend on the optimizer iteration, which the checkpoint restores as well
    part = build()
    tg, vg = gens()
    part.fit(x=tg, validation_data=vg, epochs=2, verbose=0)
    part.save(tmp_path / "ck.hdf5")
    part.close()
    resumed = load_model(tmp_path / "ck.hdf5", precision="fp32")
    assert resumed._compiled is not None and resumed._compiled["resume"][0] == 4       # 2 epochs x 2 batches
    tg, vg = gens()
    resumed.fit(x=tg, validation_data=vg, epochs=4, initial_epoch=2, verbose=0)
    w_res = resumed.get_weights()
    resumed.close()
    worst = max(float(np.abs(a - b).max()) for a, b in zip(w_full, w_res))
    # fp32 atomics reorder sums between runs; pre-BN biases move by +-lr per step on rounding-noise gradients
    assert worst <= 5e-3, worst
    close = np.mean([np.allclose(a, b, atol=2e-4) for a, b in zip(w_full, w_res)])
    assert close >= 0.6, close


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
