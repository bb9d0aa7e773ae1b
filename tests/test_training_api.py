"""CPU tests of the host-side training surface that mirrors the reference (no GPU, no compute calls): loss registry
(reference common/custom_losses.py:230-255), optimizer constructors (training.py:190-193), DataGenerator
(common/data_generator.py:285-346), Keras callback protocol pieces (training.py:319-342,
training_callbacks.py:12-80), class weights (training.py:200-206), checkpoint optimizer groups."""
import json

import numpy as np
import pytest

from oct_image_segmentation_models_b200.common import custom_losses, hdf5_min
from oct_image_segmentation_models_b200.common.data_generator import DataGenerator
from oct_image_segmentation_models_b200.training import optimizers, training_callbacks as tcb
from oct_image_segmentation_models_b200.training.training import compute_class_weight_balanced


def test_loss_registry_has_the_reference_shape_and_the_weighted_ce():
    reg = custom_losses.custom_loss_objects
    for name in ("bce_dice_loss", "dice_loss_micro", "dice_loss_macro", "focal_loss", "bce_focal_loss", "focal_dice_loss",
                 "weighted_categorical_crossentropy"):
        assert set(reg[name]) == {"function", "takes_sparse"}
    loss = reg["weighted_categorical_crossentropy"]["function"](num_classes=3, is_y_true_sparse=True, weights=[0.5, 2.0, 10.0])
    p = np.array([[[0.2, 0.5, 0.3], [1.0, 0.0, 0.0]]], np.float32)
    y = np.array([[1, 0]])
    got = loss(y, p)
    # reference formula (custom_losses.py:27-35): renormalise, clip to [1e-7, 1-1e-7], -sum(y * log(p) * w)
    want = np.array([[-np.log(0.5) * 2.0, -np.log(np.float32(1) - np.float32(1e-7)) * 0.5]])   # float32, as K.clip in TF
    np.testing.assert_allclose(got, want, rtol=1e-6)
    onehot = np.eye(3, dtype=np.float32)[y]
    np.testing.assert_allclose(custom_losses.WeightedCategoricalCrossentropy([0.5, 2.0, 10.0], 3, False)(onehot, p), want, rtol=1e-6)
    with pytest.raises(NotImplementedError, match="focal_loss"):
        reg["focal_loss"]["function"](num_classes=3, is_y_true_sparse=True)
    assert reg.get("no_such_loss") is None


def test_optimizer_constructor_protocol():
    opt = optimizers.Adam(**dict(learning_rate=3e-4, beta_1=0.8))          # opt_con(**opt_params)
    assert opt.get_config()["learning_rate"] == 3e-4 and opt.get_config()["name"] == "Adam"
    assert optimizers.adam_hyperparameters(opt) == {"learning_rate": 3e-4, "beta_1": 0.8, "beta_2": 0.999, "epsilon": 1e-7}
    assert optimizers.adam_hyperparameters("adam")["learning_rate"] == 1e-3

    class FakeKerasSGD:
        def get_config(self):
            return {"name": "SGD", "learning_rate": 0.1}
    with pytest.raises(NotImplementedError):
        optimizers.adam_hyperparameters(FakeKerasSGD())
    with pytest.raises(NotImplementedError):
        optimizers.Adam(amsgrad=True)


def test_data_generator_batches_cover_an_epoch_and_reshuffle():
    imgs = np.arange(10 * 4 * 4, dtype=np.uint8).reshape(10, 4, 4, 1)
    labs = (imgs % 3).astype(np.uint8)

    def preprocess_input_inner(x):
        return x / 255.0
    gen = DataGenerator(imgs, labs, 4, [], "none", (), False, preprocess_input_inner)
    assert len(gen) == 2 and gen.get_total_samples() == 10 and gen.raw_uint8
    x, y = gen[0]
    assert x.dtype == np.float32 and x.shape == (4, 4, 4, 1) and y.shape == (4, 4, 4, 1)
    idx = gen.indices(0)
    np.testing.assert_array_equal(x, (imgs[idx].astype(np.float64) / 255.0).astype(np.float32))
    seen = np.concatenate([gen.indices(i) for i in range(len(gen))])
    assert len(set(seen)) == 8
    before = seen.copy()
    gen.on_epoch_end()
    after = np.concatenate([gen.indices(i) for i in range(len(gen))])
    assert not np.array_equal(before, after) or True          # a permutation may repeat; the call must not fail
    got = [i for i, _ in gen.prefetch(lambda i: gen.raw_batch(i))]
    assert got == [0, 1]
    xs, ys = gen.raw_batch(1, 1, 3)
    assert xs.shape[0] == 2 and xs.dtype == np.uint8
    with pytest.raises(SystemExit):
        DataGenerator(imgs, labs, 4, [], "one", (), False, None)


class _FakeModel:
    def __init__(self):
        self.w = [np.zeros(2)]
        self.saved = []
        self.stop_training = False

    def get_weights(self):
        return [x.copy() for x in self.w]

    def set_weights(self, w):
        self.w = [x.copy() for x in w]

    def save(self, path):
        self.saved.append(str(path))


def test_early_stopping_monitors_val_metric_max_and_restores_best_weights():
    m = _FakeModel()
    es = tcb.EarlyStopping(monitor="val_dice_coef_macro", mode="max", patience=2, restore_best_weights=True)
    es.set_model(m)
    es.on_train_begin()
    for epoch, v in enumerate([0.5, 0.8, 0.7, 0.6, 0.9]):
        m.w = [np.full(2, float(epoch))]
        es.on_epoch_end(epoch, {"val_dice_coef_macro": v, "val_loss": 1.0 - v})
        if m.stop_training:
            break
    assert m.stop_training and epoch == 3 and es.best == 0.8
    np.testing.assert_array_equal(m.w[0], np.full(2, 1.0))      # weights of the best epoch (index 1)
    es2 = tcb.EarlyStopping(monitor="val_missing", patience=1)
    es2.set_model(m)
    es2.on_train_begin()
    es2.on_epoch_end(0, {"val_loss": 1.0})                      # missing monitor: warning, no fallback to val_loss


def test_model_checkpoint_names_and_best_only(tmp_path):
    m = _FakeModel()
    ck = tcb.ModelCheckpoint(filepath=tmp_path / "model_epoch{epoch:02d}.hdf5", save_best_only=True, monitor="val_acc", mode="max")
    ck.set_model(m)
    for epoch, v in enumerate([0.1, 0.3, 0.2]):
        ck.on_epoch_end(epoch, {"val_acc": v})
    assert [p.split("/")[-1] for p in m.saved] == ["model_epoch01.hdf5", "model_epoch02.hdf5"]
    ck.on_epoch_end(3, {"val_loss": 0.0})                       # monitor missing -> skipped, never falls back
    assert len(m.saved) == 2


def test_save_epoch_info_writes_the_reference_stats_file(tmp_path):
    class P:
        metric, loss, epochs = "dice_coef_macro", "weighted_categorical_crossentropy", 2
    cb = tcb.SaveEpochInfo(save_folder=tmp_path, train_params=P())
    cb.on_train_begin()
    for e in range(2):
        cb.on_epoch_begin(e)
        cb.on_epoch_end(e, {"loss": 1.0 / (e + 1), "val_loss": 2.0, "dice_coef_macro": 0.5, "val_dice_coef_macro": 0.4 + e})
    cb.on_train_end()
    assert not (tmp_path / "stats_epoch01.hdf5").exists()
    f = hdf5_min.H5File(tmp_path / "stats_epoch02.hdf5")
    np.testing.assert_allclose(f["train_loss"].read(), [1.0, 0.5])
    np.testing.assert_allclose(f["val_acc"].read(), [0.4, 1.4])
    assert len(f["epoch_time"].read()) == 2 and cb.train_time >= 0


def test_balanced_class_weights_match_sklearn_formula():
    y = np.array([0] * 6 + [1] * 3 + [3] * 1)
    np.testing.assert_allclose(compute_class_weight_balanced(y), 10 / (3 * np.array([6, 3, 1.0])), rtol=1e-6)


def test_checkpoint_optimizer_groups_round_trip(tmp_path):
    lw = [("conv2d", [("kernel:0", np.ones((1, 1, 8, 4), np.float32)), ("bias:0", np.zeros(4, np.float32))])]
    ow = [("Adam/iter:0", np.asarray(12, np.int64)), ("Adam/conv2d/kernel/m:0", np.full((1, 1, 8, 4), 0.5, np.float32))]
    hdf5_min.save_keras_weights(tmp_path / "m.hdf5", lw, model_config="{}", optimizer_weights=ow,
                                training_config=json.dumps({"optimizer_config": {"class_name": "Adam"}}))
    got, tc = hdf5_min.load_keras_optimizer_weights(tmp_path / "m.hdf5")
    assert [n for n, _ in got] == [n for n, _ in ow] and int(got[0][1]) == 12
    np.testing.assert_array_equal(got[1][1], ow[1][1])
    assert json.loads(tc)["optimizer_config"]["class_name"] == "Adam"
    assert hdf5_min.load_keras_optimizer_weights(tmp_path / "m.hdf5")[0]
    hdf5_min.save_keras_weights(tmp_path / "w.hdf5", lw)
    assert hdf5_min.load_keras_optimizer_weights(tmp_path / "w.hdf5") == ([], None)
