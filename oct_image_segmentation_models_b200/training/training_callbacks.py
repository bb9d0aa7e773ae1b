"""Keras callback protocol used by the reference's training loop (training/training.py:319-345,
training/training_callbacks.py:12-80): `Callback` base class, `ModelCheckpoint`, `EarlyStopping` and the reference's own
`SaveEpochInfo` (per-epoch `stats_epochNN.hdf5`, written with the built-in HDF5 writer; the PNG plot is out of scope)."""
import logging as log
import os
import time
from pathlib import Path

import numpy as np


class Callback:
    def __init__(self):
        self.model = None
        self.params = {}

    def set_model(self, model):
        self.model = model

    def set_params(self, params):
        self.params = params

    def on_train_begin(self, logs=None): ...
    def on_train_end(self, logs=None): ...
    def on_epoch_begin(self, epoch, logs=None): ...
    def on_epoch_end(self, epoch, logs=None): ...
    def on_train_batch_begin(self, batch, logs=None): ...
    def on_train_batch_end(self, batch, logs=None): ...


def _monitor_op(mode, monitor):
    if mode not in ("auto", "min", "max"):
        mode = "auto"
    if mode == "auto":
        mode = "max" if ("acc" in monitor or "dice" in monitor or monitor.startswith("fmeasure")) else "min"
    return (np.greater, -np.inf) if mode == "max" else (np.less, np.inf)


class ModelCheckpoint(Callback):
    """tf.keras.callbacks.ModelCheckpoint as the reference configures it (training.py:319-326): `filepath` with an
    `{epoch:02d}` field, `save_best_only`, `monitor`, `mode`.  A missing monitor key is a warning and no file, as in Keras."""

    def __init__(self, filepath, monitor="val_loss", save_best_only=False, mode="auto", save_weights_only=False, **_):
        super().__init__()
        self.filepath = str(filepath)
        self.monitor, self.save_best_only = monitor, save_best_only
        self.monitor_op, self.best = _monitor_op(mode, monitor)
        self.saved = []

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        path = self.filepath.format(epoch=epoch + 1, **logs)
        if self.save_best_only:
            cur = logs.get(self.monitor)
            if cur is None:
                log.warning(f"Can save best model only with {self.monitor} available, skipping.")
                return
            if not self.monitor_op(cur, self.best):
                return
            self.best = cur
        self.model.save(path)
        self.saved.append(path)


class EarlyStopping(Callback):
    """tf.keras.callbacks.EarlyStopping (training.py:335-342): monitor, mode, patience, restore_best_weights."""

    def __init__(self, monitor="val_loss", min_delta=0, patience=0, mode="auto", restore_best_weights=False, **_):
        super().__init__()
        self.monitor, self.patience, self.restore_best_weights = monitor, patience, restore_best_weights
        self.min_delta = abs(min_delta)
        self.monitor_op, self._init_best = _monitor_op(mode, monitor)
        if self.monitor_op is np.greater:
            self.min_delta *= 1
        else:
            self.min_delta *= -1
        self.stopped_epoch = 0
        self.best_weights = None

    def on_train_begin(self, logs=None):
        self.wait, self.stopped_epoch, self.best, self.best_weights, self.best_epoch = 0, 0, self._init_best, None, 0

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            log.warning(f"Early stopping conditioned on metric `{self.monitor}` which is not available. "
                        f"Available metrics are: {','.join(sorted(logs or {}))}")
            return
        if self.restore_best_weights and self.best_weights is None:
            self.best_weights = self.model.get_weights()
        self.wait += 1
        if self.monitor_op(cur - self.min_delta, self.best):
            self.best, self.best_epoch, self.wait = cur, epoch, 0
            if self.restore_best_weights:
                self.best_weights = self.model.get_weights()
        if self.wait >= self.patience and epoch > 0:
            self.stopped_epoch = epoch
            self.model.stop_training = True
            if self.restore_best_weights and self.best_weights is not None:
                self.model.set_weights(self.best_weights)


class SaveEpochInfo(Callback):
    """reference training/training_callbacks.py:12-80: keeps loss / metric curves and epoch times, rewrites
    `stats_epochNN.hdf5` every epoch (datasets train_acc, val_acc, train_loss, val_loss, epoch_time) and removes the
    previous epoch's file."""

    def __init__(self, save_folder: Path, train_params):
        super().__init__()
        self.acc_name, self.loss_name = train_params.metric, train_params.loss
        self.save_folder = Path(save_folder)
        self.num_epochs = train_params.epochs
        self.train_time = -1
        self.on_train_begin()

    def on_train_begin(self, logs=None):
        self.train_losses, self.train_accs, self.val_losses, self.val_accs, self.epoch_times = [], [], [], [], []
        self.start_time = time.time()

    def on_train_end(self, logs=None):
        self.train_time = time.time() - self.start_time

    def on_epoch_begin(self, epoch, logs=None):
        self.start_epoch_time = time.time()

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.train_losses.append(logs.get("loss"))
        self.train_accs.append(logs.get(self.acc_name))
        self.val_losses.append(logs.get("val_loss"))
        self.val_accs.append(logs.get("val_" + self.acc_name))
        self.epoch_times.append(time.time() - self.start_epoch_time)
        from ..common.hdf5_min import H5Writer

        def arr(v):
            return np.asarray([np.nan if x is None else x for x in v], np.float64)
        with H5Writer(self.save_folder / f"stats_epoch{epoch + 1:02d}.hdf5") as f:
            f.create_dataset("train_acc", data=arr(self.train_accs))
            f.create_dataset("val_acc", data=arr(self.val_accs))
            f.create_dataset("train_loss", data=arr(self.train_losses))
            f.create_dataset("val_loss", data=arr(self.val_losses))
            f.create_dataset("epoch_time", data=arr(self.epoch_times))
        prev = self.save_folder / f"stats_epoch{epoch:02d}.hdf5"
        if prev.is_file():
            try:
                os.remove(prev)
            except OSError:
                pass
