"""`train_model(TrainingParams, mlflow_params=None)` (reference training/training.py:135-408).

Same flow: load train/val arrays, derive num_classes and the image geometry from the data
(:176-179), build the model, compile (optimizer + loss), iterate epochs over shuffled global
batches, validate, checkpoint `model_epochNN.hdf5` + `model_config.json`, early-stop.  The
compute of every step (forward, weighted CE, backward, all-reduce, Adam) is one liboctseg call.
Data-parallel runs are one process per GPU (torchrun): each rank takes its shard of every global
batch, the library all-reduces gradients over NCCL (MirroredStrategy semantics: per-replica BN
statistics, loss scaled by the global batch).  MLflow logging and plots are out of scope.
"""
import json
import logging as log
import time
from pathlib import Path

import numpy as np

from .. import parallel
from ..models import get_model_class
from .training_parameters import TrainingParams


def _load_dataset(path: Path):
    """train_images / train_labels / val_images / val_labels: the reference's HDF5 dataset keys
    (common/dataset_loader.py:9-22), read with the built-in minimal HDF5 reader (or from an .npz)."""
    from ..common import dataset_loader as dl
    ds = dl.open_dataset(path)
    tr_i, tr_l = dl.load_training_data(ds)
    va_i, va_l = dl.load_validation_data(ds)
    return tr_i, tr_l, va_i, va_l


def _class_weights(training_params: TrainingParams, train_labels, num_classes):
    cw = training_params.class_weight
    if cw is None:
        return np.ones(num_classes, np.float32)
    if isinstance(cw, str) and cw == "balanced":
        # sklearn.utils.class_weight.compute_class_weight("balanced"), as training.py:200-210
        counts = np.bincount(train_labels.reshape(-1).astype(np.int64), minlength=num_classes)
        return (train_labels.size / (num_classes * np.maximum(counts, 1))).astype(np.float32)
    return np.asarray(cw, np.float32)


def train_model(training_params: TrainingParams, mlflow_params=None, rank: int = 0, world: int = 1):
    train_images, train_labels, val_images, val_labels = _load_dataset(training_params.training_dataset_path)
    num_classes = len(np.unique(train_labels))                    # training.py:176
    H, W, C = train_images.shape[1], train_images.shape[2], train_images.shape[3]
    try:
        model_class = get_model_class(training_params.model_architecture)
    except ValueError as e:
        log.error(e)
        exit(1)
    container = model_class(input_channels=C, num_classes=num_classes, image_height=H, image_width=W,
                            **training_params.model_hyperparameters)
    model = container.build_model(device=rank % max(1, _device_count()))
    if world > 1:   # every replica starts from rank 0's initial weights (MirroredStrategy mirrors variables)
        import torch.distributed as dist
        box = [model.get_weights() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        model.set_weights(box[0])
    cw = _class_weights(training_params, train_labels, num_classes)
    opt = dict(learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7)
    opt.update({k: v for k, v in training_params.opt_params.items() if k in opt})
    per = parallel.split_global_batch(training_params.batch_size, world)
    model.engine.train_begin(cw, dropout_rate=0.5, dropout_seed=1234 + rank, global_batch=training_params.batch_size,
                             **opt)
    parallel.init_training_comm(model.engine)
    out_dir = training_params.results_location
    if rank == 0:
        out_dir.mkdir(parents=True, exist_ok=True)
        with open(out_dir / "model_config.json", "w") as f:       # training.py:50-51
            json.dump(container.get_config(), f)
    monitor, mode = training_params.model_save_monitor
    best, best_epoch, history = None, -1, []
    rng = np.random.default_rng(0)                                # same order on every rank
    n_train = len(train_images) // training_params.batch_size * training_params.batch_size
    for epoch in range(training_params.epochs):
        t0 = time.time()
        order = rng.permutation(len(train_images)) if training_params.shuffle else np.arange(len(train_images))
        losses = []
        for g0 in range(0, n_train, training_params.batch_size):
            a, b = parallel.shard_range(training_params.batch_size, rank, world)
            idx = np.sort(order[g0 + a:g0 + b])
            losses.append(model.engine.train_step(train_images[idx], train_labels[idx]))
        if world > 1:      # replicas: mean of the per-replica BN moving statistics, one loss for everybody
            import torch.distributed as dist
            parallel.sync_bn_moving_stats(model, [nm for nm, _ in model.engine.param_specs], dist)
        probs = model.predict(val_images)
        lab = val_labels.reshape(val_labels.shape[:3]).astype(np.int64)
        pt = np.clip(np.take_along_axis(probs, lab[..., None], -1)[..., 0], 1e-7, 1 - 1e-7)
        local_loss = float(np.sum(losses) / max(1, len(losses)))      # already scaled by the GLOBAL batch
        logs = {"loss": parallel.allreduce_sum_scalar(local_loss) if world > 1 else local_loss,
                "val_loss": float(np.mean(-cw[lab] * np.log(pt))),
                "val_acc": float((probs.argmax(-1) == lab).mean()), "epoch_time": time.time() - t0}
        history.append(logs)
        cur = logs.get(monitor, logs["val_loss"])
        improved = best is None or (cur > best if mode == "max" else cur < best)
        if improved:
            best, best_epoch = cur, epoch
        if rank == 0 and (improved or not training_params.model_save_best):
            model.save(out_dir / f"model_epoch{epoch + 1:02d}.hdf5")   # training.py:319-326 naming
        log.info(f"epoch {epoch + 1}: {logs}")
        if training_params.early_stopping and epoch - best_epoch >= training_params.patience:
            break
    return model, history


def _device_count() -> int:
    from .. import _native as nat
    return nat.load().octseg_device_count()
