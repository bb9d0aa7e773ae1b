"""`train_model(TrainingParams, mlflow_params=None)` (reference training/training.py:135-408), same flow and the same
objects: load train/val arrays, derive num_classes and the image geometry from the data (:176-179), build the
optimizer with `opt_con(**opt_params)` (:190-193), look the loss up in `custom_loss_objects` (:195-217; class weights
"balanced" over train+val labels as :200-206) and the monitor metric in `training_monitor_metric_objects` (:219-224),
build or load the model (:236-266), then `model.compile(...)`, callbacks `ModelCheckpoint("model_epoch{epoch:02d}.hdf5")`
+ `SaveEpochInfo` + `EarlyStopping(monitor=f"val_{metric}", mode="max", restore_best_weights)` (:319-342), two
`DataGenerator`s (:363-383) and `model.fit(x=train_gen, validation_data=val_gen, epochs, callbacks)` (:401-407).

Every train step (forward, weighted CE, backward, all-reduce, Adam) is one liboctseg call made by `B200Model.fit`.
Data-parallel runs are one process per GPU (torchrun): each rank takes its shard of every global batch, the library
all-reduces gradients over NCCL (MirroredStrategy semantics: per-replica BN statistics, loss scaled by the global
batch); file-writing callbacks run on rank 0 only.  MLflow logging and plots are out of scope.
"""
import json
import logging as log
from pathlib import Path

import numpy as np

from ..common import custom_losses, custom_metrics
from ..common.data_generator import DataGenerator
from ..common.utils import get_timestamp
from ..models import get_model_class
from ..models.keras_like import _dist_info, load_model
from . import training_callbacks
from .training_parameters import TrainingParams


def _load_dataset(path: Path):
    """train_images / train_labels / val_images / val_labels: the reference's HDF5 dataset keys
    (common/dataset_loader.py:9-22), read with the built-in minimal HDF5 reader (or from an .npz)."""
    from ..common import dataset_loader as dl
    ds = dl.open_dataset(path)
    tr_i, tr_l = dl.load_training_data(ds)
    va_i, va_l = dl.load_validation_data(ds)
    return tr_i, tr_l, va_i, va_l


def compute_class_weight_balanced(labels: np.ndarray) -> np.ndarray:
    """sklearn.utils.class_weight.compute_class_weight("balanced", classes=np.unique(y), y=y.flatten()) =
    n_samples / (n_classes * bincount(y)) over the classes PRESENT in y (reference training.py:200-206)."""
    y = np.asarray(labels).reshape(-1).astype(np.int64)
    classes = np.unique(y)
    counts = np.bincount(y, minlength=int(classes.max()) + 1)[classes]
    return (y.size / (len(classes) * counts.astype(np.float64))).astype(np.float32)


def _save_training_params_file(folder: Path, model_summary: str, model_config: dict, dataset_md5: str, c_weight,
                               timestamp: str, tp: TrainingParams, optimizer):
    """text twin of the reference's params.hdf5 attributes (training.py:40-132): one json next to the checkpoints"""
    cfg = {"timestamp": timestamp, "model_config": model_config, "training_dataset": str(tp.training_dataset_path),
           "training_dataset_md5": dataset_md5, "loss": tp.loss, "metric": tp.metric, "epochs": tp.epochs,
           "batch_size": tp.batch_size, "class_weight": None if c_weight is None else [float(x) for x in c_weight],
           "optimizer": getattr(tp.opt_con, "__name__", str(tp.opt_con)),
           "opt_params": optimizer.get_config() if hasattr(optimizer, "get_config") else {},
           "model_summary": model_summary}
    with open(folder / "params.json", "w") as f:
        json.dump(cfg, f, indent=1, default=str)
    with open(folder / "model_config.json", "w") as f:            # what load_model_and_config reads (utils.py:60-62)
        json.dump(model_config, f)


def train_model(training_params: TrainingParams, mlflow_params=None):
    if mlflow_params:
        log.error("MLflow logging is outside the accelerated path; pass mlflow_params=None")
        exit(1)
    rank, world, _ = _dist_info()
    train_images, train_labels, val_images, val_labels = _load_dataset(training_params.training_dataset_path)
    num_classes = len(np.unique(train_labels))                    # training.py:176
    log.info(f"Detected {num_classes} classes")
    _, image_height, image_width, input_channels = train_images.shape

    optimizer_con = training_params.opt_con
    if isinstance(optimizer_con, str):
        if optimizer_con.lower() != "adam":
            log.error(f"Optimizer '{optimizer_con}' not found. Exiting...")
            exit(1)
        from .optimizers import Adam as optimizer_con
    optimizer = optimizer_con(**training_params.opt_params)      # training.py:190-193

    loss = custom_losses.custom_loss_objects.get(training_params.loss)
    if loss is None:
        log.error(f"Loss '{training_params.loss}' not found. Exiting...")
        exit(1)
    if training_params.class_weight == "balanced":
        c_weight = compute_class_weight_balanced(np.concatenate((train_labels, val_labels)))
    elif isinstance(training_params.class_weight, (list, tuple, np.ndarray)):
        c_weight = np.asarray(training_params.class_weight, np.float32)
    else:
        c_weight = None
    sparse_labels = loss["takes_sparse"]
    try:
        kwargs = dict(training_params.loss_fn_kwargs)
        if c_weight is not None and training_params.loss == "weighted_categorical_crossentropy":
            kwargs.setdefault("weights", c_weight)
        loss_fn = loss["function"](num_classes=num_classes, is_y_true_sparse=sparse_labels, **kwargs)
    except NotImplementedError as e:
        log.error(f"{e}. Exiting...")
        exit(1)

    metric = custom_metrics.training_monitor_metric_objects.get(training_params.metric)
    if metric is None:
        log.error(f"Metric '{training_params.metric}' not found. Exiting...")
        exit(1)
    metric_fn = metric(sparse_labels, num_classes)

    if training_params.initial_model:
        log.info(f"Starting training from model: {training_params.initial_model}")
        model = load_model(Path(training_params.initial_model), device=rank % max(1, _device_count()))
        model_config = dict(input_channels=input_channels, num_classes=num_classes, image_height=image_height,
                            image_width=image_width, **{k: v for k, v in model.spec_kwargs.items()
                                                        if k not in ("input_channels", "num_classes")})
        preprocess = get_model_class(model.name)(**model_config).get_preprocess_input_fn()
        if model._compiled is None:      # a weights-only file: compile as a fresh run would
            model.compile(optimizer=optimizer, loss=loss_fn, metrics=[metric_fn])
    else:
        log.info(f"Starting training from scratch {training_params.model_architecture} model")
        try:
            model_class = get_model_class(training_params.model_architecture)
        except ValueError as e:
            log.error(e)
            exit(1)
        container = model_class(input_channels=input_channels, num_classes=num_classes, image_height=image_height,
                                image_width=image_width, **training_params.model_hyperparameters)
        model = container.build_model(device=rank % max(1, _device_count()))
        model.compile(optimizer=optimizer, loss=loss_fn, metrics=[metric_fn])
        model_config = container.get_config()
        preprocess = container.get_preprocess_input_fn()

    monitor = training_params.model_save_monitor
    timestamp = get_timestamp()
    save_foldername = training_params.results_location / Path(timestamp + "_" + training_params.model_architecture)
    callbacks_list = []
    if rank == 0:
        save_foldername.mkdir(parents=True, exist_ok=True)
        callbacks_list.append(training_callbacks.ModelCheckpoint(
            filepath=save_foldername / "model_epoch{epoch:02d}.hdf5", save_best_only=training_params.model_save_best,
            monitor=monitor[0], mode=monitor[1]))
        callbacks_list.append(training_callbacks.SaveEpochInfo(save_folder=save_foldername, train_params=training_params))
    if training_params.early_stopping:
        callbacks_list.append(training_callbacks.EarlyStopping(
            monitor=f"val_{training_params.metric}", mode="max", patience=training_params.patience,
            restore_best_weights=training_params.restore_best_weights))
    if rank == 0:
        summary = []
        model.summary(print_fn=summary.append)
        from ..common.utils import md5
        _save_training_params_file(save_foldername, "\n".join(summary), model_config,
                                   md5(training_params.training_dataset_path), c_weight, timestamp, training_params,
                                   optimizer)

    bs = training_params.batch_size
    train_gen = DataGenerator(train_images, train_labels, bs, [], "none", (), False, preprocess,
                              shuffle=training_params.shuffle)
    val_gen = DataGenerator(val_images, val_labels, bs, [], "none", (), False, preprocess, shuffle=False)
    for what, gen in (("training", train_gen), ("validation", val_gen)):
        if bs > gen.get_total_samples():
            log.error(f"The batch size ({bs}) cannot be larger than the number of {what} samples "
                      f"({gen.get_total_samples()})")
            exit(1)
    history = model.fit(x=train_gen, validation_data=val_gen, epochs=training_params.epochs, callbacks=callbacks_list,
                        verbose=1)
    model.results_folder = save_foldername
    return model, history


def _device_count() -> int:
    from .. import _native as nat
    return nat.load().octseg_device_count()
