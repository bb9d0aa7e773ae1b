"""`TrainingParams` with the reference's constructor signature
(training/training_parameters.py:50-135).  `opt_con` is an optimizer *constructor*, called as
`opt_con(**opt_params)` exactly as the reference does (training/training.py:190-193): pass
`oct_image_segmentation_models_b200.training.optimizers.Adam` (or tf.keras.optimizers.Adam where TensorFlow exists --
anything whose instance has a Keras-Adam `get_config()`); the string "adam" is accepted as a shorthand."""
import logging as log
from pathlib import Path
from typing import Callable, Optional, Union


class TrainingParams:
    def __init__(
        self,
        model_architecture: str,
        training_dataset_path: Path,
        initial_model: Optional[Path],
        results_location: Path,
        opt_con: Union[Callable, str],
        loss: str,
        metric: str,
        epochs: int,
        batch_size: int,
        model_hyperparameters: dict = {},
        opt_params: dict = {},
        loss_fn_kwargs: dict = {},
        augmentations: list = [],
        aug_mode: str = "none",
        aug_probs: tuple = (),
        aug_fly: bool = False,
        aug_val: bool = True,
        shuffle: bool = True,
        model_save_best: bool = True,
        model_save_monitor: tuple = ("val_acc", "max"),
        class_weight: Union[None, list, str] = None,
        channels_last: bool = True,
        early_stopping: bool = True,
        restore_best_weights: bool = True,
        patience: int = 50,
    ) -> None:
        self.model_architecture = model_architecture
        self.training_dataset_path = Path(training_dataset_path)
        self.training_dataset_name = self.training_dataset_path.stem
        self.initial_model = initial_model
        self.results_location = Path(results_location)
        self.opt_con = opt_con
        self.opt_params = dict(opt_params)
        self.loss = loss
        self.loss_fn_kwargs = dict(loss_fn_kwargs)
        self.metric = metric
        self.epochs = epochs
        self.batch_size = batch_size
        self.model_hyperparameters = dict(model_hyperparameters)
        self.augmentations = list(augmentations)
        self.aug_mode = aug_mode
        self.aug_probs = aug_probs
        self.aug_fly = aug_fly
        self.aug_val = aug_val
        self.shuffle = shuffle
        self.model_save_best = model_save_best
        self.model_save_monitor = model_save_monitor
        self.class_weight = class_weight
        self.channels_last = channels_last
        self.early_stopping = early_stopping
        self.restore_best_weights = restore_best_weights
        self.patience = patience
        if aug_mode != "none" or augmentations:
            # reference: common/augmentation.py (host-side, out of the hot path -- SURVEY section 2 #13)
            log.error("augmentations are outside the B200 hot path; pass aug_mode='none'")
            exit(1)
        # `loss` is looked up in common.custom_losses.custom_loss_objects by train_model (reference training.py:195-198);
        # names the reference registers but this package has no kernel for fail there with the reason
