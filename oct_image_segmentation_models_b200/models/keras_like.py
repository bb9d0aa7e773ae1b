"""Duck-typed stand-in for the Keras `Functional` model the reference passes around.

Consumers in the reference use exactly: `.name`, `.output.shape[-1]`,
`.predict(x, verbose=2, batch_size=1)` (prediction/prediction.py:61,75; evaluation/evaluation.py:99,129),
and for training `.compile/.fit/.save/.summary` (training/training.py:262,345,400-407).
All arithmetic is delegated to the CUDA engine; nothing here computes on the CPU.
"""
import json
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

from .. import _native as nat


class _Output:
    def __init__(self, num_classes: int):
        self.shape = (None, None, None, num_classes)


class B200Model:
    """`model = UNet(**config).build_model()`; weights live on the GPU."""

    def __init__(self, name: str, spec_kwargs: dict, precision: Optional[str] = None, device: int = 0,
                 init_seed: Optional[int] = None):
        import os
        self.name = name
        self.spec_kwargs = dict(spec_kwargs)
        # default = the mode that meets the reference contract (probabilities within 1e-4, identical boundaries): fp32
        # accuracy on the tensor cores; "bf16" / "fp16" trade that for 2x throughput
        self.precision = precision or os.environ.get("OCTSEG_PRECISION", "fp32")
        from ..engine import UNetEngine   # late import: engine imports models.unet_spec
        self.engine = UNetEngine(precision=self.precision, device=device, **spec_kwargs)
        self.device = device
        # fit() runs on a second handle when the training precision differs from the inference precision: the default
        # object predicts in the exact fp32 mode and trains in bf16 on the tensor cores (13.8 k instead of ~0.7 k
        # samples/s); OCTSEG_TRAIN_PRECISION=fp32 trains on the fp32 kernels, as the reference's arithmetic does
        self.train_precision = os.environ.get("OCTSEG_TRAIN_PRECISION", "bf16" if self.precision in ("fp32", "fp16") else self.precision)
        self._train_engine = None
        self.output = _Output(spec_kwargs["num_classes"])
        self.input_channels = spec_kwargs["input_channels"]
        self._compiled = None
        self.stop_training = False
        self.history = None
        # Keras initialises a freshly built model: glorot_uniform kernels, zero biases,
        # BatchNormalization (gamma, beta, moving_mean, moving_variance) = (1, 0, 0, 1)
        from ..common.synthetic import synthetic_weights
        import os as _os
        if init_seed is None and _os.environ.get("OCTSEG_INIT_SEED"):
            init_seed = int(_os.environ["OCTSEG_INIT_SEED"])
        seed = int(np.random.SeedSequence().entropy % (2 ** 31)) if init_seed is None else init_seed
        self.engine.set_weights(synthetic_weights(seed=seed, random_bn_stats=False, **spec_kwargs))

    # ---- Keras-compatible inference --------------------------------------------
    def predict(self, x, verbose=0, batch_size=None, **_):
        """x: [N,H,W,C] ALREADY preprocessed by `get_preprocess_input_fn()` (x/255.0, float64 in
        the reference, cast to float32 exactly as Keras does).  uint8 input is accepted too and
        takes the fused on-device preprocessing path (bit-identical result)."""
        x = np.asarray(x)
        if x.dtype == np.uint8:
            probs, _ = self.engine.predict(x)
            return probs
        x32 = np.ascontiguousarray(x, dtype=np.float32)
        return self.engine.predict_preprocessed(x32)

    def predict_raw(self, images_u8: np.ndarray, want_labels: bool = False):
        """Batched fast path on raw uint8 B-scans: (probs, labels)."""
        return self.engine.predict(images_u8, want_probs=True, want_labels=want_labels)

    # ---- weights -------------------------------------------------------------------
    def get_weights(self) -> List[np.ndarray]:
        return self.engine.get_weights()

    def set_weights(self, weights: Sequence[np.ndarray]):
        self.engine.set_weights(weights)
        if self._train_engine is not None:
            self._train_engine.set_weights(weights)

    def _trainer(self):
        """handle the train steps run on (created on first use, seeded with the current weights)"""
        if self.train_precision == self.precision:
            return self.engine
        if self._train_engine is None:
            from ..engine import UNetEngine
            self._train_engine = UNetEngine(precision=self.train_precision, device=self.device, **self.spec_kwargs)
            self._train_engine.set_weights(self.engine.get_weights())
        return self._train_engine

    def _pull_trained_weights(self):
        """after training steps: the inference handle follows the training handle (weights + BN moving statistics)"""
        if self._train_engine is not None:
            self.engine.set_weights(self._train_engine.get_weights())

    def count_params(self) -> int:
        return int(sum(int(np.prod(s)) for _, s in self.engine.param_specs))

    def summary(self, print_fn=print):
        print_fn(f'Model: "{self.name}"  ({self.precision}, liboctseg / sm_100a)')
        for name, shape in self.engine.param_specs:
            print_fn(f"  {name:44s} {tuple(shape)}")
        print_fn(f"Total params: {self.count_params():,}")

    # ---- persistence ---------------------------------------------------------------
    def _layer_weights(self):
        """[(layer_name, [(weight_name, array)])] in Keras model.layers order (weighted layers only)."""
        out, cur = [], None
        for (name, _), w in zip(self.engine.param_specs, self.get_weights()):
            layer, wname = name.split("/", 1)
            if cur is None or cur[0] != layer:
                cur = (layer, [])
                out.append(cur)
            cur[1].append((wname, w))
        return out

    def save(self, path, include_optimizer: bool = True, **_):
        """`.hdf5` / `.h5`: Keras-2.x weight layout (root attrs + model_weights/<layer>/<layer>/<w>:0, layer_names /
        weight_names attributes) written by the built-in minimal HDF5 writer, as the reference's ModelCheckpoint files
        are (training/training.py:319-326); a compiled model that has trained also gets `optimizer_weights`
        (Adam/iter:0, then the m and the v slot of every trainable variable -- Keras' optimizer.weights order) and a
        `training_config` attribute, so `load_model()` of this package resumes exactly.  The file is
        `load_weights`-compatible with Keras; its `model_config` attribute is a stub naming the graph
        hyper-parameters (not Keras' layer graph), so tf.keras.models.load_model cannot rebuild the model from it.
        Other suffixes: .npz."""
        path = Path(path)
        if path.suffix.lower() in (".hdf5", ".h5"):
            from ..common import hdf5_min
            cfg = json.dumps({"class_name": "Functional", "config": {"name": self.name},
                              "octseg_b200": {k: (list(v) if isinstance(v, tuple) else v)
                                              for k, v in self.spec_kwargs.items()}})
            opt_w, tc = None, ""
            if include_optimizer and self._compiled and self._compiled.get("started"):
                it, ms, vs = self._trainer().get_optimizer_state()
                names = [n for n, _ in self.engine.param_specs]
                train = [i for i, n in enumerate(names) if "moving_" not in n]
                opt_w = [("Adam/iter:0", np.asarray(it, np.int64))]
                opt_w += [(f"Adam/{names[i][:-2]}/m:0", ms[i]) for i in train]
                opt_w += [(f"Adam/{names[i][:-2]}/v:0", vs[i]) for i in train]
                tc = json.dumps({"loss": self._compiled["loss"].get_config(), "metrics": self._compiled["metric_names"],
                                 "optimizer_config": {"class_name": "Adam", "config": self._compiled["opt"]}})
            hdf5_min.save_keras_weights(path, self._layer_weights(), model_config=cfg, optimizer_weights=opt_w,
                                        training_config=tc)
            return
        arrays = {f"w{i:03d}": w for i, w in enumerate(self.get_weights())}
        with open(path, "wb") as f:
            np.savez(f, __model_name__=np.frombuffer(self.name.encode(), dtype=np.uint8), **arrays)

    # ---- Keras-compatible training (reference training/training.py:262-266, 401-407) ----------------------
    def compile(self, optimizer=None, loss=None, metrics=None, **_):
        """optimizer: training.optimizers.Adam (or any Keras-Adam-like object with get_config()); loss: a loss object
        from common.custom_losses (or its registry name); metrics: [monitor metric factories' results or names]."""
        from ..common import custom_losses
        from ..training.optimizers import adam_hyperparameters
        K = self.spec_kwargs["num_classes"]
        if loss is None or isinstance(loss, str):
            entry = custom_losses.custom_loss_objects.get(loss or "categorical_crossentropy")
            if entry is None:
                raise ValueError(f"Loss '{loss}' not found")
            loss = entry["function"](num_classes=K, is_y_true_sparse=entry["takes_sparse"])
        if not isinstance(loss, custom_losses.WeightedCategoricalCrossentropy):
            raise NotImplementedError("compile(loss=...): the accelerated train step implements the (weighted) categorical "
                                      "cross-entropy objects of common.custom_losses")
        names = []
        for m in (metrics or []):
            nm = m if isinstance(m, str) else getattr(m, "__name__", str(m))
            if nm not in ("dice_coef_macro", "dice_coef_micro", "acc", "accuracy"):
                raise NotImplementedError(f"metric '{nm}': available on the device: dice_coef_macro, dice_coef_micro, acc")
            names.append("acc" if nm == "accuracy" else nm)
        self._compiled = {"opt": adam_hyperparameters(optimizer), "optimizer": optimizer, "loss": loss,
                          "metric_names": names, "started": False, "resume": None}

    def _metrics_from_counts(self, counts_per_batch, loss_sums, pixels):
        """Keras aggregation: every metric is the mean over batches of its per-batch value.  counts: list of int64
        [b,K,3] arrays (one per batch) of (intersection, predicted, true) pixel counts (thresholded at 0.5, as
        reference common/custom_metrics.py:19-77)."""
        out = {"loss": float(np.sum(loss_sums) / max(1, pixels))}
        eps = 1e-5
        macro, micro, acc = [], [], []
        for c in counts_per_batch:
            c = c.astype(np.float64)
            macro.append(float(np.mean((2.0 * c[..., 0] + eps) / (c[..., 1] + c[..., 2] + eps))))
            micro.append(float(2.0 * c[..., 0].sum() / max(1.0, c[..., 1].sum() + c[..., 2].sum())))
            acc.append(float(c[..., 0].sum() / max(1.0, c[..., 2].sum())))   # confidently (p > 0.5) correct pixels
        out["dice_coef_macro"], out["dice_coef_micro"], out["acc"] = (float(np.mean(macro)), float(np.mean(micro)),
                                                                      float(np.mean(acc)))
        return out

    def evaluate_sequence(self, seq, max_batches=None, rank: int = 0, world: int = 1, dist=None):
        """loss + monitor metrics of a Sequence in inference mode, computed on the device (octseg_evaluate_host);
        with several ranks the batches are dealt round-robin and the per-batch results gathered."""
        cw = self._compiled["loss"].weights if self._compiled else None
        nb = len(seq) if max_batches is None else min(len(seq), max_batches)
        mine = []
        for i in range(rank, nb, world):
            if hasattr(seq, "raw_batch") and getattr(seq, "raw_uint8", False):
                x, y = seq.raw_batch(i)
                counts, ls = self.engine.evaluate_counts(x, _sparse(y, self.spec_kwargs["num_classes"]), cw)
            else:
                x, y = seq[i]
                counts, ls = self.engine.evaluate_counts(np.asarray(x, np.float32), _sparse(y, self.spec_kwargs["num_classes"]),
                                                         cw, preprocessed=True)
            mine.append((i, counts, float(ls.sum()), int(np.prod(np.asarray(y).shape[:3]))))
        if world > 1:
            box = [None] * world
            dist.all_gather_object(box, mine)
            mine = sorted((t for part in box for t in part), key=lambda t: t[0])
        return self._metrics_from_counts([m[1] for m in mine], [m[2] for m in mine], sum(m[3] for m in mine))

    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=1, callbacks=None, validation_data=None,
            shuffle=True, initial_epoch=0, **_):
        """Keras `Model.fit` for the call the reference makes (`x` / `validation_data` = Sequence objects yielding
        (preprocessed images, labels), Keras callback protocol with on_epoch_end(epoch, logs{"loss", "val_loss",
        <metric>, "val_<metric>"})).  One liboctseg train step per global batch; under torch.distributed every rank
        takes its shard of each batch and the library all-reduces the gradients (MirroredStrategy semantics).
        logs["loss"] is Keras' running mean of the batch losses; logs[<metric>] is evaluated after the epoch, in
        inference mode, on the first (up to 4) training batches (Keras averages train-mode batch metrics instead)."""
        import time
        from .. import parallel
        from ..common.data_generator import DataGenerator
        if not self._compiled:
            raise RuntimeError("You must compile your model before training/testing. Use `model.compile(optimizer, loss)`.")
        rank, world, dist = _dist_info()
        if not hasattr(x, "__getitem__") or isinstance(x, np.ndarray):
            if y is None or batch_size is None:
                raise ValueError("fit(x=array) needs y and batch_size")
            x = DataGenerator(np.asarray(x), np.asarray(y), batch_size, shuffle=shuffle)
        if isinstance(validation_data, tuple):
            validation_data = DataGenerator(np.asarray(validation_data[0]), np.asarray(validation_data[1]),
                                            batch_size or x.batch_size, shuffle=False)
        K = self.spec_kwargs["num_classes"]
        first_x, _ = x[0] if not getattr(x, "raw_uint8", False) else x.raw_batch(0)
        global_batch = len(first_x)
        per = parallel.split_global_batch(global_batch, world)
        a, b = parallel.shard_range(global_batch, rank, world)
        assert b - a == per
        comp = self._compiled
        if not comp["started"]:
            if world > 1:      # every replica starts from rank 0's weights (MirroredStrategy mirrors variables)
                box = [self.get_weights() if rank == 0 else None]
                dist.broadcast_object_list(box, src=0)
                self.set_weights(box[0])
            trainer = self._trainer()
            trainer.train_begin(comp["loss"].weights, dropout_rate=0.5, dropout_seed=1234 + rank,
                                global_batch=global_batch, **comp["opt"])
            parallel.init_training_comm(trainer, dist)
            if comp["resume"] is not None:
                trainer.set_optimizer_state(*comp["resume"])
            comp["started"] = True
        cbs = list(callbacks or [])
        for cb in cbs:
            cb.set_model(self)
            cb.set_params({"epochs": epochs, "steps": len(x), "verbose": verbose})
        self.stop_training = False
        history = {}
        for cb in cbs:
            cb.on_train_begin({})
        names = [nm for nm, _ in self.engine.param_specs]
        trainer = self._trainer()
        for epoch in range(initial_epoch, epochs):
            for cb in cbs:
                cb.on_epoch_begin(epoch, {})
            t0 = time.time()

            def fetch(i):
                if getattr(x, "raw_uint8", False):
                    xi, yi = x.raw_batch(i, a, b)
                    return xi, _sparse(yi, K), False
                xi, yi = x[i]
                return np.asarray(xi[a:b], np.float32), _sparse(np.asarray(yi)[a:b], K), True

            it = x.prefetch(fetch) if hasattr(x, "prefetch") else ((i, fetch(i)) for i in range(len(x)))
            losses = []
            for i, (xi, yi, pre) in it:
                losses.append(trainer.train_step(xi, yi, preprocessed=pre))
                for cb in cbs:
                    cb.on_train_batch_end(i, {"loss": losses[-1]})
            if hasattr(x, "on_epoch_end"):
                x.on_epoch_end()
            local = float(np.mean(losses)) if losses else float("nan")     # each = this rank's share of the global mean
            logs = {"loss": parallel.allreduce_sum_scalar(local, dist) if world > 1 else local}
            self._pull_trained_weights()
            if world > 1:
                parallel.sync_bn_moving_stats(self, names, dist)
            if comp["metric_names"]:
                tr = self.evaluate_sequence(x, max_batches=4, rank=rank, world=world, dist=dist)
                for nm in comp["metric_names"]:
                    logs[nm] = tr[nm]
            if validation_data is not None:
                va = self.evaluate_sequence(validation_data, rank=rank, world=world, dist=dist)
                logs["val_loss"] = va["loss"]
                for nm in comp["metric_names"]:
                    logs["val_" + nm] = va[nm]
            logs["epoch_time"] = time.time() - t0
            for k, v in logs.items():
                history.setdefault(k, []).append(v)
            if verbose and rank == 0:
                print(f"Epoch {epoch + 1}/{epochs} - " + " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()), flush=True)
            for cb in cbs:
                cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        for cb in cbs:
            cb.on_train_end({})
        self.history = type("History", (), {"history": history, "epoch": list(range(initial_epoch, initial_epoch + len(history.get("loss", []))))})()
        return self.history

    def load_weights(self, path):
        self.set_weights(read_weight_file(path)[1])

    def close(self):
        self.engine.close()
        if self._train_engine is not None:
            self._train_engine.close()
            self._train_engine = None


def _sparse(y, num_classes):
    """labels as class ids [B,H,W]: accepts sparse [B,H,W] / [B,H,W,1] or one-hot [B,H,W,K]"""
    y = np.asarray(y)
    if y.ndim == 4 and y.shape[-1] == num_classes and num_classes > 1:
        return y.argmax(-1).astype(np.uint8)
    return y.reshape(y.shape[:3]).astype(np.uint8)


def _dist_info():
    """(rank, world, torch.distributed or None) of the running job"""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.get_rank(), dist.get_world_size(), dist
    except ImportError:
        pass
    return 0, 1, None


def load_model(path, precision=None, device: int = 0, model_config=None):
    """Rebuild a model from a file written by `B200Model.save` (or a Keras `.hdf5` with a sibling
    model_config.json): weights, and -- when the file carries them -- the compile() configuration and the
    optimizer state, so that training resumes where the checkpoint was taken (reference utils.load_model,
    used for `initial_model`, training/training.py:236-238)."""
    from . import get_model_class
    from ..common import custom_losses, hdf5_min
    from ..training.optimizers import Adam
    path = Path(path)
    name, weights = read_weight_file(path)
    spec = None
    if model_config is None:
        try:
            _, cfg = hdf5_min.load_keras_weights(path)
            spec = json.loads(cfg).get("octseg_b200") if cfg else None
        except Exception:  # noqa: BLE001 -- not an HDF5 file / no stub: fall back to the sibling json
            spec = None
        if spec is None:
            with open(path.parent / "model_config.json") as f:
                model_config = json.load(f)
    if spec is not None:
        model = B200Model(name, {k: (tuple(v) if isinstance(v, list) else v) for k, v in spec.items()},
                          precision=precision, device=device)
    else:
        model = get_model_class(name)(**model_config).build_model(precision=precision, device=device)
    model.set_weights(weights)
    try:
        opt_w, tc = hdf5_min.load_keras_optimizer_weights(path)
    except Exception:  # noqa: BLE001
        opt_w, tc = [], None
    if tc:
        t = json.loads(tc)
        K = model.spec_kwargs["num_classes"]
        loss = custom_losses.WeightedCategoricalCrossentropy(t["loss"].get("weights"), K)
        model.compile(optimizer=Adam(**{k: v for k, v in t["optimizer_config"]["config"].items()
                                        if k in ("learning_rate", "beta_1", "beta_2", "epsilon")}),
                      loss=loss, metrics=t.get("metrics") or [])
        if opt_w:
            names = [n for n, _ in model.engine.param_specs]
            d = dict(opt_w)
            zeros = [np.zeros(s, np.float32) for _, s in model.engine.param_specs]
            ms = [d.get(f"Adam/{n[:-2]}/m:0", z) for n, z in zip(names, zeros)]
            vs = [d.get(f"Adam/{n[:-2]}/v:0", z) for n, z in zip(names, zeros)]
            model._compiled["resume"] = (int(np.asarray(d.get("Adam/iter:0", 0))), ms, vs)
    return model


def read_weight_file(path):
    """(model_name, [weights in Keras get_weights() order]) from a Keras-style HDF5 or an .npz container.
    HDF5 layers are mapped by ORDER and TYPE, never by exact name (Keras auto-names depend on a
    process-global counter)."""
    path = Path(path)
    with open(path, "rb") as fh:
        magic = fh.read(8)
    if magic.startswith(b"\x89HDF"):
        from ..common import hdf5_min
        layers, cfg = hdf5_min.load_keras_weights(path)
        name = "unet"
        if cfg:
            try:
                name = json.loads(cfg)["config"]["name"]
            except (ValueError, KeyError, TypeError):
                pass
        order = {"kernel": 0, "bias": 1, "gamma": 0, "beta": 1, "moving_mean": 2, "moving_variance": 3}
        conv_set, bn_set = {"kernel", "bias"}, {"gamma", "beta", "moving_mean", "moving_variance"}
        weights = []
        for lname, ws in layers:
            kinds = {kv[0].split("/")[-1].split(":")[0] for kv in ws}
            if ws and kinds != conv_set and kinds != bn_set:
                raise ValueError(f"{path}: layer '{lname}' holds {sorted(kinds)}; a U-Net file has Conv2D layers "
                                 "(kernel, bias) and BatchNormalization layers (gamma, beta, moving_mean, moving_variance) only")
            ws = sorted(ws, key=lambda kv: order.get(kv[0].split("/")[-1].split(":")[0], 99))
            weights += [np.asarray(a, np.float32) for _, a in ws]
        return name, weights
    with np.load(path) as z:
        name = bytes(z["__model_name__"]).decode() if "__model_name__" in z.files else "unet"
        n = len([k for k in z.files if k.startswith("w")])
        return name, [z[f"w{i:03d}"] for i in range(n)]
