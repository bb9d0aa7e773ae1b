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
        self.precision = precision or os.environ.get("OCTSEG_PRECISION", "bf16")
        from ..engine import UNetEngine   # late import: engine imports models.unet_spec
        self.engine = UNetEngine(precision=self.precision, device=device, **spec_kwargs)
        self.output = _Output(spec_kwargs["num_classes"])
        self.input_channels = spec_kwargs["input_channels"]
        self._compiled = None
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

    def save(self, path, **_):
        """`.hdf5` / `.h5`: Keras-2.x weight layout (root attrs + model_weights/<layer>/<layer>/<w>:0,
        layer_names / weight_names attributes) written by the built-in minimal HDF5 writer, which is what
        ModelCheckpoint produces in the reference (training/training.py:319-326).  Other suffixes: .npz."""
        path = Path(path)
        if path.suffix.lower() in (".hdf5", ".h5"):
            from ..common import hdf5_min
            cfg = json.dumps({"class_name": "Functional", "config": {"name": self.name},
                              "octseg_b200": {k: (list(v) if isinstance(v, tuple) else v)
                                              for k, v in self.spec_kwargs.items()}})
            hdf5_min.save_keras_weights(path, self._layer_weights(), model_config=cfg)
            return
        arrays = {f"w{i:03d}": w for i, w in enumerate(self.get_weights())}
        with open(path, "wb") as f:
            np.savez(f, __model_name__=np.frombuffer(self.name.encode(), dtype=np.uint8), **arrays)

    def load_weights(self, path):
        self.set_weights(read_weight_file(path)[1])

    def close(self):
        self.engine.close()


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
        weights = []
        for _, ws in layers:
            ws = sorted(ws, key=lambda kv: order.get(kv[0].split("/")[-1].split(":")[0], 99))
            weights += [np.asarray(a, np.float32) for _, a in ws]
        return name, weights
    with np.load(path) as z:
        name = bytes(z["__model_name__"]).decode() if "__model_name__" in z.files else "unet"
        n = len([k for k in z.files if k.startswith("w")])
        return name, [z[f"w{i:03d}"] for i in range(n)]
