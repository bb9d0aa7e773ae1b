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

from ..engine import UNetEngine
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
        self.engine = UNetEngine(precision=self.precision, device=device, **spec_kwargs)
        self.output = _Output(spec_kwargs["num_classes"])
        self.input_channels = spec_kwargs["input_channels"]
        self._compiled = None
        # Keras initialises a freshly built model: glorot_uniform kernels, zero biases,
        # BatchNormalization (gamma, beta, moving_mean, moving_variance) = (1, 0, 0, 1)
        from ..common.synthetic import synthetic_weights
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
    def save(self, path, **_):
        """Weights container next to `model_config.json`.  (.npz payload; a Keras-compatible
        HDF5 writer is the f-1 'next' row of SURVEY section 8.)"""
        path = Path(path)
        arrays = {f"w{i:03d}": w for i, w in enumerate(self.get_weights())}
        names = json.dumps([n for n, _ in self.engine.param_specs])
        with open(path, "wb") as f:
            np.savez(f, __names__=np.frombuffer(names.encode(), dtype=np.uint8),
                     __model_name__=np.frombuffer(self.name.encode(), dtype=np.uint8), **arrays)

    def load_weights(self, path):
        with np.load(Path(path)) as z:
            n = len([k for k in z.files if k.startswith("w")])
            self.set_weights([z[f"w{i:03d}"] for i in range(n)])

    def close(self):
        self.engine.close()
