"""`DataGenerator`: the Keras `Sequence` the reference feeds to `model.fit` (common/data_generator.py:285-346,
constructed at training/training.py:363-383).  Same constructor arguments, `__len__` = floor(samples / batch_size),
one shuffled pass per epoch, batches of (preprocessed float32 images [B,H,W,C], labels).  Augmentation (host-side
skimage code in the reference) is outside the accelerated path: only aug_mode "none" is accepted.

Batches are assembled by index gather; `prefetch()` returns an iterator that assembles batch i+1 on a worker thread
while the GPU runs step i."""
import logging as log
import queue
import threading
from typing import Callable, List, Tuple

import numpy as np


class DataGenerator:
    def __init__(self, images: np.ndarray, labels: np.ndarray, batch_size: int, aug_fn_args: List[Tuple] = (),
                 aug_mode: str = "none", aug_probs: Tuple = (), aug_fly: bool = False,
                 preprocess_input_fn: Callable = None, shuffle: bool = True, seed: int = 0):
        if aug_mode != "none" or len(aug_fn_args):
            log.error(f"Augmentation mode '{aug_mode}' is outside the accelerated path (only 'none'). Exiting...")
            exit(1)
        if images.ndim != 4 or len(images) != len(labels):
            raise ValueError("images must be [N,H,W,C] with one label map per image")
        self.images, self.labels = images, labels
        self.batch_size = int(batch_size)
        self.preprocess_input_fn = preprocess_input_fn
        self.shuffle = shuffle
        self.total_samples = len(images)
        self.num_batches = self.total_samples // self.batch_size
        self._rng = np.random.default_rng(seed)
        self._order = np.arange(self.total_samples)
        # raw uint8 B-scans with the U-Net's x/255 preprocessing can skip the host-side float conversion: the device
        # applies the identical float32(x)/255 (bit-equal, tests/test_oracle.py)
        self.raw_uint8 = images.dtype == np.uint8 and getattr(preprocess_input_fn, "__name__", "") == "preprocess_input_inner"
        self.on_epoch_end()

    def get_total_samples(self) -> int:
        return self.total_samples

    def __len__(self) -> int:
        return self.num_batches

    def indices(self, i: int) -> np.ndarray:
        if not 0 <= i < self.num_batches:
            raise IndexError(i)
        return np.sort(self._order[i * self.batch_size:(i + 1) * self.batch_size])

    def raw_batch(self, i: int, lo: int = 0, hi: int = None):
        """(uint8 / raw images, labels) of samples [lo, hi) of batch i, no preprocessing (data-parallel shards)."""
        idx = self.indices(i)[lo:hi]
        return self.images[idx], self.labels[idx]

    def __getitem__(self, i: int):
        x, y = self.raw_batch(i)
        x = x.astype(np.float64)
        if self.preprocess_input_fn is not None:
            x = self.preprocess_input_fn(x)
        return np.asarray(x, np.float32), y

    def on_epoch_end(self):
        if self.shuffle:
            self._order = self._rng.permutation(self.total_samples)

    def prefetch(self, fetch: Callable, depth: int = 2):
        """Iterate fetch(0), fetch(1), ... fetch(len-1), computed `depth` batches ahead on a worker thread."""
        q: "queue.Queue" = queue.Queue(maxsize=depth)

        def work():
            try:
                for i in range(self.num_batches):
                    q.put((i, fetch(i)))
                q.put(None)
            except BaseException as ex:  # noqa: BLE001 -- re-raised in the consumer
                q.put(ex)

        threading.Thread(target=work, daemon=True).start()
        while True:
            item = q.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            yield item
