"""HDF5 dataset readers (reference common/dataset_loader.py:9-33).  `hdf5_data_file` is an
`hdf5_min.H5File` (or anything indexable that returns arrays / objects with `.read()`)."""
from pathlib import Path
from typing import List, Tuple

import numpy as np


def _arr(x):
    return x.read() if hasattr(x, "read") else np.asarray(x)


def open_dataset(path):
    """HDF5 (the reference's format) or .npz with the same keys."""
    path = Path(path)
    with open(path, "rb") as fh:
        magic = fh.read(8)
    if magic.startswith(b"\x89HDF"):
        from .hdf5_min import H5File
        return H5File(path)
    return np.load(path)


def load_training_data(hdf5_data_file):
    return _arr(hdf5_data_file["train_images"]), _arr(hdf5_data_file["train_labels"])


def load_validation_data(hdf5_data_file):
    return _arr(hdf5_data_file["val_images"]), _arr(hdf5_data_file["val_labels"])


def load_testing_data(hdf5_data_file) -> Tuple[np.ndarray, np.ndarray, List[Path]]:
    test_images = _arr(hdf5_data_file["test_images"])
    test_labels = _arr(hdf5_data_file["test_labels"])
    test_image_paths = [Path(str(x, "ascii") if isinstance(x, (bytes, np.bytes_)) else str(x))
                        for x in _arr(hdf5_data_file["test_images_source"])]
    return test_images, test_labels, test_image_paths
