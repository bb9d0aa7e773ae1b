"""`Dataset` container (reference common/dataset.py:10-32)."""
from pathlib import Path
from typing import List, Optional

import numpy as np


class Dataset:
    def __init__(self, images: np.ndarray, image_masks: Optional[np.ndarray], image_names: List[Path],
                 image_output_dirs: List[Path]):
        if not isinstance(images, np.ndarray):
            raise TypeError("images must be a numpy array")
        self.images = images
        self.image_masks = image_masks
        self.image_names = image_names
        self.image_output_dirs = image_output_dirs
