"""Deterministic synthetic B-scans, labels and weights (no dataset / checkpoint is
reachable offline).  Definitions follow SURVEY.md section 8(d).

B-scan i: K-1 smooth boundaries b_j(x) = H*c_j + A_j*sin(2*pi*x/lambda_j + phi_j),
per-region mean intensity, multiplicative Rayleigh speckle, additive Gaussian
noise, clipped to uint8.  Label = number of boundaries at or above the pixel (the
boundary pixel belongs to the lower region -- reference convention,
oct_image_segmentation_models/min_path_processing/utils.py:5-12).
"""
from typing import List, Tuple

import numpy as np

from ..models.unet_spec import unet_param_specs

_REGION_MEANS = (30.0, 140.0, 90.0, 50.0, 110.0, 70.0, 160.0, 40.0)


def synthetic_boundaries(i: int, height: int, width: int, num_classes: int = 4) -> np.ndarray:
    """int32 [K-1, W] row index of the first pixel of each lower region."""
    rng = np.random.default_rng(1234 + i)
    nb = num_classes - 1
    centers = np.linspace(0.2, 0.8, nb) if nb > 1 else np.array([0.5])
    x = np.arange(width, dtype=np.float64)
    b = np.zeros((nb, width), dtype=np.float64)
    for j in range(nb):
        amp = rng.uniform(0.01, 0.03) * height
        lam = rng.uniform(0.5, 2.0) * width
        phi = rng.uniform(0, 2 * np.pi)
        b[j] = height * centers[j] + amp * np.sin(2 * np.pi * x / lam + phi)
    b = np.clip(np.rint(b), 1, height - 2).astype(np.int32)
    # keep boundaries strictly ordered top to bottom
    for j in range(1, nb):
        b[j] = np.maximum(b[j], b[j - 1] + 1)
    return np.clip(b, 1, height - 1)


def synthetic_bscan(i: int, height: int, width: int, num_classes: int = 4
                    ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(image u8 [H,W,1], label u8 [H,W,1], boundaries int32 [K-1,W])."""
    rng = np.random.default_rng(1234 + i)
    b = synthetic_boundaries(i, height, width, num_classes)
    rows = np.arange(height, dtype=np.int32)[:, None]
    label = np.zeros((height, width), dtype=np.uint8)
    for j in range(b.shape[0]):
        label += (rows >= b[j][None, :]).astype(np.uint8)
    means = np.asarray(_REGION_MEANS, dtype=np.float64)[label % len(_REGION_MEANS)]
    # consume the same leading draws as synthetic_boundaries so the two stay in step
    rng2 = np.random.default_rng(99991 + i)
    speckle = rng2.rayleigh(scale=0.25, size=(height, width)) / (0.25 * np.sqrt(np.pi / 2))
    img = means * speckle + rng2.normal(0.0, 8.0, size=(height, width))
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return img[..., None], label[..., None], b


def synthetic_batch(start: int, n: int, height: int, width: int, num_classes: int = 4
                    ) -> Tuple[np.ndarray, np.ndarray]:
    """(images u8 [n,H,W,1], labels u8 [n,H,W,1]) for B-scans start..start+n-1."""
    imgs = np.empty((n, height, width, 1), dtype=np.uint8)
    labs = np.empty((n, height, width, 1), dtype=np.uint8)
    for k in range(n):
        imgs[k], labs[k], _ = synthetic_bscan(start + k, height, width, num_classes)
    return imgs, labs


def fast_random_batch(seed: int, n: int, height: int, width: int) -> np.ndarray:
    """Cheap uniform-u8 images for throughput runs where content is irrelevant."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n, height, width, 1), dtype=np.uint8)


def synthetic_weights(seed: int = 42, random_bn_stats: bool = True, **spec_kwargs
                      ) -> List[np.ndarray]:
    """Keras-order float32 weights: Glorot-uniform kernels (Keras default), zero
    bias; BN gamma~U(.5,1.5), beta~N(0,.1), mean~N(0,.1), var~U(.5,1.5) so that
    BN folding is exercised (SURVEY.md 8(d) 'Weights (i)').  With
    random_bn_stats=False BN is at its Keras initial state (1,0,0,1), which is
    the right starting point for training."""
    rng = np.random.default_rng(seed)
    out: List[np.ndarray] = []
    for name, shape in unet_param_specs(**spec_kwargs):
        leaf = name.split("/")[-1]
        if leaf == "kernel:0":
            kh, kw, cin, cout = shape
            limit = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
            w = rng.uniform(-limit, limit, size=shape)
        elif leaf == "bias:0":
            w = np.zeros(shape)
        elif leaf == "gamma:0":
            w = rng.uniform(0.5, 1.5, size=shape) if random_bn_stats else np.ones(shape)
        elif leaf == "beta:0":
            w = rng.normal(0, 0.1, size=shape) if random_bn_stats else np.zeros(shape)
        elif leaf == "moving_mean:0":
            w = rng.normal(0, 0.1, size=shape) if random_bn_stats else np.zeros(shape)
        elif leaf == "moving_variance:0":
            w = rng.uniform(0.5, 1.5, size=shape) if random_bn_stats else np.ones(shape)
        else:  # pragma: no cover
            raise AssertionError(name)
        out.append(np.asarray(w, dtype=np.float32))
    return out
