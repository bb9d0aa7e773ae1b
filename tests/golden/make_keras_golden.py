"""Pins oracle/unet_oracle.py (and common/hdf5_min.py) against the REAL reference: run this where the reference's
own stack is installed (tensorflow==2.9.0 + h5py, e.g. the reference's docker image, docker/Dockerfile:1) with the
reference checkout on PYTHONPATH:

    PYTHONPATH=/path/to/oct-image-segmentation-models:/path/to/this/repo python tests/golden/make_keras_golden.py

It builds the reference's own Keras graph (oct_image_segmentation_models.models.unet.UNet(...).build_model(),
reference models/unet.py:106-153), loads this repo's deterministic synthetic weights into it with
model.set_weights() (Keras order == octseg_param_info order), and stores

    tests/golden/keras_golden.npz        images, model.predict() probabilities (inference), one training-mode
                                         forward/backward: loss of the reference's weighted_categorical_crossentropy
                                         (common/custom_losses.py:11-37), every trainable gradient, the BN moving
                                         statistics after that step
    tests/golden/keras_golden_model.hdf5 model.save() output -- a real Keras/h5py file for common/hdf5_min.py

tests/test_keras_golden.py activates as soon as those two files exist and holds the oracle (CPU) and the CUDA path
(GPU) to them.  This container has neither TensorFlow nor h5py, so the fixture could not be generated here:
until it is, the network oracle stays "parity unpinned" (DESIGN.md section 2)."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch, synthetic_weights  # noqa: E402

CFG = dict(input_channels=1, num_classes=4, start_neurons=8, pool_layers=2, conv_layers=2)
N, H, W = 4, 64, 64
CLASS_W = [0.5, 1.0, 2.0, 1.0]


def main():
    import tensorflow as tf
    from oct_image_segmentation_models.common.custom_losses import weighted_categorical_crossentropy
    from oct_image_segmentation_models.models.unet import UNet

    tf.keras.utils.set_random_seed(0)
    unet = UNet(image_height=H, image_width=W, **CFG)
    model = unet.build_model()
    weights = synthetic_weights(seed=42, **CFG)
    assert [tuple(w.shape) for w in model.get_weights()] == [tuple(w.shape) for w in weights], "Keras weight order changed"
    model.set_weights(weights)
    imgs, labs = synthetic_batch(10, N, H, W, CFG["num_classes"])
    x = unet.get_preprocess_input_fn()(imgs)               # x / 255.0 in float64, cast by Keras
    probs = model.predict(x, verbose=0, batch_size=1)
    out = {"images": imgs, "labels": labs, "probs": probs.astype(np.float32)}

    # one training-mode pass: batch-statistics BN; dropout made deterministic by an explicit mask on the
    # bottleneck output (Keras Dropout(0.5) scales kept units by 2)
    loss_fn = weighted_categorical_crossentropy(np.asarray(CLASS_W, np.float32))
    y = tf.one_hot(labs[..., 0], CFG["num_classes"])
    drop = [l for l in model.layers if isinstance(l, tf.keras.layers.Dropout)][0]
    rng = np.random.default_rng(9)
    cmid = CFG["start_neurons"] << CFG["pool_layers"]
    mask = (rng.random((N, H >> CFG["pool_layers"], W >> CFG["pool_layers"], cmid)) < 0.5).astype(np.uint8)
    drop.call = lambda inputs, training=None: inputs * tf.constant(mask.astype(np.float32) * 2.0)   # noqa: E731
    with tf.GradientTape() as tape:
        p_train = model(tf.constant(x, tf.float32), training=True)
        loss = tf.reduce_mean(loss_fn(y, p_train))          # Keras SUM_OVER_BATCH_SIZE over all N*H*W pixels
    grads = tape.gradient(loss, model.trainable_variables)
    out["dropout_mask"] = mask
    out["train_loss"] = np.float64(loss.numpy())
    names = [v.name for v in model.trainable_variables]
    out["grad_names"] = np.asarray(names)
    for i, g in enumerate(grads):
        out[f"grad{i:03d}"] = g.numpy()
    for i, w in enumerate(model.get_weights()):             # includes the updated moving statistics
        out[f"after{i:03d}"] = w
    np.savez_compressed(HERE / "keras_golden.npz", **out)
    model.set_weights(weights)
    model.save(str(HERE / "keras_golden_model.hdf5"))
    print("wrote keras_golden.npz and keras_golden_model.hdf5; tf", tf.__version__)


if __name__ == "__main__":
    main()
