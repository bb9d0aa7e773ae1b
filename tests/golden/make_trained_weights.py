"""Trains the CPU oracle for a few hundred Keras-Adam steps on synthetic B-scans so that the
network output is layered (SURVEY.md 8(d) 'Weights (ii)'); stores the weights, a held-out
batch, the oracle probabilities and the oracle min-path boundaries.
Run: python tests/golden/make_trained_weights.py   (about 2-3 minutes of CPU)"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch, synthetic_weights  # noqa: E402
from oracle import postproc  # noqa: E402
from oracle.unet_oracle import KerasAdam, OracleUNet  # noqa: E402

CFG = dict(input_channels=1, num_classes=4, start_neurons=8, pool_layers=2, conv_layers=2)
H = W = 64


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    w0 = synthetic_weights(seed=7, random_bn_stats=False, **CFG)
    net = OracleUNet(w0, **CFG)
    opt = KerasAdam(lr=2e-3)
    imgs, labs = synthetic_batch(100, 96, H, W)
    rng = np.random.default_rng(0)
    t0 = time.time()
    for step in range(260):
        idx = rng.choice(96, size=12, replace=False)
        loss, grads, stats, _ = net.loss_and_grads(imgs[idx], labs[idx], [0.5, 1.0, 2.0, 1.0])
        opt.step(net.params, grads)
        net.apply_bn_moving_update(stats)
        if step % 20 == 0:
            print(f"step {step} loss {loss:.4f} ({time.time() - t0:.0f}s)")
    # BN moving stats need more than 260 steps at momentum 0.99 -- settle them on the training set
    for _ in range(300):
        idx = rng.choice(96, size=12, replace=False)
        _, _, stats, _ = net.loss_and_grads(imgs[idx], labs[idx], [0.5, 1.0, 2.0, 1.0])
        net.apply_bn_moving_update(stats)
    test_imgs, test_labs = synthetic_batch(1000, 6, H, W)
    probs = net.predict(test_imgs)
    acc = float((probs.argmax(-1) == test_labs[..., 0]).mean())
    segs = np.stack([postproc.boundaries_from_probs(probs[i:i + 1]) for i in range(len(test_imgs))])
    print("held-out pixel accuracy", acc)
    weights = net.get_weights()
    np.savez_compressed(Path(__file__).parent / "trained_small_unet.npz",
                        images=test_imgs, labels=test_labs, probs=probs.astype(np.float32), segs=segs,
                        **{f"w{i:03d}": w for i, w in enumerate(weights)})
    print("saved; boundaries sample", segs[0][:, :8])


if __name__ == "__main__":
    main()
