"""Golden for BASELINE configs[4] / the north_star boundary gate at full size: the DEFAULT U-Net (start_neurons 8,
4 pools) trained on synthetic 512x512 B-scans so that its output is layered, plus the CPU oracle's answers on
256 held-out B-scans.

  stage 1 (needs a B200; run under gpurun):   python tests/golden/make_trained_default_golden.py train
      trains with the GPU trainer (any weights would do -- the trainer is only the cheapest way to get a
      confident net) and writes tests/golden/trained_default_unet_weights.npz
  stage 2 (CPU, about two minutes):            python tests/golden/make_trained_default_golden.py oracle
      runs oracle/unet_oracle.py + oracle/postproc.py + oracle/min_path.py on synthetic_batch(0, 256, 512, 512)
      and writes tests/golden/trained_default_unet_golden.npz (argmax labels, min-path boundaries, probabilities
      of rows 192..319 of the first B-scan)

The images are not stored: synthetic_batch() is deterministic."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch, synthetic_weights  # noqa: E402

CFG = dict(input_channels=1, num_classes=4)
H = W = 512
N_EVAL = 256
HERE = Path(__file__).resolve().parent
OUT_DIR = Path(sys.argv[2]) if len(sys.argv) > 2 else HERE


def train():
    from oct_image_segmentation_models_b200.engine import UNetEngine
    tr_i, tr_l = synthetic_batch(100000, 128, H, W)
    eng = UNetEngine(precision="bf16", **CFG)
    eng.set_weights(synthetic_weights(seed=11, random_bn_stats=False, **CFG))
    eng.train_begin([0.5, 1.0, 2.0, 1.0], learning_rate=2e-3, dropout_rate=0.5, dropout_seed=3, global_batch=16)
    rng = np.random.default_rng(0)
    for s in range(800):
        idx = rng.choice(len(tr_i), 16, replace=False)
        loss = eng.train_step(tr_i[idx], tr_l[idx])
        if s % 100 == 0:
            print(f"train step {s} loss {loss:.4f}", flush=True)
    # BatchNorm moving statistics (momentum 0.99) need more steps than the weights: settle them at learning rate 0
    eng.train_begin([0.5, 1.0, 2.0, 1.0], learning_rate=0.0, dropout_rate=0.5, dropout_seed=4, global_batch=16)
    for s in range(600):
        idx = rng.choice(len(tr_i), 16, replace=False)
        eng.train_step(tr_i[idx], tr_l[idx])
    weights = eng.get_weights()
    ev_i, ev_l = synthetic_batch(0, 16, H, W)
    _, lab = eng.predict(ev_i, want_probs=False, want_labels=True)
    print("held-out pixel accuracy (inference mode)", float((lab == ev_l[..., 0]).mean()), flush=True)
    eng.close()
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT_DIR / "trained_default_unet_weights.npz", **{f"w{i:03d}": w for i, w in enumerate(weights)})
    print("final loss", loss, "-> trained_default_unet_weights.npz")


def oracle():
    import torch
    from oracle import postproc
    from oracle.unet_oracle import OracleUNet
    torch.set_num_threads(8)
    g = np.load(HERE / "trained_default_unet_weights.npz")
    weights = [g[f"w{i:03d}"] for i in range(len(g.files))]
    net = OracleUNet(weights, **CFG)
    imgs, labs = synthetic_batch(0, N_EVAL, H, W)
    labels = np.empty((N_EVAL, H, W), np.uint8)
    segs = np.empty((N_EVAL, CFG["num_classes"] - 1, W), np.uint16)
    probs_band = None
    t0 = time.time()
    for i0 in range(0, N_EVAL, 4):
        p = net.predict(imgs[i0:i0 + 4])
        if i0 == 0:
            probs_band = p[0, 192:320].astype(np.float32)
        labels[i0:i0 + 4] = p.argmax(-1)
        for k in range(p.shape[0]):
            segs[i0 + k] = postproc.boundaries_from_probs(p[k:k + 1])
        if i0 % 32 == 0:
            print(f"{i0} / {N_EVAL}  ({time.time() - t0:.0f} s)", flush=True)
    acc = float((labels == labs[..., 0]).mean())
    print("oracle pixel accuracy vs truth", acc)
    np.savez_compressed(HERE / "trained_default_unet_golden.npz", labels=labels, segs=segs, probs_band=probs_band,
                        accuracy=np.float64(acc))


if __name__ == "__main__":
    {"train": train, "oracle": oracle}[sys.argv[1]]()
