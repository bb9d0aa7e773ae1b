"""A few training steps of BASELINE configs[2] (default U-Net, 512x256x1, bf16) and nothing else -- the
command profiled for profiles/r1_launches_train_step.csv and r1_wgrad_ncu_full_summary.csv.
usage: python tools/train_steps.py [batch=64] [steps=4]   (OCTSEG_TRAIN_GRAPH=0 so every launch is visible)"""
import os
import sys
from pathlib import Path

import numpy as np

os.environ.setdefault("OCTSEG_TRAIN_GRAPH", "0")
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oct_image_segmentation_models_b200.common.synthetic import fast_random_batch, synthetic_weights  # noqa: E402
from oct_image_segmentation_models_b200.engine import UNetEngine  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = dict(input_channels=1, num_classes=4)
eng = UNetEngine(precision="bf16", **cfg)
eng.set_weights(synthetic_weights(seed=3, random_bn_stats=False, **cfg))
eng.train_begin([0.5, 1.0, 2.0, 1.0], learning_rate=1e-3, dropout_rate=0.5, global_batch=batch)
imgs = fast_random_batch(7, batch, 512, 256)
labs = np.random.default_rng(1).integers(0, 4, size=(batch, 512, 256), dtype=np.uint8)
for i in range(steps):
    print("step", i, "loss", eng.train_step(imgs, labs), flush=True)
eng.close()
