"""Diagnostic: train N steps eagerly twice and once with CUDA-graph replay; per-tensor relative L2 differences."""
import os
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch, synthetic_weights
from oct_image_segmentation_models_b200.engine import UNetEngine
from oct_image_segmentation_models_b200.models.unet_spec import unet_param_specs

cfg = dict(input_channels=1, num_classes=4, start_neurons=8, pool_layers=2, conv_layers=2)
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
weights = synthetic_weights(seed=11, random_bn_stats=False, **cfg)
imgs, labs = synthetic_batch(77, 8, 64, 64)
names = [nm for nm, _ in unet_param_specs(**cfg)]


def run(mode, dual="1"):
    os.environ["OCTSEG_TRAIN_GRAPH"] = mode
    if dual == "0":
        os.environ["OCTSEG_NO_DUAL_STREAM"] = "1"
    else:
        os.environ.pop("OCTSEG_NO_DUAL_STREAM", None)
    eng = UNetEngine(precision=prec, **cfg)
    eng.set_weights(weights)
    eng.train_begin([0.5, 1, 2, 1], learning_rate=1e-3, dropout_rate=0.5, dropout_seed=123, global_batch=8)
    losses = [eng.train_step(imgs, labs) for _ in range(steps)]
    w = eng.get_weights()
    g = eng.get_grads()
    eng.close()
    return losses, w, g


def cmp(tag, A, B):
    print(tag, "losses", np.round(A[0], 6).tolist(), np.round(B[0], 6).tolist())
    for what, k in (("w", 1), ("g", 2)):
        worst = []
        for nm, a, b in zip(names, A[k], B[k]):
            if a is None or b is None or (nm.endswith("bias:0") and nm != names[-1]):
                continue
            err = np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-12)
            worst.append((err, nm))
        worst.sort(reverse=True)
        print("  ", what, [(f"{e:.2e}", n) for e, n in worst[:5]])


e1 = run("0")
e2 = run("0")
g1 = run("1")
s1 = run("0", dual="0")
g2 = run("1", dual="0")
cmp("eager vs eager", e1, e2)
cmp("graph vs eager", g1, e1)
cmp("eager single-stream vs eager", s1, e1)
cmp("graph single-stream vs eager", g2, e1)
