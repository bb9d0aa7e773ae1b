import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch, synthetic_weights
from oct_image_segmentation_models_b200.models.unet_spec import unet_param_specs
from oct_image_segmentation_models_b200.engine import UNetEngine
cfg = dict(input_channels=1, num_classes=4, start_neurons=16, pool_layers=3, conv_layers=2)
n, h, w = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
weights = synthetic_weights(seed=13, random_bn_stats=True, **cfg)
imgs, labs = synthetic_batch(40, n, h, w, 4)
P = 3; cmid = 16 << P
mask = (np.random.default_rng(9).random((n, h >> P, w >> P, cmid)) < 0.5).astype(np.uint8)
names = [nm for nm, _ in unet_param_specs(**cfg)]
def run(tc_off, rowpair_off):
    os.environ["OCTSEG_DISABLE_TC"] = "1" if tc_off else "0"
    os.environ["OCTSEG_DISABLE_ROWPAIR"] = "1" if rowpair_off else "0"
    eng = UNetEngine(precision="bf16", **cfg); eng.set_weights(weights)
    eng.train_begin([0.5, 1, 2, 1], global_batch=n)
    loss = eng.train_step(imgs, labs, dropout_mask=mask); g = eng.get_grads(); eng.close()
    return loss, g
ref = run(True, True)
for tag, rp in (("tc+rowpair", False), ("tc no rowpair", True)):
    got = run(False, rp)
    errs = []
    for nm, a, b in zip(names, got[1], ref[1]):
        if a is None or b is None or "moving" in nm or (nm.endswith("bias:0") and nm != names[-1]): continue
        errs.append((round(float(np.linalg.norm((a-b).ravel())/max(np.linalg.norm(b.ravel()),1e-12)),3), nm))
    print(tag, "loss", got[0], ref[0], "worst", max(errs)); print(sorted(errs, reverse=True)[:8])
