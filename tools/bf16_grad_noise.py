"""Per-tensor relative gradient error of the bf16 training step vs the fp32 oracle (diagnostic)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.unet_oracle import OracleUNet
from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch, synthetic_weights
from oct_image_segmentation_models_b200.models.unet_spec import unet_param_specs
from oct_image_segmentation_models_b200.engine import UNetEngine

cfg = dict(input_channels=1, num_classes=4, start_neurons=8, pool_layers=2, conv_layers=2)
for n, h, w in ((4, 32, 32), (8, 64, 64), (8, 128, 128)):
    weights = synthetic_weights(seed=3, random_bn_stats=True, **cfg)
    imgs, labs = synthetic_batch(40, n, h, w)
    names = [nm for nm, _ in unet_param_specs(**cfg)]
    ora = OracleUNet(weights, **cfg)
    loss_ref, grads_ref, _, _ = ora.loss_and_grads(imgs, labs, [0.5, 1, 2, 1])
    out = []
    for prec in ("fp32", "bf16"):
        eng = UNetEngine(precision=prec, **cfg)
        eng.set_weights(weights)
        eng.train_begin([0.5, 1, 2, 1], dropout_rate=0.0, global_batch=n)
        loss = eng.train_step(imgs, labs)
        g = eng.get_grads()
        errs = []
        for nm, a, r in zip(names, g, grads_ref):
            if r is None or (nm.endswith("bias:0") and nm != names[-1]):
                continue
            r = r.numpy()
            errs.append((nm, float(np.abs(a - r).max() / max(np.abs(r).max(), 1e-12))))
        out.append((prec, loss, errs))
        eng.close()
    print(f"== {n}x{h}x{w} loss ref {loss_ref:.5f} fp32 {out[0][1]:.5f} bf16 {out[1][1]:.5f}")
    for (nm, e32), (_, e16) in zip(out[0][2], out[1][2]):
        print(f"  {nm:36s} fp32 {e32:.2e}  bf16 {e16:.2e}")
