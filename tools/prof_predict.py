"""Nothing but K device-resident predict steps -- the command profiled for the launch lists and ncu --set full
captures under profiles/ (see profiles/README.md).
usage: python tools/prof_predict.py <bf16|fp16|fp32> [steps=3] [default|wide]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oct_image_segmentation_models_b200 import _native as nat  # noqa: E402
from oct_image_segmentation_models_b200.common.synthetic import fast_random_batch, synthetic_weights  # noqa: E402
from oct_image_segmentation_models_b200.engine import UNetEngine  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wide = len(sys.argv) > 3 and sys.argv[3] == "wide"
cfg = dict(input_channels=1, num_classes=4, start_neurons=64) if wide else dict(input_channels=1, num_classes=4)
n, h, w = (8, 1024, 512) if wide else (64, 512, 512)
eng = UNetEngine(precision=prec, **cfg)
eng.set_weights(synthetic_weights(seed=4 if wide else 42, **cfg))
x = torch.from_numpy(fast_random_batch(1, n, h, w)).cuda()
out = torch.empty((n, h, w, 4), dtype=torch.float32, device="cuda")
for _ in range(steps):
    eng.predict_device(x.data_ptr(), nat.U8, n, h, w, out.data_ptr(), None, None)
eng.synchronize()
print("launches", eng.launch_count(), "checksum", float(out[0, :2, :2].sum()))
eng.close()
