"""Device-resident timing of BASELINE configs[1] (64 x 512x512 predict) in the given precisions.
usage: python tools/predict_bench.py [bf16 fp32 ...] ; OCTSEG_LIB selects an alternative liboctseg build."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oct_image_segmentation_models_b200 import _native as nat  # noqa: E402
from oct_image_segmentation_models_b200.common.synthetic import fast_random_batch, synthetic_weights  # noqa: E402
from oct_image_segmentation_models_b200.engine import UNetEngine  # noqa: E402

cfg = dict(input_channels=1, num_classes=4)
n, h, w = 64, 512, 512
x = torch.from_numpy(fast_random_batch(1, n, h, w)).cuda()
out = torch.empty((n, h, w, 4), dtype=torch.float32, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
for prec in (sys.argv[1:] or ["bf16", "fp32"]):
    eng = UNetEngine(precision=prec, **cfg)
    eng.set_weights(synthetic_weights(seed=42, **cfg))
    for _ in range(5):
        eng.predict_device(x.data_ptr(), nat.U8, n, h, w, out.data_ptr(), None, st.cuda_stream)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(30):
            eng.predict_device(x.data_ptr(), nat.U8, n, h, w, out.data_ptr(), None, st.cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 30)
    print(f"{prec}: {best:.4f} ms/step = {n / best * 1e3:.0f} B-scans/s  checksum {float(out[0, :2, :2].sum()):.4f}")
    eng.close()
