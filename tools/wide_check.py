"""Wide U-Net (start_neurons 64, BASELINE configs[3]) parity at a small size + whole-net timing at 1024x512."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.unet_oracle import OracleUNet
from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch, synthetic_weights, fast_random_batch
from oct_image_segmentation_models_b200.engine import UNetEngine
from oct_image_segmentation_models_b200.models.unet_spec import unet_blocks
import torch

cfg = dict(input_channels=1, num_classes=4, start_neurons=64)
w = synthetic_weights(seed=4, **cfg)
imgs, _ = synthetic_batch(0, 1, 64, 64)
ref = OracleUNet(w, **cfg).predict(imgs)
eng = UNetEngine(precision="bf16", **cfg)
eng.set_weights(w)
p, l = eng.predict(imgs, want_labels=True)
rel = np.abs(p - ref) / np.maximum(ref, 1e-3)
print("wide parity 64x64: max rel", rel.max(), "argmax agree", (l == ref.argmax(-1)).mean(),
      "tc layers", [i for i in range(23) if eng.layer_uses_tensor_core(i, 1024, 512)])
n, H, W = 8, 1024, 512
x = torch.from_numpy(fast_random_batch(3, n, H, W)).cuda()
out = torch.empty((n, H, W, 4), dtype=torch.float32, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for _ in range(3):
    eng.predict_device(x.data_ptr(), 0, n, H, W, out.data_ptr(), None, st.cuda_stream)
torch.cuda.synchronize()
eng.set_profiling(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.predict_device(x.data_ptr(), 0, n, H, W, out.data_ptr(), None, st.cuda_stream)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
flop = 872.751e9 * n
print(f"wide net 8x1024x512: {ms:.2f} ms/step  {n / ms * 1e3:.1f} B-scans/s  {flop / ms / 1e9:.1f} TFLOP/s")
bt = eng.block_times_ms()
blocks = unet_blocks(**cfg)
for b, t in zip(blocks, bt):
    lh, lw = H >> b.level, W >> b.level
    fl = 2.0 * n * lh * lw * b.cin * b.cout * b.kh * b.kw
    print(f"  block {b.index:2d} {b.role:4s} {b.cin:4d}->{b.cout:4d} k{b.kh} @{lh}x{lw}: {t:.3f} ms  {fl / t / 1e9:7.1f} TFLOP/s")
