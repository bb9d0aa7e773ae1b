"""Device-resident timing of the BASELINE configs[2] train step (default U-Net, 512x256x1, bf16) on one GPU.
usage: python tools/train_bench.py [per-GPU batch=256] [steps=10]; OCTSEG_TRAIN_PROFILE=1 adds the phase split."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oct_image_segmentation_models_b200 import _native as nat  # noqa: E402
from oct_image_segmentation_models_b200.common.synthetic import fast_random_batch, synthetic_weights  # noqa: E402
from oct_image_segmentation_models_b200.engine import UNetEngine  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = dict(input_channels=1, num_classes=4)
eng = UNetEngine(precision="bf16", **cfg)
eng.set_weights(synthetic_weights(seed=7, random_bn_stats=False, **cfg))
eng.train_begin([0.5, 1.0, 2.0, 1.0], learning_rate=1e-3, dropout_rate=0.5, dropout_seed=1234, global_batch=batch)
imgs = torch.from_numpy(fast_random_batch(77, batch, 512, 256)).cuda()
labs = torch.from_numpy(np.random.default_rng(5).integers(0, 4, size=(batch, 512, 256), dtype=np.uint8)).cuda()
loss = torch.zeros(1, dtype=torch.float32, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)


def step():
    eng.train_step_device(imgs.data_ptr(), nat.U8, labs.data_ptr(), batch, 512, 256, loss.data_ptr(), st.cuda_stream)


for _ in range(3):
    step()
torch.cuda.synchronize()
l0 = eng.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"batch {batch}: {ms:.3f} ms/step = {batch / ms * 1e3:.0f} samples/s, {(eng.launch_count() - l0) / steps:.0f} launches/step, "
      f"loss {loss.item():.4f}; HBM roofline (227 MB/sample, 6534 GB/s): {227e6 * batch / (ms * 1e-3) / 6534e9:.3f}")
eng.close()
