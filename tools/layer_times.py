"""Per-layer timing of the tcgen05 conv at BASELINE cfg2 shapes with pipeline wait counters
(OCTSEG_TC_DEBUG=1).  Run under gpurun."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("OCTSEG_TC_DEBUG", "1")
from oct_image_segmentation_models_b200.engine import UNetEngine  # noqa: E402
from oct_image_segmentation_models_b200.models.unet_spec import unet_blocks  # noqa: E402
from oct_image_segmentation_models_b200.common.synthetic import synthetic_weights  # noqa: E402

cfg = dict(input_channels=1, num_classes=4)
if len(sys.argv) > 1 and sys.argv[1] == "wide":
    cfg["start_neurons"] = 64
eng = UNetEngine(precision="bf16", **cfg)
eng.set_weights(synthetic_weights(seed=1, **cfg))
blocks = unet_blocks(**cfg)
N, H, W = (64, 512, 512) if cfg.get("start_neurons", 8) == 8 else (8, 1024, 512)
if os.environ.get("OCTSEG_LT_SHAPE"):            # e.g. OCTSEG_LT_SHAPE=32,512,256: the per-GPU batch of the 8-GPU training run
    N, H, W = (int(v) for v in os.environ["OCTSEG_LT_SHAPE"].split(","))
rng = np.random.default_rng(0)
sel = [int(a) for a in sys.argv[2:]] if len(sys.argv) > 2 else list(range(1, len(blocks) - 1))
for idx in sel:
    b = blocks[idx]
    lh, lw = H >> b.level, W >> b.level
    ih, iw = (lh // 2, lw // 2) if b.upsample_before else (lh, lw)
    n = N
    x = rng.random((n, ih, iw, b.cin), dtype=np.float32)
    _, ms = eng.debug_conv_block(idx, x, path=1, timed=True)
    byt = n * (ih * iw * b.cin + lh * lw * b.cout) * 2
    flops = 2.0 * n * lh * lw * b.cin * b.cout * b.kh * b.kw
    print(f"block {idx:2d} {b.role:4s} {b.cin:4d}->{b.cout:4d} k{b.kh} in {ih}x{iw}: {ms:.3f} ms  {byt / ms / 1e6:7.0f} GB/s  "
          f"{flops / ms / 1e9:7.1f} TFLOP/s", flush=True)
