"""First-light check of the fp32-accurate tensor-core path (error-compensated fp16 pairs): per-block error against
a float64 block, whole-net error against the CPU oracle and the FFMA path, and timing at BASELINE cfg2.
Run under gpurun:  python tools/split_check.py"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oct_image_segmentation_models_b200.common.synthetic import fast_random_batch, synthetic_batch, synthetic_weights  # noqa: E402
from oct_image_segmentation_models_b200.engine import UNetEngine  # noqa: E402
from oct_image_segmentation_models_b200.models.unet_spec import unet_blocks  # noqa: E402
from oracle.unet_oracle import OracleUNet  # noqa: E402

CFG = dict(input_channels=1, num_classes=4)


def main():
    from test_gpu_parity import oracle_block
    w = synthetic_weights(seed=42, **CFG)
    e32 = UNetEngine(precision="fp32", **CFG)
    e32.set_weights(w)
    blocks = unet_blocks(**CFG)
    rng = np.random.default_rng(0)
    for idx in (1, 3, 5, 8, 9, 10, 11, 13, 19, 20, 21):
        b = blocks[idx]
        x = np.maximum(rng.normal(0.3, 1.0, size=(2, 32, 24, b.cin)), 0).astype(np.float32)
        ref = oracle_block(b, w, x)
        try:
            got = e32.debug_conv_block(idx, x, path=1)
            cc = e32.debug_conv_block(idx, x, path=0)
            print(f"block {idx:2d} k{b.kh} {b.cin:3d}->{b.cout:3d} ups {int(b.upsample_before)}: split-TC err {np.abs(got - ref).max():.3e}  "
                  f"FFMA err {np.abs(cc - ref).max():.3e}  (scale {np.abs(ref).max():.2f})", flush=True)
        except Exception as ex:  # noqa: BLE001
            print(f"block {idx}: FAILED {ex}", flush=True)
    imgs, _ = synthetic_batch(10, 2, 256, 256)
    ref = OracleUNet(w, **CFG).predict(imgs)
    p, l = e32.predict(imgs, want_labels=True)
    rel = np.abs(p - ref) / np.maximum(ref, 1e-3)
    print(f"whole net 256x256 split-TC: uses TC {e32.layer_uses_tensor_core(5, 256, 256)}  max rel {rel.max():.3e}  "
          f"argmax agree {(l == ref.argmax(-1)).mean():.6f}", flush=True)
    os.environ["OCTSEG_FP32_PATH"] = "cuda"
    ecc = UNetEngine(precision="fp32", **CFG)
    ecc.set_weights(w)
    pc, lc = ecc.predict(imgs, want_labels=True)
    relc = np.abs(pc - ref) / np.maximum(ref, 1e-3)
    print(f"whole net 256x256 FFMA    : max rel {relc.max():.3e}  argmax agree {(lc == ref.argmax(-1)).mean():.6f}; "
          f"TC vs FFMA max rel {(np.abs(p - pc) / np.maximum(pc, 1e-3)).max():.3e}", flush=True)
    del os.environ["OCTSEG_FP32_PATH"]
    # timing, cfg2
    import torch
    from oct_image_segmentation_models_b200 import _native as nat
    n = 64
    x = torch.from_numpy(fast_random_batch(1, n, 512, 512)).cuda()
    out = torch.empty((n, 512, 512, 4), dtype=torch.float32, device="cuda")
    for name, eng in (("split-TC", e32), ("FFMA", ecc)):
        for _ in range(3):
            eng.predict_device(x.data_ptr(), nat.U8, n, 512, 512, out.data_ptr(), None, None)
        eng.synchronize()
        steps = 20 if name == "split-TC" else 3
        t0 = time.perf_counter()
        for _ in range(steps):
            eng.predict_device(x.data_ptr(), nat.U8, n, 512, 512, out.data_ptr(), None, None)
        eng.synchronize()
        dt = (time.perf_counter() - t0) / steps
        print(f"cfg2 fp32 {name}: {dt * 1e3:.3f} ms / 64 B-scans = {n / dt:.0f} B-scans/s", flush=True)
    e32.set_profiling(True)
    e32.predict_device(x.data_ptr(), nat.U8, n, 512, 512, out.data_ptr(), None, None)
    e32.synchronize()
    bt = e32.block_times_ms()
    print("block ms:", " ".join(f"{i}:{t:.3f}" for i, t in enumerate(bt)), " sum", sum(bt))
    e32.close()
    ecc.close()


if __name__ == "__main__":
    main()
