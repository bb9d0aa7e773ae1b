"""Randomised shape / config stress of the tensor-core predict path: bf16 and fp16 engines against the fp32
engine (itself checked against the oracle to 1e-4 in tests/) on many random (config, batch, height, width)
draws, including ragged tiles, odd widths and every epilogue specialisation; also predict_maps vs predict
labels and chunked vs unchunked host pipeline.  usage: python tools/stress_parity.py [draws=60] [seed=0]"""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oct_image_segmentation_models_b200.common.synthetic import synthetic_weights  # noqa: E402
from oct_image_segmentation_models_b200.engine import UNetEngine  # noqa: E402

draws = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = 0.0
worst_split, n_split = 0.0, 0
for d in range(draws):
    P = int(rng.integers(1, 5))
    cfg = dict(input_channels=1, num_classes=int(rng.integers(2, 9)), start_neurons=int(rng.choice([8, 8, 16, 32])),
               pool_layers=P, conv_layers=int(rng.integers(1, 4)))
    q = 1 << P
    h = int(rng.integers(1, 1 + 224 // q)) * q
    w = int(rng.integers(1, 1 + 256 // q)) * q
    n = int(rng.integers(1, 6))
    weights = synthetic_weights(seed=int(rng.integers(1 << 30)), **cfg)
    imgs = rng.integers(0, 256, size=(n, h, w, 1), dtype=np.uint8)
    e32 = UNetEngine(precision="fp32", **cfg); e32.set_weights(weights)
    ref, lab_ref = e32.predict(imgs, want_labels=True)
    on_tc = e32.layer_uses_tensor_core(1, h, w)
    e32.close()
    if on_tc:
        # the fp32 mode has two implementations: tcgen05 on fp16 (hi, lo') pairs (just used) and FFMA; same contract
        os.environ["OCTSEG_FP32_PATH"] = "cuda"
        e32c = UNetEngine(precision="fp32", **cfg); e32c.set_weights(weights)
        ref_cc, lab_cc = e32c.predict(imgs, want_labels=True)
        e32c.close()
        del os.environ["OCTSEG_FP32_PATH"]
        rel = float((np.abs(ref - ref_cc) / np.maximum(ref_cc, 1e-3)).max())
        n_split += 1
        worst_split = max(worst_split, rel)
        if rel > 5e-5 or (lab_ref != lab_cc).mean() > 1e-4 or not np.isfinite(ref).all():
            print("FAIL fp32 tcgen05 vs FFMA", d, cfg, (n, h, w), "max rel", rel, "label mismatch", float((lab_ref != lab_cc).mean()))
            sys.exit(1)
    for prec in ("bf16", "fp16"):
        os.environ["OCTSEG_MICROBATCH"] = str(int(rng.integers(1, n + 1)))
        e = UNetEngine(precision=prec, **cfg); e.set_weights(weights)
        p, lab = e.predict(imgs, want_labels=True)
        lab2, maps = e.predict_maps(imgs)
        e.close()
        err = float(np.abs(p - ref).max())
        worst = max(worst, err)
        agree = float((lab == lab_ref).mean())
        ok = err <= 4e-2 and np.array_equal(lab, p.argmax(-1)) and np.array_equal(lab, lab2) and np.isfinite(p).all()
        if not ok or agree < 0.97:
            print("FAIL", d, cfg, (n, h, w), prec, "max abs err", err, "argmax agreement", agree)
            sys.exit(1)
    print(f"draw {d:3d} {cfg} {(n, h, w)} ok", flush=True)
print("ALL OK, worst |p - p_fp32| =", worst, "; fp32 tcgen05 vs FFMA on", n_split, "draws, worst relative difference", worst_split)
