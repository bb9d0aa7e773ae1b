"""GPU first-light diagnostics (run under gpurun).  Prints per-layer parity of the CUDA-core
and tcgen05 conv paths against the CPU oracle, then whole-net parity in fp32 and bf16.
Not a test: tests/ holds the pass/fail versions.  Writes gpurun_out/first_light.log."""
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oct_image_segmentation_models_b200.engine import UNetEngine  # noqa: E402
from oct_image_segmentation_models_b200.models.unet_spec import unet_blocks  # noqa: E402
from oct_image_segmentation_models_b200.common.synthetic import synthetic_weights, synthetic_batch  # noqa: E402
from oracle.unet_oracle import OracleUNet, BN_EPS  # noqa: E402


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def oracle_block(b, weights, x_nhwc, quant_w=False):
    i = 6 * b.index
    w, bias, gamma, beta, mean, var = [torch.from_numpy(np.asarray(t, dtype=np.float32)).double() for t in weights[i:i + 6]]
    if quant_w:
        w = torch.from_numpy(bf16_round(weights[i])).double()
    x = torch.from_numpy(x_nhwc).double().permute(0, 3, 1, 2)
    if b.upsample_before:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    kh, kw = b.kh, b.kw
    pt, pl = (kh - 1) // 2, (kw - 1) // 2
    x = F.pad(x, (pl, kw - 1 - pl, pt, kh - 1 - pt))
    z = F.conv2d(x, w.permute(3, 2, 0, 1), bias)
    inv = torch.rsqrt(var + BN_EPS) * gamma
    y = torch.relu((z - mean[None, :, None, None]) * inv[None, :, None, None] + beta[None, :, None, None])
    return y.permute(0, 2, 3, 1).numpy()


def describe_err(name, got, ref):
    err = np.abs(got - ref)
    scale = np.abs(ref).max() + 1e-12
    bad = ~np.isfinite(got)
    print(f"  {name}: max|err| {np.nanmax(err):.4e}  rel-to-max {np.nanmax(err) / scale:.3e}  "
          f"mean|err| {np.nanmean(err):.3e}  nonfinite {int(bad.sum())}  ref max {scale:.3f}")
    return float(np.nanmax(err) / scale), int(bad.sum())


def err_pattern(got, ref):
    """Where are the big errors? (rows / cols / channels histogram)"""
    err = np.abs(np.nan_to_num(got, nan=1e9) - ref)
    thr = 0.05 * (np.abs(ref).max() + 1e-12)
    bad = err > thr
    if not bad.any():
        return
    n, h, w, c = got.shape
    print(f"    bad fraction {bad.mean():.4f}; by image {bad.mean(axis=(1, 2, 3)).round(3)}")
    print(f"    by row (first 24)  {bad.mean(axis=(0, 2, 3))[:24].round(2)}")
    print(f"    by col (first 24)  {bad.mean(axis=(0, 1, 3))[:24].round(2)}")
    print(f"    by chan (first 16) {bad.mean(axis=(0, 1, 2))[:16].round(2)}")
    idx = np.argwhere(bad)[:5]
    for i in idx:
        print(f"    e.g. {tuple(i)} got {got[tuple(i)]:.4f} ref {ref[tuple(i)]:.4f}")


def main():
    cfg = dict(input_channels=1, num_classes=4)
    weights = synthetic_weights(seed=42, **cfg)
    blocks = unet_blocks(**cfg)
    rng = np.random.default_rng(7)
    t0 = time.time()
    eng = UNetEngine(precision="bf16", device=0, **cfg)
    eng.set_weights(weights)
    eng32 = UNetEngine(precision="fp32", device=0, **cfg)
    eng32.set_weights(weights)
    print(f"engines up in {time.time() - t0:.1f}s")

    sizes = {1: (2, 32, 24), 2: (2, 32, 24), 3: (2, 32, 24), 5: (2, 32, 16), 7: (2, 16, 16), 9: (1, 16, 16),
             10: (2, 16, 8), 11: (2, 32, 16), 12: (2, 16, 16), 13: (2, 16, 16), 16: (2, 16, 16),
             19: (2, 32, 24), 20: (2, 32, 24), 21: (2, 48, 40)}
    summary = []
    for idx, (n, h, w) in sizes.items():
        b = blocks[idx]
        x = np.maximum(rng.normal(0.3, 1.0, size=(n, h, w, b.cin)), 0).astype(np.float32)
        xq = bf16_round(x)
        ref = oracle_block(b, weights, xq)
        refq = oracle_block(b, weights, xq, quant_w=True)
        print(f"block {idx} {b.role} k{b.kh}x{b.kw} {b.cin}->{b.cout} ups={b.upsample_before} on {n}x{h}x{w}")
        try:
            d32 = eng32.debug_conv_block(idx, x, path=0)
            describe_err("fp32 direct vs oracle(fp32 x)", d32, oracle_block(b, weights, x))
        except Exception as e:  # noqa: BLE001
            print("  fp32 direct FAILED:", e)
        try:
            d = eng.debug_conv_block(idx, x, path=0)
            describe_err("bf16 direct", d, ref)
        except Exception as e:  # noqa: BLE001
            print("  bf16 direct FAILED:", e)
        try:
            t = eng.debug_conv_block(idx, x, path=1)
            r, nf = describe_err("bf16 tcgen05 (vs bf16-weight oracle)", t, refq)
            summary.append((idx, r, nf))
            if r > 0.02 or nf:
                err_pattern(t, refq)
        except Exception as e:  # noqa: BLE001
            print("  tcgen05 FAILED:", e)
            summary.append((idx, -1.0, -1))
        sys.stdout.flush()
    print("tcgen05 summary (block, rel err, nonfinite):", summary)

    # whole net
    imgs, _ = synthetic_batch(0, 2, 64, 64)
    ref = OracleUNet(weights, **cfg).predict(imgs)
    for name, e in (("fp32", eng32), ("bf16", eng)):
        try:
            probs, labels = e.predict(imgs, want_labels=True)
            rel = np.abs(probs - ref) / np.maximum(ref, 1e-3)
            print(f"whole net {name}: max rel {rel.max():.3e} max abs {np.abs(probs - ref).max():.3e} "
                  f"argmax agree {(labels == ref.argmax(-1)).mean():.5f} "
                  f"tc layers {[i for i in range(len(blocks)) if e.layer_uses_tensor_core(i, 64, 64)]}")
        except Exception as ex:  # noqa: BLE001
            print(f"whole net {name} FAILED:", ex)
    # timing of big layers, tcgen05 vs direct
    for idx, (n, h, w) in {1: (16, 512, 512), 3: (16, 256, 256), 7: (16, 64, 64), 20: (16, 512, 512),
                           19: (16, 256, 256)}.items():
        b = blocks[idx]
        x = np.maximum(rng.normal(0.3, 1.0, size=(n, h, w, b.cin)), 0).astype(np.float32)
        for path in (0, 1):
            try:
                _, ms = eng.debug_conv_block(idx, x, path=path, timed=True)
                oh, ow = (2 * h, 2 * w) if b.upsample_before else (h, w)
                byt = n * (h * w * b.cin + oh * ow * b.cout) * 2
                print(f"time block {idx} {b.cin}->{b.cout} {n}x{h}x{w} path {path}: {ms:.3f} ms  "
                      f"{byt / ms / 1e6:.0f} GB/s algorithmic")
            except Exception as ex:  # noqa: BLE001
                print(f"time block {idx} path {path} FAILED:", ex)
        sys.stdout.flush()


if __name__ == "__main__":
    main()
