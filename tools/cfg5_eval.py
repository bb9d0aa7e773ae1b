"""BASELINE configs[4] end to end (scaled by --n): train the default U-Net on synthetic B-scans with the
GPU trainer, predict N synthetic 512x512 B-scans (bf16 and fp32), argmax -> boundary maps (numpy, reference
semantics) -> min-path (native C++), and compare boundaries with the CPU oracle chain on a subset.
Run under gpurun:  python tools/cfg5_eval.py --n 2000 --oracle 48"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oct_image_segmentation_models_b200.common import utils  # noqa: E402
from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch, synthetic_weights  # noqa: E402
from oct_image_segmentation_models_b200.engine import UNetEngine  # noqa: E402
from oct_image_segmentation_models_b200.min_path_processing import graph_search  # noqa: E402


def boundaries(labels_u8, K=4, threads=16):
    """labels [n,H,W] -> uint16 [n,K-1,W] with the reference chain (one-hot -> maps -> min-path)."""
    n, H, W = labels_u8.shape
    out = np.zeros((n, K - 1, W), np.uint16)
    for i0 in range(0, n, 64):
        lab = labels_u8[i0:i0 + 64]
        cat = np.transpose(utils.to_categorical(lab, K), (0, 3, 1, 2))
        maps = utils.convert_predictions_to_maps_semantic(cat, bg_ilm=True, bg_csi=False)      # [m,K-1,H,W]
        maps_t = np.ascontiguousarray(np.transpose(maps, (0, 1, 3, 2))).reshape(-1, W, H)
        seg = graph_search.segment_maps(maps_t, None, None, n_threads=threads, return_prob_maps=False)[0]
        out[i0:i0 + len(lab)] = seg.reshape(len(lab), K - 1, W)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--oracle", type=int, default=32)
    ap.add_argument("--train-steps", type=int, default=400)
    ap.add_argument("--size", type=int, default=512)
    args = ap.parse_args()
    H = W = args.size
    cfg = dict(input_channels=1, num_classes=4)
    # ---- train on the GPU (bf16 tensor-core trainer), synthetic set
    t0 = time.time()
    tr_i, tr_l = synthetic_batch(100000, 128, H, W)
    eng = UNetEngine(precision="bf16", **cfg)
    eng.set_weights(synthetic_weights(seed=11, random_bn_stats=False, **cfg))
    eng.train_begin([0.5, 1.0, 2.0, 1.0], learning_rate=2e-3, dropout_rate=0.5, dropout_seed=3, global_batch=16)
    rng = np.random.default_rng(0)
    for s in range(args.train_steps):
        idx = rng.choice(len(tr_i), 16, replace=False)
        loss = eng.train_step(tr_i[idx], tr_l[idx])
        if s % 100 == 0:
            print(f"train step {s} loss {loss:.4f}", flush=True)
    weights = eng.get_weights()
    print(f"trained {args.train_steps} steps in {time.time() - t0:.1f}s, final loss {loss:.4f}", flush=True)
    # ---- evaluate
    imgs, labs = synthetic_batch(0, args.n, H, W)
    res = {}
    for prec in ("bf16", "fp16", "fp32"):
        e = eng if prec == "bf16" else UNetEngine(precision=prec, **cfg)
        if prec != "bf16":
            e.set_weights(weights)
        t1 = time.time()
        labels = np.empty((args.n, H, W), np.uint8)
        maps_t = np.empty((args.n, 3, W, H), np.uint8)
        for i0 in range(0, args.n, 64):
            l, m = e.predict_maps(imgs[i0:i0 + 64], transposed=True)     # argmax + boundary maps on the GPU
            labels[i0:i0 + len(l)] = l
            maps_t[i0:i0 + len(l)] = m
        t_pred = time.time() - t1
        t2 = time.time()
        segs = graph_search.segment_maps(maps_t.reshape(-1, W, H), None, None, n_threads=os.cpu_count() or 1,
                                         return_prob_maps=False)[0].reshape(args.n, 3, W)
        t_path = time.time() - t2
        if prec == "bf16":   # the numpy chain on the same labels gives the same boundaries (first 64 images)
            assert np.array_equal(segs[:64], boundaries(labels[:64], threads=os.cpu_count() or 1))
        res[prec] = dict(labels=labels, segs=segs, t_pred=t_pred, t_path=t_path,
                         acc=float((labels == labs[..., 0]).mean()))
        print(f"{prec}: predict {args.n / t_pred:.0f} B-scans/s (labels + boundary maps, host API), native min-path "
              f"{args.n / t_path:.0f} B-scans/s on {os.cpu_count()} cores, pixel acc vs truth {res[prec]['acc']:.4f}", flush=True)
        if prec != "bf16":
            e.close()
    for lo in ("bf16", "fp16"):
        d = np.abs(res[lo]["segs"].astype(np.int32) - res["fp32"]["segs"].astype(np.int32))
        agree_lab = float((res[lo]["labels"] == res["fp32"]["labels"]).mean())
        print(f"{lo} vs fp32 over all {args.n}: argmax agreement {agree_lab:.5f}, boundary positions identical "
              f"{float((d == 0).mean()):.5f}, max |delta| {int(d.max())} rows, B-scans with every boundary identical "
              f"{int((d.reshape(args.n, -1).max(1) == 0).sum())}")
    # ---- CPU oracle on a subset
    from oracle import postproc
    from oracle.unet_oracle import OracleUNet
    m = args.oracle
    ora = OracleUNet(weights, **cfg)
    t3 = time.time()
    probs = np.concatenate([ora.predict(imgs[i:i + 4]) for i in range(0, m, 4)])
    t_cpu = time.time() - t3
    ref_lab = probs.argmax(-1).astype(np.uint8)
    ref_segs = np.stack([postproc.boundaries_from_probs(probs[i:i + 1]) for i in range(m)])
    out = {"n": args.n, "oracle_subset": m, "cpu_oracle_bscans_per_s": m / t_cpu}
    for prec in ("fp32", "fp16", "bf16"):
        dd = np.abs(res[prec]["segs"][:m].astype(np.int32) - ref_segs.astype(np.int32))
        out[prec] = {"argmax_agreement_vs_oracle": float((res[prec]["labels"][:m] == ref_lab).mean()),
                     "boundaries_identical_frac": float((dd == 0).mean()), "max_row_delta": int(dd.max()),
                     "images_fully_identical": int((dd.reshape(m, -1).max(1) == 0).sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
