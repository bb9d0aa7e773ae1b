"""Multi-GPU training parity (run under torchrun on N GPUs of one box):
all-reduced gradients of the data-parallel step == sum over shards of the oracle's gradients
(each shard with its own BN statistics, loss scaled by the GLOBAL batch), and every rank ends
with identical weights.  Prints PASS/FAIL on rank 0."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.unet_oracle import OracleUNet  # noqa: E402
from oct_image_segmentation_models_b200 import parallel  # noqa: E402
from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch, synthetic_weights  # noqa: E402
from oct_image_segmentation_models_b200.engine import UNetEngine  # noqa: E402
from oct_image_segmentation_models_b200.models.unet_spec import unet_param_specs  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = dict(input_channels=1, num_classes=4, start_neurons=8, pool_layers=2, conv_layers=2)
    G, H, W = 4 * world, 32, 32
    cw = [0.5, 1.0, 2.0, 1.0]
    weights = synthetic_weights(seed=5, **cfg)
    imgs, labs = synthetic_batch(300, G, H, W)
    per = parallel.split_global_batch(G, world)
    s, e = parallel.shard_range(G, rank, world)
    assert e - s == per
    eng = UNetEngine(precision="fp32", device=local, **cfg)
    eng.set_weights(weights)
    eng.train_begin(cw, dropout_rate=0.0, global_batch=G)
    parallel.init_training_comm(eng)
    loss_local = eng.train_step(imgs[s:e], labs[s:e])
    grads = eng.get_grads()
    # oracle: sum of per-shard gradients, each shard normalised by the GLOBAL pixel count
    names = [n for n, _ in unet_param_specs(**cfg)]
    total = None
    loss_ref = 0.0
    for r in range(world):
        a, b = parallel.shard_range(G, r, world)
        l, g, _, _ = OracleUNet(weights, **cfg).loss_and_grads(imgs[a:b], labs[a:b], cw, loss_scale_pixels=G * H * W)
        loss_ref += l
        total = g if total is None else [None if x is None else x + y for x, y in zip(total, g)]
    worst = 0.0
    for nm, a, r in zip(names, grads, total):
        if r is None or (nm.endswith("bias:0") and nm != names[-1]):
            continue
        r = r.numpy()
        worst = max(worst, float(np.abs(a - r).max() / max(np.abs(r).max(), 1e-12)))
    lt = torch.tensor([loss_local], device="cuda", dtype=torch.float64)
    dist.all_reduce(lt)
    # identical weights on every rank after the step
    flat = torch.from_numpy(np.concatenate([w.ravel() for w, nm in zip(eng.get_weights(), names) if "moving" not in nm])).cuda()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(flat, ref))
    flags = torch.tensor([1.0 if same else 0.0, worst], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    worst_t = torch.tensor([worst], device="cuda")
    dist.all_reduce(worst_t, op=dist.ReduceOp.MAX)
    # four more steps: from the second identical call the step (NCCL all-reduce included) is replayed from a
    # captured CUDA graph; ranks must stay bit-identical and the summed loss must fall
    losses = []
    for _ in range(4):
        t = torch.tensor([eng.train_step(imgs[s:e], labs[s:e])], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        losses.append(t.item())
    flat = torch.from_numpy(np.concatenate([w.ravel() for w, nm in zip(eng.get_weights(), names) if "moving" not in nm])).cuda()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    g_same = torch.tensor([1.0 if torch.equal(flat, ref) else 0.0], device="cuda")
    dist.all_reduce(g_same, op=dist.ReduceOp.MIN)
    graph_ok = g_same.item() == 1.0 and losses[-1] < lt.item()
    if rank == 0:
        print(f"graph-replayed steps: losses {[round(x, 5) for x in losses]} ranks identical {g_same.item() == 1.0}")
        ok = graph_ok and flags[0].item() == 1.0 and worst_t.item() < 1e-2 and abs(lt.item() - loss_ref) < 1e-4 * max(1, abs(loss_ref))
        print(f"world {world}: worst rel grad err {worst_t.item():.2e}, loss sum {lt.item():.6f} vs oracle {loss_ref:.6f}, "
              f"weights identical across ranks {flags[0].item() == 1.0} -> {'PASS' if ok else 'FAIL'}")
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
