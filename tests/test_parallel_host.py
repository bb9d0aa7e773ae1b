"""CPU tests of the N>1 host logic with the gloo backend (world_size 2): sharding covers every
B-scan exactly once and in order; the data-parallel gradient (sum over shards, loss scaled by
the global batch, per-shard BN statistics) equals what the training step's all-reduce must
produce -- computed here with the oracle standing in for the GPU kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oct_image_segmentation_models_b200 import parallel


def test_shard_range_partitions_in_order():
    for n in (0, 1, 7, 64, 10_000):
        for world in (1, 2, 3, 8):
            cuts = [parallel.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for (a, b), (c, d) in zip(cuts, cuts[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_split_global_batch():
    assert parallel.split_global_batch(256, 8) == 32
    with pytest.raises(ValueError):
        parallel.split_global_batch(10, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle.unet_oracle import OracleUNet
    from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch, synthetic_weights
    cfg = dict(input_channels=1, num_classes=3, start_neurons=8, pool_layers=1, conv_layers=1)
    G, H, W = 4, 16, 16
    weights = synthetic_weights(seed=11, **cfg)
    imgs, labs = synthetic_batch(7, G, H, W, 3)
    a, b = parallel.shard_range(G, rank, world)
    # ---- inference sharding: gather label maps back in image order
    probs = OracleUNet(weights, **cfg).predict(imgs[a:b])
    labels = parallel.gather_in_order(probs.argmax(-1).astype(np.uint8), dist)
    # ---- training: local gradient scaled by the global batch, then SUM all-reduce
    loss, grads, _, _ = OracleUNet(weights, **cfg).loss_and_grads(imgs[a:b], labs[a:b], [1.0, 2.0, 0.5],
                                                                 loss_scale_pixels=G * H * W)
    flat = torch.cat([g.reshape(-1) for g in grads if g is not None])
    dist.all_reduce(flat)
    lt = torch.tensor([loss], dtype=torch.float64)
    dist.all_reduce(lt)
    if rank == 0:
        np.savez(os.path.join(out_dir, "r0.npz"), labels=labels, flat=flat.numpy(), loss=lt.numpy())
    dist.destroy_process_group()


def test_gloo_world2_sharded_predict_and_gradient_sum(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "r0.npz")
    from oracle.unet_oracle import OracleUNet
    from oct_image_segmentation_models_b200.common.synthetic import synthetic_batch, synthetic_weights
    cfg = dict(input_channels=1, num_classes=3, start_neurons=8, pool_layers=1, conv_layers=1)
    weights = synthetic_weights(seed=11, **cfg)
    imgs, labs = synthetic_batch(7, 4, 16, 16, 3)
    # unsharded inference == gathered sharded inference (no cross-image coupling at inference)
    full = OracleUNet(weights, **cfg).predict(imgs).argmax(-1)
    assert np.array_equal(got["labels"], full)
    # data-parallel gradient == sum of per-shard gradients (per-replica BN statistics)
    total, loss = None, 0.0
    for r in range(world):
        a, b = parallel.shard_range(4, r, world)
        l, g, _, _ = OracleUNet(weights, **cfg).loss_and_grads(imgs[a:b], labs[a:b], [1.0, 2.0, 0.5],
                                                             loss_scale_pixels=4 * 16 * 16)
        f = torch.cat([x.reshape(-1) for x in g if x is not None])
        total = f if total is None else total + f
        loss += l
    np.testing.assert_allclose(got["flat"], total.numpy(), rtol=1e-5, atol=1e-8)
    assert abs(float(got["loss"][0]) - loss) < 1e-6


class _FakeModel:
    def __init__(self, weights):
        self.w = [np.array(x) for x in weights]

    def get_weights(self):
        return [x.copy() for x in self.w]

    def set_weights(self, ws):
        self.w = [np.array(x) for x in ws]


def _worker_state(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names = ["conv2d/kernel:0", "batch_normalization/gamma:0", "batch_normalization/moving_mean:0",
             "batch_normalization/moving_variance:0"]
    model = _FakeModel([np.full((2, 2), 5.0, np.float32), np.full(3, 1.0 + rank, np.float32),
                        np.arange(3, dtype=np.float32) + 10 * rank, np.full(3, 2.0 * (rank + 1), np.float32)])
    parallel.sync_bn_moving_stats(model, names, dist)
    total = parallel.allreduce_sum_scalar(0.25 * (rank + 1), dist)
    np.savez(os.path.join(out_dir, f"s{rank}.npz"), *model.get_weights(), total=total)
    dist.destroy_process_group()


def test_gloo_world2_replica_state_helpers(tmp_path):
    """BN moving statistics are averaged over replicas, everything else is left alone (the gradient all-reduce
    already keeps trainable weights identical); logged scalars are summed; every rank sees the same result."""
    world = 2
    mp.spawn(_worker_state, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"s{k}.npz") for k in range(world)]
    for k in range(world):
        assert np.array_equal(r[k]["arr_0"], np.full((2, 2), 5.0, np.float32))
        assert np.array_equal(r[k]["arr_1"], np.full(3, 1.0 + k, np.float32))          # per-rank value untouched
        assert np.allclose(r[k]["arr_2"], np.arange(3) + 5.0)                          # mean of +0 and +10
        assert np.allclose(r[k]["arr_3"], 3.0)                                         # mean of 2 and 4
        assert abs(float(r[k]["total"]) - 0.75) < 1e-12
    # single process: both helpers are no-ops
    assert parallel.allreduce_sum_scalar(1.5) == 1.5
