"""One-process-per-GPU plumbing (torch.distributed is used only for rendezvous / host-side
exchange; the data path has no collective for inference and exactly one -- the gradient
all-reduce inside liboctseg -- for training; reference: tf.distribute.MirroredStrategy,
training/training.py:185)."""
from typing import List, Optional, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n_items over `world` ranks; the first n_items % world ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def split_global_batch(global_batch: int, world: int) -> int:
    """Per-replica batch under synchronous data parallelism (Keras splits the global batch evenly)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by {world} replicas")
    return global_batch // world


def init_training_comm(engine, dist=None):
    """Create the NCCL communicator inside the library: rank 0 makes the unique id, the
    initialised torch.distributed group (any backend) broadcasts it."""
    if dist is None:
        import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        engine.comm_init(bytes(128), 0, 1)
        return
    rank, world = dist.get_rank(), dist.get_world_size()
    box: List[Optional[bytes]] = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    engine.comm_init(box[0], rank, world)


def predict_sharded(engine, images: np.ndarray, rank: int, world: int, want_labels: bool = True):
    """Each rank predicts its contiguous shard of B-scans (no collective).  Returns
    (start, stop, probs, labels) for this rank; callers that need everything on one host gather
    the label maps (16x smaller than the probabilities) with `gather_in_order`."""
    start, stop = shard_range(len(images), rank, world)
    if stop == start:
        return start, stop, None, None
    probs, labels = engine.predict(images[start:stop], want_probs=True, want_labels=want_labels)
    return start, stop, probs, labels


def gather_in_order(local: Optional[np.ndarray], dist=None) -> Optional[List[np.ndarray]]:
    """all_gather_object of per-rank arrays, concatenated in rank (= original image) order."""
    if dist is None:
        import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None if local is None else local
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    parts = [p for p in out if p is not None]
    return np.concatenate(parts, axis=0) if parts else None


def allreduce_sum_scalar(value: float, dist=None) -> float:
    """Sum of a host scalar over the ranks (loss / metric logging; the reference reduces its logged values
    across replicas the same way).  Backend-agnostic (object gather), not on the data path."""
    if dist is None:
        import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, float(value))
    return float(sum(out))


def mean_over_ranks(arrays: List[np.ndarray], dist=None) -> List[np.ndarray]:
    """Element-wise mean over the ranks of a list of small host arrays (every rank gets the result)."""
    if dist is None:
        import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [np.asarray(a) for a in arrays]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, [np.asarray(a, dtype=np.float64) for a in arrays])
    world = len(out)
    return [(sum(o[i] for o in out) / world).astype(np.asarray(arrays[i]).dtype) for i in range(len(arrays))]


def sync_bn_moving_stats(model, names: List[str], dist=None) -> None:
    """Per-replica BatchNorm moving statistics (every replica normalises with the statistics of its own
    shard, as under tf.distribute.MirroredStrategy) are averaged over the replicas before validation and
    checkpointing, so that every rank evaluates and saves the same model.  `model` needs get_weights() /
    set_weights() in Keras order; `names` are the Keras weight names in the same order."""
    if dist is None:
        import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    weights = model.get_weights()
    idx = [i for i, nm in enumerate(names) if "moving_mean" in nm or "moving_variance" in nm]
    if not idx:
        return
    merged = mean_over_ranks([weights[i] for i in idx], dist)
    for i, m in zip(idx, merged):
        weights[i] = m
    model.set_weights(weights)
