"""Slice-axis data parallelism over the GPUs of one box (one process per GPU, torch.distributed).

The reference has no multi-GPU path (single device, lungmask/mask.py:118-121).  Everything per slice -
preprocess, forward, reshape (utils.py:48-51, mask.py:173-187,196-202) - shards over contiguous slice
ranges with no communication; the 3-D post-processing (utils.py:293-358) needs the whole label volume, so
the uint8 argmax volume (S x 256 x 256 bytes) is all-gathered once (NCCL over NVLink on GPUs, gloo in the
CPU tests) and post-processing + reshape run replicated on every rank.
"""
import numpy as np


def shard_range(S: int, rank: int, world: int):
    """Contiguous slice range [lo, hi) of `rank`; ranges have ceil(S/world) slices except the tail."""
    per = -(-S // world)
    lo = min(S, rank * per)
    return lo, min(S, lo + per)


def all_gather_slices(local: np.ndarray, S: int, rank: int, world: int, group=None, device=None) -> np.ndarray:
    """All-gathers per-rank slabs (first axis = this rank's slices) into the full (S, ...) array."""
    import torch
    import torch.distributed as dist
    per = -(-S // world)
    pad = np.zeros((per,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy()[:S]


def apply_sharded(engine, slot: int, volume: np.ndarray, rank: int, world: int, group=None, device=None,
                  postprocess: bool = True) -> np.ndarray:
    """One volume over `world` ranks: returns the full (S,H,W) uint8 result on every rank, identical to the
    single-GPU result (slices are independent up to the post-processing, which runs on the gathered volume)."""
    S, H, W = volume.shape
    lo, hi = shard_range(S, rank, world)
    if hi > lo:
        resized, boxes = engine.preprocess(volume[lo:hi])
        labels = engine.forward(slot, resized)
    else:
        labels = np.zeros((0, 256, 256), np.uint8)
        boxes = np.zeros((0, 4), np.int32)
    if world > 1:
        labels = all_gather_slices(labels, S, rank, world, group, device)
        boxes = all_gather_slices(boxes.astype(np.int32), S, rank, world, group, device)
    masks = engine.postprocess(labels) if postprocess else labels
    return engine.reshape_masks(masks, boxes, H, W)
