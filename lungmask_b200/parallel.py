"""Slice-axis data parallelism over the GPUs of one box (one process per GPU).

The reference has no multi-GPU path (single device, lungmask/mask.py:118-121).  Everything per slice -
preprocess, forward, reshape (utils.py:48-51, mask.py:173-187,196-202) - shards over contiguous slice
ranges with no communication; the 3-D post-processing (utils.py:293-358) needs the whole label volume, so
the uint8 argmax volume (S x 256 x 256 bytes) and the crop boxes are all-gathered once per volume and
post-processing + reshape run replicated on every rank.

Two implementations of that gather:
  * `connect` + `apply_sharded_device` - the engine's own device-side gather: every rank's engine owns a gather block,
    the peers map it through CUDA IPC (the 64-byte handles are exchanged here over torch.distributed - host
    plumbing), and the engine pushes its slab into all peers' blocks over NVLink (csrc/shard.cu).  The volume never
    leaves the device between the stages.
  * `apply_sharded` - stage-level C-ABI calls with numpy arrays in between and `all_gather_into_tensor`
    (NCCL on GPUs, gloo in the CPU tests): the reference implementation the device path is checked against.
"""
import numpy as np


def shard_range(S: int, rank: int, world: int):
    """Contiguous slice range [lo, hi) of `rank`; ranges have ceil(S/world) slices except the tail.
    (The engine computes the same split: csrc/engine.cu shard_range.)"""
    per = -(-S // world)
    lo = min(S, rank * per)
    return lo, min(S, lo + per)


def connect(engine, rank: int, world: int, max_slices: int, group=None):
    """Collective: allocates every rank's gather block and maps the peers' blocks.  `engine` is this rank's
    lungmask_b200._native.Engine.  Needs an initialised torch.distributed process group when world > 1."""
    engine.shard_init(rank, world, max_slices)
    if world == 1:
        return engine
    import torch.distributed as dist
    handles = [None] * world
    dist.all_gather_object(handles, engine.shard_export(), group=group)   # 64 bytes per rank, host side
    engine.shard_connect(handles)
    dist.barrier(group=group)   # nobody pushes before every rank has mapped every block
    return engine


def apply_sharded_device(engine, slot: int, volume: np.ndarray, postprocess: bool = True, want_output: bool = True):
    """One volume over the ranks of a `connect`ed engine group (collective call).  Every rank passes the same volume,
    copies only its slab to its GPU, and gets the whole (S,H,W) uint8 result (None with want_output=False)."""
    return engine.apply_volume_sharded(slot, volume, postprocess=postprocess, want_output=want_output)


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
    """One volume over `world` ranks through the stage-level entry points: returns the full (S,H,W) uint8 result on
    every rank, identical to the single-GPU result (slices are independent up to the post-processing, which runs on
    the gathered volume)."""
    if volume.dtype != np.int16:
        raise TypeError("apply_sharded expects an int16 HU volume (see lungmask_b200.mask._to_int16_volume)")
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
