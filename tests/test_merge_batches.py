"""CPU emulation of the batch schedule of merge_loop_mc_kernel (lungmask_b200/csrc/postproc.cu) against the sequential
merge loop of utils.postprocessing (utils.py:310-339).

The device processes the area-ordered region list in batches: the longest prefix of upcoming candidates whose extents
are pairwise separated by a voxel (cut as well at a region that a record growth could turn into a candidate), decides
all members' targets on the SAME snapshot in parallel, then applies the area / record arithmetic in order; a merge that
lifts a skipped (< skip_below) region inside the batch's span over the threshold truncates the batch there.  This file
states that schedule in numpy and checks it against the plain sequential loop: identical region maps, areas and records
on random, speckled and blob label volumes, with spare labels and skip_below variants."""
import numpy as np
from scipy import ndimage

from oracle.standins import cc_label


def setup(lab, spare, skip):
    reg = cc_label(lab)
    R = reg.max()
    area = np.bincount(reg.ravel(), minlength=R + 1).astype(np.int64)
    value = np.zeros(R + 1, int)
    for r in range(1, R + 1):
        value[r] = lab[reg == r][0]
    order = sorted(range(1, R + 1), key=lambda r: (area[r], r))
    record = {}
    for r in order:
        if area[r] > record.get(value[r], 0):
            record[value[r]] = area[r]
    bbox = {}
    for r in range(1, R + 1):
        idx = np.argwhere(reg == r)
        bbox[r] = [idx[:, 0].min(), idx[:, 0].max() + 1, idx[:, 1].min(), idx[:, 1].max() + 1, idx[:, 2].min(), idx[:, 2].max() + 1]
    return reg, R, area, value, order, record, bbox

def decide(reg, r, spare):
    m = reg == r
    ring = ndimage.binary_dilation(m) & ~m
    ids, counts = np.unique(reg[ring], return_counts=True)
    target, best = r, 0
    for n, c in zip(ids, counts):
        if n != 0 and n != r and c > best and n not in spare:
            best, target = c, n
    return target

def apply(reg, area, value, record, bbox, r, t):
    moved = 0
    if t != r:
        moved = area[r]
        reg[reg == r] = t
        b, c = bbox[t], bbox[r]
        bbox[t] = [min(b[0], c[0]), max(b[1], c[1]), min(b[2], c[2]), max(b[3], c[3]), min(b[4], c[4]), max(b[5], c[5])]
    tv = value[t]
    if area[t] == record.get(tv, 0):
        record[tv] = record.get(tv, 0) + moved
    area[t] += moved

def is_cand(r, area, value, record, spare, skip):
    return (area[r] < record.get(value[r], 0) or value[r] in spare) and area[r] >= skip

def sequential(lab, spare, skip):
    reg, R, area, value, order, record, bbox = setup(lab, spare, skip)
    for r in order:
        if is_cand(r, area, value, record, spare, skip):
            apply(reg, area, value, record, bbox, r, decide(reg, r, spare))
    return reg, area, record

def separated(a, b):
    return any(a[2 * ax] >= b[2 * ax + 1] + 1 or b[2 * ax] >= a[2 * ax + 1] + 1 for ax in range(3))

def batched(lab, spare, skip, B=8, stats=None):
    reg, R, area, value, order, record, bbox = setup(lab, spare, skip)
    pos_of = {r: i for i, r in enumerate(order)}
    k = 0
    while k < R:
        batch, pos = [], k
        while pos < R and len(batch) < B:
            r = order[pos]
            if is_cand(r, area, value, record, spare, skip):
                if any(not separated(bbox[r], bbox[m]) for _, m in batch):
                    break
                batch.append((pos, r))
            elif area[r] >= skip and batch:
                break
            pos += 1
        end = pos
        if not batch:
            k = end
            continue
        if stats is not None:
            stats.append(len(batch))
        targets = [decide(reg, r, spare) for _, r in batch]       # parallel: all on the same snapshot
        trunc = None
        nxt = end
        for (p, r), t in zip(batch, targets):
            if trunc is not None and p > trunc:
                break
            before = area[t]
            apply(reg, area, value, record, bbox, r, t)
            if t != r and before < skip <= area[t] and p < pos_of[t] < end:
                if trunc is None or pos_of[t] < trunc:
                    trunc = pos_of[t]
        k = trunc if trunc is not None else end
    return reg, area, record



def _check(lab, B, stats=None):
    K = int(lab.max()) + 1
    for spare in ([], [K - 1], [1, 2]):
        for skip in (3, 1, 2):
            a = sequential(lab, spare, skip)
            b = batched(lab, spare, skip, B=B, stats=stats)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2], (lab.shape, spare, skip)


def test_batched_schedule_equals_sequential_loop_small_volumes():
    rng = np.random.default_rng(1)
    for trial in range(60):
        shape = (int(rng.integers(1, 5)), int(rng.integers(4, 14)), int(rng.integers(4, 14)))
        K = int(rng.integers(2, 5))
        if trial % 3 == 0:
            lab = rng.integers(0, K, size=shape)
        elif trial % 3 == 1:
            lab = np.where(rng.random(shape) < 0.7, 1, rng.integers(0, K, size=shape))
        else:
            lab = (rng.random(shape) < 0.5).astype(int) * rng.integers(1, K, size=shape)
        _check(lab.astype(np.uint8), B=int(rng.integers(2, 12)))


def test_batched_schedule_equals_sequential_loop_speckled_lungs():
    rng = np.random.default_rng(7)
    sizes = []
    for trial in range(6):
        shape = (int(rng.integers(2, 5)), 24, 24)
        lab = np.zeros(shape, np.uint8)
        lab[:, 4:20, 4:12] = 1
        lab[:, 4:20, 12:20] = 2
        n = int(0.06 * lab.size)
        lab[rng.integers(0, shape[0], n), rng.integers(0, 24, n), rng.integers(0, 24, n)] = rng.integers(0, 3, n)
        _check(lab, B=64, stats=sizes)
    assert max(sizes) > 8      # the batches really are batches
