"""CPU emulation of the union-find labelling kernels of lungmask_b200/csrc/postproc.cu (ccl_init_kernel +
ccl_merge_kernel), statement by statement, against scipy.ndimage.label:
  * ccl_init_kernel links the x-runs inside every 32-voxel segment of the box-linear thread order without atomics
    (parent = first voxel of the run within the segment); lane 0 of the merge kernel joins the segments;
  * the pruned neighbour rule (rule = 1) of the 26-connected kernel: per backward row with a = (x-1), b = (x),
    c = (x+1): b same label -> union with b unless the left voxel carries the label; otherwise union with c, and with
    a only if the left voxel does not carry the label;
  * the pruned face rule of the 6- / 4-connected kernels (no union with the upper / previous-slice neighbour when the
    left voxel and its own upper / previous-slice neighbour carry the label too).
The partition - and with min-index roots the raster-order ids the device derives from it - must equal the one from
all backward neighbours, on random, homogeneous-with-holes and blob volumes, on the full volume and on sub-boxes."""
import numpy as np
from scipy import ndimage


def _emulate(vol, box, conn, rule):
    S, H, W = vol.shape
    z0, z1, y0, y1, x0, x1 = box
    bw, bh = x1 - x0, y1 - y0
    n = (z1 - z0) * bh * bw
    parent = np.full(vol.size, -1, dtype=np.int64)

    def voxel(t):
        x = x0 + t % bw
        r = t // bw
        return z0 + r // bh, y0 + r % bh, x

    idx = lambda z, y, x: (z * H + y) * W + x

    def find(i):
        while parent[i] != i:
            i = parent[i]
        return i

    def union(a, b):
        a, b = find(a), find(b)
        if a != b:
            parent[max(a, b)] = min(a, b)

    # ccl_init_kernel: segments of 32 consecutive t
    for t0 in range(0, n, 32):
        lefts = []
        for lane in range(32):
            t = t0 + lane
            if t >= n:
                lefts.append(False)
                continue
            z, y, x = voxel(t)
            v = vol[z, y, x]
            lefts.append(bool(v) and lane > 0 and x > x0 and vol[z, y, x - 1] == v)
        for lane in range(32):
            t = t0 + lane
            if t >= n:
                continue
            z, y, x = voxel(t)
            if not vol[z, y, x]:
                continue
            start = lane
            while lefts[start]:
                start -= 1
            parent[idx(z, y, x)] = idx(z, y, x) - (lane - start)
    # ccl_merge_kernel
    for t in range(n):
        z, y, x = voxel(t)
        v = vol[z, y, x]
        if not v:
            continue
        i = idx(z, y, x)
        left = x > x0 and vol[z, y, x - 1] == v
        if left and t % 32 == 0:
            union(i, i - 1)
        if conn == 26:
            for dy, dz in ((-1, 0), (-1, -1), (0, -1), (1, -1)):
                yy, zz = y + dy, z + dz
                if zz < z0 or yy < y0 or yy >= y1:
                    continue
                j = idx(zz, yy, x)
                sb = vol[zz, yy, x] == v
                sa = x > x0 and vol[zz, yy, x - 1] == v
                sc = x + 1 < x1 and vol[zz, yy, x + 1] == v
                if rule:
                    if sb:
                        if not left:
                            union(i, j)
                        continue
                    if sa and not left:
                        union(i, j - 1)
                    if sc:
                        union(i, j + 1)
                else:
                    if sa:
                        union(i, j - 1)
                    if sb:
                        union(i, j)
                    if sc:
                        union(i, j + 1)
        else:
            if y > y0 and vol[z, y - 1, x] == v and not (left and vol[z, y - 1, x - 1] == v):
                union(i, i - W)
            if conn == 6 and z > z0 and vol[z - 1, y, x] == v and not (left and vol[z - 1, y, x - 1] == v):
                union(i, i - H * W)
    roots = np.full(vol.size, -1, dtype=np.int64)
    for t in range(n):
        z, y, x = voxel(t)
        if vol[z, y, x]:
            roots[idx(z, y, x)] = find(idx(z, y, x))
    return roots.reshape(vol.shape)


def _scipy_roots(vol, box, conn):
    """min-index root of every foreground voxel's component (equal value, connectivity `conn`) inside the box"""
    z0, z1, y0, y1, x0, x1 = box
    S, H, W = vol.shape
    sub = vol[z0:z1, y0:y1, x0:x1]
    if conn == 26:
        st = np.ones((3, 3, 3), int)
    else:
        st = ndimage.generate_binary_structure(3, 1)
        if conn == 4:
            st[0] = 0
            st[2] = 0
    lin = np.arange(vol.size).reshape(vol.shape)[z0:z1, y0:y1, x0:x1]
    roots = np.full(vol.shape, -1, dtype=np.int64)
    out = roots[z0:z1, y0:y1, x0:x1]
    for v in np.unique(sub):
        if v == 0:
            continue
        lab, k = ndimage.label(sub == v, structure=st)
        if k:
            mins = ndimage.minimum(lin, lab, index=np.arange(1, k + 1))
            m = lab > 0
            out[m] = np.asarray(mins, dtype=np.int64)[lab[m] - 1]
    return roots


def _volumes():
    rng = np.random.default_rng(0)
    for trial in range(36):
        shape = (int(rng.integers(1, 5)), int(rng.integers(1, 7)), int(rng.integers(1, 75)))  # rows longer than two segments
        if trial % 3 == 0:      # salt-and-pepper labels
            vol = rng.integers(0, 4, size=shape)
        elif trial % 3 == 1:    # mostly homogeneous with holes
            vol = np.where(rng.random(shape) < 0.85, 1, rng.integers(0, 3, size=shape))
        else:                   # two labels, blobs
            vol = (rng.random(shape) < 0.6).astype(int) * (1 + (rng.random(shape) < 0.3))
        yield trial, vol, rng


def test_pruned_26_connected_rule_and_run_links():
    for trial, vol, rng in _volumes():
        S, H, W = vol.shape
        full = (0, S, 0, H, 0, W)
        want = _scipy_roots(vol, full, 26)
        assert np.array_equal(_emulate(vol, full, 26, rule=0), want), (trial, vol.shape)
        assert np.array_equal(_emulate(vol, full, 26, rule=1), want), (trial, vol.shape)


def test_pruned_face_rules_on_boxes():
    for trial, vol, rng in _volumes():
        S, H, W = vol.shape
        x0 = int(rng.integers(0, max(1, W // 3)))
        box = (0, S, int(rng.integers(0, H)), H, x0, int(rng.integers(x0 + 1, W + 1)))
        box = (box[0], box[1], min(box[2], H - 1), box[3], box[4], box[5])
        binv = (vol > 0).astype(int)
        assert np.array_equal(_emulate(binv, box, 6, rule=1), _scipy_roots(binv, box, 6)), (trial, box)
        for z in range(S):
            b2 = (z, z + 1, box[2], box[3], box[4], box[5])
            assert np.array_equal(_emulate(binv, b2, 4, rule=1), _scipy_roots(binv, b2, 4)), (trial, b2)
