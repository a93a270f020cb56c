"""CPU proof-by-enumeration of the reduced neighbour set used by ccl_merge_kernel<26>(reduced = true) in
lungmask_b200/csrc/postproc.cu: when the left neighbour (z, y, x-1) carries the same label, a voxel only needs to be
united with it and with its four backward neighbours at dx = +1; the other eight backward neighbours are backward
neighbours of the left voxel.  The partition (and therefore the minimum-index roots the device ranks into skimage's
ids) must equal the one built from all 13 backward neighbours, on random and on blob-like label volumes."""
import numpy as np

BACKWARD = [(dz, dy, dx) for dz in (-1, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)
            if not (dz == 0 and (dy > 0 or (dy == 0 and dx >= 0)))]
assert len(BACKWARD) == 13


def _partition(vol, reduced):
    S, H, W = vol.shape
    parent = np.arange(vol.size)

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    def union(a, b):
        a, b = find(a), find(b)
        if a != b:
            parent[max(a, b)] = min(a, b)

    idx = lambda z, y, x: (z * H + y) * W + x
    for z in range(S):
        for y in range(H):
            for x in range(W):
                v = vol[z, y, x]
                if not v:
                    continue
                left = x > 0 and vol[z, y, x - 1] == v
                for dz, dy, dx in BACKWARD:
                    if reduced and left and dx != 1 and (dz, dy, dx) != (0, 0, -1):
                        continue
                    zz, yy, xx = z + dz, y + dy, x + dx
                    if zz < 0 or yy < 0 or yy >= H or xx < 0 or xx >= W:
                        continue
                    if vol[zz, yy, xx] == v:
                        union(idx(z, y, x), idx(zz, yy, xx))
    roots = np.array([find(i) if vol.flat[i] else -1 for i in range(vol.size)])
    return roots


def test_reduced_neighbour_set_gives_the_same_components():
    rng = np.random.default_rng(0)
    for trial in range(60):
        shape = tuple(rng.integers(1, 7, size=3))
        if trial % 3 == 0:      # salt-and-pepper labels
            vol = rng.integers(0, 4, size=shape)
        elif trial % 3 == 1:    # mostly homogeneous with holes
            vol = np.where(rng.random(shape) < 0.8, 1, rng.integers(0, 3, size=shape))
        else:                   # two labels, blobs
            vol = (rng.random(shape) < 0.55).astype(int) * (1 + (rng.random(shape) < 0.3))
        full = _partition(vol, reduced=False)
        red = _partition(vol, reduced=True)
        assert np.array_equal(full, red), (trial, shape)
