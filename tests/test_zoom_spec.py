"""oracle.restate's explicit statement of scipy.ndimage.zoom (what the CUDA resize kernels implement)
against scipy itself, including the sizes whose last sample falls one ulp outside the input."""
import numpy as np
from scipy import ndimage

from oracle import restate


def test_zoom_helpers_match_scipy():
    rng = np.random.default_rng(0)
    overshoot = [n for n in range(2, 600) if (n - 1) * (np.float64(255) / (n - 1)) > 255][:8]
    assert 12 in overshoot  # 256 -> 12 samples: the reference's reshape_mask zeroes the last row / column
    sizes = sorted(set(list(range(2, 40)) + [97, 128, 200, 216, 255, 256, 257, 300, 414, 512] + overshoot))
    for n in sizes:
        a = rng.integers(-1024, 600, size=(n, n + 3)).astype(np.int16)
        assert np.array_equal(ndimage.zoom(a, np.asarray([256, 256]) / np.asarray(a.shape), order=1), restate.zoom_linear_to_int(a, 256, 256))
        assert np.array_equal(ndimage.zoom(a, 128 / np.asarray(a.shape), order=0), restate.zoom_nearest(a, 128, 128))
        b = rng.integers(1, 3, size=(256, 256)).astype(np.uint8)
        assert np.array_equal(ndimage.zoom(b, np.asarray([n, n + 3]) / 256, order=0), restate.zoom_nearest(b, n, n + 3))
        c = rng.integers(0, 2, size=(128, 128)).astype(bool)
        assert np.array_equal(ndimage.zoom(c, np.asarray([n, n + 3]) / 128, order=0), restate.zoom_nearest(c, n, n + 3))


def test_rounding_is_half_away_from_zero():
    a = np.array([[1, 2], [-1, -2]], dtype=np.int16)
    z = ndimage.zoom(a, [1, 1.5], order=1)          # middle column lands exactly on x.5
    assert z.tolist() == [[1, 2, 2], [-1, -2, -2]]
    assert np.array_equal(z, restate.zoom_linear_to_int(a, 2, 3))
