"""Host logic of the orientation handling (lungmask_b200/orient.py) against explicit numpy constructions: an LPS volume
is re-sampled into every one of the 48 axis-aligned orientations by building the image's axes by hand, and the code /
(perm, flip) pair must bring it back."""
import itertools

import numpy as np
import pytest

from lungmask_b200 import orient


def _direction(code):
    d = np.zeros((3, 3))
    for c, ch in enumerate(code):
        r = ("LPS".index(ch) if ch in "LPS" else "RAI".index(ch))
        d[r, c] = 1.0 if ch in "LPS" else -1.0
    return d


def _all_codes():
    for p in itertools.permutations(range(3)):
        for signs in itertools.product((0, 1), repeat=3):
            yield "".join(("LPS" if s == 0 else "RAI")[r] for r, s in zip(p, signs))


def _resample(lps, code):
    """array (z,y,x) of the image whose axis c runs along physical axis phys[c] (negative side first if the letter is R/A/I),
    built element by element from the LPS array - independent of orient.py's transposes"""
    phys = [("LPS".index(ch) if ch in "LPS" else "RAI".index(ch)) for ch in code]
    size_phys = lps.shape[::-1]                      # LPS array is (z,y,x): physical sizes (x,y,z)
    size_img = [size_phys[phys[c]] for c in range(3)]
    out = np.empty(size_img[::-1], dtype=lps.dtype)
    for idx in np.ndindex(*size_img):                # idx = image index (i_x, i_y, i_z) in the new orientation
        pc = [0, 0, 0]
        for c in range(3):
            n = size_img[c]
            pc[phys[c]] = idx[c] if code[c] in "LPS" else n - 1 - idx[c]
        out[idx[2], idx[1], idx[0]] = lps[pc[2], pc[1], pc[0]]
    return out


def test_codes_from_directions():
    assert orient.orientation_from_direction(np.eye(3).ravel()) == "LPS"
    for code in _all_codes():
        assert orient.orientation_from_direction(_direction(code).ravel()) == code
    tilt = _direction("RAS") + 0.1 * np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]])   # slightly oblique
    assert orient.orientation_from_direction(tilt) == "RAS"


def test_all_48_orientations_round_trip():
    rng = np.random.default_rng(0)
    lps = rng.integers(-1000, 1000, size=(3, 4, 5)).astype(np.int16)
    for code in _all_codes():
        native = _resample(lps, code)
        assert np.array_equal(orient.to_lps(native, code), lps), code
        assert np.array_equal(orient.from_lps(lps, code), native), code
    with pytest.raises(ValueError):
        orient.array_transform_to_lps("LLS")
