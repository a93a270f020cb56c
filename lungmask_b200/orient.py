"""Orientation handling of LMInferer._inference for SimpleITK inputs (lungmask/mask.py:157-164,204-208), restated without
SimpleITK: the reference re-orients the image to DICOM "LPS" with sitk.DICOMOrient before the network and back to
the input's orientation afterwards.  For the array that sitk.GetArrayFromImage returns (axes z, y, x) such a
re-orientation is an axis permutation plus flips; both directions run on the device (lm_apply_volume_oriented).

Conventions (ITK's DICOMOrientation): the direction matrix holds, per image axis (column), the unit vector of that
axis in the patient's LPS frame (+x = Left, +y = Posterior, +z = Superior); an orientation code names, per image
axis x, y, z, the anatomical side the axis INCREASES toward ("LPS" = identity direction)."""
import numpy as np

_POS, _NEG = "LPS", "RAI"


def orientation_from_direction(direction) -> str:
    """Three-letter code of a 3x3 direction-cosine matrix (row-major 9-tuple as sitk's GetDirection returns, or 3x3).
    Oblique matrices snap to the closest axis-aligned orientation: the largest |cosine| is assigned first, then the
    largest among the remaining rows / columns (always a permutation)."""
    d = np.asarray(direction, dtype=np.float64).reshape(3, 3)
    m = np.abs(d).copy()
    code = [None, None, None]
    for _ in range(3):
        r, c = np.unravel_index(np.argmax(m), m.shape)
        code[c] = (_POS if d[r, c] > 0 else _NEG)[r]
        m[r, :] = -1.0
        m[:, c] = -1.0
    return "".join(code)


def array_transform_to_lps(code: str):
    """(perm, flip) taking the numpy array (axes z, y, x) of an image with orientation `code` to the array of the same
    image re-oriented to LPS:  lps = np.transpose(a, perm), then np.flip along every axis k with flip[k]."""
    code = code.upper()
    if len(code) != 3 or any(ch not in _POS + _NEG for ch in code):
        raise ValueError("not an orientation code: %r" % code)
    phys = [(_POS.index(ch) if ch in _POS else _NEG.index(ch)) for ch in code]   # physical axis of image axis c
    if sorted(phys) != [0, 1, 2]:
        raise ValueError("orientation code %r repeats an anatomical axis" % code)
    perm, flip = [0, 0, 0], [False, False, False]
    for j in range(3):                 # output image axis j = physical axis j
        c = phys.index(j)              # the input image axis that runs along it
        k = 2 - j                      # array axis of output image axis j (arrays are indexed z, y, x)
        perm[k] = 2 - c
        flip[k] = code[c] in _NEG
    return tuple(perm), tuple(flip)


def to_lps(a: np.ndarray, code: str) -> np.ndarray:
    """numpy statement of sitk.GetArrayFromImage(sitk.DICOMOrient(image, "LPS")) for an image of orientation `code`."""
    perm, flip = array_transform_to_lps(code)
    out = np.transpose(a, perm)
    for k in range(3):
        if flip[k]:
            out = np.flip(out, axis=k)
    return np.ascontiguousarray(out)


def from_lps(a_lps: np.ndarray, code: str) -> np.ndarray:
    """Inverse of to_lps: the LPS array back in the orientation `code` (sitk.DICOMOrient(image_lps, code))."""
    perm, flip = array_transform_to_lps(code)
    out = a_lps
    for k in range(3):
        if flip[k]:
            out = np.flip(out, axis=k)
    inv = np.argsort(perm)
    return np.ascontiguousarray(np.transpose(out, inv))
