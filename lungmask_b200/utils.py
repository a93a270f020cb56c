"""Host-side mirror of the hot-path functions of lungmask/utils.py, backed by the CUDA engine.

Same names and argument meaning as the reference (`preprocess`, `simple_bodymask`, `crop_and_resize`,
`reshape_mask`, `postprocessing`, `keep_largest_connected_component`, `bbox_3D`; lungmask/utils.py:32-129,272-404) so that the reference's
own known-answer tests (tests/test_utils.py:58-107,124-159) run unchanged against them.  Image I/O
(`read_dicoms`, `load_input_image`; utils.py:132-269) is re-exported from lungmask_b200/io.py, a SimpleITK-free reader
that returns `Volume` objects (array + geometry) instead of sitk images.
"""
import os

import numpy as np

from . import _native
from .io import load_input_image, read_dicoms  # noqa: F401  (reference names, utils.py:132,235)

_engine = None


def _eng():
    """A small default engine for the function-style API (device LOCAL_RANK or 0)."""
    global _engine
    if _engine is None:
        _engine = _native.Engine(device=int(os.environ.get("LOCAL_RANK", "0")), batch_capacity=4)
    return _engine


def _int16(img):
    img = np.asarray(img)
    if img.dtype == np.int16:
        return img
    if np.issubdtype(img.dtype, np.integer) or img.dtype == bool:
        return np.clip(img, -32768, 32767).astype(np.int16)
    raise TypeError("integer HU data expected, got %s" % img.dtype)


def preprocess(img, resolution=(192, 192)):
    """utils.py:32-52 -> (resized (S,res0,res1) int16, list of boxes)."""
    resized, boxes = _eng().preprocess(_int16(img), int(resolution[0]), int(resolution[1]), clip=True)
    return resized, [b for b in boxes.astype(np.int64)]


def simple_bodymask(img):
    """utils.py:55-82 -> boolean body mask of one slice."""
    return _eng().simple_bodymask(_int16(img)).astype(bool)


def crop_and_resize(img, width=192, height=192):
    """utils.py:85-111 -> (resized slice, bbox)."""
    resized, boxes = _eng().preprocess(_int16(img)[None], int(width), int(height), clip=False)
    return resized[0], boxes[0].astype(np.int64)


def reshape_mask(mask, tbox, origsize):
    """utils.py:114-129 -> mask resampled (nearest) into `tbox` of a zero canvas of `origsize`."""
    m = np.ascontiguousarray(np.asarray(mask), dtype=np.uint8)[None]
    box = np.asarray(tbox, dtype=np.int32)[None]
    return _eng().reshape_masks(m, box, int(origsize[0]), int(origsize[1]))[0].astype(np.float64)


def postprocessing(label_image, spare=[], disable_tqdm=False, skip_below=3):
    """utils.py:272-358 on a (S,H,W) label volume."""
    return _eng().postprocess(np.asarray(label_image), spare=spare, skip_below=skip_below)


def keep_largest_connected_component(mask):
    """utils.py:390-404 -> boolean mask of the largest full-connectivity component (2-D or 3-D input)."""
    m = (np.asarray(mask) != 0).astype(np.uint8)
    if m.ndim == 2:
        return _eng().keep_largest_component(m[None])[0].astype(bool)
    return _eng().keep_largest_component(m).astype(bool)


def bbox_3D(labelmap, margin=2):
    """utils.py:361-387 (host helper; the device merge loop tracks region extents itself)."""
    labelmap = np.asarray(labelmap)
    out = []
    for ax in range(labelmap.ndim):
        hit = np.flatnonzero(labelmap.any(axis=tuple(a for a in range(labelmap.ndim) if a != ax)))
        out += [max(int(hit[0]) - margin, 0), min(int(hit[-1]) + margin + 1, labelmap.shape[ax])]
    return np.array(out)


def get_DICOM_tags_to_keep():
    """utils.get_DICOM_tags_to_keep (utils.py:407-414) lists the tags the reference copies from the input series into the
    output image.  This build never copies DICOM tags (lungmask_b200/io.py: what `--removemetadata` asks for is always the
    case), so the list is empty."""
    return np.array([], dtype=str)
