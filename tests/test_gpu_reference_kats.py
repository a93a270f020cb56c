"""The reference's own known-answer tests (tests/test_utils.py:58-63,73-107,124-159 of JoHof/lungmask),
asserted against the CUDA-backed functions of lungmask_b200.utils (same names, same arguments)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_bbox_3D():
    from lungmask_b200.utils import bbox_3D
    m = np.zeros((10, 10, 10), dtype=np.uint8)
    m[2:8, 3:7, 4:6] = 1
    assert tuple(bbox_3D(m, margin=2)) == (0, 10, 1, 9, 2, 8)


def _kat_slice():
    img = np.full((10, 10), dtype=np.int16, fill_value=-1000)
    img[2:8, 3:7] = 1
    img[9, 9] = 1
    return img


def test_simple_bodymask():
    from lungmask_b200.utils import simple_bodymask
    assert np.sum(simple_bodymask(_kat_slice())) == 24


def test_crop_and_resize():
    from lungmask_b200.utils import crop_and_resize
    cropped, bb = crop_and_resize(_kat_slice(), width=20, height=20)
    assert tuple(bb) == (2, 3, 8, 7)
    assert cropped.shape == (20, 20)
    assert np.sum(cropped) == 400


def test_preprocess():
    from lungmask_b200.utils import preprocess
    img = np.tile(_kat_slice(), (2, 1, 1))
    cropped, bb = preprocess(img, resolution=[20, 20])
    for sl, bb_ in zip(cropped, bb):
        assert tuple(bb_) == (2, 3, 8, 7)
        assert sl.shape == (20, 20)
        assert np.sum(sl) == 400


def test_reshape_mask():
    from lungmask_b200.utils import reshape_mask
    msk = np.full((10, 10), dtype=np.uint8, fill_value=1)
    out = reshape_mask(msk, (2, 2, 22, 22), origsize=(30, 30))
    assert out.shape == (30, 30)
    assert np.sum(out) == 400


def test_postprocessing():
    from lungmask_b200.utils import postprocessing
    label_image = np.zeros((1, 6, 6), dtype=np.uint8)
    label_image[0] = np.asarray([[0, 0, 0, 0, 0, 0], [0, 1, 1, 2, 2, 0], [0, 2, 0, 3, 1, 0], [0, 4, 4, 4, 0, 0],
                                 [0, 4, 0, 4, 0, 0], [0, 4, 4, 4, 0, 0]])
    res_gt = [[0, 0, 0, 0, 0, 0], [0, 1, 1, 2, 2, 0], [0, 1, 0, 3, 2, 0], [0, 4, 4, 4, 0, 0], [0, 4, 0, 4, 0, 0],
              [0, 4, 4, 4, 0, 0]]
    vol = np.tile(label_image, (2, 1, 1))
    assert np.all(postprocessing(vol, spare=[], disable_tqdm=True, skip_below=1)[0] == res_gt)
    assert postprocessing(vol, spare=[3], disable_tqdm=True, skip_below=1)[0][2, 3] == 2
    assert postprocessing(vol, spare=[3], disable_tqdm=True, skip_below=3)[0][2, 1] == 0
