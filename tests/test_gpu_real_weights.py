"""The reference's own end-to-end known answers (tests/test_mask.py:30-60, tests/test_cli.py:11-20): voxel histograms
of LMInferer.apply on tests/testdata (two identical 512x512 slices = tests/golden/ct_slice_512.npz twice, identity
orientation, rescale slope 1 / intercept 0) with the RELEASED weights.  The weights are downloaded by the reference at
run time (mask.py:22-35) and there is no network here, so these tests are skipped unless the files are present:

    LM_WEIGHTS_DIR=/path/with/unet_r231-d5d2fc3d.pth [unet_ltrclobes-3a07043d.pth]   (or torch.hub's checkpoint dir)

When they are, the bar is the reference's: exact histograms."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

R231 = "unet_r231-d5d2fc3d.pth"
LOBES = "unet_ltrclobes-3a07043d.pth"
HIST_R231 = [423000, 64752, 36536]                      # tests/test_mask.py:36,48 and tests/test_cli.py:20
HIST_FUSED = [423000, 13334, 23202, 23834, 40918]       # tests/test_mask.py:58-60


def _find(name):
    dirs = [os.environ.get("LM_WEIGHTS_DIR", "")]
    try:
        import torch
        dirs.append(os.path.join(torch.hub.get_dir(), "checkpoints"))
    except Exception:
        pass
    for d in dirs:
        p = os.path.join(d, name) if d else ""
        if p and os.path.exists(p):
            return p
    return None


@pytest.fixture(scope="module")
def testvol(ct_slice):
    return np.stack([ct_slice, ct_slice])


def test_released_r231_histogram(testvol):
    p = _find(R231)
    if p is None:
        pytest.skip("released weights %s not available (no network); set LM_WEIGHTS_DIR" % R231)
    from lungmask_b200 import LMInferer
    res = LMInferer(modelpath=p, tqdm_disable=True).apply(testvol)
    assert np.unique(res, return_counts=True)[1].tolist() == HIST_R231
    # a path to the R231 weights with modelname LTRCLobes: the name is ignored, 3 classes come out (tests/test_mask.py:38-48)
    res = LMInferer(modelname="LTRCLobes", modelpath=p, tqdm_disable=True).apply(testvol)
    assert np.unique(res, return_counts=True)[1].tolist() == HIST_R231


def test_released_fusion_histogram(testvol):
    p3, p6 = _find(R231), _find(LOBES)
    if p3 is None or p6 is None:
        pytest.skip("released weights %s / %s not available (no network); set LM_WEIGHTS_DIR" % (R231, LOBES))
    from lungmask_b200 import LMInferer
    res = LMInferer(modelname="LTRCLobes", modelpath=p6, fillmodel="R231", fillmodel_path=p3, tqdm_disable=True).apply(testvol)
    assert np.unique(res, return_counts=True)[1].tolist() == HIST_FUSED
