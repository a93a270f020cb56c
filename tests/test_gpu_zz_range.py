"""The fp16 operand format saturates at 65504: a weight or an activation beyond that must make the C ABI fail
(LM_ERR_RANGE, include/lungmask_b200.h "Numerics") instead of returning a mask computed from infinities."""
import numpy as np
import pytest

from oracle import restate, synth

pytestmark = pytest.mark.gpu


def _model(sd):
    from lungmask_b200.mask import NativeModel
    return NativeModel(sd)


def test_weight_beyond_fp16_range_is_refused(engine):
    from lungmask_b200 import _native
    sd = dict(synth.random_state_dict(3, seed=5, head_gain=0.3))
    w = sd["down_path.1.block.0.weight"].clone()
    w[0, 0, 0, 0] = 1.0e6
    sd["down_path.1.block.0.weight"] = w
    m = _model(sd)
    with pytest.raises(_native.NativeError, match="fp16 operand range"):
        engine.load_weights(2, m.blob, m.n_classes)


def test_activation_beyond_fp16_range_is_reported_and_engine_recovers(engine):
    from lungmask_b200 import _native
    good = synth.random_state_dict(3, seed=5, head_gain=0.3)
    bad = dict(good)
    bad["down_path.0.block.2.weight"] = good["down_path.0.block.2.weight"] * 1.0e7   # BatchNorm gamma of the stem
    vol = synth.phantom(2, seed=3)
    resized, _ = restate.preprocess(vol, resolution=[256, 256])
    m = _model(bad)
    engine.load_weights(2, m.blob, m.n_classes)          # the weights themselves are in range
    with pytest.raises(_native.NativeError, match="fp16 operand range"):
        engine.forward(2, resized)
    m = _model(good)
    engine.load_weights(2, m.blob, m.n_classes)
    labels = engine.forward(2, resized)                  # the flag does not stick
    assert labels.shape == resized.shape and labels.dtype == np.uint8
