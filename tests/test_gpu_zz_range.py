"""The fp16 operand format saturates at 65504.  The engine keeps a power-of-two scale per activation tensor (and per
layer's weights): when a value leaves the range, the tensor's scale is lowered and the forward runs again - exact
(a power of two changes no significand).  The reference has no range limit; neither has the engine (an error remains
only for non-finite weights and for activations beyond 1e14)."""
import numpy as np
import pytest

from oracle import restate, synth

pytestmark = pytest.mark.gpu
P = 2.0 ** 20


def _model(sd):
    from lungmask_b200.mask import NativeModel
    return NativeModel(sd)


@pytest.fixture(scope="module")
def base(engine):
    sd = synth.random_state_dict(3, seed=5, head_gain=0.3)
    vol = synth.phantom(2, seed=3)
    resized, _ = restate.preprocess(vol, resolution=[256, 256])
    m = _model(sd)
    engine.load_weights(2, m.blob, m.n_classes)
    labels, scores = engine.forward(2, resized, return_scores=True)
    return sd, resized, labels, scores


def test_weights_beyond_fp16_range_are_scaled(engine, base):
    """down_path.1.block.0: weights and bias times 2^20 (weights of about 1e5 > 65504), its BatchNorm statistics scaled to
    match (mean x 2^20, var x 2^40, eps x 2^40 is not expressible - so compare with the oracle, not bit for bit)."""
    sd, resized, labels0, scores0 = base
    big = dict(sd)
    big["down_path.1.block.0.weight"] = sd["down_path.1.block.0.weight"] * P
    big["down_path.1.block.0.bias"] = sd["down_path.1.block.0.bias"] * P
    big["down_path.1.block.2.running_mean"] = sd["down_path.1.block.2.running_mean"] * P
    big["down_path.1.block.2.running_var"] = sd["down_path.1.block.2.running_var"] * (P * P)
    m = _model(big)
    engine.load_weights(2, m.blob, m.n_classes)
    _, want = restate.forward_volume(restate.normalise(resized), big, batch_size=2, return_scores=True)
    _, got = engine.forward(2, resized, return_scores=True)
    err = float(np.abs(got - want).max())
    print("2^20-scaled layer: max|dscore| vs oracle = %.3e" % err)
    assert err <= 1e-4
    from lungmask_b200 import _native
    bad = dict(sd)
    w = sd["down_path.1.block.0.weight"].clone()
    w[0, 0, 0, 0] = float("inf")
    bad["down_path.1.block.0.weight"] = w
    with pytest.raises(_native.NativeError, match="non-finite"):
        engine.load_weights(2, _model(bad).blob, 3)


def test_activation_beyond_fp16_range_is_rescaled(engine, base):
    """Stem BatchNorm gamma / beta times 2^20 and the next convolution's weights times 2^-20: the same function in fp32,
    with activations of about 1e7 in between.  The engine must notice the overflow, lower that tensor's scale, run again
    and return the scores of the unscaled network (to an ulp: power-of-two scales are exact, only the fp16 planes'
    subnormal threshold sits elsewhere relative to the values); the scale sticks to the slot."""
    sd, resized, labels0, scores0 = base
    big = dict(sd)
    big["down_path.0.block.2.weight"] = sd["down_path.0.block.2.weight"] * P
    big["down_path.0.block.2.bias"] = sd["down_path.0.block.2.bias"] * P
    big["down_path.0.block.3.weight"] = sd["down_path.0.block.3.weight"] * (1.0 / P)
    m = _model(big)
    engine.load_weights(2, m.blob, m.n_classes)
    labels, scores = engine.forward(2, resized, return_scores=True)
    assert float(np.abs(scores - scores0).max()) <= 2e-6 and np.array_equal(labels, labels0)
    a0 = engine.read_activation(0, 2)      # stem output: unscaled on the way out
    assert np.isfinite(a0).all() and float(np.abs(a0).max()) > 65504.0
    labels2, scores2 = engine.forward(2, resized, return_scores=True)   # no re-run needed any more, same bits
    assert np.array_equal(scores2, scores)
    out = engine.apply_volume(2, synth.phantom(3, 150, 170, seed=4))    # whole path with a rescaled tensor
    m0 = _model(sd)
    engine.load_weights(2, m0.blob, m0.n_classes)                       # new weights reset the scales
    assert np.array_equal(out, engine.apply_volume(2, synth.phantom(3, 150, 170, seed=4)))
