"""Opt-in checks of the kernels that were written after round 1's GPU budget was spent (conv_tc_pair.cu, stem_kernel_v2).
They are skipped unless LM_TEST_EXPERIMENTAL=1: each must reproduce the validated default path BIT FOR BIT (same
arithmetic in the same order, only the work assignment differs) before it may become a default."""
import os

import numpy as np
import pytest

from oracle import restate, synth

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("LM_TEST_EXPERIMENTAL", "0") in ("", "0"), reason="set LM_TEST_EXPERIMENTAL=1")]


def _forward(engine, resized, **options):
    for k, v in options.items():
        engine.set_option(k, v)
    try:
        return engine.forward(3, resized, return_scores=True)
    finally:
        for k in options:
            engine.set_option(k, 0)


@pytest.fixture(scope="module")
def setup(engine):
    from lungmask_b200.mask import NativeModel
    sd = synth.random_state_dict(3, seed=13, head_gain=0.3)
    m = NativeModel(sd)
    engine.load_weights(3, m.blob, m.n_classes)
    vol = synth.phantom(6, seed=21)      # 6 slices on a capacity-4 engine: a full wave and a 2-slice tail
    resized, _ = restate.preprocess(vol, resolution=[256, 256])
    return resized, _forward(engine, resized)


@pytest.mark.parametrize("option", ["stem_v2", "upsample_v2", "cta_pairs"])
def test_experimental_kernel_is_bit_identical(engine, setup, option):
    resized, (labels, scores) = setup
    l2, s2 = _forward(engine, resized, **{option: 1})
    assert np.array_equal(labels, l2)
    assert np.array_equal(scores, s2)
