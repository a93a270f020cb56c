"""GPU parity, stage by stage: the CUDA engine (through the C ABI) against the CPU oracle on the same
seeded inputs.  Integer stages must be bit-exact."""
import numpy as np
import pytest

from oracle import restate, synth

pytestmark = pytest.mark.gpu


def _volumes(ct_slice):
    vols = [("ct512", np.stack([ct_slice, ct_slice[::-1].copy(), np.roll(ct_slice, 37, 1), ct_slice.T.copy()]))]
    for seed, shape in enumerate([(6, 256, 256), (4, 300, 414), (3, 97, 200), (2, 512, 512), (2, 64, 64), (2, 130, 127),
                                  (2, 14, 27), (1, 10, 10)]):
        vols.append(("phantom%s" % (shape,), synth.phantom(*shape, seed=seed)))
    empty = np.full((2, 100, 120), -1000, np.int16)   # no body at all -> full-frame box
    vols.append(("air", empty))
    full = np.full((1, 90, 90), 100, np.int16)        # everything is body
    vols.append(("solid", full))
    return vols


def test_preprocess_bit_exact(engine, ct_slice):
    for name, vol in _volumes(ct_slice):
        want, want_boxes = restate.preprocess(vol, resolution=[256, 256])
        got, got_boxes = engine.preprocess(vol)
        assert np.array_equal(np.asarray(want_boxes, dtype=np.int64).reshape(-1, 4), got_boxes.astype(np.int64)), name
        assert np.array_equal(want, got), name


def test_bodymask_bit_exact(engine, ct_slice):
    for name, vol in _volumes(ct_slice):
        for sl in vol[:2]:
            want = restate.simple_bodymask(sl)
            got = engine.simple_bodymask(sl)
            assert np.array_equal(want.astype(np.uint8), got), name


def test_resize_other_resolutions(engine):
    vol = synth.phantom(2, 200, 216, seed=3)
    for res in ([192, 192], [20, 33], [256, 128]):
        want, wb = restate.preprocess(vol, resolution=res)
        got, gb = engine.preprocess(vol, res[0], res[1])
        assert np.array_equal(want, got) and np.array_equal(np.asarray(wb), gb)


def test_reshape_bit_exact(engine):
    rng = np.random.default_rng(0)
    # include target sizes whose last sample lands one ulp outside (12, 200, 216, ...: zero last row/col)
    for (H, W, box) in [(512, 512, (95, 3, 509, 509)), (300, 414, (53, 31, 247, 383)), (64, 64, (12, 5, 52, 59)),
                        (40, 40, (10, 10, 22, 24)), (256, 256, (20, 30, 220, 246)), (230, 240, (7, 9, 223, 225)),
                        (30, 30, (2, 2, 22, 22)), (300, 300, (0, 0, 300, 300)), (20, 20, (3, 4, 4, 5))]:
        S = 3
        masks = rng.integers(0, 6, size=(S, 256, 256)).astype(np.uint8)
        boxes = np.tile(np.asarray(box, dtype=np.int32), (S, 1))
        want = np.asarray([restate.reshape_mask(masks[i], box, (H, W)) for i in range(S)], dtype=np.uint8)
        got = engine.reshape_masks(masks, boxes, H, W)
        assert np.array_equal(want, got), (H, W, box)


@pytest.mark.parametrize("S,K,speckle", [(12, 3, 2e-3), (9, 6, 2e-3), (1, 3, 2e-3), (2, 6, 1e-3), (5, 3, 0.0), (40, 6, 5e-4)])
def test_postprocess_bit_exact(engine, S, K, speckle):
    lab = synth.label_noise_volume(S, K, seed=S + K, speckle=speckle)
    want = restate.postprocessing(lab)
    got = engine.postprocess(lab)
    assert np.array_equal(want, got)
    for spare in ([K - 1], [1, 2]):
        assert np.array_equal(restate.postprocessing(lab, spare=spare), engine.postprocess(lab, spare=spare)), spare
    assert np.array_equal(restate.postprocessing(lab, skip_below=1), engine.postprocess(lab, skip_below=1))


def test_postprocess_edge_cases(engine):
    z = np.zeros((3, 32, 48), np.uint8)
    assert np.array_equal(engine.postprocess(z), z)
    one = np.ones((2, 16, 16), np.uint8)  # no background at all: the first unique value is skipped (utils.py:355)
    assert np.array_equal(restate.postprocessing(one), engine.postprocess(one))
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 4, size=(4, 24, 24)).astype(np.uint8)  # pure speckle: hundreds of tiny regions
    assert np.array_equal(restate.postprocessing(noise), engine.postprocess(noise))
    assert np.array_equal(restate.postprocessing(noise, spare=[3]), engine.postprocess(noise, spare=[3]))
    single = rng.integers(0, 3, size=(1, 40, 40)).astype(np.uint8)  # single slice: area_closing branch
    assert np.array_equal(restate.postprocessing(single), engine.postprocess(single))


def test_postprocess_non_square(engine):
    lab = synth.label_noise_volume(6, 6, seed=2, speckle=1e-3, H=200, W=312)
    assert np.array_equal(restate.postprocessing(lab), engine.postprocess(lab))
    assert np.array_equal(restate.postprocessing(lab, spare=[6]), engine.postprocess(lab, spare=[6]))


def test_keep_largest_component_bit_exact(engine):
    from lungmask_b200.utils import keep_largest_connected_component
    rng = np.random.default_rng(3)
    for shape in ((6, 40, 52), (1, 64, 64), (3, 17, 9)):
        m = rng.random(shape) < 0.45
        want = restate.keep_largest_connected_component(m)
        assert np.array_equal(want, engine.keep_largest_component(m.astype(np.uint8)).astype(bool)), shape
    m2 = rng.random((33, 47)) < 0.5
    assert np.array_equal(restate.keep_largest_connected_component(m2), keep_largest_connected_component(m2))
    two = np.zeros((1, 8, 8), np.uint8)   # two components of equal area: argsort(...)[-1] keeps the later one
    two[0, 1, 1:4] = 1
    two[0, 6, 2:5] = 1
    assert np.array_equal(restate.keep_largest_connected_component(two != 0), engine.keep_largest_component(two).astype(bool))


def test_ccl_rules_agree(engine):
    """The pruned neighbour rule of the 26-connected labelling (default) against probing all 13 backward neighbours
    (lm_set_option("ccl_rule", 0)): identical post-processing on clean, speckled and pure-noise label volumes (the CPU
    emulation of both kernels against scipy is tests/test_ccl_neighbour_rule.py)."""
    rng = np.random.default_rng(7)
    vols = [synth.label_noise_volume(12, 3, seed=5, speckle=0.0), synth.label_noise_volume(12, 6, seed=15, speckle=2e-3),
            synth.label_noise_volume(7, 3, seed=25, speckle=2e-2, H=130, W=97),   # rows that are not a multiple of 32 voxels
            rng.integers(0, 4, size=(5, 33, 47)).astype(np.uint8)]
    for lab in vols:
        want = engine.postprocess(lab)
        engine.set_option("ccl_rule", 0)
        try:
            got = engine.postprocess(lab)
        finally:
            engine.set_option("ccl_rule", 1)
        assert np.array_equal(got, want)
        assert np.array_equal(want, restate.postprocessing(lab))


def test_region_table_overflow_is_rerun(engine):
    """The region count stays on the device; when it exceeds the table capacity the device raises a flag and the call
    is repeated with larger tables (lm_set_option("post_region_capacity", n) shrinks them for this test)."""
    rng = np.random.default_rng(9)
    noise = rng.integers(0, 4, size=(4, 40, 40)).astype(np.uint8)   # thousands of regions
    want = restate.postprocessing(noise)
    assert np.array_equal(engine.postprocess(noise), want)
    engine.set_option("post_region_capacity", 64)
    try:
        assert np.array_equal(engine.postprocess(noise), want)          # overflows, grows, runs again
        assert np.array_equal(engine.postprocess(noise, spare=[3]), restate.postprocessing(noise, spare=[3]))
    finally:
        engine.set_option("post_region_capacity", 1 << 16)


def test_sort_paths(engine):
    """region_sort_kernel: the shared-memory path (<= 4096 regions) and the global-memory path give the reference's
    stable ascending-area order (checked through the bit-exact post-processing of volumes on either side)."""
    rng = np.random.default_rng(13)
    small = rng.integers(0, 3, size=(2, 30, 30)).astype(np.uint8)       # a few hundred regions
    big = rng.integers(0, 5, size=(6, 64, 64)).astype(np.uint8)         # > 4096 regions
    for lab in (small, big):
        assert np.array_equal(engine.postprocess(lab), restate.postprocessing(lab))


def test_merge_kernels_agree(engine):
    """Q4 on one CTA (lm_set_option("merge_ctas", 1)) and on one CTA per SM in batches of independent candidates
    (default): the same post-processing, bit for bit, and both equal the oracle - thousands of regions, spare labels,
    skip_below = 1 (every speckle is a candidate)."""
    rng = np.random.default_rng(21)
    vols = [synth.label_noise_volume(16, 3, seed=31, speckle=2e-2),                   # lungs + dense speckle
            synth.label_noise_volume(10, 6, seed=32, speckle=5e-3, H=200, W=312),
            rng.integers(0, 4, size=(6, 48, 48)).astype(np.uint8)]                     # pure noise
    for lab in vols:
        for kw in ({}, {"spare": [int(lab.max())]}, {"skip_below": 1}):
            want = restate.postprocessing(lab, **kw)
            assert np.array_equal(engine.postprocess(lab, **kw), want), kw
            engine.set_option("merge_ctas", 1)
            try:
                assert np.array_equal(engine.postprocess(lab, **kw), want), kw
            finally:
                engine.set_option("merge_ctas", 0)
