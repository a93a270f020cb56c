"""One-hop parity: the CUDA engine (through the C ABI) against the fixtures generated from the UNMODIFIED reference
(tests/golden/*, written by oracle/make_golden.py in the build container).  No oracle restatement in between for the
integer stages; the fp32 forward is compared within the stated tolerance."""
import json
import os
import zlib

import numpy as np
import pytest

from oracle import make_golden, synth
from _parity import TOL, dice_min

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(engine, slot, K):
    from lungmask_b200.mask import NativeModel
    sd = synth.random_state_dict(K, seed=10 + K)   # the fixtures' weights (make_golden.py)
    m = NativeModel(sd)
    engine.load_weights(slot, m.blob, m.n_classes)
    return sd


def test_preprocess_golden(engine):
    """utils.preprocess of the reference: boxes and the CRC32 of every resized slice (utils.py:32-52)."""
    for c in json.load(open(os.path.join(GOLD, "preprocess.json"))):
        vol = make_golden.pre_input(c["kind"], tuple(c["shape"]) if c["shape"] else None, c["seed"])
        out, boxes = engine.preprocess(vol)
        assert str(out.dtype) == c["dtype"]
        assert boxes.astype(int).tolist() == c["boxes"], c["kind"]
        assert [int(zlib.crc32(np.ascontiguousarray(s).tobytes())) for s in out] == c["crc32"], c["kind"]


def test_postprocess_golden(engine):
    """utils.postprocessing of the reference, bit for bit (utils.py:272-358)."""
    g = np.load(os.path.join(GOLD, "postprocess.npz"))
    for i, (S, K, seed, sp) in enumerate(make_golden.POST_CASES):
        lab = g[f"in{i}"]
        assert np.array_equal(engine.postprocess(lab), g[f"out{i}_plain"]), i
        assert np.array_equal(engine.postprocess(lab, spare=[K - 1]), g[f"out{i}_spare"]), i
        assert np.array_equal(engine.postprocess(lab, skip_below=1), g[f"out{i}_skip1"]), i


def test_forward_golden(engine):
    """resunet.UNet scores of the reference (sub-sampled) within 1e-4 (resunet.py:58-70)."""
    g = np.load(os.path.join(GOLD, "forward.npz"))
    for K in (3, 6):
        _load(engine, 0, K)
        vol = synth.phantom(2, seed=21)
        resized, _ = engine.preprocess(vol)
        _, scores = engine.forward(0, resized, return_scores=True)
        err = float(np.abs(scores[:, :, 3::8, 5::8] - g[f"scores_K{K}"]).max())
        print("K=%d max|dscore| vs the reference's UNet = %.3e (score range %.1f..%.1f)" % (K, err, g[f"scores_K{K}"].min(), g[f"scores_K{K}"].max()))
        # the fixtures use head gain 1 (scores down to about -30); the bound is stated relative to that range
        assert err <= max(TOL, 1e-5 * float(np.abs(g[f"scores_K{K}"]).max()))


def test_fusion_glue_golden(engine):
    """mask.py:228-232 on the reference's own res_l / res_r: integer work, bit-exact."""
    g = np.load(os.path.join(GOLD, "fusion.npz"))
    for tag in ("pp", "nopp"):
        pre, spare = engine.fuse(g[f"res_l_{tag}"], g[f"res_r_{tag}"])
        assert spare == int(g[f"res_l_{tag}"].max()) + 1
        assert np.array_equal(engine.postprocess(pre, spare=[spare]), g[f"fused_{tag}"]), tag


def test_end_to_end_golden(engine):
    """LMInferer.apply of the reference (force_cpu) on seeded volumes: single model and fusion.  Random weights give
    speckled maps full of argmax near-ties, so voxel identity is asserted through Dice here (the explain-every-voxel
    protocol runs against the oracle in test_gpu_forward.py / test_gpu_fullsize.py)."""
    g = np.load(os.path.join(GOLD, "fusion.npz"))
    for K in (3, 6):
        _load(engine, 0, K)
        out = engine.apply_volume(0, synth.phantom(4, 200, 216, seed=30 + K))
        ref = g[f"apply_K{K}"]
        d = int((out != ref).sum())
        print("K=%d voxels differing from the reference's apply(): %d of %d, dice(min) %.6f" % (K, d, ref.size, dice_min(out, ref)))
        assert d <= 2e-3 * ref.size
    _load(engine, 0, 6)
    _load(engine, 1, 3)
    vf = synth.phantom(4, 200, 216, seed=36)
    for tag, pp in (("pp", True), ("nopp", False)):
        out = engine.apply_fused(0, 1, vf, postprocess=pp)
        ref = g[f"fused_{tag}"]
        d = int((out != ref).sum())
        print("fused (%s) voxels differing from the reference: %d of %d, dice(min) %.6f" % (tag, d, ref.size, dice_min(out, ref)))
        assert d <= 5e-3 * ref.size
