"""Shared parity protocol of the GPU tests, smoke() and bench.py (TEST INFRASTRUCTURE, see oracle/__init__.py).

north_star: label maps bit-exact, scores within 1e-4.  An fp32 forward cannot be bit-reproduced by a different
summation order (the reference itself differs between oneDNN and cuDNN), so "bit-exact labels" is checked as:
  1. pre-processing bit-exact;
  2. scores within TOL of the oracle's;
  3. every argmax flip sits on a pixel whose oracle top-2 margin is <= 2*TOL (undecidable in fp32);
  4. the integer stages are bit-exact in isolation (device post-processing of the ORACLE's labels == oracle);
  5. every differing OUTPUT voxel is explained by step 3: the oracle's integer stages applied to the ENGINE's
     argmax volume reproduce the engine's output bit for bit; with zero flips the outputs are identical.
"""
import numpy as np

from . import restate

TOL = 1e-4


def dice_min(a, b):
    """Minimum Dice over the labels present in either volume (1.0 when both are empty)."""
    out = []
    for v in np.union1d(np.unique(a), np.unique(b)):
        if v == 0:
            continue
        x, y = a == v, b == v
        out.append(2.0 * (x & y).sum() / max(1, x.sum() + y.sum()))
    return float(min(out)) if out else 1.0


def reshape_all(masks, boxes, hw):
    return np.asarray([restate.reshape_mask(masks[i], boxes[i], hw) for i in range(masks.shape[0])], dtype=np.uint8)


def explain_inference(engine, slot, vol, sd, batch, postprocess=True, tol=TOL, want=None, taps=None):
    """Runs the protocol above for LMInferer._inference (mask.py:141-210); returns a report dict."""
    if want is None:
        taps = {}
        want = restate.inference(vol, sd, batch_size=batch, volume_postprocessing=postprocess, taps=taps)
    got = engine.apply_volume(slot, vol, postprocess=postprocess)
    resized, boxes = engine.preprocess(vol)
    assert np.array_equal(resized, taps["resized"]), "pre-processing (resized slices) differs"
    assert np.array_equal(boxes.astype(np.int64), np.asarray(taps["boxes"], dtype=np.int64).reshape(-1, 4)), "crop boxes differ"
    labels, scores = engine.forward(slot, resized, return_scores=True)
    err = float(np.abs(scores - taps["scores"]).max())
    assert err <= tol, "scores differ from the oracle by %.3e (> %.1e)" % (err, tol)
    top2 = np.sort(taps["scores"], axis=1)[:, -2:]
    margin = top2[:, 1] - top2[:, 0]
    flip = labels != taps["labels"]
    assert not np.any(flip & (margin > 2 * tol)), "an argmax flip on a pixel with a decidable margin"
    if postprocess:
        assert np.array_equal(engine.postprocess(taps["labels"]), taps["post"]), "device post-processing of the oracle's labels"
        post_g = restate.postprocessing(labels)
    else:
        post_g = labels
    expect = reshape_all(post_g, boxes, vol.shape[1:])
    assert np.array_equal(got, expect), "output voxels that the argmax flips do not explain: %d" % int((got != expect).sum())
    nflip = int(flip.sum())
    if nflip == 0:
        assert np.array_equal(got, want)
    return {"score_err": err, "label_flips": nflip, "near_ties": int((margin <= 2 * tol).sum()),
            "voxels_differing": int((got != want).sum()), "voxels": int(want.size), "dice": dice_min(got, want),
            "score_range": [float(taps["scores"].min()), float(taps["scores"].max())], "got": got, "want": want}


def fmt(rep):
    return ("max|dscore| %.2e, argmax flips %d (near-ties %d), output voxels differing %d of %d, dice(min) %.6f"
            % (rep["score_err"], rep["label_flips"], rep["near_ties"], rep["voxels_differing"], rep["voxels"], rep["dice"]))


def explain_fused(engine, slot_base, slot_fill, vol, sd_base, sd_fill, batch, postprocess=True, tol=TOL):
    """The protocol for LMInferer.apply with a fill model (mask.py:223-232).  `postprocess` is
    volume_postprocessing: it reaches the inner inferences only; the fusion post-processing always runs."""
    got = engine.apply_fused(slot_base, slot_fill, vol, postprocess=postprocess)
    rl = explain_inference(engine, slot_base, vol, sd_base, batch, postprocess, tol)
    rr = explain_inference(engine, slot_fill, vol, sd_fill, batch, postprocess, tol)
    # stage isolation: fusion glue and original-resolution post-processing, bit-exact on the oracle's res_l / res_r
    want_pre, want_spare = restate.fuse_pre(rl["want"], rr["want"])
    pre, spare = engine.fuse(rl["want"], rr["want"])
    assert spare == int(want_spare) and np.array_equal(pre, want_pre), "fusion glue (mask.py:228-230)"
    want = restate.postprocessing(want_pre, spare=[want_spare])
    assert np.array_equal(engine.postprocess(pre, spare=[spare]), want), "fusion post-processing (mask.py:232)"
    # the engine's fused volume == the oracle's fusion of the ENGINE's inner results (differences come from argmax flips only)
    assert np.array_equal(got, restate.fuse(rl["got"], rr["got"])), "fused output not explained by the inner results"
    flips = rl["label_flips"] + rr["label_flips"]
    if flips == 0:
        assert np.array_equal(got, want)
    return {"score_err": max(rl["score_err"], rr["score_err"]), "label_flips": flips, "near_ties": rl["near_ties"] + rr["near_ties"],
            "voxels_differing": int((got != want).sum()), "voxels": int(want.size), "dice": dice_min(got, want),
            "got": got, "want": want}
