"""Compact GPU diagnostics (development aid): per-layer forward error and post-processing diffs."""
import sys, os, traceback
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lungmask_b200 import _native
from lungmask_b200.mask import NativeModel
from oracle import restate, synth

eng = _native.Engine(0, 4)

def fwd(K):
    sd = synth.random_state_dict(K, seed=10 + K)
    m = NativeModel(sd)
    eng.load_weights(0, m.blob, m.n_classes)
    vol = synth.phantom(2, seed=33)
    resized, _ = restate.preprocess(vol, resolution=[256, 256])
    taps = {}
    with torch.inference_mode():
        want = restate.unet_forward(torch.as_tensor(restate.normalise(resized)[:, None], dtype=torch.float32), sd, taps=taps).numpy()
    labels, scores = eng.forward(0, resized, return_scores=True)
    ids = {"S0": 1, "P0": 2, "S1": 4, "P1": 5, "S2": 7, "P2": 8, "S3": 10, "P3": 11, "B4": 13, "U0": 15, "E0": 17,
           "U1": 19, "E1": 21, "U2": 23, "E2": 25, "U3": 27}
    for name, aid in ids.items():
        got = eng.read_activation(aid, 2)
        w = taps[name].permute(0, 2, 3, 1).numpy()
        print("K=%d %-3s rel err %.3e  (max|x| %.3f)" % (K, name, np.abs(got - w).max() / (np.abs(w).max() + 1e-12), np.abs(w).max()), flush=True)
    print("K=%d scores max err %.3e" % (K, np.abs(scores - want).max()), flush=True)

def post():
    rng = np.random.default_rng(5)
    rng.integers(0, 4, size=(4, 24, 24)); 
    single = np.random.default_rng(7).integers(0, 3, size=(1, 40, 40)).astype(np.uint8)
    cases = [("single", single, {}), ("s2clean", synth.label_noise_volume(2, 3, seed=5, speckle=0.0), {}),
             ("s12", synth.label_noise_volume(12, 3, seed=15, speckle=2e-3), {}),
             ("s5skip1", synth.label_noise_volume(5, 3, seed=8, speckle=0.0), {"skip_below": 1})]
    for name, lab, kw in cases:
        taps = {}
        want = restate.postprocessing(lab, taps=taps, **kw)
        for stage, key in ((2, "regions0"), (3, "regions1"), (1, "mapped"), (0, None)):
            eng.set_option("post_debug_stage", stage)
            got = eng.postprocess(lab, **kw)
            w = want if key is None else (taps[key] & 255).astype(np.uint8)
            d = got != w
            print("post %-8s stage %d (%s): differing voxels %d of %d; R=%d" % (name, stage, key, d.sum(), d.size, taps["regions0"].max()), flush=True)
            if d.any():
                idx = np.argwhere(d)
                print("     first diffs", idx[:4].tolist(), "want", w[d][:6].tolist(), "got", got[d][:6].tolist())
        eng.set_option("post_debug_stage", 0)

def chunks():
    sd = synth.random_state_dict(3, seed=13, head_gain=0.3)
    m = NativeModel(sd)
    eng.load_weights(0, m.blob, m.n_classes)
    vol = synth.phantom(2, seed=33)
    resized, _ = restate.preprocess(vol, resolution=[256, 256])
    with torch.inference_mode():
        want = restate.unet_forward(torch.as_tensor(restate.normalise(resized)[:, None], dtype=torch.float32), sd).numpy()
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        ex = restate.unet_forward(torch.as_tensor(restate.normalise(resized)[:, None], dtype=torch.float64), sd64).numpy()
    print("oracle fp32 vs fp64 scores: max %.3e mean %.3e" % (np.abs(want - ex).max(), np.abs(want - ex).mean()))
    import time
    for ck, ckw in ((1, 1), (1, 2), (2, 2)):
        eng.set_option("chunk_kb", ck); eng.set_option("chunk_kb_wide", ckw)
        labels, scores = eng.forward(0, resized, return_scores=True)
        print("chunk_kb=%d wide=%d: engine vs fp32 oracle max %.3e mean %.3e | vs fp64 max %.3e mean %.3e" % (
            ck, ckw, np.abs(scores - want).max(), np.abs(scores - want).mean(), np.abs(scores - ex).max(), np.abs(scores - ex).mean()), flush=True)
    for K in (3, 6):
        sdk = synth.random_state_dict(K, seed=10 + K, head_gain=0.3)
        mk = NativeModel(sdk); eng.load_weights(0, mk.blob, mk.n_classes)
        vol5 = synth.phantom(5, seed=21); r5, _ = restate.preprocess(vol5, resolution=[256, 256])
        wl, ws = restate.forward_volume(restate.normalise(r5), sdk, batch_size=2, return_scores=True)
        for ck, ckw in ((1, 1), (1, 2), (2, 2)):
            eng.set_option("chunk_kb", ck); eng.set_option("chunk_kb_wide", ckw)
            l, sc = eng.forward(0, r5, return_scores=True)
            print("pytest-net K=%d chunk=%d wide=%d: max|dscore| %.3e" % (K, ck, ckw, np.abs(sc - ws).max()), flush=True)
    eng.set_option("chunk_kb", 1)

for f, a in ((chunks, ()),):
    try:
        f(*a)
    except Exception:
        traceback.print_exc()
