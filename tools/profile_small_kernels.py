"""Run one small volume through the engine (for `ncu --set full` over the non-convolution kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lungmask_b200 import _native
from lungmask_b200.mask import NativeModel
from oracle import synth

sd = synth.random_state_dict(3, seed=1, head_gain=0.3)
m = NativeModel(sd)
eng = _native.Engine(0, 37)
eng.load_weights(0, m.blob, m.n_classes)
vol = synth.phantom(74, 320, 320, seed=3)
eng.apply_volume(0, vol)          # warm-up (not profiled: ncu skips the first launches)
out = eng.apply_volume(0, vol)
print("labels", [int((out == v).sum()) for v in range(3)])
