"""Host-side mirror of lungmask/mask.py for the B200 engine.

Same public names, arguments and error behaviour as the reference (`MODEL_URLS`, `get_model`,
`LMInferer`, deprecated `apply` / `apply_fused`; lungmask/mask.py:22-35,38-68,71-232,235-279), but the
work behind `LMInferer.apply` is the CUDA engine in liblungmask_b200.so.  PyTorch is used only to
read a .pth state_dict into a flat fp32 blob (lungmask_b200.h: lm_load_weights).
"""
import os
import warnings
from typing import Optional, Union

import numpy as np

from . import _native, orient
from .logger import logger

# model name -> (release URL, number of classes); lungmask/mask.py:22-35
_RELEASES = "https://github.com/JoHof/lungmask/releases/download/v0.0/"
MODEL_URLS = {
    "R231": (_RELEASES + "unet_r231-d5d2fc3d.pth", 3),
    "LTRCLobes": (_RELEASES + "unet_ltrclobes-3a07043d.pth", 6),
    "R231CovidWeb": (_RELEASES + "unet_r231covid-0de78a7e.pth", 3),
}

_CH = [64, 128, 256, 512, 1024]


def _conv3x3_prefixes():
    """The 18 Conv3x3+BN pairs in execution order (resunet.py:58-67): (conv prefix, bn prefix)."""
    out = []
    for i in range(5):
        out += [(f"down_path.{i}.block.0", f"down_path.{i}.block.2"), (f"down_path.{i}.block.3", f"down_path.{i}.block.5")]
    for j in range(4):
        p = f"up_path.{j}.conv_block.block"
        out += [(p + ".0", p + ".2"), (p + ".3", p + ".5")]
    return out


class NativeModel:
    """What `get_model` returns here: the live tensors of a reference state_dict flattened in the
    order lm_load_weights expects, plus the class count.  The dead `residual_*` tensors and the BN
    `num_batches_tracked` counters of the reference layout are validated for presence and dropped."""

    def __init__(self, state_dict):
        import torch

        def t(key, shape=None):
            if key not in state_dict:
                raise KeyError("state_dict is missing %r (not a lungmask U-Net checkpoint?)" % key)
            a = state_dict[key].detach().to(torch.float32).cpu().contiguous().numpy()
            if shape is not None and tuple(a.shape) != tuple(shape):
                raise ValueError("%s has shape %s, expected %s" % (key, tuple(a.shape), tuple(shape)))
            return a.ravel()

        # mask.py:56: the class count is the length of the LAST tensor of the state_dict
        self.n_classes = int(len(list(state_dict.values())[-1]))
        K = self.n_classes
        parts = []
        cin = 1
        chans = []  # (cin, cout) for the 18 convs
        for i in range(5):
            chans += [(cin, _CH[i]), (_CH[i], _CH[i])]
            cin = _CH[i]
        for j in range(4):
            c = _CH[3 - j]
            chans += [(2 * c, c), (c, c)]
        for (conv, bn), (ci, co) in zip(_conv3x3_prefixes(), chans):
            parts += [t(conv + ".weight", (co, ci, 3, 3)), t(conv + ".bias", (co,)), t(bn + ".weight", (co,)),
                      t(bn + ".bias", (co,)), t(bn + ".running_mean", (co,)), t(bn + ".running_var", (co,))]
        for j in range(4):
            c = _CH[3 - j]
            parts += [t(f"up_path.{j}.up.1.weight", (c, 2 * c, 1, 1)), t(f"up_path.{j}.up.1.bias", (c,))]
        parts += [t("last.weight", (K, 64, 1, 1)), t("last.bias", (K,))]
        self.blob = np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)


def get_model(modelname: str, modelpath: Optional[str] = None) -> NativeModel:
    """lungmask/mask.py:38-68.  `modelpath` given: torch.load of that file; otherwise the released
    weights are fetched through torch.hub (needs network).  The class count always comes from the
    file, never from `modelname`."""
    import torch

    if modelpath is None:
        url, _ = MODEL_URLS[modelname]
        state_dict = torch.hub.load_state_dict_from_url(url, progress=True, map_location=torch.device("cpu"))
    else:
        state_dict = torch.load(modelpath, map_location=torch.device("cpu"))
    return NativeModel(state_dict)


def _to_int16_volume(image: np.ndarray) -> np.ndarray:
    """Integer volumes.  Integers of any width give the same result as the reference because it clips to [-1024, 600]
    before resampling (utils.py:45) and thresholds at -500 HU, so they are clipped into int16 here."""
    if image.ndim != 3:
        raise ValueError("expected a (slices, H, W) volume, got shape %s" % (image.shape,))
    if image.dtype == np.int16:
        return np.ascontiguousarray(image)
    if np.issubdtype(image.dtype, np.integer) or image.dtype == bool:
        return np.clip(image, -1024, 600).astype(np.int16)
    raise TypeError("not an integer volume: %s" % image.dtype)


def _to_engine_volume(image: np.ndarray) -> np.ndarray:
    """The array the engine receives: int16 for integer volumes; float32 / float64 volumes keep their dtype, as they do
    in the reference (utils.preprocess clips and resamples in the input dtype and mask.py:167-168 normalises in it; the
    engine has a float path for exactly that).  Other float widths are widened to float32 (the reference would compute
    in float16 / longdouble there: documented deviation)."""
    if image.ndim != 3:
        raise ValueError("expected a (slices, H, W) volume, got shape %s" % (image.shape,))
    if image.dtype in (np.float32, np.float64):
        return np.ascontiguousarray(image)
    if np.issubdtype(image.dtype, np.floating):
        logger.warning("volume dtype %s is computed as float32", image.dtype)
        return np.ascontiguousarray(image, dtype=np.float32)
    return _to_int16_volume(image)


class LMInferer:
    def __init__(
        self,
        modelname: str = "R231",
        modelpath: Optional[str] = None,
        fillmodel: Optional[str] = None,
        fillmodel_path: Optional[str] = None,
        force_cpu: bool = False,
        batch_size: int = 20,
        volume_postprocessing: bool = True,
        tqdm_disable: bool = False,
        device: Optional[int] = None,
        wave_slices: Optional[int] = None,
    ):
        """Same arguments as the reference (lungmask/mask.py:72-82) plus `device` (CUDA ordinal, default
        LOCAL_RANK or 0) and `wave_slices`.  The reference's `batch_size` only bounds memory (slices are
        independent, mask.py:172-187; the engine is batch-invariant, tests/test_gpu_forward.py); the engine
        runs the forward in waves of `wave_slices` slices, default 37 when batch_size >= 20 because
        37 x 16 tiles = 4 x 148 SMs fills every level of the U-Net with whole waves of CTAs (6 GB of
        activations), else batch_size."""
        assert modelname in MODEL_URLS, "Modelname not found. Please choose from: {}".format(MODEL_URLS.keys())
        if fillmodel is not None:
            assert fillmodel in MODEL_URLS, "Modelname not found. Please choose from: {}".format(MODEL_URLS.keys())
        if modelpath is not None:  # a path overrides the name (mask.py:104-107)
            modelname = os.path.basename(modelpath)
        if fillmodel_path is not None:
            fillmodel = os.path.basename(fillmodel_path)
        if force_cpu:
            raise RuntimeError("lungmask_b200 is a B200 (sm_100a) engine and has no CPU path; "
                               "use the reference package for force_cpu=True")
        self.fillmodel = fillmodel
        self.modelname = modelname
        self.force_cpu = force_cpu
        self.batch_size = batch_size
        self.volume_postprocessing = volume_postprocessing
        self.tqdm_disable = tqdm_disable
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device

        self.model = get_model(self.modelname, modelpath)
        if wave_slices is None:
            wave_slices = 37 if batch_size >= 20 else batch_size
            if batch_size >= 20 and os.environ.get("LM_WAVE_SLICES"):   # measurement hook: 74 = eight CTA rounds per level
                wave_slices = max(1, min(1024, int(os.environ["LM_WAVE_SLICES"])))
        self.wave_slices = wave_slices
        self.engine = _native.Engine(device=device, batch_capacity=wave_slices)
        self.engine.load_weights(0, self.model.blob, self.model.n_classes)
        self.fillmodelm = None
        if self.fillmodel is not None:
            self.fillmodelm = get_model(self.fillmodel, fillmodel_path)
            self.engine.load_weights(1, self.fillmodelm.blob, self.fillmodelm.n_classes)

    # -- SimpleITK inputs are oriented to LPS first and back afterwards (mask.py:157-164,204-208): here that is an
    #    axis permutation + flips done by the engine on the device (lungmask_b200/orient.py, lm_apply_volume_oriented)
    @staticmethod
    def _sitk():
        try:
            import SimpleITK as sitk
            return sitk
        except Exception:
            return None

    def _run(self, volume: np.ndarray, code: str = "LPS") -> np.ndarray:
        vol = _to_engine_volume(volume)
        fused = self.fillmodel is not None
        if fused:
            logger.info(f"Apply: {self.modelname}")
            logger.info(f"Apply: {self.fillmodel}")
            logger.info("Fusing results... this may take up to several minutes!")
        if vol.dtype != np.int16:   # float volume
            if code != "LPS":       # re-orient on the host (rare: float + non-LPS); the integer path does it on the device
                res = self.engine.apply_volume_float(0, orient.to_lps(vol, code), slot_fill=-1 if not fused else 1,
                                                     postprocess=self.volume_postprocessing)
                if not fused:
                    return orient.from_lps(res, code)
                raise NotImplementedError("fusion of a float volume in a non-LPS orientation: re-orient the image first")
            return self.engine.apply_volume_float(0, vol, slot_fill=1 if fused else -1, postprocess=self.volume_postprocessing)
        if code == "LPS":
            if not fused:
                return self.engine.apply_volume(0, vol, postprocess=self.volume_postprocessing)
            return self.engine.apply_fused(0, 1, vol, postprocess=self.volume_postprocessing)
        perm, flip = orient.array_transform_to_lps(code)
        return self.engine.apply_volume_oriented(0, vol, perm, flip, slot_fill=1 if fused else -1,
                                                 postprocess=self.volume_postprocessing)

    def apply_oriented(self, array: np.ndarray, direction) -> np.ndarray:
        """What `apply(sitk_image)` does, for callers without SimpleITK: `array` = sitk.GetArrayFromImage(image)
        (axes z, y, x), `direction` = image.GetDirection() (9 direction cosines).  The mask comes back in the array's
        own orientation (mask.py:157-164,204-208)."""
        return self._run(array, orient.orientation_from_direction(direction))

    def apply(self, image) -> np.ndarray:
        """Segments a volume: numpy (slices, H, W), sitk.Image or lungmask_b200.io.Volume -> uint8 labels of the same shape
        (lungmask/mask.py:212-232).  The input is not modified."""
        if isinstance(image, np.ndarray):
            return self._run(image)
        from .io import Volume
        if isinstance(image, Volume):       # what lungmask_b200.io.load_input_image returns
            return self.apply_oriented(image.array, image.GetDirection())
        sitk = self._sitk()
        if sitk is None or not isinstance(image, sitk.Image):
            raise TypeError("apply() expects a numpy array, a SimpleITK image or a lungmask_b200.io.Volume")
        return self.apply_oriented(sitk.GetArrayFromImage(image), image.GetDirection())


def apply(image, model=None, force_cpu=False, batch_size=20, volume_postprocessing=True, tqdm_disable=False):
    """Deprecated wrapper (lungmask/mask.py:235-255)."""
    warnings.warn("The function `apply` will be removed in a future version. Please use the LMInferer class!",
                  DeprecationWarning)
    inferer = LMInferer(force_cpu=force_cpu, batch_size=batch_size, volume_postprocessing=volume_postprocessing,
                        tqdm_disable=tqdm_disable)
    if model is not None:
        if not isinstance(model, NativeModel):
            model = NativeModel(model.state_dict())
        inferer.model = model
        inferer.engine.load_weights(0, model.blob, model.n_classes)
    return inferer.apply(image)


def apply_fused(image, basemodel="LTRCLobes", fillmodel="R231", force_cpu=False, batch_size=20,
                volume_postprocessing=True, tqdm_disable=False):
    """Deprecated wrapper (lungmask/mask.py:258-279)."""
    warnings.warn("The function `apply_fused` will be removed in a future version. Please use the LMInferer class!",
                  DeprecationWarning)
    inferer = LMInferer(modelname=basemodel, force_cpu=force_cpu, fillmodel=fillmodel, batch_size=batch_size,
                        volume_postprocessing=volume_postprocessing, tqdm_disable=tqdm_disable)
    return inferer.apply(image)
