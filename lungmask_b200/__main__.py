"""`lungmask INPUT OUTPUT` command line (flags as lungmask/__main__.py:26-76).

File I/O needs SimpleITK exactly as in the reference (it is outside the accelerated path); `.npy`
volumes are accepted in addition so the CLI is usable on machines without it.
"""
import argparse
import os
import sys

import numpy as np

from .logger import logger
from .mask import LMInferer

__version__ = "0.1.0+b200"


def _path(string):
    if os.path.exists(string):
        return string
    sys.exit(f"File not found: {string}")


def build_parser():
    p = argparse.ArgumentParser(prog="lungmask")
    p.add_argument("input", metavar="input", type=_path, help="Path to the input image: file, DICOM directory or .npy volume")
    p.add_argument("output", metavar="output", type=str, help="Filepath for output lungmask")
    p.add_argument("--modelname", choices=["R231", "LTRCLobes", "LTRCLobes_R231", "R231CovidWeb"], default="R231")
    p.add_argument("--modelpath", type=str, default=None, help="spcifies the path to the trained model")
    p.add_argument("--cpu", action="store_true", help="not supported by the B200 engine (kept for flag compatibility)")
    p.add_argument("--nopostprocess", action="store_true", help="Deactivates postprocessing")
    p.add_argument("--noHU", action="store_true", help=argparse.SUPPRESS)
    p.add_argument("--batchsize", type=int, default=20, help="Number of slices processed simultaneously")
    p.add_argument("--noprogress", action="store_true", help="If set, no tqdm progress bar will be shown")
    p.add_argument("--version", action="version", version=__version__)
    p.add_argument("--removemetadata", action="store_true", help="Do not keep study/patient metadata of the input")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    batchsize = 1 if args.cpu else args.batchsize  # __main__.py:81-83
    logger.info("Load model")
    is_npy = os.path.isfile(args.input) and args.input.endswith(".npy")
    if is_npy:
        image, sitk_image = np.load(args.input), None
    else:
        try:
            import SimpleITK as sitk
        except ImportError:
            sys.exit("SimpleITK is required to read this input (only .npy volumes work without it)")
        if os.path.isfile(args.input):
            sitk_image = sitk.ReadImage(args.input)
        else:
            names = sitk.ImageSeriesReader.GetGDCMSeriesFileNames(args.input)
            if not names:
                sys.exit("No dicoms found!")
            sitk_image = sitk.ReadImage(names)
        image = sitk_image
    if args.modelname == "LTRCLobes_R231":  # __main__.py:95-107
        assert args.modelpath is None, "Modelpath can not be specified for LTRCLobes_R231 fusion"
        inferer = LMInferer(modelname="LTRCLobes", force_cpu=args.cpu, fillmodel="R231", batch_size=batchsize,
                            volume_postprocessing=not args.nopostprocess, tqdm_disable=args.noprogress)
    else:
        inferer = LMInferer(modelname=args.modelname, modelpath=args.modelpath, force_cpu=args.cpu, batch_size=batchsize,
                            volume_postprocessing=not args.nopostprocess, tqdm_disable=args.noprogress)
    result = inferer.apply(image)
    logger.info(f"Save result to: {args.output}")
    if sitk_image is None or args.output.endswith(".npy"):
        np.save(args.output, result)
    else:
        import SimpleITK as sitk
        out = sitk.GetImageFromArray(result)
        out.CopyInformation(sitk_image)
        sitk.WriteImage(out, args.output)


if __name__ == "__main__":
    main()
