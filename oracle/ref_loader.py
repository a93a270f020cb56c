"""Import the UNMODIFIED reference modules from /root/reference (this container only).

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py and by the `not gpu` tests that validate
`oracle.restate` against the reference itself.  /root/reference does not exist on the GPU box, so
nothing on a `-m gpu` test, smoke() or bench.py path may call this.
"""
import importlib
import os
import sys

from . import standins

REFERENCE_ROOT = os.environ.get("LUNGMASK_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lungmask", "mask.py"))


_cached = None


def load():
    """Returns a namespace with the reference's `mask`, `utils`, `resunet` modules (verbatim)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    standins.install()
    # The reference imports itself as `lungmask`; this repo ships a same-named drop-in shim, so
    # park whatever is registered under that name while the reference is being imported.
    parked = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "lungmask" or k.startswith("lungmask.")}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        mask = importlib.import_module("lungmask.mask")
        utils = importlib.import_module("lungmask.utils")
        resunet = importlib.import_module("lungmask.resunet")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "lungmask" or k.startswith("lungmask.")]:
            del sys.modules[k]
        sys.modules.update(parked)

    class Ref:
        pass

    ref = Ref()
    ref.mask, ref.utils, ref.resunet = mask, utils, resunet
    _cached = ref
    return ref
