"""`lungmask` logger (same name and defaults as the reference's lungmask/logger.py:1-13)."""
import logging
import sys

logger = logging.getLogger("lungmask")
if not logger.handlers:
    _h = logging.StreamHandler(sys.stdout)
    _h.setFormatter(logging.Formatter("%(name)s %(asctime)s %(message)s", "%Y-%m-%d %H:%M:%S"))
    logger.addHandler(_h)
    logger.setLevel(logging.INFO)
