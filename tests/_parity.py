"""The parity protocol lives with the oracle (oracle/parity.py); the tests import it from here."""
from oracle.parity import TOL, dice_min, explain_fused, explain_inference, fmt, reshape_all  # noqa: F401
