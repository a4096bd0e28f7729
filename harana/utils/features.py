"""`harana.utils.features` drop-in (reference harana/utils/features.py:21-216): backed by libfsvc.so."""
from svcc23_fastsvc_b200.features import F0Statistics, SignalGenerator  # noqa: F401
