"""Drop-in `harana.utils` namespace: only the excitation helpers on the path into the generator
(reference harana/utils/features.py); the reference's file/HDF5/checkpoint utilities stay the reference's."""
from .features import F0Statistics, SignalGenerator  # noqa: F401
