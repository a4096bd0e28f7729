"""`harana.utils.features` drop-in (reference harana/utils/features.py:21-216): ``SignalGenerator`` runs the batched
CUDA kernel for CUDA tensors and the reference's torch ops for host tensors (DataLoader workers)."""
from svcc23_fastsvc_b200 import dropin as _dropin
from svcc23_fastsvc_b200.features import F0Statistics, SignalGenerator  # noqa: F401

__all__ = _dropin.adopt_shadowed(globals(), __package__, __file__, "features", ["F0Statistics", "SignalGenerator"])
