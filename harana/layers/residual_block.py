"""`harana.layers.residual_block` drop-in (reference harana/layers/residual_block.py:27-48 are ours)."""
from svcc23_fastsvc_b200 import dropin as _dropin
from svcc23_fastsvc_b200.layers import Conv1d, Conv1d1x1  # noqa: F401

__all__ = _dropin.adopt_shadowed(globals(), __package__, __file__, "residual_block", ["Conv1d", "Conv1d1x1"])
