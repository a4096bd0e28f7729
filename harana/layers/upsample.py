"""`harana.layers.upsample` drop-in (reference harana/layers/upsample.py:21-106 are ours; :109-242 the reference's)."""
from svcc23_fastsvc_b200 import dropin as _dropin
from svcc23_fastsvc_b200.layers import Conv1d1x3, Conv2d1x3, Squeeze2d, Stretch2d  # noqa: F401

__all__ = _dropin.adopt_shadowed(globals(), __package__, __file__, "upsample",
                                 ["Stretch2d", "Squeeze2d", "Conv1d1x3", "Conv2d1x3"])
