"""`harana.models.fastsvc` drop-in: same import path as the reference module (tacotron2.py:22 imports
FastSVCFiLMNet from here).  The four generator classes are ours; the discriminators that live in the reference's
fastsvc.py (:386-1143) are re-exported from it when it is importable."""
from svcc23_fastsvc_b200 import dropin as _dropin
from svcc23_fastsvc_b200.generator import (  # noqa: F401
    FastSVCDownsampleNet,
    FastSVCFiLMNet,
    FastSVCGenerator,
    FastSVCUpsampleNet,
)

__all__ = _dropin.adopt_shadowed(
    globals(), __package__, __file__, "fastsvc",
    ["FastSVCGenerator", "FastSVCUpsampleNet", "FastSVCDownsampleNet", "FastSVCFiLMNet"])
