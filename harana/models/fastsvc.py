"""`harana.models.fastsvc` drop-in: same import path as the reference module
(tacotron2.py:22 imports FastSVCFiLMNet from here)."""
from svcc23_fastsvc_b200.generator import (  # noqa: F401
    FastSVCDownsampleNet,
    FastSVCFiLMNet,
    FastSVCGenerator,
    FastSVCUpsampleNet,
)

__all__ = ["FastSVCGenerator", "FastSVCUpsampleNet", "FastSVCDownsampleNet", "FastSVCFiLMNet"]
