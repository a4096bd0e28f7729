"""Drop-in `harana.layers` namespace (preprocess_fastsvc.py:35 imports Stretch2d from here)."""
from .upsample import *  # noqa: F401,F403
from .residual_block import *  # noqa: F401,F403
