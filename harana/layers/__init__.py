"""Drop-in `harana.layers` namespace (preprocess_fastsvc.py:35 imports Stretch2d from here); the layer classes the
generator does not use are the reference's own when it is importable."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

from .upsample import *  # noqa: E402,F401,F403
from .residual_block import *  # noqa: E402,F401,F403

from svcc23_fastsvc_b200 import dropin as _dropin  # noqa: E402

_dropin.import_siblings(globals(), __name__, __file__)
