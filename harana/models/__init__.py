"""Drop-in `harana.models` namespace: the classes the reference's scripts look up by name
(train_fastsvc.py:700-713, utils/utils.py:266-275).  The FastSVC generator and its blocks are backed by
libfsvc.so; every other model class (discriminators, Tacotron2, HN-uSFGAN) is the reference's own, re-exported
when the reference package is importable further down ``sys.path`` (svcc23_fastsvc_b200/dropin.py)."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

from .fastsvc import *  # noqa: E402,F401,F403

from svcc23_fastsvc_b200 import dropin as _dropin  # noqa: E402

_dropin.import_siblings(globals(), __name__, __file__)
