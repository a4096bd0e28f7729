"""Drop-in `harana.utils` namespace: the excitation helpers on the path into the generator are ours (reference
harana/utils/features.py); the reference's file / HDF5 / checkpoint utilities (``read_hdf5``, ``load_model``,
``make_non_pad_mask`` ...) are re-exported from the reference when it is importable."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

from .features import *  # noqa: E402,F401,F403

from svcc23_fastsvc_b200 import dropin as _dropin  # noqa: E402

_dropin.import_siblings(globals(), __name__, __file__)
