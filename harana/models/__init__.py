"""Drop-in `harana.models` namespace: the classes the reference's scripts look up
by name (train_fastsvc.py:700-713, utils/utils.py:266-275), backed by libfsvc.so."""
from .fastsvc import *  # noqa: F401,F403
