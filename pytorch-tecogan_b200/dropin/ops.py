"""Drop-in for the reference's ``code/ops.py``: put this directory on sys.path ahead of
``./code`` (reference main.py:11 does ``sys.path.insert(1, './code')``) and the unmodified
``main.py`` / ``train.py`` pick the B200 implementation up through their star imports."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from tecogan_b200.ops import *  # noqa: F401,F403,E402
from tecogan_b200.ops import np, torch, nn, F  # noqa: F401,E402  (names the reference's star-import chain supplies)
