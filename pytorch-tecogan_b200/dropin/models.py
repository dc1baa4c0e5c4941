"""Drop-in for the reference's ``code/models.py`` (see dropin/ops.py)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from tecogan_b200.models import *  # noqa: F401,F403,E402
from tecogan_b200.models import generator, discriminator, f_net, residual_block, discriminator_block  # noqa: F401,E402
from tecogan_b200.ops import np, torch, nn, F  # noqa: F401,E402
