"""Drop-in for the reference's ``code/train.py`` (see dropin/ops.py): ``from train import FRVSR_Train`` (main.py:25)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from tecogan_b200.train import *  # noqa: F401,F403,E402
from tecogan_b200.train import TecoGAN, FRVSR_Train, EMA, VGG19_slim, Network, VGG_MEAN, identity  # noqa: F401,E402
