"""tecogan_b200 — B200-native (sm_100a) implementation of the TecoGAN recurrent x4 VSR hot path.

    from tecogan_b200 import models, ops, train   # mirrors of the reference's code/models.py, code/ops.py, code/train.py

All compute goes through libtecogan_b200.so (C ABI, include/tecogan_b200.h).  There is no
CPU or library fallback: importing works anywhere, computing requires the built library and a
B200-class GPU.
"""
from . import _native  # noqa: F401
from . import ops, models, parallel, train  # noqa: F401

__all__ = ["ops", "models", "train", "parallel", "_native"]
