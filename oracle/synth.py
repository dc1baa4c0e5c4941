"""Deterministic synthetic weights / inputs (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference never seeds anything (``--rand_seed`` is parsed, never applied:
/root/reference/main.py:34).  To make oracle, golden fixtures, GPU tests and bench agree on
the same numbers on every box, parameters and inputs come from numpy's PCG64 stream, which
is stable across numpy versions, rather than from torch's RNG.

Parameter scale mimics PyTorch's default Conv2d / ConvTranspose2d / Linear initialisation
(kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias), which
is what the reference uses (code/ops.py:45-63).  ``gain`` > 1 gives the
"stress" variant SURVEY.md H4.6 asks for (logits far from 0 so the sigmoid does not hide
errors).
"""
import math

import numpy as np


def det_uniform(shape, seed, lo=0.0, hi=1.0):
    """U[lo,hi) float32 array from PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(int(seed)))
    n = int(np.prod(shape)) if len(shape) else 1
    x = rng.random(n, dtype=np.float32)
    return (x * np.float32(hi - lo) + np.float32(lo)).reshape(shape).astype(np.float32)


def _fan_in(name, shape):
    if len(shape) == 4:
        # Conv2d [Cout,Cin,kh,kw] -> Cin*kh*kw ; ConvTranspose2d [Cin,Cout,kh,kw] -> torch uses
        # size(1)*kh*kw as "fan_in" for its default init as well.
        return shape[1] * shape[2] * shape[3]
    if len(shape) == 2:
        return shape[1]
    return None


def fill_state_dict(state_dict, seed=1, gain=1.0):
    """Return {name: np.ndarray} with deterministic values for every tensor in state_dict.

    Biases use the fan_in of the weight that precedes them (torch default).  BatchNorm
    weight/bias (tensors with a sibling running_mean) get U(0.5,1.5)/U(-0.1,0.1) so that
    the affine part is exercised; running stats keep their defaults.
    """
    out = {}
    last_fan = 1
    bn_prefixes = {k[: -len("running_mean")] for k in state_dict if k.endswith("running_mean")}
    for i, (name, t) in enumerate(state_dict.items()):
        shape = tuple(t.shape)
        s = seed * 100003 + i
        prefix = name[: name.rfind(".") + 1]
        if name.endswith("num_batches_tracked") or name.endswith("running_mean") \
                or name.endswith("running_var"):
            out[name] = np.array(t.detach().cpu().numpy(), copy=True)
            continue
        if prefix in bn_prefixes:
            if name.endswith("weight"):
                out[name] = det_uniform(shape, s, 0.5, 1.5)
            else:
                out[name] = det_uniform(shape, s, -0.1, 0.1)
            continue
        fan = _fan_in(name, shape)
        if fan is not None:
            last_fan = fan
            b = gain / math.sqrt(fan)
            out[name] = det_uniform(shape, s, -b, b)
        else:
            b = 1.0 / math.sqrt(last_fan)
            out[name] = det_uniform(shape, s, -b, b)
    return out


def clip_inputs(n, t, h, w, seed=1234, hi=1.0):
    """Synthetic LR clip [n,t,3,h,w] U[0,hi).  hi=0.25 keeps the (x*4) 'flow' grid inside
    [0,1) so the warp really gathers."""
    return det_uniform((n, t, 3, h, w), seed, 0.0, hi)
