"""taichi.math stand-in (see taichi/__init__.py) -- test infrastructure only."""
import math as _m
import numpy as _np


def isnan(x):
    return bool(_np.isnan(x))


def sqrt(x):
    with _np.errstate(all="ignore"):
        return _np.sqrt(x)


def clamp(x, lo, hi):
    return _np.minimum(_np.maximum(x, lo), hi)


pi = _m.pi
vec3 = None  # filled in by taichi/__init__.py
vec2 = None
