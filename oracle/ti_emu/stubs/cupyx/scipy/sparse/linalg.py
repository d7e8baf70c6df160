"""cupyx.scipy.sparse.linalg stand-in (test infrastructure only).

The reference calls cupyx `spsolve` (cuSOLVER sparse QR, fp64) at
code/engine/sparse_solver.py:103.  SciPy's SuperLU (fp64 direct solve) stands in.
LAST_SOLVE records the most recent system so the golden generator can store it."""
import numpy as _np
import scipy.sparse.linalg as _sla
import torch as _torch
from torch.utils import dlpack as _dl

LAST_SOLVE = {}


class _Res:
    def __init__(self, x):
        self.x = x

    def toDlpack(self):
        return _dl.to_dlpack(_torch.from_numpy(_np.ascontiguousarray(self.x)))


def spsolve(H, b):
    H = H.tocsc()
    x = _sla.spsolve(H, _np.asarray(b, dtype=_np.float64))
    LAST_SOLVE["H"] = H
    LAST_SOLVE["b"] = _np.array(b, dtype=_np.float64)
    LAST_SOLVE["x"] = _np.array(x)
    return _Res(x)


def cg(H, b, x0, tol=1e-6):
    x, info = _sla.cg(H, b, x0=x0, rtol=tol)
    return _Res(x), info
