"""Pure-Python stand-in for the `taichi` package -- TEST INFRASTRUCTURE ONLY.

Purpose: execute the UNMODIFIED reference sources (read from /root/reference/code
through oracle/ti_emu/hook.py) serially, so that they can emit golden vectors for
the parity tests.  Taichi itself (a JIT) cannot be installed in this image.  This
module only implements the subset of the Taichi API that the hot-path files use
(SURVEY.md section 8c lists it).  Nothing in the product package imports it.

Semantics reproduced:
  * fields are dense numpy arrays, zero initialised (Taichi zero-inits fields);
  * `field[idx]` of a vector/matrix field is a *view* (so `f[i][j] = v` and
    `f[i].x += v` write through) while plain assignment `a = f[i]` is turned
    into a value copy by the import hook (Taichi kernels have value semantics);
  * kernels/funcs are ordinary Python functions run serially (parallel-for
    nondeterminism disappears; atomics are rewritten by the hook);
  * default_fp = f64, default_ip = i32.
"""
import math as _math
import numpy as _np

from . import math  # noqa: F401  (taichi.math)

f64 = _np.float64
f32 = _np.float32
i32 = _np.int32
i64 = _np.int64
u8 = _np.uint8
cpu = "cpu"
gpu = "gpu"
cuda = "cuda"


def _dt(dt):
    if dt is float:
        return _np.float64
    if dt is int:
        return _np.int32
    return dt


def init(*a, **k):
    return None


def data_oriented(cls):
    return cls


def func(f):
    return f


def _to_np(a):
    try:
        import torch
        if isinstance(a, torch.Tensor):
            return a.detach().numpy()  # shares memory on CPU
    except ImportError:  # pragma: no cover
        pass
    return a


def kernel(f):
    import functools

    @functools.wraps(f)
    def w(*args, **kw):
        args = tuple(_to_np(a) for a in args)
        kw = {k: _to_np(v) for k, v in kw.items()}
        r = f(*args, **kw)
        return r
    return w


def template():
    return None


def static(x, *rest):
    if rest:
        return (x,) + rest
    return x


def loop_config(**k):
    return None


def _rng(a):
    if isinstance(a, (tuple, list)) or (isinstance(a, _np.ndarray) and a.ndim == 1):
        return range(int(a[0]), int(a[1]))
    return range(int(a))


def ndrange(*args):
    import itertools
    rs = [_rng(a) for a in args]
    if len(rs) == 1:
        return rs[0]
    return itertools.product(*rs)


def grouped(it):
    for t in it:
        if isinstance(t, tuple):
            yield Vector(list(t))
        else:
            yield Vector([t])


# --------------------------------------------------------------------------
# vector / matrix values
class _Arr(_np.ndarray):
    """ndarray subclass with the Taichi Vector/Matrix method names."""

    def __new__(cls, data, dt=None):
        a = _np.array(data, dtype=_dt(dt)) if dt is not None else _np.array(data)
        if a.dtype == _np.int64:
            a = a.astype(_np.int32)
        if a.dtype == _np.bool_:
            a = a.astype(_np.int32)
        return a.view(cls)

    # value helpers ------------------------------------------------------
    def dot(self, o):
        return _np.float64(_np.dot(_np.asarray(self), _np.asarray(o)))

    def cross(self, o):
        a = _np.asarray(self)
        b = _np.asarray(o)
        return _np.array([a[1] * b[2] - a[2] * b[1],
                          a[2] * b[0] - a[0] * b[2],
                          a[0] * b[1] - a[1] * b[0]]).view(_Arr)

    def norm(self):
        a = _np.asarray(self, dtype=_np.float64)
        return _np.float64(_math.sqrt(float((a * a).sum())))

    def normalized(self):
        with _np.errstate(all="ignore"):
            return (_np.asarray(self) / self.norm()).view(_Arr)

    def outer_product(self, o):
        return _np.outer(_np.asarray(self), _np.asarray(o)).view(_Arr)

    def transpose(self):
        return _np.asarray(self).T.copy().view(_Arr)

    def inverse(self):
        return _np.linalg.inv(_np.asarray(self, dtype=_np.float64)).view(_Arr)

    def determinant(self):
        a = _np.asarray(self, dtype=_np.float64)
        if a.shape == (2, 2):
            return a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]
        if a.shape == (3, 3):
            return (a[0, 0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1])
                    - a[0, 1] * (a[1, 0] * a[2, 2] - a[1, 2] * a[2, 0])
                    + a[0, 2] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]))
        return _np.linalg.det(a)

    def trace(self):
        return _np.trace(_np.asarray(self))

    # swizzles -----------------------------------------------------------
    @property
    def x(self):
        return self[0]

    @x.setter
    def x(self, v):
        self[0] = v

    @property
    def y(self):
        return self[1]

    @y.setter
    def y(self, v):
        self[1] = v

    @property
    def z(self):
        return self[2]

    @z.setter
    def z(self, v):
        self[2] = v

    def __iter__(self):
        if self.ndim == 1:
            return iter([self[i] for i in range(self.shape[0])])
        return super().__iter__()

    def __bool__(self):
        if self.size == 1:
            return bool(self.reshape(-1)[0])
        return bool(_np.all(self))


class Vector(_Arr):
    def __new__(cls, data, dt=None):
        return _Arr.__new__(_Arr, data, dt)

    @staticmethod
    def field(n, dtype, shape=None, **kw):
        return Field(_dt(dtype), shape, (n,))


class Matrix(_Arr):
    def __new__(cls, data, dt=None):
        return _Arr.__new__(_Arr, data, dt)

    @staticmethod
    def field(n, m, dtype, shape=None, **kw):
        return Field(_dt(dtype), shape, (n, m))

    @staticmethod
    def identity(dt, n):
        return _np.eye(n, dtype=_dt(dt)).view(_Arr)

    @staticmethod
    def zero(dt, n, m=None):
        return _np.zeros((n, m) if m else (n,), dtype=_dt(dt)).view(_Arr)

    @staticmethod
    def cols(cs):
        return _np.stack([_np.asarray(c) for c in cs], axis=1).view(_Arr)

    @staticmethod
    def rows(rs):
        return _np.stack([_np.asarray(r) for r in rs], axis=0).view(_Arr)


class _VecType:
    def __init__(self, n, dt):
        self.n, self.dt = n, _dt(dt)

    def __call__(self, *a):
        if len(a) == 1 and not isinstance(a[0], (list, tuple, _np.ndarray)):
            return _np.full((self.n,), a[0], dtype=self.dt).view(_Arr)
        if len(a) == 1:
            return _np.array(a[0], dtype=self.dt).view(_Arr)
        return _np.array(a, dtype=self.dt).view(_Arr)


class _MatType:
    def __init__(self, n, m, dt):
        self.n, self.m, self.dt = n, m, _dt(dt)

    def __call__(self, *a):
        if len(a) == 1 and not isinstance(a[0], (list, tuple, _np.ndarray)):
            return _np.full((self.n, self.m), a[0], dtype=self.dt).view(_Arr)
        return _np.array(a[0] if len(a) == 1 else a, dtype=self.dt).reshape(self.n, self.m).view(_Arr)


class types:  # noqa: N801
    @staticmethod
    def ndarray(*a, **k):
        return None

    @staticmethod
    def vector(n, dt):
        return _VecType(n, dt)

    @staticmethod
    def matrix(n, m, dt):
        return _MatType(n, m, dt)


math.vec3 = _VecType(3, _np.float64)
math.vec2 = _VecType(2, _np.float64)


# --------------------------------------------------------------------------
# fields
class Field:
    def __init__(self, dtype, shape, eshape=()):
        self.dtype = _dt(dtype)
        self.eshape = tuple(eshape)
        self.arr = None
        self.shape = None
        if shape is not None:
            self._alloc(shape)

    def _alloc(self, shape):
        if isinstance(shape, (int, _np.integer)):
            shape = (int(shape),)
        shape = tuple(int(s) for s in shape)
        self.shape = shape
        self.arr = _np.zeros(shape + self.eshape, dtype=self.dtype)

    def _ix(self, idx):
        if idx is None:
            return ()
        if isinstance(idx, tuple):
            return tuple(int(i) for i in idx)
        if isinstance(idx, _np.ndarray) and idx.ndim == 1:
            return tuple(int(i) for i in idx)
        return (int(idx),)

    def __getitem__(self, idx):
        ix = self._ix(idx)
        if len(ix) != len(self.shape):
            raise IndexError(f"field of rank {len(self.shape)} indexed with {ix}")
        for i, s in zip(ix, self.shape):
            if i < 0 or i >= s:
                raise IndexError(f"field index {ix} out of range {self.shape}")
        v = self.arr[ix]
        if self.eshape:
            return v.view(_Arr)
        return v

    def __setitem__(self, idx, val):
        ix = self._ix(idx)
        for i, s in zip(ix, self.shape):
            if i < 0 or i >= s:
                raise IndexError(f"field index {ix} out of range {self.shape}")
        self.arr[ix] = val

    def __iter__(self):
        if len(self.shape) == 1:
            return iter(range(self.shape[0]))
        import itertools
        return itertools.product(*[range(s) for s in self.shape])

    def fill(self, v):
        self.arr[...] = v

    def to_numpy(self, dtype=None):
        a = self.arr.copy()
        return a.astype(dtype) if dtype is not None else a

    def from_numpy(self, a):
        self.arr[...] = _np.asarray(a).reshape(self.arr.shape)

    def to_torch(self, device=None):
        import torch
        return torch.from_numpy(self.arr.copy())

    def from_torch(self, t):
        self.from_numpy(t.detach().cpu().numpy())

    def copy_from(self, o):
        self.arr[...] = o.arr


def field(dtype, shape=None, **kw):
    return Field(_dt(dtype), shape, ())


class _Axes:
    def __init__(self, n):
        self.n = n


i = _Axes(1)
ij = _Axes(2)
ijk = _Axes(3)


class _Dense:
    def __init__(self, shape):
        self.shape = shape

    def place(self, *fields):
        for f in fields:
            f._alloc(self.shape)


class _Root:
    def dense(self, axes, shape):
        return _Dense(shape)


root = _Root()


# --------------------------------------------------------------------------
# scalar / elementwise functions
def _wrap(r):
    if isinstance(r, _np.ndarray) and r.ndim > 0:
        return r.view(_Arr)
    return r


def abs(x):  # noqa: A001
    return _wrap(_np.abs(x))


def sqrt(x):
    with _np.errstate(all="ignore"):
        return _wrap(_np.sqrt(x))


def sin(x):
    return _wrap(_np.sin(x))


def cos(x):
    return _wrap(_np.cos(x))


def acos(x):
    with _np.errstate(all="ignore"):
        return _wrap(_np.arccos(x))


def exp(x):
    return _wrap(_np.exp(x))


def log(x):
    with _np.errstate(all="ignore"):
        return _wrap(_np.log(x))


def floor(x, dt=None):
    r = _np.floor(x)
    if dt is not None:
        if isinstance(r, _np.ndarray) and r.ndim > 0:
            return r.astype(_dt(dt)).view(_Arr)
        return int(r)
    return _wrap(r)


def cast(x, dt):
    dt = _dt(dt)
    if isinstance(x, _np.ndarray) and x.ndim > 0:
        if _np.issubdtype(dt, _np.integer):
            return _np.trunc(x).astype(dt).view(_Arr)
        return x.astype(dt).view(_Arr)
    if _np.issubdtype(dt, _np.integer):
        return int(x)  # truncation toward zero, as C casts do
    return dt(x)


def max(*a):  # noqa: A001
    r = a[0]
    for b in a[1:]:
        r = _np.maximum(r, b)
    return _wrap(r)


def min(*a):  # noqa: A001
    r = a[0]
    for b in a[1:]:
        r = _np.minimum(r, b)
    return _wrap(r)


def svd(A):
    U, s, Vt = _np.linalg.svd(_np.asarray(A, dtype=_np.float64))
    return U.view(_Arr), _np.diag(s).view(_Arr), Vt.T.copy().view(_Arr)


def sym_eig(A):
    w, v = _np.linalg.eigh(_np.asarray(A, dtype=_np.float64))
    return w.view(_Arr), v.view(_Arr)


# helpers the import hook rewrites atomics / assignments into -----------------
def _emu_val(x):
    """value-copy semantics for `a = field[i]` inside kernels"""
    if isinstance(x, _np.ndarray):
        return x.copy() if not isinstance(x, _Arr) else x.copy().view(_Arr)
    if isinstance(x, tuple):
        return tuple(_emu_val(e) for e in x)
    return x


def _emu_atomic(op, obj, idx, v):
    old = obj[idx]
    if isinstance(old, _np.ndarray):
        old = old.copy()
    if op == "add":
        obj[idx] = old + v
    elif op == "max":
        obj[idx] = _np.maximum(old, v)
    elif op == "min":
        obj[idx] = _np.minimum(old, v)
    else:  # pragma: no cover
        raise ValueError(op)
    return old


def _emu_combine(op, old, v):
    if op == "add":
        return old + v
    if op == "max":
        return _np.maximum(old, v)
    if op == "min":
        return _np.minimum(old, v)
    raise ValueError(op)  # pragma: no cover


def atomic_add(*a):  # pragma: no cover - always rewritten by the hook
    raise RuntimeError("atomic op reached un-rewritten code")


atomic_max = atomic_min = atomic_add
