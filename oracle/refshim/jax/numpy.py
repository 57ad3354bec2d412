"""jax.numpy stand-in on torch CPU tensors (TEST INFRASTRUCTURE ONLY; see ../README.md).

Only the names the reference touches through `import jax.numpy as np`
(trajectory_experiments.py:124-134,145-152,184-203,217,279,284,288;
scone_trajectory_model.py:46-56,63-71,81-90,118,125) are provided.
dtype policy mirrors JAX with x64 disabled: floats -> float32, ints -> int64 (index-safe).
"""
import os as _os
import numpy as _onp
import torch as _torch

_X64 = _os.environ.get('REFSHIM_X64', '0') == '1'
_FDT = _torch.float64 if _X64 else _torch.float32
_torch.set_grad_enabled(True)


def _as_tensor(x):
    if isinstance(x, _torch.Tensor):
        if x.is_floating_point() and x.dtype != _FDT:
            return x.to(_FDT)
        return x
    a = _onp.asarray(x)
    if a.dtype == object:
        raise TypeError('object array')
    if a.dtype.kind == 'f':
        return _torch.from_numpy(_onp.ascontiguousarray(a)).to(_FDT)
    if a.dtype.kind == 'b':
        return _torch.from_numpy(_onp.ascontiguousarray(a))
    return _torch.from_numpy(_onp.ascontiguousarray(a).astype(_onp.int64))


def array(x, dtype=None):
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], _torch.Tensor):
        return _torch.stack([_as_tensor(v) for v in x])
    return _as_tensor(x)


asarray = array


def zeros(shape, dtype=None):
    if isinstance(shape, int):
        shape = (shape,)
    return _torch.zeros(tuple(int(s) for s in shape), dtype=_FDT)


def diag(v):
    return _torch.diag(_as_tensor(v))


def append(a, b, axis=None):
    a, b = _as_tensor(a), _as_tensor(b)
    if axis is None:
        return _torch.cat([a.reshape(-1), b.reshape(-1)])
    return _torch.cat([a, b.to(a.dtype)], dim=axis)


def load(*a, **k):
    r = _onp.load(*a, **k)
    return r if r.dtype == object else _as_tensor(r)


def maximum(a, b):
    a = _as_tensor(a)
    b = _as_tensor(b).to(a.dtype) if not isinstance(b, (int, float)) else _torch.tensor(b, dtype=a.dtype)
    return _torch.maximum(a, b)


def where(c, a, b):
    return _torch.where(c, a, b)


def exp(x):
    return _torch.exp(_as_tensor(x))


def tanh(x):
    return _torch.tanh(_as_tensor(x))


def sum(x, axis=None):
    t = _as_tensor(x)
    if t.dtype == _torch.bool:
        t = t.to(_torch.int64)
    return t.sum() if axis is None else t.sum(dim=axis)


def _f(t):
    return t.to(_FDT) if not t.is_floating_point() else t


def mean(x, axis=None):
    t = _f(_as_tensor(x))
    return t.mean() if axis is None else t.mean(dim=axis)


def average(x, axis=None):
    if isinstance(x, (list, tuple)):
        x = _onp.asarray([float(v) for v in x])
    return mean(x, axis)


def argmax(x, axis=None):
    t = _as_tensor(x)
    return t.argmax() if axis is None else t.argmax(dim=axis)


class _Linalg:
    @staticmethod
    def norm(x):
        if isinstance(x, (list, tuple)):
            x = _torch.stack([_as_tensor(v) for v in x])     # same stacking JAX does for a list
        return _torch.sqrt(_torch.sum(_as_tensor(x) ** 2))


linalg = _Linalg()
