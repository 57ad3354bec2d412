"""Torch-backed stand-in for the tiny part of JAX the reference uses (TEST INFRASTRUCTURE ONLY).

Call sites in the reference: jax.vmap (scone_trajectory_model.py:256), jax.grad (:291,:307),
jax.jit (:289, dead code).  See oracle/refshim/README.md.
"""
import functools
import numpy as _onp
import torch as _torch
from . import numpy  # noqa: F401  (jax.numpy)
from .numpy import _as_tensor


class _NpFriendly(_torch.Tensor):
    """Tensor whose NumPy ufunc results stay ndarrays (as they do for jax arrays), so the reference's
    host-side bookkeeping `onp.mean(onp.abs(g[i]))` (scone_trajectory_model.py:308-309) works."""
    def __array_wrap__(self, arr, context=None, return_scalar=False):
        return arr

    def __format__(self, spec):
        return format(self.item(), spec) if self.dim() == 0 else object.__format__(self, spec)


def jit(f):
    return f


def vmap(fun, in_axes=0):
    """Per-sample loop + stack; unmapped (None) array arguments are converted to tensors once."""
    @functools.wraps(fun)
    def batched(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(axes) == len(args), 'in_axes / args mismatch'
        conv = []
        n = None
        for a, ax in zip(args, axes):
            if callable(a):
                conv.append(a)
                continue
            if isinstance(a, (list, tuple)) and ax is None:
                conv.append([_as_tensor(x) for x in a])
                continue
            t = _as_tensor(a)
            conv.append(t)
            if ax is not None:
                assert ax == 0
                n = t.shape[0] if n is None else n
                assert t.shape[0] == n
        outs = []
        for i in range(n):
            call = [c if ax is None else c[i] for c, ax in zip(conv, axes)]
            outs.append(fun(*call))
        return _torch.stack(outs)
    return batched


def grad(fun, argnums=0):
    assert argnums == 0

    def gradfun(first, *rest):
        is_list = isinstance(first, (list, tuple)) or (isinstance(first, _onp.ndarray) and first.dtype == object)
        leaves = [_as_tensor(w).detach().clone().requires_grad_(True) for w in (first if is_list else [first])]
        out = fun(leaves if is_list else leaves[0], *rest)
        gs = _torch.autograd.grad(out, leaves, allow_unused=True)
        gs = [(_torch.zeros_like(l) if g is None else g) for g, l in zip(gs, leaves)]
        gs = [g.detach().as_subclass(_NpFriendly) for g in gs]
        return gs if is_list else gs[0]
    return gradfun
