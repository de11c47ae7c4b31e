"""Stand-in for the part of jax 0.4.28 the reference touches, on torch fp64 CPU tensors.

``jax.numpy`` arrays are ``torch.Tensor`` (float64 / int64: the reference enables x64,
dataset.py:18, trainer.py:32).  ``jax.vmap`` is ``torch.vmap`` (a real batching transform: the
reference's per-pair kernel code is traced once and runs on whole axes, exactly as under JAX),
``jax.value_and_grad`` is reverse-mode ``torch.autograd`` over the parameter leaves of a gpjax
``Module``, and the ``lax`` control-flow primitives run as Python control flow.

Four methods are added to / wrapped on ``torch.Tensor`` for the process that installs the shim,
because the reference calls them on arrays:
  ``.astype(dtype)``, ``.at[i].set(v)`` (out-of-bounds updates are dropped, as in JAX),
  ``.repeat(n, None)`` in its numpy meaning (model.py:147), and integer-array indexing with JAX's
  gather rule (negative indices wrap, out-of-range indices clamp: SURVEY.md Q6).
"""
from __future__ import annotations

import math
import sys
import types

import numpy as np
import torch

F64 = torch.float64
I64 = torch.int64


def _dtype(dt):
    if dt is None:
        return None
    if dt is int:
        return I64
    if dt is float:
        return F64
    if dt is bool:
        return torch.bool
    if isinstance(dt, torch.dtype):
        return dt
    return {np.dtype("float64"): F64, np.dtype("int64"): I64, np.dtype("int32"): I64,
            np.dtype("bool"): torch.bool}[np.dtype(dt)]


def asarray(obj, dtype=None):
    """jnp.array / jnp.asarray: nested lists and tuples of tensors, numpy arrays and scalars."""
    dt = _dtype(dtype)
    if isinstance(obj, torch.Tensor):
        out = obj
    elif isinstance(obj, np.ndarray):
        a = obj
        if a.dtype.kind == "f":
            a = a.astype(np.float64)
        elif a.dtype.kind in "iu":
            a = a.astype(np.int64)
        out = torch.from_numpy(np.ascontiguousarray(a))
    elif isinstance(obj, (list, tuple)):
        parts = [asarray(o) for o in obj]
        if not parts:
            out = torch.zeros(0, dtype=F64)
        else:
            rt = parts[0].dtype
            for p in parts[1:]:
                rt = torch.promote_types(rt, p.dtype)
            out = torch.stack([p.to(rt) for p in parts])
    elif isinstance(obj, bool):
        out = torch.tensor(obj)
    elif isinstance(obj, (int, np.integer)):
        out = torch.tensor(int(obj), dtype=I64)
    elif isinstance(obj, (float, np.floating)):
        out = torch.tensor(float(obj), dtype=F64)
    elif hasattr(obj, "to_dense"):
        out = obj.to_dense()
    else:
        raise TypeError(f"refshim jnp.array: unsupported {type(obj)}")
    if dt is not None and out.dtype != dt:
        out = out.to(dt)
    return out


def _t(x):
    return x if isinstance(x, torch.Tensor) or hasattr(x, "__torch_function__") else asarray(x)


# ----------------------------------------------------------------------------------------------
# torch.Tensor additions
# ----------------------------------------------------------------------------------------------
class _AtIndex:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, value):
        out = self.arr.clone()
        idx = self.idx
        if isinstance(idx, int) and out.dim() >= 1 and not (-out.shape[0] <= idx < out.shape[0]):
            return out  # JAX drops out-of-bounds scatter updates
        out[idx] = _t(value).to(out.dtype) if not isinstance(value, torch.Tensor) else value.to(out.dtype)
        return out


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self.arr, idx)


_PATCHED = False


def _patch_tensor():
    global _PATCHED
    if _PATCHED:
        return
    _PATCHED = True
    orig_repeat = torch.Tensor.repeat
    orig_getitem = torch.Tensor.__getitem__

    def astype(self, dt):
        return self.to(_dtype(dt))

    def repeat(self, *args, **kwargs):
        if (len(args) == 2 and args[1] is None) or "axis" in kwargs:
            axis = kwargs.get("axis", args[1] if len(args) == 2 else None)
            n = args[0] if args else kwargs["repeats"]
            if axis is None:
                return self.reshape(-1).repeat_interleave(int(n))
            return self.repeat_interleave(int(n), dim=axis)
        return orig_repeat(self, *args, **kwargs)

    def getitem(self, idx):
        if isinstance(idx, torch.Tensor) and idx.dtype in (I64, torch.int32) and self.dim() >= 1:
            n = self.shape[0]
            idx = torch.where(idx < 0, idx + n, idx).clamp(0, n - 1)
        return orig_getitem(self, idx)

    torch.Tensor.astype = astype
    torch.Tensor.at = property(lambda self: _At(self))
    torch.Tensor.repeat = repeat
    torch.Tensor.__getitem__ = getitem


# ----------------------------------------------------------------------------------------------
# jax.numpy
# ----------------------------------------------------------------------------------------------
def _jnp_module():
    m = types.ModuleType("jax.numpy")
    m.pi = math.pi
    m.newaxis = None
    m.float64 = F64
    m.int64 = I64
    m.ndarray = torch.Tensor
    m.array = asarray
    m.asarray = asarray

    def linspace(start, stop, num=50):
        return torch.from_numpy(np.linspace(start, stop, num, dtype=np.float64))

    def arange(*args, dtype=None):
        return torch.arange(*args, dtype=_dtype(dtype) or (I64 if all(isinstance(a, int) for a in args) else F64))

    def tile(a, reps):
        return torch.tile(_t(a), (reps,) if isinstance(reps, int) else tuple(reps))

    def repeat(a, repeats, axis=None):
        a = _t(a)
        if a.dim() == 0:
            return a.expand(int(repeats)).clone()
        if axis is None:
            return a.reshape(-1).repeat_interleave(int(repeats))
        return a.repeat_interleave(int(repeats), dim=axis)

    def stack(seq, axis=0):
        parts = [_t(s) for s in seq]
        rt = parts[0].dtype
        for p in parts[1:]:
            rt = torch.promote_types(rt, p.dtype)
        return torch.stack([p.to(rt) for p in parts], dim=axis)

    def concatenate(seq, axis=0):
        parts = [_t(s) for s in seq]
        rt = parts[0].dtype
        for p in parts[1:]:
            rt = torch.promote_types(rt, p.dtype)
        return torch.cat([p.to(rt) for p in parts], dim=axis)

    def ones(shape, dtype=None):
        return torch.ones(shape, dtype=_dtype(dtype) or F64)

    def zeros(shape, dtype=None):
        return torch.zeros(shape, dtype=_dtype(dtype) or F64)

    def _f(x):
        x = _t(x)
        return x if (not isinstance(x, torch.Tensor)) or x.is_floating_point() else x.to(F64)

    m.linspace, m.arange, m.tile, m.repeat, m.stack = linspace, arange, tile, repeat, stack
    m.concatenate, m.ones, m.zeros = concatenate, ones, zeros
    m.sqrt = lambda x: torch.sqrt(_f(x))
    m.exp = lambda x: torch.exp(_f(x))
    m.log = lambda x: torch.log(_f(x))
    m.ceil = lambda x: math.ceil(x) if isinstance(x, (int, float)) else torch.ceil(x)
    m.square = lambda x: torch.square(_t(x))
    m.divide = lambda a, b: torch.divide(_t(a), _t(b))
    m.multiply = lambda a, b: torch.multiply(_t(a), _t(b))
    m.where = lambda c, a, b: torch.where(c, _t(a), _t(b))
    m.diag = lambda a: torch.diag(_t(a))
    m.matmul = lambda a, b: torch.matmul(_t(a), _t(b))
    m.atleast_1d = lambda a: torch.atleast_1d(_t(a))
    m.sum = lambda a, axis=None: torch.sum(_t(a)) if axis is None else torch.sum(_t(a), dim=axis)
    m.abs = lambda a: torch.abs(_t(a))
    return m


# ----------------------------------------------------------------------------------------------
# transforms
# ----------------------------------------------------------------------------------------------
def vmap(fun, in_axes=0, out_axes=0):
    return torch.vmap(fun, in_dims=in_axes, out_dims=out_axes)


def value_and_grad(fun):
    """Differentiate ``fun`` with respect to the parameter leaves of its first argument (a gpjax
    ``Module`` stand-in).  Returns (value, Module-shaped gradient), like jax.value_and_grad on a
    pytree (trainer.py:126)."""
    def wrapped(tree, *args, **kwargs):
        names = tree._leaf_names()
        leaves = [getattr(tree, n).detach().clone().requires_grad_(True) for n in names]
        traced = tree.replace(**dict(zip(names, leaves)))
        out = fun(traced, *args, **kwargs)
        grads = torch.autograd.grad(out, leaves, allow_unused=True)
        grads = [torch.zeros_like(l) if g is None else g for g, l in zip(grads, leaves)]
        return out.detach(), tree.replace(**dict(zip(names, grads)))
    return wrapped


def _index_tree(xs, i):
    if isinstance(xs, (tuple, list)):
        return tuple(_index_tree(x, i) for x in xs)
    return xs[i]


def _tree_len(xs):
    if isinstance(xs, (tuple, list)):
        return _tree_len(xs[0])
    return len(xs)


def scan(f, init, xs, length=None):
    carry, ys = init, []
    for i in range(_tree_len(xs) if xs is not None else length):
        carry, y = f(carry, _index_tree(xs, i) if xs is not None else None)
        ys.append(y)
    return carry, torch.stack([_t(y) for y in ys])


def register():
    _patch_tensor()
    jax = types.ModuleType("jax")
    jax.__path__ = []
    jnp = _jnp_module()
    jax.numpy = jnp
    jax.vmap = vmap
    jax.value_and_grad = value_and_grad
    jax.grad = lambda fun: (lambda *a, **k: value_and_grad(fun)(*a, **k)[1])
    jax.jit = lambda fun, *a, **k: fun
    jax.Array = torch.Tensor

    config = types.SimpleNamespace(update=lambda *a, **k: None)
    jax.config = config

    random = types.ModuleType("jax.random")
    random.PRNGKey = lambda seed: torch.tensor([0, int(seed)], dtype=I64)
    random.split = lambda key, num=2: torch.zeros((num, 2), dtype=I64)
    jax.random = random

    lax = types.ModuleType("jax.lax")
    lax.cond = lambda pred, tf, ff, *ops: tf(*ops) if bool(pred) else ff(*ops)
    lax.scan = scan
    lax.stop_gradient = lambda x: x.detach()
    jax.lax = lax

    scipy = types.ModuleType("jax.scipy")
    scipy.__path__ = []
    special = types.ModuleType("jax.scipy.special")
    special.erf = lambda x: torch.special.erf(_t(x))
    scipy.special = special
    jax.scipy = scipy

    for name, mod in (("jax", jax), ("jax.numpy", jnp), ("jax.random", random), ("jax.lax", lax),
                      ("jax.scipy", scipy), ("jax.scipy.special", special)):
        sys.modules[name] = mod
