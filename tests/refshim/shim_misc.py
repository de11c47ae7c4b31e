"""Stand-ins for cola-ml 0.0.5, optax 0.1.9, the tfp bijectors, jaxtyping and matplotlib ([3P],
semantics restated from the pinned versions' published behaviour).

* cola: ``Dense`` wraps a matrix; sums / differences / scalar products of operators and arrays
  stay operators (``to_dense()`` gives the matrix back); ``PSD`` is an annotation; ``I_like(A)``
  is the identity of A's shape; ``inv`` and ``solve`` have dense Cholesky semantics (the matrices
  the reference passes are symmetric positive definite: model.py:446-450, :497-504).
* optax: ``adam(lr)`` = scale_by_adam(b1=.9, b2=.999, eps=1e-8, eps_root=0) then scale(-lr);
  ``apply_updates`` adds leaf by leaf.
* tfp: ``Softplus`` (forward log(1+e^x), inverse x + log(-expm1(-x))), ``Sigmoid(low, high)``
  (forward low + (high-low)*sigmoid(x), inverse logit((y-low)/(high-low))).
"""
from __future__ import annotations

import sys
import types

import torch

from shim_gpjax import tree_map


# ----------------------------------------------------------------------------------------------
# cola
# ----------------------------------------------------------------------------------------------
def _dense(x):
    return x.A if isinstance(x, Dense) else x


class Dense:
    def __init__(self, A):
        self.A = _dense(A)

    shape = property(lambda self: self.A.shape)
    dtype = property(lambda self: self.A.dtype)
    T = property(lambda self: Dense(self.A.T))

    def to_dense(self):
        return self.A

    def __add__(self, other):
        return Dense(self.A + _dense(other))

    __radd__ = __add__
    __iadd__ = __add__

    def __sub__(self, other):
        return Dense(self.A - _dense(other))

    def __rsub__(self, other):
        return Dense(_dense(other) - self.A)

    def __mul__(self, c):
        return Dense(self.A * _dense(c))

    __rmul__ = __mul__

    def __matmul__(self, other):
        out = self.A @ _dense(other)
        return Dense(out) if isinstance(other, Dense) else out

    def __neg__(self):
        return Dense(-self.A)

    @classmethod
    def __torch_function__(cls, func, types_, args=(), kwargs=None):
        # an array on the left of an operator expression (``array += operator``, model.py:461)
        unwrap = lambda a: a.A if isinstance(a, Dense) else a
        out = func(*[unwrap(a) for a in args], **{k: unwrap(v) for k, v in (kwargs or {}).items()})
        return Dense(out) if isinstance(out, torch.Tensor) and out.dim() == 2 else out


def I_like(A):
    n = A.shape[-1]
    return Dense(torch.eye(n, dtype=torch.float64))


def _chol_inverse(A):
    return torch.cholesky_inverse(torch.linalg.cholesky(A))


def inv(A):
    return Dense(_chol_inverse(_dense(A)))


def solve(A, b):
    return torch.cholesky_solve(_dense(b), torch.linalg.cholesky(_dense(A)))


# ----------------------------------------------------------------------------------------------
# optax
# ----------------------------------------------------------------------------------------------
class GradientTransformation:
    def __init__(self, init, update):
        self.init, self.update = init, update


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    def init(params):
        zeros = tree_map(torch.zeros_like, params)
        return {"count": 0, "mu": zeros, "nu": tree_map(torch.zeros_like, params)}

    def update(grads, state, params=None):
        mu = tree_map(lambda g, m: (1 - b1) * g + b1 * m, grads, state["mu"])
        nu = tree_map(lambda g, v: (1 - b2) * (g * g) + b2 * v, grads, state["nu"])
        count = state["count"] + 1
        c1, c2 = 1 - b1 ** count, 1 - b2 ** count
        updates = tree_map(lambda m, v: -learning_rate * ((m / c1) / (torch.sqrt(v / c2 + eps_root) + eps)), mu, nu)
        return updates, {"count": count, "mu": mu, "nu": nu}

    return GradientTransformation(init, update)


def apply_updates(params, updates):
    return tree_map(lambda p, u: p + u, params, updates)


# ----------------------------------------------------------------------------------------------
# tfp bijectors
# ----------------------------------------------------------------------------------------------
class Identity:
    def forward(self, x):
        return x

    def inverse(self, y):
        return y


class Softplus:
    def forward(self, x):
        return torch.nn.functional.softplus(x, beta=1.0, threshold=1e9)

    def inverse(self, y):
        return y + torch.log(-torch.expm1(-y))


class Sigmoid:
    def __init__(self, low=0.0, high=1.0):
        self.low, self.high = low, high

    def forward(self, x):
        return self.low + (self.high - self.low) * torch.sigmoid(x)

    def inverse(self, y):
        u = (y - self.low) / (self.high - self.low)
        return torch.log(u) - torch.log1p(-u)


# ----------------------------------------------------------------------------------------------
# inert placeholders
# ----------------------------------------------------------------------------------------------
class _Subscriptable:
    def __class_getitem__(cls, item):
        return cls


class _Inert(types.ModuleType):
    """matplotlib placeholder: every attribute is a callable that returns another placeholder."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _InertObj()


class _InertObj:
    def __call__(self, *a, **k):
        return _InertObj()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _InertObj()

    def __getitem__(self, key):
        return ["C0", "C1", "C2", "C3"] if key == "color" else _InertObj()

    def __iter__(self):
        return iter(())


def register():
    cola = types.ModuleType("cola")
    cola.__path__ = []
    ops = types.ModuleType("cola.ops")
    ops.Dense, ops.LinearOperator, ops.I_like = Dense, Dense, I_like
    cola.ops, cola.PSD, cola.inv, cola.solve = ops, (lambda A: A if isinstance(A, Dense) else Dense(A)), inv, solve
    sys.modules["cola"], sys.modules["cola.ops"] = cola, ops

    optax = types.ModuleType("optax")
    optax.adam, optax.apply_updates, optax.GradientTransformation = adam, apply_updates, GradientTransformation
    sys.modules["optax"] = optax

    chain = ["tensorflow_probability", "tensorflow_probability.substrates",
             "tensorflow_probability.substrates.jax", "tensorflow_probability.substrates.jax.bijectors"]
    mods = []
    for name in chain:
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod
        mods.append(mod)
    for parent, child, name in zip(mods, mods[1:], chain[1:]):
        setattr(parent, name.rsplit(".", 1)[1], child)
    mods[-1].Softplus, mods[-1].Sigmoid, mods[-1].Identity = Softplus, Sigmoid, Identity

    jaxtyping = types.ModuleType("jaxtyping")
    for name in ("Float", "Num", "Int", "Array", "Bool"):
        setattr(jaxtyping, name, type(name, (_Subscriptable,), {}))
    sys.modules["jaxtyping"] = jaxtyping

    mpl = _Inert("matplotlib")
    mpl.__path__ = []
    plt = _Inert("matplotlib.pyplot")
    mpl.pyplot = plt
    mpl.rcParams = _InertObj()
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
