"""Stand-in for the part of gpjax 0.8.2 the reference touches (semantics restated, [3P]).

* ``gpjax.base.Module`` / ``param_field`` / ``static_field``: a dataclass whose parameter leaves
  carry a bijector and a ``trainable`` flag in the field metadata; ``constrain`` / ``unconstrain``
  map every leaf through ``bijector.forward`` / ``.inverse``; ``stop_gradient`` detaches the
  non-trainable leaves; ``replace`` copies the object and overwrites attributes (it does not go
  through ``__init__``, which is why the reference can replace ``init=False`` fields,
  trainer.py:158).
* ``gpjax.Dataset(X, y)``; ``gpjax.objectives.AbstractObjective`` (``constant`` = -1 when
  ``negative``; calling the objective calls ``step``).
* ``gpjax.distributions.GaussianDistribution(loc, scale)``: ``log_prob`` in the dense Cholesky
  form  -1/2 [ n log 2 pi + log det S + (y - mu)^T S^-1 (y - mu) ]  (SURVEY.md Q9 defines the
  dense semantics as the oracle), ``mean``, ``variance`` = diag(S), ``stddev`` = sqrt(diag(S)).
* ``gpjax.scan.vscan``: ``lax.scan`` without the progress bar.
"""
from __future__ import annotations

import copy
import dataclasses
import math
import sys
import types

import torch

import shim_jax


def param_field(default=dataclasses.MISSING, *, bijector=None, trainable=True, metadata=None, **kwargs):
    from shim_misc import Identity
    meta = dict(metadata or {})
    meta.update({"bijector": bijector if bijector is not None else Identity(), "trainable": trainable,
                 "pytree_node": True})
    if default is not dataclasses.MISSING:
        kwargs["default"] = default
    return dataclasses.field(metadata=meta, **kwargs)


def static_field(default=dataclasses.MISSING, **kwargs):
    meta = dict(kwargs.pop("metadata", None) or {})
    meta["pytree_node"] = False
    if default is not dataclasses.MISSING:
        kwargs["default"] = default
    return dataclasses.field(metadata=meta, **kwargs)


class Module:
    def _leaf_fields(self):
        return [f for f in dataclasses.fields(self) if "bijector" in f.metadata]

    def _leaf_names(self):
        return [f.name for f in self._leaf_fields()]

    def replace(self, **kwargs):
        for key in kwargs:
            if key not in vars(self):
                raise ValueError(f"'{key}' is not a field of {type(self).__name__}")
        out = copy.copy(self)
        out.__dict__.update(kwargs)
        return out

    def constrain(self):
        return self.replace(**{f.name: f.metadata["bijector"].forward(getattr(self, f.name))
                               for f in self._leaf_fields()})

    def unconstrain(self):
        return self.replace(**{f.name: f.metadata["bijector"].inverse(getattr(self, f.name))
                               for f in self._leaf_fields()})

    def stop_gradient(self):
        return self.replace(**{f.name: getattr(self, f.name).detach()
                               for f in self._leaf_fields() if not f.metadata["trainable"]})


def tree_map(fn, tree, *rest):
    """jax.tree_util.tree_map restricted to Modules (leaves = parameter fields) and tensors."""
    if isinstance(tree, Module):
        return tree.replace(**{n: fn(getattr(tree, n), *[getattr(r, n) for r in rest])
                               for n in tree._leaf_names()})
    return fn(tree, *rest)


@dataclasses.dataclass
class Dataset:
    X: torch.Tensor = None
    y: torch.Tensor = None

    def __post_init__(self):
        if self.X is not None and self.y is not None and self.X.shape[0] != self.y.shape[0]:
            raise ValueError("Inputs, X, and outputs, y, must have the same number of rows.")

    @property
    def n(self):
        return self.X.shape[0]


@dataclasses.dataclass
class AbstractObjective(Module):
    negative: bool = static_field(False)
    constant: torch.Tensor = static_field(init=False, repr=False)

    def __post_init__(self):
        self.constant = torch.tensor(-1.0 if self.negative else 1.0, dtype=torch.float64)

    def __hash__(self):
        return hash(tuple(sorted(k for k in vars(self))))

    def __call__(self, *args, **kwargs):
        return self.step(*args, **kwargs)

    def step(self, *args, **kwargs):
        raise NotImplementedError


class GaussianDistribution:
    def __init__(self, loc=None, scale=None):
        self.loc = loc
        self.scale = scale

    def _dense(self):
        return self.scale.to_dense() if hasattr(self.scale, "to_dense") else self.scale

    def mean(self):
        return self.loc

    def covariance(self):
        return self._dense()

    def variance(self):
        return torch.diagonal(self._dense())

    def stddev(self):
        return torch.sqrt(torch.diagonal(self._dense()))

    def log_prob(self, y):
        mu, sigma = self.loc, self._dense()
        n = mu.shape[-1]
        chol = torch.linalg.cholesky(sigma)
        diff = (y - mu).reshape(-1, 1)
        half = torch.linalg.solve_triangular(chol, diff, upper=False)
        logdet = 2.0 * torch.sum(torch.log(torch.diagonal(chol)))
        return -0.5 * (n * math.log(2.0 * math.pi) + logdet + torch.sum(half * half))


def register():
    gpx = types.ModuleType("gpjax")
    gpx.__path__ = []
    base = types.ModuleType("gpjax.base")
    base.Module, base.param_field, base.static_field = Module, param_field, static_field
    dataset = types.ModuleType("gpjax.dataset")
    dataset.Dataset = Dataset
    objectives = types.ModuleType("gpjax.objectives")
    objectives.AbstractObjective = AbstractObjective
    distributions = types.ModuleType("gpjax.distributions")
    distributions.GaussianDistribution = GaussianDistribution
    typing = types.ModuleType("gpjax.typing")
    typing.Array = torch.Tensor
    typing.ScalarFloat = torch.Tensor
    typing.KeyArray = torch.Tensor
    scan = types.ModuleType("gpjax.scan")
    scan.vscan = lambda f, init, xs, length=None, **kw: shim_jax.scan(f, init, xs, length)
    gpx.base, gpx.dataset, gpx.objectives, gpx.distributions = base, dataset, objectives, distributions
    gpx.typing, gpx.scan = typing, scan
    gpx.Dataset, gpx.Module = Dataset, Module
    for name, mod in (("gpjax", gpx), ("gpjax.base", base), ("gpjax.dataset", dataset),
                      ("gpjax.objectives", objectives), ("gpjax.distributions", distributions),
                      ("gpjax.typing", typing), ("gpjax.scan", scan)):
        sys.modules[name] = mod
