"""Minimal stand-ins for the third-party types the reference's call sites use (gpjax 0.8.2,
cola-ml 0.0.5, optax 0.1.9 -- none installable here).  Only the members the reference
actually touches on the hot path are provided:

  gpx.Dataset(X, y)                       main.py:38, objectives.py:64
  GaussianDistribution(loc, scale)        model.py:463 -> .mean(), .stddev(), .variance()   (plotter.py:62-63)
  cola Dense / PSD wrapper                model.py:414 -> .to_dense()
  optax.adam(lr) -> init / update          main.py:45, trainer.py:127-128, 199
  optax.apply_updates                      trainer.py:128
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, NamedTuple, Optional

import numpy as np


def _np(a) -> np.ndarray:
    try:
        import torch

        if isinstance(a, torch.Tensor):
            return a.detach().cpu().numpy()
    except Exception:  # pragma: no cover
        pass
    return np.asarray(a)


@dataclass
class Dataset:
    """gpjax.Dataset: X (N, D) inputs, y (N, Q) outputs, same leading dimension."""

    X: Optional[Any] = None
    y: Optional[Any] = None

    def __post_init__(self):
        if self.X is not None and self.y is not None:
            X, y = _np(self.X), _np(self.y)
            if X.ndim != 2 or y.ndim != 2:
                raise ValueError(f"Inputs, X, and outputs, y, must both be 2-dimensional. Got X.ndim={X.ndim} and y.ndim={y.ndim}.")
            if X.shape[0] != y.shape[0]:
                raise ValueError(f"Inputs, X, and outputs, y, must have the same number of rows. Got X.shape[0]={X.shape[0]} and y.shape[0]={y.shape[0]}.")

    @property
    def n(self) -> int:
        return int(_np(self.X).shape[0])


class DenseOperator:
    """cola.PSD(Dense(K)) stand-in: holds the dense (device) matrix; `.to_dense()` returns it."""

    def __init__(self, dense):
        self._dense = dense

    def to_dense(self):
        return self._dense

    @property
    def shape(self):
        return tuple(self._dense.shape)

    def __add__(self, other):
        o = other.to_dense() if isinstance(other, DenseOperator) else other
        return DenseOperator(self._dense + o)

    __radd__ = __add__


class GaussianDistribution:
    """gpjax.distributions.GaussianDistribution with a diagonal or dense covariance.

    `scale` may be a 1-D array (the diagonal; what latent_predict keeps, model.py:460-461) or a
    2-D covariance.
    """

    def __init__(self, loc, scale):
        self.loc = _np(loc).astype(np.float64).reshape(-1)
        self.scale = _np(scale).astype(np.float64)

    def mean(self) -> np.ndarray:
        return self.loc

    def variance(self) -> np.ndarray:
        return self.scale if self.scale.ndim == 1 else np.diag(self.scale)

    def stddev(self) -> np.ndarray:
        return np.sqrt(self.variance())

    def covariance(self) -> np.ndarray:
        return np.diag(self.scale) if self.scale.ndim == 1 else self.scale


# ---- optax.adam ---------------------------------------------------------------------------------
class AdamState(NamedTuple):
    count: int
    mu: np.ndarray
    nu: np.ndarray


@dataclass(frozen=True)
class GradientTransformation:
    """optax.GradientTransformation for adam: `init(params)`, `update(grads, state, params)`."""

    learning_rate: float
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8
    name: str = "adam"

    def init(self, params) -> AdamState:
        p = np.asarray(_flat(params), dtype=np.float64)
        return AdamState(0, np.zeros_like(p), np.zeros_like(p))

    def update(self, grads, state: AdamState, params=None):
        g = np.asarray(_flat(grads), dtype=np.float64)
        count = state.count + 1
        mu = self.b1 * state.mu + (1.0 - self.b1) * g
        nu = self.b2 * state.nu + (1.0 - self.b2) * g * g
        mu_hat = mu / (1.0 - self.b1**count)
        nu_hat = nu / (1.0 - self.b2**count)
        updates = -self.learning_rate * mu_hat / (np.sqrt(nu_hat) + self.eps)
        return updates, AdamState(count, mu, nu)


def adam(learning_rate: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> GradientTransformation:
    """optax.adam(learning_rate) with optax 0.1.9's defaults (eps_root = 0)."""
    return GradientTransformation(float(learning_rate), float(b1), float(b2), float(eps))


def _flat(params):
    if hasattr(params, "pack_unconstrained_leaves"):
        return params.pack_unconstrained_leaves()
    return params


def apply_updates(params, updates):
    """optax.apply_updates on the flat leaf vector (or on a model exposing `with_leaves`)."""
    if hasattr(params, "with_leaves"):
        return params.with_leaves(np.asarray(_flat(params)) + np.asarray(updates))
    return np.asarray(params) + np.asarray(updates)
