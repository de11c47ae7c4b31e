"""``CustomConjMLL``: the (negative) conjugate marginal log-likelihood of the LFM.

Mirror of the reference's ``src/objectives.py:19-78`` (itself a variant of gpjax's ConjugateMLL that
takes the custom model).  ``objective(model, train_data)`` evaluates
``constant * log N(y; mean_function(X), gram(X) + jitter I + obs_stddev^2 I)`` with the Cholesky,
triangular solves and log-det on the B200; ``value_and_grad`` is what ``jax.value_and_grad`` of the
trainer's loss (``src/trainer.py:126``) returns, computed by the fused CUDA path.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import ops
from .gpx_compat import Dataset
from .model import ExactLFM


class CustomConjMLL:
    """``CustomConjMLL(negative=True)`` -> callable ``(model, Dataset) -> scalar``."""

    def __init__(self, negative: bool = False):
        self.negative = bool(negative)
        self.constant = -1.0 if negative else 1.0  # gpjax AbstractObjective.constant
        self._plan = None       # CUDA-graph evaluation plan of the last (data set, jitter) seen by value_and_grad
        self._plan_key = None

    def __call__(self, model: ExactLFM, train_data: Dataset) -> float:
        return self.step(model, train_data)

    def step(self, model: ExactLFM, train_data: Dataset) -> float:
        """constant * log p(y | X, theta) for a CONSTRAINED model (reference objectives.py:64-78)."""
        val, info = ops.nlml(train_data.X, train_data.y, model.pack(), model.jitter, model.num_genes)
        nl = float(val.item())  # NaN when Sigma is not positive definite, like JAX's Cholesky
        return -self.constant * nl

    def value_and_grad(self, model_unconstrained: ExactLFM, train_data: Dataset) -> Tuple[float, np.ndarray]:
        """Loss and gradient w.r.t. the UNCONSTRAINED leaves [d, s, b, l, obs_stddev]:
        jax.value_and_grad(lambda m: objective(m.constrain(), data)) (reference trainer.py:86-103,126)."""
        n = train_data.n
        if n <= 8192:
            # launch-bound sizes: the evaluation is captured once per data set as a CUDA graph and replayed per step
            key = (id(train_data), float(model_unconstrained.jitter), model_unconstrained.num_genes)
            if self._plan_key != key:
                if self._plan is not None:
                    self._plan.close()
                self._plan = ops.NlmlGradPlan(train_data.X, train_data.y, model_unconstrained.num_genes,
                                              model_unconstrained.jitter, unconstrained=True)
                self._plan_key = key
            out, info = self._plan(model_unconstrained.pack())
        else:
            out, info = ops.nlml_grad_unc(train_data.X, train_data.y, model_unconstrained.pack(),
                                          model_unconstrained.jitter, model_unconstrained.num_genes)
        out = out.cpu().numpy()
        s = -self.constant
        return s * float(out[0]), s * out[1:]
