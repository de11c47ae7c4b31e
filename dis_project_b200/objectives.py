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


def _content_token(a):
    """Cheap identity of an array's CONTENTS: (address, version counter) for torch tensors, shape plus two exact
    64-bit checksums of the bit patterns for host arrays (about 10 us at N = 4000)."""
    try:
        import torch

        if isinstance(a, torch.Tensor):
            return ("torch", a.data_ptr(), a._version, tuple(a.shape), str(a.device))
    except ImportError:  # pragma: no cover
        pass
    h = np.ascontiguousarray(a, dtype=np.float64)
    bits = h.reshape(-1).view(np.uint64)
    ramp = np.arange(1, bits.size + 1, dtype=np.uint64)
    return ("host", h.shape, int(np.bitwise_xor.reduce(bits)) if bits.size else 0, int((bits * ramp).sum(dtype=np.uint64)) if bits.size else 0)


class CustomConjMLL:
    """``CustomConjMLL(negative=True)`` -> callable ``(model, Dataset) -> scalar``."""

    def __init__(self, negative: bool = False, variances=None):
        """`variances` (optional, (N,) or (N,1): the third return value of dataset_3d) switches to the
        heteroscedastic objective Sigma = K + diag(variances) + jitter I + obs_stddev^2 I -- the training
        convention of the reference's GPyTorch twin (src/gpytorch_alfi/model_alfi.py:294-299).  None = the
        reference's GPJax objective (objectives.py:70-73)."""
        self.negative = bool(negative)
        self.variances = None if variances is None else np.ascontiguousarray(np.asarray(variances, dtype=np.float64).reshape(-1))
        self.constant = -1.0 if negative else 1.0  # gpjax AbstractObjective.constant
        self._plan = None       # CUDA-graph evaluation plan of the last (data set, jitter) seen by value_and_grad
        self._plan_key = None
        self._plan_data = None  # strong reference to the Dataset the plan was built from

    def __call__(self, model: ExactLFM, train_data: Dataset) -> float:
        return self.step(model, train_data)

    def step(self, model: ExactLFM, train_data: Dataset) -> float:
        """constant * log p(y | X, theta) for a CONSTRAINED model (reference objectives.py:64-78)."""
        val, info = ops.nlml(train_data.X, train_data.y, model.pack(), model.jitter, model.num_genes,
                             variances=self.variances)
        nl = float(val.item())  # NaN when Sigma is not positive definite, like JAX's Cholesky
        return -self.constant * nl

    def value_and_grad(self, model_unconstrained: ExactLFM, train_data: Dataset) -> Tuple[float, np.ndarray]:
        """Loss and gradient w.r.t. the UNCONSTRAINED leaves [d, s, b, l, obs_stddev]:
        jax.value_and_grad(lambda m: objective(m.constrain(), data)) (reference trainer.py:86-103,126)."""
        n = train_data.n
        if n <= 8192:
            # launch-bound sizes: the evaluation is captured once per data set as a CUDA graph and replayed per step
            # The plan owns device copies of X and y, so it is only valid for THIS data object with THESE contents:
            # the Dataset is held strongly (an id() can be reused after garbage collection) and compared with `is`,
            # and a content token catches in-place edits of X / y.
            key = (float(model_unconstrained.jitter), model_unconstrained.num_genes, n,
                   _content_token(train_data.X), _content_token(train_data.y),
                   None if self.variances is None else _content_token(self.variances))
            if self._plan_data is not train_data or self._plan_key != key:
                if self._plan is not None:
                    self._plan.close()
                self._plan = ops.NlmlGradPlan(train_data.X, train_data.y, model_unconstrained.num_genes,
                                              model_unconstrained.jitter, unconstrained=True, variances=self.variances)
                self._plan_key = key
                self._plan_data = train_data
            out, info = self._plan(model_unconstrained.pack())
        else:
            out, info = ops.nlml_grad_unc(train_data.X, train_data.y, model_unconstrained.pack(),
                                          model_unconstrained.jitter, model_unconstrained.num_genes,
                                          variances=self.variances)
        out = out.cpu().numpy()
        s = -self.constant
        return s * float(out[0]), s * out[1:]
