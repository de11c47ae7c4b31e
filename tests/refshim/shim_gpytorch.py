"""Stand-in for the part of gpytorch the reference's GPyTorch twin touches (``src/gpytorch_alfi/{model_alfi,
dataset_alfi}.py``), on real torch.  TEST INFRASTRUCTURE ONLY (tests/golden/make_ref_twin_golden.py).

Semantics restated from gpytorch's published behaviour [3P]:
* ``constraints.Positive`` = softplus transform (inverse ``v + log(-expm1(-v))``); ``constraints.Interval(a, b)`` =
  ``a + (b - a) sigmoid(x)``; ``constraints.GreaterThan(lb)`` = ``softplus(x) + lb``.
* ``Module.register_parameter / register_constraint / initialize`` as on ``gpytorch.Module``.
* ``kernels.Kernel.__call__(x1, x2=None)`` evaluates ``forward`` lazily (``.evaluate()`` / ``.to_dense()`` give the matrix).
* ``likelihoods.GaussianLikelihood``: ``noise = softplus(raw_noise) + 1e-4``, ``raw_noise`` initialised to 0;
  ``likelihood(mvn)`` adds ``noise I`` to the covariance.
* ``models.ExactGP.__call__`` in training mode returns ``forward(train_inputs)`` (the prior at the training inputs).
* ``mlls.ExactMarginalLogLikelihood(likelihood, model)(output, target)`` = ``likelihood(output).log_prob(target) / N``.
* ``distributions.MultivariateNormal.log_prob``: dense Cholesky.
"""
from __future__ import annotations

import math
import sys
import types

import torch


def _inv_softplus(v):
    return v + torch.log(-torch.expm1(-v))


class GreaterThan(torch.nn.Module):
    def __init__(self, lower_bound=0.0):
        super().__init__()
        self.lower_bound = float(lower_bound)

    def transform(self, x):
        return torch.nn.functional.softplus(x) + self.lower_bound

    def inverse_transform(self, v):
        return _inv_softplus(v - self.lower_bound)


class Positive(GreaterThan):
    def __init__(self):
        super().__init__(0.0)


class Interval(torch.nn.Module):
    def __init__(self, lower_bound, upper_bound):
        super().__init__()
        self.lower_bound, self.upper_bound = float(lower_bound), float(upper_bound)

    def transform(self, x):
        return self.lower_bound + (self.upper_bound - self.lower_bound) * torch.sigmoid(x)

    def inverse_transform(self, v):
        u = (v - self.lower_bound) / (self.upper_bound - self.lower_bound)
        return torch.log(u) - torch.log1p(-u)


class Module(torch.nn.Module):
    def __init__(self, **kwargs):
        super().__init__()

    def register_parameter(self, name, parameter=None, prior=None):
        super().register_parameter(name, parameter)

    def register_constraint(self, param_name, constraint):
        self.add_module(param_name + "_constraint", constraint)

    def initialize(self, **kwargs):
        for name, val in kwargs.items():
            p = getattr(self, name)
            with torch.no_grad():
                p.copy_(torch.as_tensor(val, dtype=p.dtype).reshape(p.shape))
        return self


class _Lazy:
    def __init__(self, fn):
        self._fn, self._val = fn, None

    def evaluate(self):
        if self._val is None:
            self._val = self._fn()
        return self._val

    to_dense = evaluate


class Kernel(Module):
    def forward(self, x1, x2, **params):
        raise NotImplementedError

    def __call__(self, x1, x2=None, **params):
        return _Lazy(lambda: self.forward(x1, x1 if x2 is None else x2, **params))


class Mean(Module):
    def __call__(self, x):
        return self.forward(x)


class MultivariateNormal:
    def __init__(self, mean, covariance_matrix):
        self.mean = mean
        self._cov = covariance_matrix

    @property
    def covariance_matrix(self):
        return self._cov.evaluate() if isinstance(self._cov, _Lazy) else self._cov

    def log_prob(self, value):
        cov = self.covariance_matrix
        diff = (value - self.mean).reshape(-1, 1).to(cov.dtype)
        chol = torch.linalg.cholesky(cov)
        half = torch.linalg.solve_triangular(chol, diff, upper=False)
        n = diff.shape[0]
        return -0.5 * (n * math.log(2.0 * math.pi) + 2.0 * torch.log(torch.diagonal(chol)).sum() + (half * half).sum())


class MultitaskMultivariateNormal(MultivariateNormal):
    @classmethod
    def from_batch_mvn(cls, batch_mvn, task_dim=-1):
        return batch_mvn


class GaussianLikelihood(Module):
    def __init__(self):
        super().__init__()
        self.register_parameter(name="raw_noise", parameter=torch.nn.Parameter(torch.zeros(1)))
        self.noise_constraint = GreaterThan(1e-4)

    @property
    def noise(self):
        return self.noise_constraint.transform(self.raw_noise)

    def __call__(self, mvn):
        cov = mvn.covariance_matrix
        return MultivariateNormal(mvn.mean, cov + self.noise.to(cov.dtype) * torch.eye(cov.shape[-1], dtype=cov.dtype))


class ExactGP(Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        self.train_inputs = (train_inputs,)
        self.train_targets = train_targets
        self.likelihood = likelihood

    def __call__(self, *args):
        return self.forward(*args)


class ExactMarginalLogLikelihood(torch.nn.Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood, self.model = likelihood, model

    def forward(self, output, target):
        return self.likelihood(output).log_prob(target) / target.shape[-1]


def register():
    g = types.ModuleType("gpytorch")
    g.__path__ = []
    sub = {}
    for name in ("constraints", "kernels", "means", "likelihoods", "models", "distributions", "mlls"):
        m = types.ModuleType("gpytorch." + name)
        m.__path__ = []
        sub[name] = m
        setattr(g, name, m)
        sys.modules["gpytorch." + name] = m
    sub["constraints"].Positive, sub["constraints"].Interval, sub["constraints"].GreaterThan = Positive, Interval, GreaterThan
    sub["constraints"].Constraint = torch.nn.Module
    sub["kernels"].Kernel, sub["means"].Mean = Kernel, Mean
    sub["likelihoods"].GaussianLikelihood = GaussianLikelihood
    sub["models"].ExactGP = ExactGP
    sub["distributions"].MultivariateNormal = MultivariateNormal
    sub["distributions"].MultitaskMultivariateNormal = MultitaskMultivariateNormal
    eml = types.ModuleType("gpytorch.mlls.exact_marginal_log_likelihood")
    eml.ExactMarginalLogLikelihood = ExactMarginalLogLikelihood
    sub["mlls"].exact_marginal_log_likelihood = eml
    sub["mlls"].ExactMarginalLogLikelihood = ExactMarginalLogLikelihood
    sys.modules["gpytorch.mlls.exact_marginal_log_likelihood"] = eml
    g.Module = Module
    sys.modules["gpytorch"] = g
