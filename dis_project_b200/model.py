"""``ExactLFM``: the single-input-motif latent force model, with the numerics on the B200.

Mirror of the reference's ``src/model.py:29-514`` public surface.  Hyper-parameters live on the host
as float64 numpy arrays (they are 3G+2 numbers); every covariance / mean / posterior evaluation is a
CUDA kernel behind the C-ABI of ``include/lfm_b200.h``.  There is no CPU implementation in this class.
"""
from __future__ import annotations

import copy
from typing import Any, Callable, Optional

import numpy as np

from . import ops
from .dataset import JaxP53Data, dataset_3d
from .gpx_compat import DenseOperator, GaussianDistribution

L_LOW, L_HIGH = 0.5, 3.5  # Sigmoid(low, high) bijector of the lengthscale (reference model.py:111)


def _softplus(x):
    return np.logaddexp(0.0, x)


def _softplus_inv(y):
    return y + np.log(-np.expm1(-y))


def _to_host(a) -> np.ndarray:
    try:
        import torch

        if isinstance(a, torch.Tensor):
            return a.detach().cpu().numpy().astype(np.float64)
    except Exception:  # pragma: no cover
        pass
    return np.asarray(a, dtype=np.float64)


class ExactLFM:
    """GP latent force model of Lawrence et al. (2006) on the p53 network.

    Same constructor keywords, attributes and methods as the reference class:
    ``jitter`` (static), ``obs_stddev``, ``data``, ``num_genes``, ``true_d``, ``true_s``, ``true_b``,
    ``l`` and ``mean_function / kernel / kernel_xx / kernel_xf / kernel_ff / h / gamma /
    cross_covariance / gram / latent_predict`` plus the gpjax ``Module`` members the trainer relies
    on (``constrain / unconstrain / stop_gradient / replace``).
    """

    def __init__(self, jitter: float = 1e-6, obs_stddev: Any = 1.0, data: Optional[JaxP53Data] = None,
                 num_genes: int = 5):
        self.jitter = float(np.asarray(jitter))
        self.obs_stddev = np.asarray(obs_stddev, dtype=np.float64)
        self.num_genes = int(num_genes)
        if data is None:
            # reference: default_factory=lambda: JaxP53Data() (model.py:70); the CSVs may be absent
            try:
                data = JaxP53Data()
            except FileNotFoundError:
                data = None
        self.data = data
        # reference __post_init__ (model.py:99-108)
        self.initial_decays = np.full(self.num_genes, 0.4)
        self.initial_sensitivities = np.full(self.num_genes, 1.0)
        self.initial_basals = np.full(self.num_genes, 0.05)
        self.true_d = self.initial_decays.copy()
        self.true_s = self.initial_sensitivities.copy()
        self.true_b = self.initial_basals.copy()
        self.initial_lengthscale = np.asarray(2.5)
        self.l = np.asarray(2.5)

    # ---- gpjax.Module surface --------------------------------------------------------------------
    def replace(self, **kwargs) -> "ExactLFM":
        new = copy.copy(self)
        for k, v in kwargs.items():
            if not hasattr(new, k):
                raise ValueError(f"'{k}' is not a field of ExactLFM")
            setattr(new, k, np.asarray(v, dtype=np.float64) if k != "data" else v)
        return new

    def pack(self) -> np.ndarray:
        """Leaves as the C-ABI vector [true_d(G), true_s(G), true_b(G), l, obs_stddev]."""
        return np.concatenate([np.asarray(self.true_d, dtype=np.float64).reshape(-1),
                               np.asarray(self.true_s, dtype=np.float64).reshape(-1),
                               np.asarray(self.true_b, dtype=np.float64).reshape(-1),
                               np.asarray(self.l, dtype=np.float64).reshape(1),
                               np.asarray(self.obs_stddev, dtype=np.float64).reshape(1)])

    # used by gpx_compat.apply_updates / GradientTransformation
    pack_unconstrained_leaves = pack

    def with_leaves(self, theta) -> "ExactLFM":
        theta = _to_host(theta).reshape(-1)
        G = self.num_genes
        if theta.shape[0] != 3 * G + 2:
            raise ValueError(f"expected {3 * G + 2} leaves, got {theta.shape[0]}")
        return self.replace(true_d=theta[:G].copy(), true_s=theta[G:2 * G].copy(), true_b=theta[2 * G:3 * G].copy(),
                            l=np.asarray(theta[3 * G]), obs_stddev=np.asarray(theta[3 * G + 1]))

    def constrain(self) -> "ExactLFM":
        """Bijector forward on every leaf: softplus, and Sigmoid(0.5, 3.5) on l."""
        th = self.pack()
        G = self.num_genes
        out = _softplus(th)
        out[3 * G] = L_LOW + (L_HIGH - L_LOW) * 0.5 * (1.0 + np.tanh(0.5 * th[3 * G]))
        return self.with_leaves(out)

    def unconstrain(self) -> "ExactLFM":
        """Bijector inverse on every leaf."""
        th = self.pack()
        G = self.num_genes
        out = _softplus_inv(th)
        u = (th[3 * G] - L_LOW) / (L_HIGH - L_LOW)
        out[3 * G] = np.log(u) - np.log1p(-u)
        return self.with_leaves(out)

    def stop_gradient(self) -> "ExactLFM":
        """All leaves are trainable in the reference, so this is the identity (trainer.py:102)."""
        return self

    # ---- mean and kernels (device) ---------------------------------------------------------------
    def mean_function(self, x):
        """(B/D) per positional block times the flag column (reference model.py:124-149); (N, 1)."""
        return ops.mean_function(x, self.pack(), self.num_genes)

    def _pair(self, t, t_prime, force_flags=None) -> float:
        a = _to_host(t).reshape(1, 3).copy()
        b = _to_host(t_prime).reshape(1, 3).copy()
        if force_flags is not None:
            a[0, 2], b[0, 2] = force_flags
        return float(ops.cross_covariance(a, b, self.pack(), self.num_genes).item())

    def kernel(self, t, t_prime) -> float:
        """Flag-switched kernel between two (time, gene, flag) points (reference model.py:152-195)."""
        return self._pair(t, t_prime)

    def kernel_xx(self, t, t_prime) -> float:
        """Gene-gene covariance, eq. 5 of Lawrence et al. (reference model.py:197-235)."""
        return self._pair(t, t_prime, (1.0, 1.0))

    def kernel_xf(self, t, t_prime) -> float:
        """Gene-latent cross covariance; the flag of `t` says which argument is the latent point
        (reference model.py:237-282)."""
        a = _to_host(t).reshape(3)
        if a[2] == 0:
            return self._pair(t, t_prime, (0.0, 1.0))
        return self._pair(t, t_prime, (1.0, 0.0))

    def kernel_ff(self, t, t_prime) -> float:
        """Latent RBF prior with the reference's 2*l denominator (reference model.py:284-312)."""
        return self._pair(t, t_prime, (0.0, 0.0))

    def h(self, j, k, t1, t2):
        """Convolution term h_jk(t1, t2) (reference model.py:315-365); scalars or arrays."""
        out = ops.h_terms(np.asarray(j, dtype=np.float64), np.asarray(k, dtype=np.float64),
                          np.asarray(t1, dtype=np.float64), np.asarray(t2, dtype=np.float64), self.pack(),
                          self.num_genes)
        out = out.cpu().numpy()
        return float(out[0]) if out.size == 1 else out

    def gamma(self, k):
        """gamma_k = D_k l / 2 (reference model.py:367-369)."""
        return np.asarray(self.true_d)[np.asarray(k, dtype=np.int64)] * np.asarray(self.l) / 2

    def cross_covariance(self, kernel: Optional[Callable], x, y):
        """Dense N x M cross covariance on the device.  `kernel` is accepted for signature
        compatibility; it must be this model's own `kernel` (reference model.py:372-394)."""
        self._check_kernel(kernel)
        return ops.cross_covariance(x, y, self.pack(), self.num_genes)

    def gram(self, kernel: Optional[Callable], x) -> DenseOperator:
        """Gram matrix as a PSD dense operator (reference model.py:396-414)."""
        self._check_kernel(kernel)
        return DenseOperator(ops.gram(x, self.pack(), self.num_genes))

    def _check_kernel(self, kernel) -> None:
        if kernel is None:
            return
        owner = getattr(kernel, "__self__", None)
        if getattr(kernel, "__name__", "") != "kernel" or not isinstance(owner, ExactLFM):
            raise NotImplementedError("only ExactLFM.kernel is implemented on the device (no generic kernel callables)")

    # ---- predictions -----------------------------------------------------------------------------
    def latent_predict(self, test_inputs, train_data: JaxP53Data) -> GaussianDistribution:
        """Posterior of the latent force at `test_inputs` (reference model.py:420-463): noise model
        K + diag(measurement variances) + jitter I, variance 1 + 2 jitter - k^T Sigma^-1 k."""
        x, y, variances = dataset_3d(train_data)
        t = _to_host(test_inputs)
        if t.shape[0] % self.num_genes and np.any(t[:, 2] != 0):
            raise ValueError("mean_function: test rows must be divisible by num_genes (reference model.py:145-149)")
        mean, var, info = ops.latent_posterior(x, y, variances, self.pack(), self.jitter, t, self.num_genes)
        return GaussianDistribution(mean, var)

    def multi_gene_predict(self, test_inputs, train_data: JaxP53Data) -> GaussianDistribution:
        """Posterior of the gene expressions at `test_inputs` (reference model.py:465-514): noise model
        K + diag(measurement variances) + obs_stddev^2 I (no jitter), full predictive covariance
        K_tt - K_tx Sigma^-1 K_xt + jitter I."""
        x, y, variances = dataset_3d(train_data)
        t = _to_host(test_inputs)
        mean, cov, var, info = ops.gene_posterior(x, y, variances, self.pack(), self.jitter, t, self.num_genes)
        return GaussianDistribution(mean, cov)

    # ---- the GPyTorch twin's posteriors (reference src/gpytorch_alfi/model_alfi.py:68-150) ------------------
    # Same two C-ABI entry points, the twin's conventions: no mean function (basal rates play no part), training
    # covariance K_xx + 1e-4 I + diag(measurement variances) without the likelihood noise (model_alfi.py:282-299),
    # K_ff + 1e-3 I (K_ff of the twin), and `jitter` added to the predictive variances at the end.
    TWIN_KERNEL_JITTER = 1e-4

    def _twin_theta(self, sigma: float) -> np.ndarray:
        theta = self.pack()
        G = self.num_genes
        theta[2 * G:3 * G] = 0.0
        theta[3 * G + 1] = sigma
        return theta

    def predict_f(self, pred_t, train_data: JaxP53Data, jitter: float = 1e-3):
        """Twin's ``ExactLFM.predict_f(pred_t, jitter)`` (model_alfi.py:109-150): latent force at the times
        ``pred_t``; returns (mean (T*,), variance (T*,)) -- the twin keeps the diagonal too (:141-143)."""
        x, y, variances = dataset_3d(train_data)
        t = _to_host(pred_t).reshape(-1)
        xs = np.stack((t, -np.ones_like(t), np.zeros_like(t)), axis=-1)
        kj = self.TWIN_KERNEL_JITTER
        mean, var, info = ops.latent_posterior(x, y, variances, self._twin_theta(0.0), kj, xs, self.num_genes)
        var = var.cpu().numpy() - 2.0 * kj + 1e-3 + float(jitter)
        return mean.cpu().numpy(), var

    def predict_m(self, pred_t, train_data: JaxP53Data, jitter: float = 1e-5):
        """Twin's ``ExactLFM.predict_m(pred_t, jitter)`` (model_alfi.py:68-107): gene expressions at the times
        ``pred_t``; returns (mean (T*, G), variance (T*, G)) as the twin's batch of diagonal normals (:100-107)."""
        x, y, variances = dataset_3d(train_data)
        t = _to_host(pred_t).reshape(-1)
        G, T = self.num_genes, t.size
        xg = np.stack((np.tile(t, G), np.repeat(np.arange(G), T).astype(np.float64), np.ones(G * T)), axis=-1)
        kj = self.TWIN_KERNEL_JITTER
        mean, _, var, info = ops.gene_posterior(x, y, variances, self._twin_theta(float(np.sqrt(kj))), kj, xg, G,
                                                full_cov=False)
        return mean.cpu().numpy().reshape(G, T).T, (var.cpu().numpy() + float(jitter)).reshape(G, T).T
