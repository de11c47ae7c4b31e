"""JAX leg of the drop-in boundary: the XLA FFI custom calls of ``csrc/lfm_xla_ffi.cc`` registered as JAX FFI targets,
and the objective wrapped in ``jax.custom_vjp`` so that the reference's ``jax.value_and_grad(self.loss)``
(``src/trainer.py:126``) reaches the sm_100a kernels unchanged, inside ``jit`` / ``lax.scan`` (``trainer.py:214``).

OPTIONAL: needs an importable ``jax`` (>= 0.4.31, for ``jax.ffi``) with a CUDA backend and ``liblfm_xla_ffi.so``
(``make -C dis_project_b200/csrc xla_ffi``).  Neither exists in the build image, so nothing in this package imports this
module; it raises ImportError with the reason when its requirements are missing.  The ctypes route of ``ops.py`` is
the one the tests exercise.

Reference-side use (the whole change to ``src/trainer.py``):

    from dis_project_b200.jax_ffi import lfm_loss, pack_leaves
    class JaxTrainer:
        def loss(self, model, batch):                      # trainer.py:86-103
            model = model.stop_gradient()
            return lfm_loss(pack_leaves(model), batch.X, batch.y[:, 0], float(model.jitter), model.num_genes)

``pack_leaves(model)`` concatenates ``[true_d, true_s, true_b, l, obs_stddev]`` of the UNCONSTRAINED module: the theta
layout of ``include/lfm_b200.h``.
"""
from __future__ import annotations

import ctypes
import os

try:
    import jax
    import jax.numpy as jnp
    import numpy as np
    _ffi = jax.ffi if hasattr(jax, "ffi") else None
except ImportError as exc:  # pragma: no cover
    raise ImportError("dis_project_b200.jax_ffi needs jax >= 0.4.31 with a CUDA backend (absent from this image); "
                      "use dis_project_b200.ops (ctypes) instead") from exc
if _ffi is None:  # pragma: no cover
    raise ImportError("this jax has no jax.ffi module (need >= 0.4.31; the reference pins 0.4.28)")

_HERE = os.path.dirname(os.path.abspath(__file__))
_FFI_LIB = os.path.join(_HERE, "liblfm_xla_ffi.so")
_CORE_LIB = os.path.join(_HERE, "liblfm_b200.so")
if not os.path.exists(_FFI_LIB):  # pragma: no cover
    raise ImportError(f"{_FFI_LIB} is missing: build it with `make -C dis_project_b200/csrc xla_ffi` on a host that has "
                      "xla/ffi/api/ffi.h (jax.ffi.include_dir())")

_core = ctypes.CDLL(_CORE_LIB, mode=ctypes.RTLD_GLOBAL)
_core.lfm_nlml_workspace_bytes_tg.restype = ctypes.c_size_t
_core.lfm_nlml_workspace_bytes_tg.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int64]
_core.lfm_latent_posterior_workspace_bytes.restype = ctypes.c_size_t
_core.lfm_latent_posterior_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int64]
_core.lfm_count_distinct_times.restype = ctypes.c_int64
_core.lfm_count_distinct_times.argtypes = [ctypes.c_int64, ctypes.c_void_p]
_lib = ctypes.CDLL(_FFI_LIB)

TARGETS = {"lfm_nlml": "LfmNlml", "lfm_nlml_grad": "LfmNlmlGrad", "lfm_nlml_grad_unc": "LfmNlmlGradUnc",
           "lfm_cross_covariance": "LfmCrossCovariance", "lfm_mean_function": "LfmMeanFunction",
           "lfm_latent_posterior": "LfmLatentPosterior", "lfm_gene_posterior": "LfmGenePosterior",
           "lfm_batched_fit": "LfmBatchedFit"}
for _name, _sym in TARGETS.items():
    _ffi.register_ffi_target(_name, _ffi.pycapsule(getattr(_lib, _sym)), platform="CUDA")


def pack_leaves(model):
    """[true_d, true_s, true_b, l, obs_stddev] of a (constrained or unconstrained) reference ExactLFM."""
    return jnp.concatenate([jnp.ravel(model.true_d), jnp.ravel(model.true_s), jnp.ravel(model.true_b),
                            jnp.reshape(model.l, (1,)), jnp.reshape(model.obs_stddev, (1,))])


def distinct_times(X) -> int:
    """Static `time_grid` attribute: X is a concrete array when the trainer is built (trace time)."""
    Xh = np.ascontiguousarray(np.asarray(X), dtype=np.float64)
    return int(_core.lfm_count_distinct_times(Xh.shape[0], Xh.ctypes.data))


_EMPTY = None


def _value_and_grad_unc(theta_unc, X, y, jitter, G, time_grid, variances):
    N, P = X.shape[0], theta_unc.shape[0]
    ws = int(_core.lfm_nlml_workspace_bytes_tg(N, G, time_grid))
    var = jnp.zeros((0,), jnp.float64) if variances is None else jnp.ravel(variances)
    out, info, _ = _ffi.ffi_call(
        "lfm_nlml_grad_unc",
        (jax.ShapeDtypeStruct((P + 1,), jnp.float64), jax.ShapeDtypeStruct((1,), jnp.int32),
         jax.ShapeDtypeStruct((ws,), jnp.uint8)))(X, y, var, theta_unc, jitter=np.float64(jitter), G=np.int64(G),
                                                  time_grid=np.int64(time_grid))
    return out


def make_loss(X, y, jitter: float, G: int, variances=None):
    """``loss(theta_unc) -> scalar`` with a custom VJP: objective(model.constrain(), batch) of trainer.py:86-103 for
    fixed data; one custom call computes the value and the gradient, the backward pass only scales the stored gradient."""
    tg = distinct_times(X)

    @jax.custom_vjp
    def loss(theta_unc):
        return _value_and_grad_unc(theta_unc, X, y, jitter, G, tg, variances)[0]

    def fwd(theta_unc):
        out = _value_and_grad_unc(theta_unc, X, y, jitter, G, tg, variances)
        return out[0], out[1:]

    def bwd(grad, ct):
        return (ct * grad,)

    loss.defvjp(fwd, bwd)
    return loss


def lfm_loss(theta_unc, X, y, jitter: float, G: int, variances=None):
    """Differentiable (w.r.t. theta_unc) negative marginal log-likelihood: see make_loss."""
    return make_loss(X, y, jitter, G, variances)(theta_unc)


def latent_posterior(theta, X, y, variances, Xstar, jitter: float, G: int):
    """ExactLFM.latent_predict (src/model.py:420-463) as a custom call: (mean[T*], var[T*], info[1])."""
    N, T = X.shape[0], Xstar.shape[0]
    ws = int(_core.lfm_latent_posterior_workspace_bytes(N, G, T))
    mean, var, info, _ = _ffi.ffi_call(
        "lfm_latent_posterior",
        (jax.ShapeDtypeStruct((T,), jnp.float64), jax.ShapeDtypeStruct((T,), jnp.float64),
         jax.ShapeDtypeStruct((1,), jnp.int32), jax.ShapeDtypeStruct((ws,), jnp.uint8)))(
        X, y, jnp.ravel(variances), theta, Xstar, jitter=np.float64(jitter), G=np.int64(G))
    return mean, var, info


def cross_covariance(theta, X, Y, G: int):
    """ExactLFM.cross_covariance(kernel, x, y) (src/model.py:372-394) as a custom call."""
    return _ffi.ffi_call("lfm_cross_covariance", jax.ShapeDtypeStruct((X.shape[0], Y.shape[0]), jnp.float64))(
        X, Y, theta, G=np.int64(G))
