"""refshim: stand-ins for the reference's third-party stack, so that the reference's OWN source
files (``/root/reference/src/{dataset,model,objectives,trainer,utils}.py``) import and run here,
unmodified, on torch fp64 CPU tensors.

TEST INFRASTRUCTURE ONLY (offline fixture generation: ``tests/golden/make_ref_golden.py``).
Nothing under ``dis_project_b200/`` imports it, and no test imports it at run time on the GPU box:
the tests read the committed ``tests/golden/ref_*.json`` files it produced.

What is real and what is a stand-in:

* REAL, executed unmodified from /root/reference/src: ``dataset.py`` (CSV -> log-normal moments ->
  rescale -> (N,3) layout), ``model.py`` (ExactLFM: mean function, flag-switched kernel, k_xx,
  k_xf, k_ff, h, gamma, cross_covariance, gram, latent_predict, multi_gene_predict),
  ``objectives.py`` (CustomConjMLL.step), ``trainer.py`` (JaxTrainer: loss, step, after_epoch_jax,
  fit), ``utils.py`` (generate_test_times*, GeneExpressionPredictor.generate_test_times_pred /
  decompose_predictions*, print_hyperparams).
* STAND-IN (this package; restated from the published behaviour of the versions pinned in the
  reference's environment.yml:48-119, none of which is installable here):
    jax 0.4.28          -> ``shim_jax``    jnp on torch.float64, erf, vmap = torch.vmap,
                                           value_and_grad = torch.autograd over a Module's
                                           parameter leaves, lax.cond / scan as Python control flow
    gpjax 0.8.2         -> ``shim_gpjax``  Module (param_field / static_field, constrain,
                                           unconstrain, stop_gradient, replace), Dataset,
                                           AbstractObjective (constant = -1 when negative),
                                           GaussianDistribution (dense-Cholesky log_prob, mean,
                                           stddev, variance), scan.vscan
    cola-ml 0.0.5       -> ``shim_misc``   Dense / PSD / I_like / inv / solve, dense semantics
    optax 0.1.9         -> ``shim_misc``   adam (b1 .9, b2 .999, eps 1e-8, bias corrected), apply_updates
    tfp bijectors       -> ``shim_misc``   Softplus, Sigmoid(low, high)
    jaxtyping, matplotlib -> inert placeholders (annotations / plotting are never exercised)
    gpytorch            -> ``shim_gpytorch`` (on real torch) for the reference's GPyTorch twin,
                           ``src/gpytorch_alfi/{dataset_alfi,model_alfi}.py`` (tests/golden/make_ref_twin_golden.py)

``install()`` registers the stand-ins in ``sys.modules`` and puts the reference's ``src`` directory
on ``sys.path`` (the reference modules import each other by bare name).
"""
from __future__ import annotations

import importlib
import os
import sys

REFERENCE_SRC = os.environ.get("LFM_REFERENCE_SRC", "/root/reference/src")

_REF_MODULES = ("dataset", "model", "objectives", "trainer", "utils", "plotter")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "model.py"))


def install() -> None:
    """Register the stand-in modules and make the reference's flat modules importable."""
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import shim_jax
    import shim_gpjax
    import shim_misc

    shim_jax.register()
    shim_gpjax.register()
    shim_misc.register()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)


def load_reference():
    """Import the reference's own modules (after ``install``) and return them in a namespace."""
    if not reference_available():
        raise FileNotFoundError(f"reference source not found under {REFERENCE_SRC}")
    install()
    for name in _REF_MODULES:
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(REFERENCE_SRC):
            raise RuntimeError(f"module name {name!r} is already taken by {mod.__file__}")

    class _NS:
        pass

    ns = _NS()
    for name in _REF_MODULES:
        mod = importlib.import_module(name)
        assert mod.__file__.startswith(REFERENCE_SRC), mod.__file__
        setattr(ns, name, mod)
    return ns
