"""Generate the golden vectors in this directory.

The reference stack (jax/gpjax/cola/optax/tfp) cannot be imported in the build container, so the
vectors come from the CPU oracle (oracle/lfm_oracle.py), whose individual kernel entries are in turn
pinned against an independent 50-digit mpmath evaluation of the reference formulas
(tests/test_oracle.py).  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import lfm_oracle as o  # noqa: E402


def case(name, G, T, R, seed, theta_seed=None, Ts=50):
    x, y, var, _ = o.synthetic_problem(G, T, R, seed=seed)
    if theta_seed is None:
        p = o.Params.reference_init(G)
    else:
        rng = np.random.default_rng(theta_seed)
        p = o.Params(d=rng.uniform(0.2, 1.0, G), s=rng.uniform(0.5, 1.5, G), b=rng.uniform(0.01, 0.1, G),
                     l=float(rng.uniform(0.8, 3.2)), sigma=float(rng.uniform(0.6, 1.4)), jitter=1e-4)
    xs = o.generate_test_times(Ts)
    both = np.concatenate([x[: min(12, len(x))], xs[:6]], axis=0)
    val, grad = o.nlml_and_grad(p, x, y)
    u = o.unconstrain(p.pack())
    val_u, grad_u = o.nlml_and_grad_unc(u, x, y, p.jitter)
    mean, pvar = o.latent_predict(p, xs, x, y, var)
    theta_fit, hist = o.fit(p.pack(), x, y, p.jitter, num_iters=150) if x.shape[0] <= 105 else (None, None)
    out = {
        "name": name, "G": G, "T": T, "R": R, "jitter": p.jitter,
        "X": x.tolist(), "y": y.tolist(), "variances": var.tolist(), "theta": p.pack().tolist(),
        "Xstar": xs.tolist(), "K_rows": both.tolist(), "K_block": o.cross_covariance(p, both, both).tolist(),
        "mean_function": o.mean_function(p, x).tolist(),
        "nlml": val, "grad_constrained": grad.tolist(), "theta_unc": u.tolist(), "grad_unconstrained": grad_u.tolist(),
        "posterior_mean": mean.tolist(), "posterior_var": pvar.tolist(),
    }
    if theta_fit is not None:
        out["fit_theta"] = theta_fit.tolist()
        out["fit_history"] = hist.tolist()
    with open(os.path.join(HERE, name + ".json"), "w") as fh:
        json.dump(out, fh)
    print(name, "nlml", val)


if __name__ == "__main__":
    case("p53_rep0_n35", 5, 7, 1, seed=42)                 # main.py:32 shape, reference init
    case("p53_all_n105", 5, 7, 3, seed=42)                 # notebook.py:36 shape, reference init
    case("toy_g3_t6", 3, 6, 1, seed=7, theta_seed=8, Ts=48)     # 3 genes x 6 times, random theta (no p21 index)
    case("mid_g6_t50_n300", 6, 50, 1, seed=9, theta_seed=10, Ts=120)  # three 128-blocks in the dense path
