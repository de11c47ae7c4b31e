"""Golden vectors whose provenance is REFERENCE SOURCE EXECUTED HERE.

The reference's own files (/root/reference/src/{dataset,model,objectives,trainer,utils}.py) are
imported UNMODIFIED under ``tests/refshim`` (torch-fp64 stand-ins for jax / gpjax / cola / optax /
tfp, see its docstring for what is real and what is restated) and run on the two small CSVs
committed in ``tests/golden/ref_csv/`` (fabricated: the Barenco CSVs are not distributed with the
reference; the files have its on-disk format, dataset.py:233-245).  Everything below is produced by
calling reference functions; this script only chooses the inputs and writes the outputs down.

    python tests/golden/make_ref_golden.py            # needs /root/reference; writes ref_*.json

Consumers: tests/test_ref_parity.py (CPU: oracle and host layer against these files;
GPU: the CUDA path against these files).  Neither needs /root/reference at run time.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
CSV_DIR = os.path.join(HERE, "ref_csv")
sys.path.insert(0, TESTS)


def fabricate_csvs(out_dir, seed=2006):
    """Two CSVs in the reference's on-disk format: Affymetrix probe ids as the index (the six the
    loader keeps, shuffled among four it must drop), columns cARP{r}-{t}hrs.CEL (in a scrambled
    order, plus one column the loader must ignore).  Values: log-expression level and its standard
    error of a SIM ODE driven by Barenco's profile, so the transformed data look like p53 data."""
    import pandas as pd
    rng = np.random.default_rng(seed)
    f = np.array([0.1845, 1.1785, 1.6160, 0.8156, 0.6862, -0.1828, 0.5131])
    B = np.array([0.0649, 0.0069, 0.0181, 0.0033, 0.0869, 0.05])
    D = np.array([0.2829, 0.3720, 0.3617, 0.8000, 0.3573, 0.5])
    S = np.array([0.9075, 0.9748, 0.9785, 1.0000, 0.9680, 1.2])
    tf = np.linspace(0, 12, 1201)
    ff = np.interp(tf, np.linspace(0, 12, 7), f)
    x = np.empty((6, tf.size))
    x[:, 0] = B / D + 0.3
    for n in range(1, tf.size):
        x[:, n] = x[:, n - 1] + (tf[1] - tf[0]) * (B + S * ff[n - 1] - D * x[:, n - 1])
    clean = np.maximum(x[:, ::200], 0.05) * np.array([3.0, 1.5, 2.0, 5.0, 2.5, 4.0])[:, None]  # (6, 7)
    probes = ["203409_at", "205780_at", "209295_at", "202284_s_at", "218346_s_at", "211300_s_at"]  # DDB2 BIK DR5 p21 SESN1 p53
    cols = [f"cARP{r}-{t}hrs.CEL" for r in range(1, 4) for t in range(0, 14, 2)]
    lin = np.concatenate([clean * np.exp(0.08 * rng.standard_normal(clean.shape)) for _ in range(3)], axis=1)
    se = rng.uniform(0.04, 0.16, size=lin.shape)
    rows = {p: (np.log(lin[i]), se[i]) for i, p in enumerate(probes)}
    for extra in ("200000_at", "200001_s_at", "AFFX-1", "217373_x_at"):
        rows[extra] = (rng.normal(5, 1, 21), rng.uniform(0.05, 0.3, 21))
    order = list(rows)
    rng.shuffle(order)
    col_order = list(cols) + ["cARP1-24hrs.CEL"]
    rng.shuffle(col_order)
    ex = pd.DataFrame({c: [rows[p][0][cols.index(c)] if c in cols else 0.0 for p in order] for c in col_order}, index=order)
    sd = pd.DataFrame({c: [rows[p][1][cols.index(c)] if c in cols else 0.1 for p in order] for c in col_order}, index=order)
    os.makedirs(out_dir, exist_ok=True)
    ex.to_csv(os.path.join(out_dir, "barencoPUMA_exprs.csv"), float_format="%.10f")
    sd.to_csv(os.path.join(out_dir, "barencoPUMA_se.csv"), float_format="%.10f")


def L(x):
    """tensor / array -> nested lists of Python floats (repr round-trips fp64 exactly)."""
    import torch
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    if hasattr(x, "to_dense"):
        return L(x.to_dense())
    return np.asarray(x).tolist()


def theta_of(model):
    """[d, s, b, l, sigma]: the theta layout of include/lfm_b200.h."""
    return L(model.true_d) + L(model.true_s) + L(model.true_b) + [float(model.l), float(model.obs_stddev)]


def grad_of(g):
    return L(g.true_d) + L(g.true_s) + L(g.true_b) + [float(g.l), float(g.obs_stddev)]


def main(out_dir=HERE):
    import refshim
    if not os.path.isdir(CSV_DIR):
        fabricate_csvs(CSV_DIR)
    work = tempfile.mkdtemp(prefix="refshim_")
    os.makedirs(os.path.join(work, "data"))
    for name in ("barencoPUMA_exprs.csv", "barencoPUMA_se.csv"):
        with open(os.path.join(CSV_DIR, name)) as src, open(os.path.join(work, "data", name), "w") as dst:
            dst.write(src.read())
    os.chdir(work)  # JaxP53Data(data_dir="data") and ExactLFM's default data factory (model.py:70)

    ref = refshim.load_reference()
    import jax
    import jax.numpy as jnp
    import gpjax as gpx
    import optax as ox
    import torch
    ds, md, ob, tr, ut = ref.dataset, ref.model, ref.objectives, ref.trainer, ref.utils
    key = jax.random.PRNGKey(42)
    provenance = {
        "provenance": "reference source executed under tests/refshim (torch fp64)",
        "reference_files": ["src/dataset.py", "src/model.py", "src/objectives.py", "src/trainer.py", "src/utils.py"],
        "generator": "tests/golden/make_ref_golden.py", "torch": torch.__version__,
    }

    # ---- dataset.py ---------------------------------------------------------------------------
    raw = ds.load_barenco_data("data")
    out = dict(provenance)
    out["load_barenco_data"] = {k: (v if k == "gene_names" else L(v)) for k, v in raw.items()}
    out["variants"] = []
    for kw in ({"replicate": 0}, {"replicate": None}, {"replicate": 2},
               {"replicate": None, "selected_genes": ["p21", "DDB2", "SESN1"]},
               {"replicate": 1, "selected_genes": ["DR5", "BIK"]}):
        d = ds.JaxP53Data(data_dir="data", **kw)
        X, y, var = ds.dataset_3d(d)
        ft, fy = ds.flatten_dataset_jax(d)
        b, s, dd = d.params_ground_truth()
        out["variants"].append({
            "kwargs": kw, "gene_names": list(d.gene_names), "selected_indices": list(d.selected_indices),
            "num_genes": d.num_genes, "len": len(d), "shape": list(d.shape), "timepoints": L(d.timepoints),
            "f_observed": L(d.f_observed), "gene_expressions": L(d.gene_expressions),
            "gene_variances": L(d.gene_variances), "X": L(X), "y": L(y), "variances": L(var),
            "flatten_t": L(ft), "flatten_y": L(fy), "B_exact": L(b), "S_exact": L(s), "D_exact": L(dd),
            "item0": [L(d[0][0]), L(d[0][1])],
        })
    errors = {}
    for name, kw in (("invalid", {"selected_genes": ["p21", "nope"]}), ("duplicate", {"selected_genes": ["p21", "p21"]}),
                     ("empty", {"selected_genes": []})):
        try:
            ds.JaxP53Data(data_dir="data", **kw)
        except ValueError as e:
            errors[name] = str(e)
    out["errors"] = errors
    out["generate_test_times_5"] = L(ut.generate_test_times(5))
    out["generate_test_times_pred_2"] = L(ut.generate_test_times_pred(2))
    with open(os.path.join(out_dir, "ref_dataset.json"), "w") as fh:
        json.dump(out, fh)
    print("ref_dataset.json")

    # ---- model / objective / trainer ------------------------------------------------------------
    def run_case(name, data_kw, num_genes, fix_params, theta_seed, t_pred=12):
        p53 = ds.JaxP53Data(data_dir="data", **data_kw)
        X, y, var = ds.dataset_3d(p53)
        train = gpx.Dataset(X, y)
        model = md.ExactLFM(jitter=jnp.array(1e-4), num_genes=num_genes)
        loss = ob.CustomConjMLL(negative=True)
        trainer = tr.JaxTrainer(model=model, objective=loss, training_data=train, optim=ox.adam(0.01),
                                key=key, num_iters=150)
        rng = np.random.default_rng(theta_seed)
        G = num_genes
        other = model.replace(true_d=jnp.array(rng.uniform(0.2, 1.0, G)), true_s=jnp.array(rng.uniform(0.5, 1.5, G)),
                              true_b=jnp.array(rng.uniform(0.01, 0.1, G)), l=jnp.array(float(rng.uniform(0.8, 3.2))),
                              obs_stddev=jnp.array(float(rng.uniform(0.6, 1.4))))
        case = dict(provenance)
        case.update({"name": name, "G": G, "N": int(X.shape[0]), "jitter": 1e-4, "fix_params": fix_params,
                     "data_kwargs": data_kw, "X": L(X), "y": L(y.reshape(-1)), "variances": L(var.reshape(-1)),
                     "points": []})
        xs = ut.generate_test_times(100 - 100 % G)   # Q3: mean_function needs len(test) divisible by num_genes
        gp = ut.GeneExpressionPredictor(model, p53, t=t_pred)
        xg = gp.generate_test_times_pred()
        mixed = jnp.concatenate([X[:: max(1, X.shape[0] // 9)][:9], xs[::17], xg[::5]])  # xg holds gene index G: JAX clamps (Q6)
        case["Xstar"], case["Xgene"], case["K_rows"] = L(xs), L(xg), L(mixed)
        for label, m in (("init", model), ("random", other)):
            unc = m.unconstrain()
            val_u, g_u = jax.value_and_grad(trainer.loss)(unc, train)            # trainer.py:126
            val_c, g_c = jax.value_and_grad(lambda mm, b: loss(mm, b))(m, train)  # objectives.py:21-78
            lat = m.latent_predict(xs, p53)                                       # model.py:420-463
            gene = m.multi_gene_predict(xg, p53)                                  # model.py:465-514
            j = jnp.array(1); k = jnp.array(G - 1)
            entries = {
                "h": [[1, G - 1, 2.0, 7.5, float(m.h(j, k, jnp.array(2.0), jnp.array(7.5)))],
                      [G - 1, 1, 9.0, 0.5, float(m.h(k, j, jnp.array(9.0), jnp.array(0.5)))],
                      [0, 0, 4.0, 4.0, float(m.h(jnp.array(0), jnp.array(0), jnp.array(4.0), jnp.array(4.0)))]],
                "gamma": [float(m.gamma(jnp.array(i))) for i in range(G)],
                "kernel_xx": float(m.kernel_xx(jnp.array([2.0, 1.0, 1.0]), jnp.array([7.5, G - 1.0, 1.0]))),
                "kernel_xf": float(m.kernel_xf(jnp.array([6.0, 1.0, 1.0]), jnp.array([2.5, -1.0, 0.0]))),
                "kernel_xf_swapped": float(m.kernel_xf(jnp.array([2.5, -1.0, 0.0]), jnp.array([6.0, 1.0, 1.0]))),
                "kernel_ff": float(m.kernel_ff(jnp.array([6.0, -1.0, 0.0]), jnp.array([2.5, -1.0, 0.0]))),
            }
            case["points"].append({
                "label": label, "theta": theta_of(m), "theta_unc": theta_of(unc),
                "nlml": float(val_c), "nlml_via_trainer_loss": float(val_u),
                "grad_constrained": grad_of(g_c), "grad_unconstrained": grad_of(g_u),
                "mean_function": L(m.mean_function(X).reshape(-1)),
                "K_block": L(m.cross_covariance(m.kernel, mixed, mixed)),         # model.py:372-394
                "gram_diag": L(torch.diagonal(m.gram(m.kernel, X).to_dense())),
                "latent_mean": L(lat.mean()), "latent_std": L(lat.stddev()),
                "gene_mean": L(gene.mean()), "gene_std": L(gene.stddev()), "gene_cov_row0": L(gene.scale.to_dense()[0]),
                "entries": entries,
            })
        trained, history = trainer.fit(fix_params=fix_params, num_steps_per_epoch=1000)   # trainer.py:162-228
        lat = trained.latent_predict(xs, p53)
        gp = ut.GeneExpressionPredictor(trained, p53, t=t_pred)
        gene = trained.multi_gene_predict(gp.generate_test_times_pred(), p53)
        dec = gp.decompose_predictions2(gene.mean()) if G == 5 else gp.decompose_predictions(gene.mean())
        case["fit"] = {"theta": theta_of(trained), "history": L(history), "latent_mean": L(lat.mean()),
                       "latent_std": L(lat.stddev()), "gene_mean": L(gene.mean()), "gene_std": L(gene.stddev()),
                       "gene_mean_decomposed": [L(d) for d in dec]}
        with open(os.path.join(out_dir, f"ref_{name}.json"), "w") as fh:
            json.dump(case, fh)
        print(f"ref_{name}.json  N={case['N']}  nlml(init)={case['points'][0]['nlml']:.12f}  "
              f"loss[0]={float(history[0]):.12f} loss[-1]={float(history[-1]):.12f}")

    run_case("p53_rep0", {"replicate": 0}, 5, True, theta_seed=11)                     # main.py:32-59
    run_case("p53_all", {"replicate": None}, 5, False, theta_seed=12)                  # notebook.py:33-75
    run_case("p53_sub3", {"replicate": None, "selected_genes": ["DDB2", "p21", "SESN1"]}, 3, True, theta_seed=13)  # ablation, Q5 out-of-bounds hook


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else HERE)
