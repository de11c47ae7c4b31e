"""Golden vectors from the reference's GPyTorch twin, EXECUTED HERE: ``src/gpytorch_alfi/{dataset_alfi,model_alfi}.py``
imported unmodified under ``tests/refshim/shim_gpytorch.py`` (a stand-in for gpytorch on real torch) and run on the CSVs
of ``tests/golden/ref_csv``, exactly as ``src/gpytorch_alfi/main_alfi.py:24-35`` and ``trainer_alfi.py:170-177`` do:

    dataset = PyTorchDataset(replicate=0, data_dir=...)
    model = ExactLFM(dataset, dataset.gene_variances.reshape(-1))
    loss = -ExactMarginalLogLikelihood(model.likelihood, model)(model(model.train_t), model.train_y.squeeze())

Also recorded per parameter point: ``model.predict_f`` / ``model.predict_m`` (model_alfi.py:68-150) on the 80 prediction
times of main_alfi.py:55-57 -- the twin's posteriors, with its own conventions (no mean function, Sigma = K_xx + 1e-4 I +
diag(variances) without the likelihood noise, K_ff + 1e-3 I, float32 K_xf).

This is the second, independent implementation of the same formulas inside the reference (block-structured Gram,
model_alfi.py:266-300) and the source of the heteroscedastic training convention
``K_xx + 1e-4 I + diag(variances) (+ likelihood noise I)`` (model_alfi.py:294-299) that ``lfm_nlml_het_tg`` implements.
NB: the twin keeps its kernel parameters in float32 (``TorchKernel(dtype=torch.float32)``, model_alfi.py:191), so its
numbers carry float32 rounding (about 1e-7 relative): the consumers compare at 1e-5.

    python tests/golden/make_ref_twin_golden.py      # needs /root/reference; writes tests/golden/ref_twin_*.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "refshim"))
TWIN = os.environ.get("LFM_REFERENCE_TWIN", "/root/reference/src/gpytorch_alfi")


def L(x):
    import torch
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float64).tolist()


def main(out_dir=HERE):
    import torch
    import shim_gpytorch
    shim_gpytorch.register()
    torch.set_default_dtype(torch.float64)     # the twin's data tensors are float64 (dataset_alfi.py:66-71); its kernel
    sys.path.insert(0, TWIN)                   # parameters stay float32 by its own default argument
    import dataset_alfi
    import model_alfi
    from gpytorch.mlls.exact_marginal_log_likelihood import ExactMarginalLogLikelihood
    assert dataset_alfi.__file__.startswith(TWIN) and model_alfi.__file__.startswith(TWIN)

    for replicate, name in ((0, "rep0"), (2, "rep2")):
        dataset = dataset_alfi.PyTorchDataset(replicate=replicate, data_dir=os.path.join(HERE, "ref_csv"))
        model = model_alfi.ExactLFM(dataset, dataset.gene_variances.reshape(-1))      # main_alfi.py:27-28
        loss_fn = ExactMarginalLogLikelihood(model.likelihood, model)                 # main_alfi.py:32
        rng = np.random.default_rng(100 + replicate)
        points = []
        for label in ("init", "random"):
            if label == "random":
                G = dataset.num_outputs
                model.covar_module.decay = torch.tensor(rng.uniform(0.2, 1.0, G))
                model.covar_module.sensitivity = torch.tensor(rng.uniform(0.5, 1.5, G))
                model.covar_module.lengthscale = torch.tensor([[float(rng.uniform(0.8, 3.2))]])
                model.mean_module.basal = torch.tensor(rng.uniform(0.01, 0.1, G))
                model.likelihood.initialize(raw_noise=model.likelihood.noise_constraint.inverse_transform(
                    torch.tensor([float(rng.uniform(0.4, 1.6))])))
            model.zero_grad()
            output = model(model.train_t)                                             # trainer_alfi.py:172
            loss = -loss_fn(output, model.train_y.squeeze())                          # trainer_alfi.py:173
            loss.backward()
            K = model.covar_module(model.train_t).evaluate()
            # the twin's posteriors exactly as main_alfi.py:55-57 calls them (80 times on [0, 13], jitter 1e-3).  K_xf()
            # REPLACES the kernel's own k_xf method by its result (model_alfi.py: `self.k_xf = K_xf`), so predict_f works once
            # per kernel object: the attribute is removed afterwards, which is what a fresh model would see.
            with torch.no_grad():
                t_predict = torch.linspace(0, 13, 80, dtype=torch.float64)
                p_f = model.predict_f(t_predict, jitter=1e-3)
                del model.covar_module.k_xf
                p_m = model.predict_m(t_predict, jitter=1e-3)
                post = {"t_predict": L(t_predict), "jitter": 1e-3,
                        "f_mean": L(p_f.mean.reshape(-1)), "f_var": L(torch.diagonal(p_f.covariance_matrix, dim1=-2, dim2=-1).reshape(-1)),
                        "m_mean": L(p_m.mean),                                                # (80, G)
                        "m_var": L(torch.diagonal(p_m.covariance_matrix, dim1=-2, dim2=-1))}  # (80, G)
                # the same unmodified methods after `model.double()` (torch.nn.Module API: the raw parameters become float64):
                # what is left of float32 is K_xf's buffer (model_alfi.py: `torch.zeros(shape, dtype=torch.float32)`), so
                # predict_m is float64 throughout and pins the conventions to ~1e-9 instead of float32 rounding
                import copy
                m64 = copy.deepcopy(model).double()
                q_f = m64.predict_f(t_predict, jitter=1e-3)
                del m64.covar_module.k_xf
                q_m = m64.predict_m(t_predict, jitter=1e-3)
                out64 = m64(m64.train_t)
                post["f64"] = {"f_mean": L(q_f.mean.reshape(-1)),
                               "f_var": L(torch.diagonal(q_f.covariance_matrix, dim1=-2, dim2=-1).reshape(-1)),
                               "m_mean": L(q_m.mean), "m_var": L(torch.diagonal(q_m.covariance_matrix, dim1=-2, dim2=-1)),
                               "loss": float(-ExactMarginalLogLikelihood(m64.likelihood, m64)(out64, m64.train_y.squeeze())),
                               "decay": L(m64.covar_module.decay), "sensitivity": L(m64.covar_module.sensitivity),
                               "basal": L(m64.mean_module.basal), "lengthscale": float(m64.covar_module.lengthscale),
                               "noise": float(m64.likelihood.noise)}
            points.append({
                "posterior": post,
                "label": label,
                "decay": L(model.covar_module.decay), "sensitivity": L(model.covar_module.sensitivity),
                "basal": L(model.mean_module.basal), "lengthscale": float(model.covar_module.lengthscale),
                "noise": float(model.likelihood.noise),
                "loss": float(loss), "K_xx": L(K), "mean": L(output.mean),
                "raw": {"decay": L(model.covar_module.raw_decay), "sensitivity": L(model.covar_module.raw_sensitivity),
                        "basal": L(model.mean_module.raw_basal), "lengthscale": L(model.covar_module.raw_lengthscale.reshape(-1)),
                        "noise": L(model.likelihood.raw_noise)},
                "grad_raw": {"decay": L(model.covar_module.raw_decay.grad), "sensitivity": L(model.covar_module.raw_sensitivity.grad),
                             "basal": L(model.mean_module.raw_basal.grad),
                             "lengthscale": L(model.covar_module.raw_lengthscale.grad.reshape(-1)),
                             "noise": L(model.likelihood.raw_noise.grad)},
            })
        case = {"provenance": "reference GPyTorch twin executed under tests/refshim/shim_gpytorch.py (real torch)",
                "reference_files": ["src/gpytorch_alfi/dataset_alfi.py", "src/gpytorch_alfi/model_alfi.py"],
                "generator": "tests/golden/make_ref_twin_golden.py", "torch": torch.__version__, "name": name,
                "replicate": replicate, "G": dataset.num_outputs, "kernel_jitter": 1e-4,
                "train_t": L(model.train_t.reshape(-1)), "train_y": L(model.train_y.reshape(-1)),
                "variances": L(np.asarray(dataset.gene_variances).reshape(-1)), "points": points}
        with open(os.path.join(out_dir, f"ref_twin_{name}.json"), "w") as fh:
            json.dump(case, fh)
        print(f"ref_twin_{name}.json", "loss(init)", points[0]["loss"], "loss(random)", points[1]["loss"])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else HERE)
