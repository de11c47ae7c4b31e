"""Golden vectors from the reference's GPyTorch twin, EXECUTED HERE: ``src/gpytorch_alfi/{dataset_alfi,model_alfi}.py``
imported unmodified under ``tests/refshim/shim_gpytorch.py`` (a stand-in for gpytorch on real torch) and run on the CSVs
of ``tests/golden/ref_csv``, exactly as ``src/gpytorch_alfi/main_alfi.py:24-35`` and ``trainer_alfi.py:170-177`` do:

    dataset = PyTorchDataset(replicate=0, data_dir=...)
    model = ExactLFM(dataset, dataset.gene_variances.reshape(-1))
    loss = -ExactMarginalLogLikelihood(model.likelihood, model)(model(model.train_t), model.train_y.squeeze())

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
            points.append({
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
