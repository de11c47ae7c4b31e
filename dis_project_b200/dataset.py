"""p53 (Barenco et al. 2006) data handling and the (N, 3) input layout of the LFM.

Mirror of the reference's ``src/dataset.py`` public surface (same names, arguments and error
behaviour): ``JaxP53Data`` (:21-210), ``load_barenco_data`` (:213-321), ``flatten_dataset_jax``
(:324-355), ``dataset_3d`` (:358-399).  Host-side only (numpy/pandas); the arrays it produces are
what the CUDA path consumes.  The Barenco CSVs are not distributed with the reference (its
``data/README.md`` points at a Drive folder), so ``JaxP53Data.synthetic`` additionally builds a
p53-shaped data object from the published kinetics (``params_ground_truth``, :189-210) for tests
and benchmarks.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np

GENE_ORDER = ["DDB2", "BIK", "DR5", "p21", "SESN1"]
PROBE_TO_GENE = {
    "203409_at": "DDB2",
    "202284_s_at": "p21",
    "218346_s_at": "SESN1",
    "205780_at": "BIK",
    "209295_at": "DR5",
    "211300_s_at": "p53",
}
# Latent force reported by Barenco et al. at t = 0, 2, ..., 12 h (reference dataset.py:111-113)
F_BARENCO = np.array([0.1845, 1.1785, 1.6160, 0.8156, 0.6862, -0.1828, 0.5131])
# Experimentally measured kinetics (reference dataset.py:201-203)
B_EXACT = np.array([0.0649, 0.0069, 0.0181, 0.0033, 0.0869])
D_EXACT = np.array([0.2829, 0.3720, 0.3617, 0.8000, 0.3573])
S_EXACT = np.array([0.9075, 0.9748, 0.9785, 1.0000, 0.9680])


def load_barenco_data(dir_path: str) -> dict:
    """Read ``barencoPUMA_exprs.csv`` / ``barencoPUMA_se.csv`` and return the rescaled log-normal
    moments as ``(3, 5, 7)`` arrays (replicate, gene, time) plus the p53 channel ``(3, 1, 7)``.

    Raises FileNotFoundError when the CSVs are in neither ``dir_path`` nor ``../data``.
    """
    import pandas as pd

    def _read(name: str):
        for base in (dir_path, "../data"):
            path = os.path.join(base, name)
            if os.path.exists(path):
                return pd.read_csv(path, index_col=0)
        raise FileNotFoundError(
            f"{name} not found in {dir_path!r} or '../data' (the Barenco CSVs are not shipped; see data/README.md)")

    exprs = _read("barencoPUMA_exprs.csv")
    stderr = _read("barencoPUMA_se.csv")
    columns = [f"cARP{rep}-{hour}hrs.CEL" for rep in (1, 2, 3) for hour in range(0, 14, 2)]
    order = GENE_ORDER + ["p53"]

    def _pick(df):
        sub = df[df.index.isin(list(PROBE_TO_GENE))][columns].rename(index=PROBE_TO_GENE)
        return sub.reindex(order).to_numpy(dtype=np.float64)

    log_mean = _pick(exprs)  # (6, 21), column = replicate-major, time-minor
    log_var = _pick(stderr) ** 2
    # moments of the log-normal in linear space
    lin_mean = np.exp(log_mean + 0.5 * log_var)
    lin_var = np.expm1(log_var) * np.exp(2.0 * log_mean + log_var)
    # every probe is rescaled by the sample std (ddof=1) of its first replicate's 7 time points
    scale = np.sqrt(np.var(lin_mean[:, :7], axis=1, ddof=1))[:, None]
    lin_mean = lin_mean / scale
    lin_var = lin_var / scale**2

    def _cube(a):  # (genes, 21) -> (replicate, genes, time)
        return np.ascontiguousarray(a.reshape(a.shape[0], 3, 7).swapaxes(0, 1))

    return {
        "gene_names": list(GENE_ORDER),
        "gene_expressions": _cube(lin_mean[:-1]),
        "gene_variances": _cube(lin_var[:-1]),
        "p53_expressions": _cube(lin_mean[-1:]),
        "p53_variances": _cube(lin_var[-1:]),
    }


class JaxP53Data:
    """Gene expressions, their variances and time points, optionally restricted to one replicate
    and/or a subset of genes.  Same constructor and attributes as the reference class."""

    def __init__(self, replicate: Optional[int] = None, data_dir: str = "data",
                 selected_genes: Optional[Sequence[str]] = None, _gene_data: Optional[dict] = None):
        gene_data = _gene_data if _gene_data is not None else load_barenco_data(data_dir)
        all_genes: List[str] = list(gene_data["gene_names"])

        assert replicate is None or 0 <= replicate < 3, "Invalid replicate number"

        expr = np.asarray(gene_data["gene_expressions"], dtype=np.float64)
        var = np.asarray(gene_data["gene_variances"], dtype=np.float64)
        if selected_genes is not None:
            selected_genes = list(selected_genes)
            unknown = set(selected_genes) - set(all_genes)
            if unknown:
                raise ValueError(f"Invalid gene names provided: {', '.join(sorted(unknown))}")
            if len(set(selected_genes)) != len(selected_genes):
                dup = sorted({g for g in selected_genes if selected_genes.count(g) > 1})
                raise ValueError(f"Duplicate genes provided: {', '.join(dup)}")
            if len(selected_genes) == 0:
                raise ValueError("Empty list of genes selected, set 'selected_genes' to None")
            # rows are taken in the dataset's own gene order (reference dataset.py:90-94), while
            # selected_indices follows the order the caller gave (:95-97)
            keep = [i for i, g in enumerate(all_genes) if g in selected_genes]
            self.selected_indices = [all_genes.index(g) for g in selected_genes]
            self.gene_names = selected_genes
            expr, var = expr[:, keep], var[:, keep]
        else:
            self.selected_indices = list(range(len(all_genes)))
            self.gene_names = all_genes

        self.gene_expressions = expr
        self.gene_variances_raw = var
        self.num_genes = len(self.gene_names)
        self.timepoints = np.linspace(0, 12, 7)
        self.f_observed = F_BARENCO.reshape(1, 1, 7)

        if replicate is None:
            reps = range(expr.shape[0])
            self.data = [(self.timepoints, expr[r, g]) for r in reps for g in range(self.num_genes)]
            self.gene_variances = np.array([var[r, g] for r in reps for g in range(self.num_genes)])
        else:
            self.gene_expressions = expr[replicate:replicate + 1]
            self.data = [(self.timepoints, self.gene_expressions[0, g]) for g in range(self.num_genes)]
            self.gene_variances = var[replicate:replicate + 1]

    # -- container protocol ---------------------------------------------------------------------
    def __getitem__(self, index):
        if index < 0 or index >= len(self.data):
            raise IndexError("Index out of range")
        return self.data[index]

    def __len__(self):
        return len(self.data)

    @property
    def shape(self):
        return np.array(self.data).shape

    def params_ground_truth(self):
        """(B_exact, S_exact, D_exact) of the selected genes."""
        idx = self.selected_indices
        return B_EXACT[idx], S_EXACT[idx], D_EXACT[idx]

    # -- synthetic stand-in for the missing CSVs ---------------------------------------------------
    @classmethod
    def synthetic(cls, replicate: Optional[int] = None, selected_genes: Optional[Sequence[str]] = None,
                  seed: int = 42) -> "JaxP53Data":
        """p53-shaped data (3 replicates x 5 genes x 7 times) generated by integrating the SIM ODE
        dx_j/dt = B_j + S_j f(t) - D_j x_j with the published kinetics and Barenco's latent profile
        (piecewise-linear f), plus replicate noise.  Used when the real CSVs are unavailable."""
        rng = np.random.default_rng(seed)
        t_fine = np.linspace(0.0, 12.0, 1201)
        f_fine = np.interp(t_fine, np.linspace(0, 12, 7), F_BARENCO)
        dt = t_fine[1] - t_fine[0]
        x = np.empty((5, t_fine.size))
        x[:, 0] = B_EXACT / D_EXACT
        for n in range(1, t_fine.size):
            x[:, n] = x[:, n - 1] + dt * (B_EXACT + S_EXACT * f_fine[n - 1] - D_EXACT * x[:, n - 1])
        clean = x[:, ::200]  # (5, 7)
        expr = np.stack([clean + 0.05 * rng.standard_normal(clean.shape) for _ in range(3)])
        var = rng.uniform(0.002, 0.02, size=expr.shape)
        gd = {"gene_names": list(GENE_ORDER), "gene_expressions": expr, "gene_variances": var}
        return cls(replicate=replicate, selected_genes=selected_genes, _gene_data=gd)


def flatten_dataset_jax(dataset):
    """(train_t, train_y): time points tiled per entry and the concatenated expressions."""
    n = len(dataset)
    train_t = np.tile(np.asarray(dataset[0][0], dtype=np.float64), n)
    train_y = np.concatenate([np.asarray(dataset[i][1], dtype=np.float64) for i in range(n)]).reshape(-1)
    return train_t, train_y


def dataset_3d(data):
    """(training_times (N,3), gene_expressions (N,1), variances (N,1)).

    Row layout [time, gene index, flag=1]: gene-major inside a replicate, replicate-major outside
    (reference dataset.py:380-391).
    """
    num_genes = data.num_genes
    entries = np.array([data[i] for i in range(len(data))], dtype=np.float64)  # (entries, 2, T)
    replicates = entries.shape[0] // num_genes
    times = entries[0, 0, :]
    T = times.shape[0]
    t_col = np.tile(times, entries.shape[0])
    g_col = np.tile(np.repeat(np.arange(num_genes), T), replicates).astype(np.float64)
    flag = np.ones(num_genes * T * replicates, dtype=np.float64)
    training_times = np.stack((t_col, g_col, flag), axis=-1)
    gene_expressions = entries[:, 1, :].reshape(-1, 1)
    variances = np.asarray(data.gene_variances, dtype=np.float64).reshape(-1, 1)
    return training_times, gene_expressions, variances
