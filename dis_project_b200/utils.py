"""Host-side helpers that define input layouts (reference ``src/utils.py:237-314``)."""
from __future__ import annotations

from typing import Optional

import numpy as np


def generate_test_times(t: Optional[int] = 100) -> np.ndarray:
    """(t, 3) latent test inputs [linspace(0, 13, t), -1, 0] (reference utils.py:268-287)."""
    times = np.linspace(0, 13, t)
    return np.stack((times, np.repeat(-1.0, t), np.repeat(0.0, t)), axis=-1)


def generate_test_times_pred(t: Optional[int] = 100, num_genes: int = 5) -> np.ndarray:
    """(t * num_genes, 3) gene-expression test inputs with gene indices 1..G and flag 1
    (reference utils.py:290-314; the off-by-one indices are the reference's)."""
    times = np.tile(np.linspace(0, 13, t), num_genes)
    genes = np.repeat(np.arange(1, num_genes + 1), t).astype(np.float64)
    return np.stack((times, genes, np.ones(times.shape[0])), axis=1)


class GeneExpressionPredictor:
    """Gene-expression predictions of a trained model (reference utils.py:40-234): `predict()` returns what
    `plot_predictions` draws; the drawing itself is `plotter.plot_gene_predictions` (SVG, no matplotlib here)."""

    def __init__(self, model, p53_data, t: Optional[int] = 100):
        self.model = model
        self.p53_data = p53_data
        self.num_genes = p53_data.num_genes
        self.gene_names = p53_data.gene_names
        self.t = t

    def generate_test_times_pred(self) -> np.ndarray:
        return generate_test_times_pred(self.t, self.num_genes)

    def decompose_predictions(self, pred) -> tuple:
        return tuple(pred[i * self.t:(i + 1) * self.t] for i in range(self.num_genes))

    def decompose_predictions2(self, pred) -> tuple:
        """Five-gene variant with blocks 3 and 4 swapped, as the reference does (utils.py:135-140)."""
        n = self.t
        g1, g2, g4, g3, g5 = pred[:n], pred[n:2 * n], pred[2 * n:3 * n], pred[3 * n:4 * n], pred[4 * n:]
        return g1, g2, g3, g4, g5

    def predict(self):
        """(test_times, means per gene, stddevs per gene) -- reference utils.py:173-182."""
        xpr_times = self.generate_test_times_pred()
        dist = self.model.multi_gene_predict(xpr_times, self.p53_data)
        split = self.decompose_predictions2 if self.num_genes == 5 else self.decompose_predictions
        return xpr_times, split(dist.mean()), split(dist.stddev())


    def plot_predictions(self, p53_data, stddev: Optional[int] = 2, save: Optional[bool] = True,
                         save_name: Optional[str] = None):
        """Plot gene expression predictions (reference utils.py:143-234): gpjax_gxpr[_<save_name>]."""
        from .plotter import plot_gene_predictions

        return plot_gene_predictions(self, p53_data, stddev=stddev, save=save, save_name=save_name)


def print_hyperparams(model, dataset, file: Optional[str] = None) -> list:
    """Table of learned B, S, D per gene plus l (reference utils.py:237-265).  Returns the rows;
    writes a CSV when `file` is given."""
    rows = [[name, float(model.true_b[i]), float(model.true_s[i]), float(model.true_d[i])]
            for i, name in enumerate(dataset.gene_names)]
    rows.append(["lengthscale", float(np.asarray(model.l)), "", ""])
    headers = ["Gene", "B", "S", "D"]
    try:
        from tabulate import tabulate

        print(tabulate(rows, headers=headers))
    except Exception:  # pragma: no cover
        print(headers)
        for r in rows:
            print(r)
    if file:
        import csv

        with open(file, "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(headers)
            w.writerows(rows)
    return rows
