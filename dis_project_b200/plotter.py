"""Figures of the reference's ``src/plotter.py:33-234`` (and the plotting half of ``src/utils.py:143-234``).

Same functions and signatures -- ``plot_lf``, ``plot_comparison_gpjax``, ``plot_gene_predictions`` (what
``GeneExpressionPredictor.plot_predictions`` draws), ``save_plot``, ``clean_legend`` -- and the same file names
(``gpjax_lf.png``, ``gpjax_lf_<name>.png``, ``gpjax_gxpr.png``, ``gpjax_comparison.png``).  matplotlib is not part of this
image, so the drawing itself goes through a small dependency-free SVG writer (``_Figure`` below): the files then carry
the extension ``.svg`` instead of ``.png``.  Host-side presentation only; every number drawn comes from the CUDA path.

Figures are written to ``PLOTS_DIR`` (default ``./plots``; the reference writes next to its own source file,
plotter.py:216-234, which is read-only here).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np

PLOTS_DIR = os.environ.get("LFM_PLOTS_DIR", os.path.join(os.getcwd(), "plots"))
COLORS = ["#1f77b4", "#ff7f0e", "#2ca02c", "#d62728"]   # colors[0] = data / truth, colors[1] = model (plotter.py:31)


def _np(a) -> np.ndarray:
    try:
        import torch

        if isinstance(a, torch.Tensor):
            return a.detach().cpu().numpy().astype(np.float64)
    except ImportError:  # pragma: no cover
        pass
    return np.asarray(a, dtype=np.float64)


class _Axes:
    def __init__(self, title: str = "", xlabel: str = "", ylabel: str = ""):
        self.title, self.xlabel, self.ylabel = title, xlabel, ylabel
        self.items: List[tuple] = []
        self.xticklabels: Optional[Sequence[str]] = None

    def fill_between(self, x, lo, hi, color, alpha=0.2, label=None):
        self.items.append(("band", _np(x), _np(lo), _np(hi), color, alpha, label))

    def plot(self, x, y, color, dashed=False, label=None, width=1.5):
        self.items.append(("line", _np(x), _np(y), color, dashed, label, width))

    def scatter(self, x, y, color, marker="o", label=None):
        self.items.append(("scatter", _np(x), _np(y), color, marker, label))

    def bar(self, x, h, width, color, label=None):
        self.items.append(("bar", _np(x), _np(h), float(width), color, label))

    def _range(self):
        xs, ys = [], []
        for it in self.items:
            if it[0] == "band":
                xs += [it[1]]; ys += [it[2], it[3]]
            elif it[0] == "bar":
                xs += [it[1] - it[3] / 2, it[1] + it[3] / 2]; ys += [it[2], np.zeros(1)]
            else:
                xs += [it[1]]; ys += [it[2]]
        x = np.concatenate([np.ravel(v) for v in xs]) if xs else np.array([0.0, 1.0])
        y = np.concatenate([np.ravel(v) for v in ys]) if ys else np.array([0.0, 1.0])
        x, y = x[np.isfinite(x)], y[np.isfinite(y)]
        x0, x1, y0, y1 = float(x.min()), float(x.max()), float(y.min()), float(y.max())
        if x1 <= x0:
            x1 = x0 + 1.0
        if y1 <= y0:
            y1 = y0 + 1.0
        pad = 0.05 * (y1 - y0)
        return x0, x1, y0 - pad, y1 + pad

    def svg(self, ox: float, oy: float, w: float, h: float) -> str:
        x0, x1, y0, y1 = self._range()
        ml, mr, mt, mb = 52.0, 110.0, 22.0, 40.0
        pw, ph = w - ml - mr, h - mt - mb
        X = lambda v: ox + ml + (np.asarray(v) - x0) / (x1 - x0) * pw
        Y = lambda v: oy + mt + ph - (np.asarray(v) - y0) / (y1 - y0) * ph
        out = [f'<rect x="{ox + ml:.1f}" y="{oy + mt:.1f}" width="{pw:.1f}" height="{ph:.1f}" fill="none" stroke="#444"/>']
        for k in range(5):   # ticks
            ty = y0 + k * (y1 - y0) / 4
            out.append(f'<text x="{ox + ml - 4:.1f}" y="{float(Y(ty)) + 3:.1f}" font-size="9" text-anchor="end">{ty:.3g}</text>')
            if self.xticklabels is None:
                tx = x0 + k * (x1 - x0) / 4
                out.append(f'<text x="{float(X(tx)):.1f}" y="{oy + mt + ph + 12:.1f}" font-size="9" text-anchor="middle">{tx:.3g}</text>')
        legend = []
        for it in self.items:
            if it[0] == "band":
                _, x, lo, hi, color, alpha, label = it
                pts = " ".join(f"{a:.2f},{b:.2f}" for a, b in zip(np.concatenate([X(x), X(x)[::-1]]), np.concatenate([Y(hi), Y(lo)[::-1]])))
                out.append(f'<polygon points="{pts}" fill="{color}" fill-opacity="{alpha}" stroke="none"/>')
            elif it[0] == "line":
                _, x, y, color, dashed, label, width = it
                pts = " ".join(f"{a:.2f},{b:.2f}" for a, b in zip(X(x), Y(y)))
                dash = ' stroke-dasharray="4 3"' if dashed else ""
                out.append(f'<polyline points="{pts}" fill="none" stroke="{color}" stroke-width="{width}"{dash}/>')
            elif it[0] == "scatter":
                _, x, y, color, marker, label = it
                for a, b in zip(np.ravel(X(x)), np.ravel(Y(y))):
                    if marker == "x":
                        out.append(f'<path d="M{a - 3:.1f},{b - 3:.1f} L{a + 3:.1f},{b + 3:.1f} M{a - 3:.1f},{b + 3:.1f} L{a + 3:.1f},{b - 3:.1f}" stroke="{color}" stroke-width="1.5"/>')
                    else:
                        out.append(f'<circle cx="{a:.1f}" cy="{b:.1f}" r="2.5" fill="{color}"/>')
            else:
                _, x, hgt, width, color, label = it
                for a, b in zip(x, hgt):
                    xa, xb = float(X(a - width / 2)), float(X(a + width / 2))
                    ya, yb = float(Y(max(b, 0.0))), float(Y(min(b, 0.0)))
                    out.append(f'<rect x="{xa:.1f}" y="{ya:.1f}" width="{xb - xa:.1f}" height="{max(yb - ya, 0.5):.1f}" fill="{color}"/>')
            if label and label not in [l for l, _ in legend]:   # clean_legend: one entry per label
                legend.append((label, it[4] if it[0] in ("band", "bar") else it[3]))
        if self.xticklabels is not None:
            for i, name in enumerate(self.xticklabels):
                out.append(f'<text x="{float(X(i)):.1f}" y="{oy + mt + ph + 12:.1f}" font-size="9" text-anchor="middle">{name}</text>')
        for i, (label, color) in enumerate(legend):
            ly = oy + mt + 10 + 13 * i
            out.append(f'<rect x="{ox + ml + pw + 8:.1f}" y="{ly - 7:.1f}" width="10" height="8" fill="{color}"/>')
            out.append(f'<text x="{ox + ml + pw + 22:.1f}" y="{ly:.1f}" font-size="9">{label}</text>')
        out.append(f'<text x="{ox + ml + pw / 2:.1f}" y="{oy + 13:.1f}" font-size="11" text-anchor="middle">{self.title}</text>')
        out.append(f'<text x="{ox + ml + pw / 2:.1f}" y="{oy + h - 6:.1f}" font-size="10" text-anchor="middle">{self.xlabel}</text>')
        out.append(f'<text x="{ox + 11:.1f}" y="{oy + mt + ph / 2:.1f}" font-size="10" text-anchor="middle" '
                   f'transform="rotate(-90 {ox + 11:.1f} {oy + mt + ph / 2:.1f})">{self.ylabel}</text>')
        return "\n".join(out)


class _Figure:
    """rows x cols grid of axes rendered as one SVG document."""

    def __init__(self, rows: int = 1, cols: int = 1, width: float = 720.0, height: float = 240.0):
        self.rows, self.cols, self.width, self.height = rows, cols, width, height
        self.axes = [_Axes() for _ in range(rows * cols)]

    def render(self) -> str:
        cw, ch = self.width / self.cols, self.height / self.rows
        body = [ax.svg((i % self.cols) * cw, (i // self.cols) * ch, cw, ch) for i, ax in enumerate(self.axes)]
        return (f'<svg xmlns="http://www.w3.org/2000/svg" width="{self.width:.0f}" height="{self.height:.0f}" '
                f'viewBox="0 0 {self.width:.0f} {self.height:.0f}" font-family="sans-serif">\n'
                f'<rect width="100%" height="100%" fill="white"/>\n' + "\n".join(body) + "\n</svg>\n")


def clean_legend(ax):
    """Duplicate legend entries are dropped at render time (reference plotter.py:196-213); kept for API parity."""
    return ax


def save_plot(plot_name: str, fig: Optional[_Figure] = None) -> str:
    """Write the figure under PLOTS_DIR (reference plotter.py:216-234) and return the path."""
    os.makedirs(PLOTS_DIR, exist_ok=True)
    base = os.path.splitext(plot_name)[0]
    path = os.path.join(PLOTS_DIR, base + ".svg")
    print(f"Saving plot to {path}")
    with open(path, "w") as fh:
        fh.write(fig.render() if fig is not None else _Figure().render())
    return path


def plot_lf(testing_times, predictive_dist, stddev: Optional[int] = 2, y_scatter=None, title: Optional[str] = None,
            save: Optional[bool] = True, save_name: Optional[str] = None):
    """Latent force with its +-`stddev` sigma band and Barenco's measured profile as crosses (fig. 1a of Lawrence et al.;
    reference plotter.py:33-113).  Returns the figure; writes gpjax_lf[_<save_name>] when `save`."""
    mean, std = _np(predictive_dist.mean()), _np(predictive_dist.stddev())
    t = _np(testing_times)[:, 0]
    fig = _Figure(1, 1, 720, 240)
    ax = fig.axes[0]
    ax.fill_between(t, mean - stddev * std, mean + stddev * std, COLORS[1], 0.2, label=f"{stddev} sigma")
    ax.plot(t, mean - stddev * std, COLORS[1], dashed=True, width=1)
    ax.plot(t, mean + stddev * std, COLORS[1], dashed=True, width=1)
    ax.plot(t, mean, COLORS[1], label="Predictive mean")
    if y_scatter is not None:
        y = _np(y_scatter).reshape(-1)
        ax.scatter(np.linspace(0, 12, len(y)), y, COLORS[0], marker="x", label="True values")
    ax.xlabel, ax.ylabel = "Time", "mRNA Expression"
    ax.title = "Latent Force Model (GPJax)" + (f" - {title}" if title is not None else "")
    if save:
        save_plot(f"gpjax_lf_{save_name}.png" if save_name is not None else "gpjax_lf.png", fig)
    return fig


def plot_comparison_gpjax(model, dataset, save: Optional[bool] = True):
    """Learned against measured basal rates, sensitivities and decay rates per gene (reference plotter.py:116-193)."""
    basal_true, sensitivity_true, decay_true = dataset.params_ground_truth()
    names = list(dataset.gene_names)
    x = np.arange(len(names), dtype=np.float64)
    fig = _Figure(1, 3, 720, 240)
    for ax, title, learned, true, ll, lt in (
            (fig.axes[0], "Basal rates", model.true_b, basal_true, "basal_rates", "B_exact"),
            (fig.axes[1], "Sensitivities", model.true_s, sensitivity_true, "kxx_sensitivities", "S_exact"),
            (fig.axes[2], "Decay rates", model.true_d, decay_true, "Calculated", "Measured")):
        ax.bar(x + 0.2, _np(learned).reshape(-1), 0.4, COLORS[1], label=ll)
        ax.bar(x - 0.2, _np(true).reshape(-1), 0.4, COLORS[0], label=lt)
        ax.title = title
        ax.xticklabels = names
    if save:
        save_plot("gpjax_comparison.png", fig)
    return fig


def plot_gene_predictions(predictor, p53_data, stddev: Optional[int] = 2, save: Optional[bool] = True,
                          save_name: Optional[str] = None):
    """Predicted expression of every gene with its band and the measurements (reference utils.py:143-234)."""
    xpr_times, means, stds = predictor.predict()
    t = _np(xpr_times)[:predictor.t, 0]
    G = predictor.num_genes
    fig = _Figure(G, 1, 720, 200.0 * G)
    expr = _np(p53_data.gene_expressions)
    for i in range(G):
        ax = fig.axes[i]
        m, s = _np(means[i]), _np(stds[i])
        ax.fill_between(t, m - stddev * s, m + stddev * s, COLORS[1], 0.2, label=f"{stddev} sigma")
        ax.plot(t, m - stddev * s, COLORS[1], dashed=True, width=1)
        ax.plot(t, m + stddev * s, COLORS[1], dashed=True, width=1)
        ax.plot(t, m, COLORS[1], label="Predictive mean")
        obs = expr[:, i].reshape(-1)
        ax.scatter(np.tile(_np(p53_data.timepoints), expr.shape[0]), obs, COLORS[0], label="True values")
        ax.title = f"{predictor.gene_names[i]} Expression Over Time"
        ax.xlabel, ax.ylabel = "Time", "Expression Level"
    if save:
        save_plot(f"gpjax_gxpr_{save_name}.png" if save_name is not None else "gpjax_gxpr.png", fig)
    return fig
