"""Entry point with the call sequence of the reference's src/main.py:30-78 on the B200 path.

Same steps, same names: load p53 data -> dataset_3d -> Dataset -> ExactLFM(jitter=1e-4) -> CustomConjMLL(negative=True)
-> adam(0.01) -> JaxTrainer(...).fit(num_steps_per_epoch=1000) -> print_hyperparams -> latent_predict ->
GeneExpressionPredictor.  Two differences, both forced by this image: the Barenco CSVs are not part of the reference
checkout, so without `--data-dir` the p53-shaped synthetic set (`JaxP53Data.synthetic`, ground-truth kinetics of
dataset.py:201-203) is fitted; matplotlib is absent, so the three figures of main.py:67-78 (`plot_lf`,
`plot_predictions`, `plot_comparison_gpjax`) are written as SVG by dis_project_b200/plotter.py, next to the same
numbers as CSV columns.

  python examples/main.py [--data-dir data] [--replicate 0] [--out-dir out]
"""
from __future__ import annotations

import argparse
import csv
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from dis_project_b200.dataset import JaxP53Data, dataset_3d  # noqa: E402
from dis_project_b200.gpx_compat import Dataset, adam  # noqa: E402
from dis_project_b200.model import ExactLFM  # noqa: E402
from dis_project_b200.objectives import CustomConjMLL  # noqa: E402
from dis_project_b200.trainer import JaxTrainer  # noqa: E402
from dis_project_b200.utils import GeneExpressionPredictor, generate_test_times, print_hyperparams  # noqa: E402
from dis_project_b200 import plotter  # noqa: E402
from dis_project_b200.plotter import plot_comparison_gpjax, plot_lf  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--data-dir", default=None, help="directory with the Barenco CSVs (dataset.py:233-243); default: synthetic")
    ap.add_argument("--replicate", type=int, default=0, help="replicate to fit (main.py:32 uses 0); -1 = all three (notebook.py:36)")
    ap.add_argument("--out-dir", default="out")
    args = ap.parse_args()
    rep = None if args.replicate < 0 else args.replicate

    # Load the data (main.py:32)
    p53_data = JaxP53Data(replicate=rep, data_dir=args.data_dir) if args.data_dir else JaxP53Data.synthetic(replicate=rep)
    # Artificially augment the data to 3D (main.py:35)
    training_times, gene_expressions, variances = dataset_3d(p53_data)
    dataset_train = Dataset(training_times, gene_expressions)
    # Model, loss, optimiser, trainer (main.py:41-55)
    custom_posterior = ExactLFM(jitter=np.array(1e-4), data=p53_data)
    loss = CustomConjMLL(negative=True)
    optimiser = adam(0.01)
    trainer = JaxTrainer(model=custom_posterior, objective=loss, training_data=dataset_train, optim=optimiser,
                         key=None, num_iters=150)
    print("Training model...")
    trained_model, training_history = trainer.fit(num_steps_per_epoch=1000)
    print(f"NLML {training_history[0]:.6f} -> {training_history[-1]:.6f} in {len(training_history)} steps")

    os.makedirs(args.out_dir, exist_ok=True)
    plotter.PLOTS_DIR = os.path.join(args.out_dir, "plots")
    print_hyperparams(trained_model, p53_data, file=os.path.join(args.out_dir, "hyperparams.csv"))

    print("Making predictions...")
    testing_times = generate_test_times()
    latent_dist = trained_model.latent_predict(testing_times, p53_data)
    mean, std = np.asarray(latent_dist.mean()), np.asarray(latent_dist.stddev())
    with open(os.path.join(args.out_dir, "latent_force.csv"), "w", newline="") as fh:   # plot_lf (plotter.py:33-101)
        w = csv.writer(fh)
        w.writerow(["t", "mean", "stddev", "lower_2sd", "upper_2sd"])
        for t, m, sd in zip(testing_times[:, 0], mean, std):
            w.writerow([t, m, sd, m - 2 * sd, m + 2 * sd])
    # Plot latent force (main.py:67-69)
    f = np.asarray(p53_data.f_observed).squeeze()
    plot_lf(testing_times, latent_dist, y_scatter=f, stddev=2)
    # Plot gene expression predictions (main.py:71-73)
    gene_predictor = GeneExpressionPredictor(trained_model, p53_data)
    gene_predictor.plot_predictions(p53_data)
    # Plot hyperparameter comparison (main.py:75-76)
    plot_comparison_gpjax(trained_model, p53_data)
    xpr_times, means, stds = gene_predictor.predict()
    t100 = xpr_times[:gene_predictor.t, 0]
    with open(os.path.join(args.out_dir, "gene_expression.csv"), "w", newline="") as fh:   # utils.py:173-234
        w = csv.writer(fh)
        w.writerow(["t"] + [f"{g}_{k}" for g in p53_data.gene_names for k in ("mean", "stddev")])
        for i, t in enumerate(t100):
            w.writerow([t] + [v for m, sd in zip(means, stds) for v in (np.asarray(m)[i], np.asarray(sd)[i])])
    basal_true, sensitivity_true, decay_true = p53_data.params_ground_truth()   # plot_comparison_gpjax (plotter.py:116-193)
    with open(os.path.join(args.out_dir, "comparison.csv"), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["gene", "B_learned", "B_exact", "S_learned", "S_exact", "D_learned", "D_exact"])
        for i, g in enumerate(p53_data.gene_names):
            w.writerow([g, float(trained_model.true_b[i]), float(np.asarray(basal_true)[i]),
                        float(trained_model.true_s[i]), float(np.asarray(sensitivity_true)[i]),
                        float(trained_model.true_d[i]), float(np.asarray(decay_true)[i])])
    print("wrote", ", ".join(sorted(os.listdir(args.out_dir))), "to", args.out_dir)


if __name__ == "__main__":
    main()
