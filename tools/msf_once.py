"""One multi_start_fit of `B` p53 restarts (150 steps, chunk 10) on the current GPU, after a warm-up -- for launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dis_project_b200.batched import make_restarts, multi_start_fit
from dis_project_b200.dataset import JaxP53Data, dataset_3d
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x, y, _ = dataset_3d(JaxP53Data.synthetic()); y = y.reshape(-1)
TH = make_restarts(np.concatenate([np.full(5, 0.4), np.ones(5), np.full(5, 0.05), [2.5, 1.0]]), B)
for it in range(2):
    r = multi_start_fit(x, y, TH, 1e-4, num_iters=150, chunk=10)
print("B", B, "best", r.best_loss, r.best_id)
