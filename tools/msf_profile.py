import sys, os, time, cProfile, pstats
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from dis_project_b200.batched import make_restarts, multi_start_fit
from dis_project_b200.dataset import JaxP53Data, dataset_3d
x, y, _ = dataset_3d(JaxP53Data.synthetic()); y = y.reshape(-1)
TH = make_restarts(np.concatenate([np.full(5, 0.4), np.ones(5), np.full(5, 0.05), [2.5, 1.0]]), 512)
for _ in range(3): multi_start_fit(x, y, TH, 1e-4, num_iters=150, chunk=10)
pr = cProfile.Profile(); pr.enable()
for _ in range(50): multi_start_fit(x, y, TH, 1e-4, num_iters=150, chunk=10)
pr.disable()
ps = pstats.Stats(pr); ps.sort_stats("cumulative").print_stats(28)
