import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dis_project_b200 import ops
from dis_project_b200.batched import make_restarts, multi_start_fit
from dis_project_b200.dataset import JaxP53Data, dataset_3d
x, y, _ = dataset_3d(JaxP53Data.synthetic()); y = y.reshape(-1)
TH = make_restarts(np.concatenate([np.full(5, 0.4), np.ones(5), np.full(5, 0.05), [2.5, 1.0]]), 4096)
X = torch.as_tensor(x).cuda(); Y = torch.as_tensor(y).cuda()
def kernel_only(B, chunk):
    st = ops.BatchedFitState(TH[:B], 5, 150)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for c in range(0, 150, chunk):
        ops.batched_fit_steps(st, X, Y, 1e-4, chunk)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for B in (148, 512, 592, 1184, 4096):
    kernel_only(B, 150)
    print("B", B, "kernel ms chunk150", round(kernel_only(B, 150), 2), "chunk10", round(kernel_only(B, 10), 2), "chunk1", round(kernel_only(B, 1), 2))
for B in (512, 4096):
    for chunk in (10, 150):
        multi_start_fit(x, y, TH[:B], 1e-4, num_iters=150, chunk=chunk)
        t0 = time.perf_counter(); multi_start_fit(x, y, TH[:B], 1e-4, num_iters=150, chunk=chunk); torch.cuda.synchronize()
        print("B", B, "chunk", chunk, "multi_start_fit wall ms", round(1e3 * (time.perf_counter() - t0), 2))
