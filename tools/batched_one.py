"""One batched-fit launch (B restarts, S steps) for profiling."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dis_project_b200 import ops
from dis_project_b200.batched import make_restarts
from dis_project_b200.dataset import JaxP53Data, dataset_3d
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10
x, y, _ = dataset_3d(JaxP53Data.synthetic()); y = y.reshape(-1)
TH = make_restarts(np.concatenate([np.full(5, 0.4), np.ones(5), np.full(5, 0.05), [2.5, 1.0]]), B)
X = torch.as_tensor(x).cuda(); Y = torch.as_tensor(y).cuda()
for it in range(2):
    st = ops.BatchedFitState(TH, 5, S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.batched_fit_steps(st, X, Y, 1e-4, S); e1.record(); torch.cuda.synchronize()
print("B", B, "steps", S, "ms", e0.elapsed_time(e1), "finite", bool(torch.isfinite(st.hist).all()))
