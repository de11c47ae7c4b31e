"""Do the restarts of config 4 cost the same?  148 copies of ONE start point (one team per SM, no contention): the kernel time
is that LFM's own 150-step time.  Printed for a sample of restarts, with the time of 148 DIFFERENT restarts beside it."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dis_project_b200 import ops
from dis_project_b200.batched import make_restarts
from dis_project_b200.dataset import JaxP53Data, dataset_3d
x, y, _ = dataset_3d(JaxP53Data.synthetic()); y = y.reshape(-1)
TH = make_restarts(np.concatenate([np.full(5, 0.4), np.ones(5), np.full(5, 0.05), [2.5, 1.0]]), 4096)
X = torch.as_tensor(x).cuda(); Y = torch.as_tensor(y).cuda()
os.environ["LFM_BATCHED_TEAM"] = "4"
def fit(th):
    st = ops.BatchedFitState(th, 5, 150)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.batched_fit_steps(st, X, Y, 1e-4, 150); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
fit(TH[:148])
print("148 different restarts:", round(min(fit(TH[:148]) for _ in range(3)), 3), "ms")
ts = []
for k in list(range(0, 64)) :
    th = np.repeat(TH[k:k + 1], 148, axis=0)
    ts.append(min(fit(th) for _ in range(2)))
ts = np.array(ts)
print("one restart x 148: min %.3f median %.3f max %.3f ms" % (ts.min(), np.median(ts), ts.max()))
print(np.round(ts, 3).tolist())
