"""Per-phase cycles of one optimiser step of the warp-per-LFM kernel (LFM 0, second step), alone and with a full GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dis_project_b200 import ops, _lib
from dis_project_b200.batched import make_restarts
from dis_project_b200.dataset import JaxP53Data, dataset_3d
x, y, _ = dataset_3d(JaxP53Data.synthetic()); y = y.reshape(-1)
names = ["A", "B", "C tables", "D build M", "E load", "E routine", "E W^TW", "E Schur", "E store", "F beta", "I grad", "J fold", "K adam"]
l = _lib.lib()
for B in (1, 1036):
    TH = make_restarts(np.concatenate([np.full(5, 0.4), np.ones(5), np.full(5, 0.05), [2.5, 1.0]]), B)
    st = ops.BatchedFitState(TH, 5, 3)
    X = torch.as_tensor(x).cuda(); Y = torch.as_tensor(y).cuda()
    stamps = torch.zeros(32, dtype=torch.int64, device="cuda")
    for _ in range(2):
        st = ops.BatchedFitState(TH, 5, 3)
        _lib.check(l.lfm_debug_batched_stamps(torch.cuda.current_stream().cuda_stream, B, 105, 5, X.data_ptr(), Y.data_ptr(),
                                              st.u.data_ptr(), st.adam.data_ptr(), 1e-4, 3, ops.unique_rows(x), ops.distinct_times(x),
                                              st.hist.data_ptr(), st.info.data_ptr(), stamps.data_ptr()), "stamps")
        torch.cuda.synchronize()
    s = stamps.cpu().numpy()[:14]
    d = np.diff(s)
    print(f"B={B}: step {s[-1]-s[0]} cycles:", ", ".join(f"{n} {v}" for n, v in zip(names, d)))
