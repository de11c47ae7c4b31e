"""Team sizes of the batched kernel (LFM_BATCHED_TEAM = 1 | 4 | 8): agreement of a 150-step fit, kernel time against the
batch size, per-phase cycles of one step."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dis_project_b200 import ops, _lib
from dis_project_b200.batched import make_restarts
from dis_project_b200.dataset import JaxP53Data, dataset_3d
x, y, _ = dataset_3d(JaxP53Data.synthetic()); y = y.reshape(-1)
TH = make_restarts(np.concatenate([np.full(5, 0.4), np.ones(5), np.full(5, 0.05), [2.5, 1.0]]), 4096)
X = torch.as_tensor(x).cuda(); Y = torch.as_tensor(y).cuda()

def fit(B, chunk=150):
    st = ops.BatchedFitState(TH[:B], 5, 150)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for c in range(0, 150, chunk):
        ops.batched_fit_steps(st, X, Y, 1e-4, chunk)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), st

ref = None
for team in (1, 4, 8):
    os.environ["LFM_BATCHED_TEAM"] = str(team)
    _, st = fit(64)
    h, th, info = st.hist.cpu().numpy(), st.theta.cpu().numpy(), st.info.cpu().numpy()
    if ref is None:
        ref = (h, th)
    print("team", team, "info any", info.any(), "finite", np.isfinite(h).all(),
          "hist rel", np.abs(h - ref[0]).max() / np.abs(ref[0]).max(), "theta rel", np.abs(th - ref[1]).max() / np.abs(ref[1]).max())
for team in (1, 4, 8):
    os.environ["LFM_BATCHED_TEAM"] = str(team)
    fit(4096)
    print("team", team, "kernel ms:", ", ".join(f"B={B} {fit(B)[0]:.2f}" for B in (1, 148, 296, 512, 592, 1024, 1184, 2048, 4096)))
names = ["A", "B", "C tables", "D build M", "E load", "E routine", "E W^TW", "E Schur", "E store", "F beta", "I grad", "J fold", "K adam"]
l = _lib.lib()
for team in (1, 4, 8):
    os.environ["LFM_BATCHED_TEAM"] = str(team)
    for B in (1, 512):
        stamps = torch.zeros(32, dtype=torch.int64, device="cuda")
        for _ in range(2):
            st = ops.BatchedFitState(TH[:B], 5, 3)
            _lib.check(l.lfm_debug_batched_stamps(torch.cuda.current_stream().cuda_stream, B, 105, 5, X.data_ptr(), Y.data_ptr(),
                                                  st.u.data_ptr(), st.adam.data_ptr(), 1e-4, 3, ops.unique_rows(x), ops.distinct_times(x),
                                                  st.hist.data_ptr(), st.info.data_ptr(), stamps.data_ptr()), "stamps")
            torch.cuda.synchronize()
        s = stamps.cpu().numpy()[:14]
        d = np.diff(s)
        print(f"team {team} B={B}: step {s[-1]-s[0]} cycles:", ", ".join(f"{n} {v}" for n, v in zip(names, d)))
os.environ.pop("LFM_BATCHED_TEAM")
print("auto:", ", ".join(f"B={B} {fit(B)[0]:.2f}" for B in (148, 512, 592, 1024, 2048, 4096)))
