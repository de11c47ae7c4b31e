"""Quick probe of the team kernel (four warps per LFM): agreement with the warp-per-LFM kernel over a 150-step fit,
kernel time at a few batch sizes, per-phase cycles of one step alone / with 512 LFMs on the GPU."""
import sys, os
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
for team in (1, 4):
    os.environ["LFM_BATCHED_TEAM"] = str(team)
    _, st = fit(64, chunk=10)
    h, th, info = st.hist.cpu().numpy(), st.theta.cpu().numpy(), st.info.cpu().numpy()
    if ref is None:
        ref = (h, th)
    print("team", team, "info any", info.any(), "finite", np.isfinite(h).all(),
          "hist rel", np.abs(h - ref[0]).max() / np.abs(ref[0]).max(), "theta rel", np.abs(th - ref[1]).max() / np.abs(ref[1]).max(),
          "hist step0 rel", np.abs(h[:, 0] - ref[0][:, 0]).max() / np.abs(ref[0][:, 0]).max())
teams = [int(t) for t in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["4"])]
for team in teams:
    os.environ["LFM_BATCHED_TEAM"] = str(team)
    fit(512)
    print("team", team, "kernel ms:", ", ".join(f"B={B} {min(fit(B)[0] for _ in range(3)):.3f}" for B in (1, 148, 296, 444, 512, 592)),
          " chunk10 B=512:", f"{min(fit(512, 10)[0] for _ in range(3)):.3f}")
names = ["A", "B", "C tables", "D build M", "E load", "E routine", "E W^TW", "E Schur", "E store", "F beta", "I grad", "J fold", "K adam"]
l = _lib.lib()
for team in teams:
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
        full = stamps.cpu().numpy()
        fx = full[16:22]
        print(f"   fine: B.stage1 {fx[0]-s[1]}, B.any+sync {fx[1]-fx[0]}, B.stage2 {fx[2]-fx[1]}, B.zz {s[2]-fx[2]};"
              f" F.matvec {fx[5]-s[9]}, F.sum {s[10]-fx[5]}; I.pairs {fx[3]-s[10]}, I.sum {fx[4]-fx[3]}, I.rows {s[11]-fx[4]}")
os.environ.pop("LFM_BATCHED_TEAM")
