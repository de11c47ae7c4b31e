"""Host-side cost of one multi_start_fit of B p53 restarts on ONE GPU (no collective): wall clock against the device-side
time (CUDA events from the first enqueued copy to the last read-back) and against the kernel alone; then a cProfile."""
import sys, os, time, json, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dis_project_b200 import batched
from dis_project_b200.batched import make_restarts, multi_start_fit
from dis_project_b200.dataset import JaxP53Data, dataset_3d
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x, y, _ = dataset_3d(JaxP53Data.synthetic()); y = y.reshape(-1)
TH = make_restarts(np.concatenate([np.full(5, 0.4), np.ones(5), np.full(5, 0.05), [2.5, 1.0]]), B)
for mode in ({"chunk": 10}, {"trace": True, "chunk": None}):
    for _ in range(3): multi_start_fit(x, y, TH, 1e-4, num_iters=150, **mode)
    wall, dev = [], []
    for _ in range(21):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = multi_start_fit(x, y, TH, 1e-4, num_iters=150, **mode)
        wall.append(1e3 * (time.perf_counter() - t0)); dev.append(r.device_ms)
    os.environ["LFM_MSF_TIMING"] = "1"
    ph = []
    for _ in range(7):
        multi_start_fit(x, y, TH, 1e-4, num_iters=150, **mode); ph.append(batched.LAST_TIMING)
    os.environ.pop("LFM_MSF_TIMING")
    print(json.dumps({"B": B, "mode": str(mode), "wall_ms_median": round(float(np.median(wall)), 3), "wall_min": round(min(wall), 3),
                      "device_ms_median": round(float(np.median(dev)), 3), "phases_ms_median": np.median(np.array(ph), axis=0).round(3).tolist()}), flush=True)
if len(sys.argv) > 2:
    pr = cProfile.Profile(); pr.enable()
    for _ in range(50): multi_start_fit(x, y, TH, 1e-4, num_iters=150, trace=True, chunk=None)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
