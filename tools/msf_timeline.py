"""Device-side timeline (CUPTI through torch.profiler) of ONE multi_start_fit of B p53 restarts on one GPU: every kernel and
copy with start / duration, to see what surrounds the fit kernel."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from dis_project_b200.batched import make_restarts, multi_start_fit
from dis_project_b200.dataset import JaxP53Data, dataset_3d
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x, y, _ = dataset_3d(JaxP53Data.synthetic()); y = y.reshape(-1)
TH = make_restarts(np.concatenate([np.full(5, 0.4), np.ones(5), np.full(5, 0.05), [2.5, 1.0]]), B)
kw = dict(num_iters=150, trace=True, chunk=None)
for _ in range(3): multi_start_fit(x, y, TH, 1e-4, **kw)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        multi_start_fit(x, y, TH, 1e-4, **kw)
path = "gpurun_out/trace_msf.json"
prof.export_chrome_trace(path)
allev = json.load(open(path))["traceEvents"]
ev = [e for e in allev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
ev = ev[len(ev) // 2:]
t0 = ev[0]["ts"]
for e in ev:
    print(f'{e["ts"] - t0:9.1f} us  +{e["dur"]:8.1f}  {e["cat"]:10s} {e["name"][:70]}')
cpu = [e for e in allev if e.get("cat") in ("cuda_runtime", "cuda_driver") and e["ts"] >= t0 - 300]
for e in cpu:
    print(f'   host {e["ts"] - t0:9.1f} us  +{e["dur"]:8.1f}  {e["name"][:60]}')
os.remove(path)
