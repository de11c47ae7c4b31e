"""Kernel timeline of one N=4000 NLML+grad evaluation (CUPTI through torch.profiler): start / end of every launch per
stream, so that the dependent chain and the idle gaps of the three-stream factorisation can be read off.
Writes gpurun_out/timeline_<tag>.json (list of [name, stream, start_us, dur_us]) and prints a summary."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from dis_project_b200 import ops
tag = sys.argv[1] if len(sys.argv) > 1 else "eager"
G, T = 50, 80
times = np.linspace(0, 12, T)
X = np.stack((np.tile(times, G), np.repeat(np.arange(G), T).astype(np.float64), np.ones(G * T)), axis=-1)
y = np.random.default_rng(1).standard_normal(G * T)
th = np.concatenate([np.full(G, 0.4), np.full(G, 1.0), np.full(G, 0.05), [2.5, 1.0]])
Xd, yd, thd = (torch.as_tensor(a).cuda() for a in (X, y, th))
plan = ops.NlmlGradPlan(Xd, yd, G, 1e-4) if tag == "graph" else None
def step():
    return plan(thd) if plan is not None else ops.nlml_grad(Xd, yd, thd, 1e-4, G)
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
        torch.cuda.synchronize()
path = f"gpurun_out/trace_{tag}.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# second evaluation only
cut = len(ev) // 2
ev = ev[cut:]
t0 = ev[0]["ts"]
rows = [[e["name"][:60], e["args"].get("stream"), round(e["ts"] - t0, 2), round(e["dur"], 2), (e["args"].get("grid") or [0])[0]] for e in ev]
json.dump(rows, open(f"gpurun_out/timeline_{tag}.json", "w"))
os.remove(path)
end = max(r[2] + r[3] for r in rows)
print("launches", len(rows), "span us", round(end, 1))
