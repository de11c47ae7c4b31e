"""Does replaying one N=4000 NLML+grad evaluation as a CUDA graph beat stream launches?"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dis_project_b200 import ops
G, T = 50, 80
times = np.linspace(0, 12, T)
X = np.stack((np.tile(times, G), np.repeat(np.arange(G), T).astype(np.float64), np.ones(G * T)), axis=-1)
y = np.random.default_rng(1).standard_normal(G * T)
th = np.concatenate([np.full(G, 0.4), np.full(G, 1.0), np.full(G, 0.05), [2.5, 1.0]])
Xd, yd, thd = (torch.as_tensor(a).cuda() for a in (X, y, th))
tg = ops.distinct_times(X)
def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
t_stream = timeit(lambda: ops.nlml_grad(Xd, yd, thd, 1e-4, G, time_grid=tg))
ref, _ = ops.nlml_grad(Xd, yd, thd, 1e-4, G, time_grid=tg)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): ops.nlml_grad(Xd, yd, thd, 1e-4, G, time_grid=tg)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out, info = ops.nlml_grad(Xd, yd, thd, 1e-4, G, time_grid=tg)
t_graph = timeit(lambda: g.replay())
torch.cuda.synchronize()
print(json.dumps({"stream_ms": t_stream, "graph_ms": t_graph, "same": bool(torch.equal(out, ref)), "info": int(info)}))
