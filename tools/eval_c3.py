"""Two N=32768 NLML+grad evaluations (profiling target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dis_project_b200 import ops
G, T = 256, 128
times = np.linspace(0, 12, T)
X = np.stack((np.tile(times, G), np.repeat(np.arange(G), T).astype(np.float64), np.ones(G * T)), axis=-1)
y = np.random.default_rng(1).standard_normal(G * T)
th = np.concatenate([np.full(G, 0.4), np.full(G, 1.0), np.full(G, 0.05), [2.5, 1.0]])
Xd, yd, thd = (torch.as_tensor(a).cuda() for a in (X, y, th))
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out, info = ops.nlml_grad(Xd, yd, thd, 1e-4, G); e1.record()
    torch.cuda.synchronize()
print("nlml", float(out[0]), "info", int(info), "ms", e0.elapsed_time(e1))
