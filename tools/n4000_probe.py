"""Device-side timing of one N=4000 NLML+grad evaluation and of potrf / potrf+potri at 4096."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dis_project_b200 import ops, _lib

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

dev = torch.device("cuda:0")
l = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
for n in (2048, 4096):
    g = torch.Generator(device=dev); g.manual_seed(0)
    A = torch.randn(n, 256, dtype=torch.float64, device=dev, generator=g)
    S0 = A @ A.T; S0.diagonal().add_(float(n)); del A
    S = S0.clone(); W = torch.zeros_like(S); Sinv = torch.zeros_like(S); info = torch.zeros(1, dtype=torch.int32, device=dev)
    tc = timeit(lambda: S.copy_(S0))
    def potrf():
        S.copy_(S0); _lib.check(l.lfm_debug_potrf_potri(st, n, S.data_ptr(), W.data_ptr(), None, info.data_ptr()), "potrf")
    def potri():
        S.copy_(S0); _lib.check(l.lfm_debug_potrf_potri(st, n, S.data_ptr(), W.data_ptr(), Sinv.data_ptr(), info.data_ptr()), "potri")
    t1 = timeit(potrf) - tc; t2 = timeit(potri) - tc
    Lref = torch.linalg.cholesky(S0)
    potrf(); torch.cuda.synchronize()
    err = float((torch.tril(S) - Lref).abs().max() / Lref.abs().max())
    print(json.dumps({"n": n, "potrf_ms": t1, "potrf_tflops": n**3 / 3 / t1 / 1e9, "potrf_potri_ms": t2,
                      "potri_only_ms": t2 - t1, "total_tflops": n**3 / t2 / 1e9, "L_err": err, "info": int(info.item())}), flush=True)
G, T = 50, 80
times = np.linspace(0, 12, T)
X = np.stack((np.tile(times, G), np.repeat(np.arange(G), T).astype(np.float64), np.ones(G * T)), axis=-1)
y = np.random.default_rng(1).standard_normal(G * T)
th = np.concatenate([np.full(G, 0.4), np.full(G, 1.0), np.full(G, 0.05), [2.5, 1.0]])
Xd, yd, thd = (torch.as_tensor(a).to(dev) for a in (X, y, th))
t = timeit(lambda: ops.nlml_grad(Xd, yd, thd, 1e-4, G))
print(json.dumps({"nlml_grad_N4000_ms": t, "evals_per_s": 1e3 / t}), flush=True)
