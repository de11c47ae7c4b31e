"""cuBLAS Dgemm (through torch.addmm / torch.matmul) on the SHAPES of the blocked factorisation -- context for the roofline
fractions of the short-K launches: what the vendor library reaches at m x m x K with K = 128 / 256 / 512 (C -= P P^T as a full
square product: twice the flops of the lower-triangle launch, so TF/s are comparable, microseconds are not)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
dev = torch.device("cuda:0")
n = 4096
A = torch.randn(n, n, dtype=torch.float64, device=dev)
def t(fn, reps=6):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))
for K in (128, 256, 512, 1024):
    for m in (1024, 2048, 3072, 3840):
        if m + K > n: continue
        P = A[n - m:, :K]
        C = A[n - m:, K:K + m]
        ms = t(lambda: torch.addmm(C, P, P.T, beta=1.0, alpha=-1.0, out=C))
        print(json.dumps({"shape": "C(m x m) -= P P^T", "m": m, "K": K, "us": ms * 1e3, "tflops": 2.0 * m * m * K / ms / 1e9}), flush=True)
for nn in (2048, 4096, 8192):
    X = torch.randn(nn, nn, dtype=torch.float64, device=dev); Y = torch.randn(nn, nn, dtype=torch.float64, device=dev)
    ms = t(lambda: torch.matmul(X, Y.T))
    print(json.dumps({"shape": "n^3", "n": nn, "us": ms * 1e3, "tflops": 2.0 * nn ** 3 / ms / 1e9}), flush=True)
