"""Quick device-side timing probe (CUDA events) for the dense kernels and the end-to-end eval."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dis_project_b200 import ops, _lib

def timeit(fn, iters=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts), float(np.median(ts))

res = {}
dev = torch.device("cuda:0")
for n in (4096, 8192):
    A = torch.randn(n, n, dtype=torch.float64, device=dev)
    B = torch.randn(n, n, dtype=torch.float64, device=dev)
    t, _ = timeit(lambda: torch.matmul(A, B.T), iters=5)
    res[f"cublas_dgemm_{n}_tflops"] = 2 * n**3 / t / 1e12
    t, _ = timeit(lambda: ops.debug_dgemm_nt(A, B), iters=5)
    res[f"lfm_dgemm_nt_{n}_tflops"] = 2 * n**3 / t / 1e12
    del A, B
print(json.dumps(res), flush=True)

sizes = [int(s) for s in (sys.argv[1] if len(sys.argv) > 1 else "4096,8192,16384").split(",")]
for n in sizes:
    g = torch.Generator(device=dev); g.manual_seed(0)
    A = torch.randn(n, 256, dtype=torch.float64, device=dev, generator=g)
    S = A @ A.T
    S.diagonal().add_(float(n))
    del A
    W = torch.zeros_like(S); Sinv = torch.zeros_like(S); info = torch.zeros(1, dtype=torch.int32, device=dev)
    S0 = S.clone()
    l = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    def potrf():
        S.copy_(S0)
        _lib.check(l.lfm_debug_potrf_potri(st, n, S.data_ptr(), W.data_ptr(), None, info.data_ptr()), "potrf")
    def potri():
        S.copy_(S0)
        _lib.check(l.lfm_debug_potrf_potri(st, n, S.data_ptr(), W.data_ptr(), Sinv.data_ptr(), info.data_ptr()), "potri")
    def copy_only():
        S.copy_(S0)
    tc, _ = timeit(copy_only, iters=3)
    t1, _ = timeit(potrf, iters=3)
    t2, _ = timeit(potri, iters=3)
    res[f"potrf_{n}_s"] = t1 - tc
    res[f"potrf_{n}_tflops"] = n**3 / 3 / (t1 - tc) / 1e12
    res[f"potrf_potri_{n}_s"] = t2 - tc
    res[f"potrf_potri_{n}_tflops"] = n**3 / (t2 - tc) / 1e12
    def torch_chol():
        torch.linalg.cholesky(S0)
    t3, _ = timeit(torch_chol, iters=2)
    res[f"cusolver_potrf_{n}_tflops"] = n**3 / 3 / t3 / 1e12
    print(json.dumps({k: v for k, v in res.items() if f"_{n}_" in k}), flush=True)
    del S, S0, W, Sinv
    torch.cuda.empty_cache()

# end-to-end nlml+grad
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lfm_oracle as o
for (G, T) in ((50, 80),) + (((256, 128),) if "big" in sys.argv else ()):
    x = o.make_inputs(G, T)
    N = x.shape[0]
    rng = np.random.default_rng(42)
    y = rng.standard_normal(N)
    p = o.Params.reference_init(G)
    X = torch.as_tensor(x).cuda(); Y = torch.as_tensor(y).cuda(); TH = torch.as_tensor(p.pack()).cuda()
    def ev():
        ops.nlml_grad(X, Y, TH, 1e-4, G)
    t, tm = timeit(ev, iters=3)
    res[f"nlml_grad_N{N}_s"] = t
    res[f"nlml_grad_N{N}_dense_tflops"] = (lambda Np: Np**3 / t / 1e12)((N + 127) // 128 * 128)
    print(json.dumps({k: v for k, v in res.items() if f"N{N}" in k}), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/perf_probe.json", "w"), indent=1)
