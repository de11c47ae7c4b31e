#!/usr/bin/env python
"""Full-size parity of BASELINE configs 3 and 5 against the CPU oracle.

    python tools/fullsize_parity.py [--out gpurun_out/fullsize_parity.json] [--tstar-sample 256] [--skip-c3]

Config 3: N = 32768 (256 genes x 128 time points) NLML + gradient: ``ops.nlml_grad`` on the B200 against
``oracle.nlml_and_grad`` (row-chunked numpy/scipy/LAPACK, a few minutes on the box's host cores).
Config 5: latent posterior at 102 400 test times from the same LFM: the GPU computes ALL of them; the oracle
computes an evenly spaced sample of them (default 256) from its own Cholesky of the N = 32768 covariance.
Tolerance of north_star: 1e-9 relative.  Writes the relative errors and timings as JSON (committed under
profiles/ by the round that ran it).  Uses oracle/ as the checker only; nothing here is a product path.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fullsize_parity.json"))
    ap.add_argument("--tstar-sample", type=int, default=256)
    ap.add_argument("--genes", type=int, default=256)
    ap.add_argument("--times", type=int, default=128)
    ap.add_argument("--tstar", type=int, default=102400)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--skip-c3", action="store_true")
    args = ap.parse_args()

    import torch
    from bench import _Inputs, JITTER
    from dis_project_b200 import ops
    from oracle import lfm_oracle as o
    from scipy.linalg import cho_solve, solve_triangular
    from scipy.linalg.lapack import dpotrf

    G, T, TS = args.genes, args.times, args.tstar
    cores = os.cpu_count() or 1
    threads = args.threads or min(cores, 24)
    dev = torch.device("cuda", 0)
    Xh, yh, thh = _Inputs.make_problem(G, T)
    N = Xh.shape[0]
    p = o.Params.unpack(thh, JITTER)
    res = {"N": N, "G": G, "T": T, "Tstar": TS, "host_cores": cores, "oracle_threads": threads, "tolerance": 1e-9,
           "inputs": "bench.py _Inputs.make_problem (seed 42), theta = reference initial state (model.py:100-114)"}
    X, y, th = (torch.as_tensor(a).to(dev) for a in (Xh, yh, thh))

    if not args.skip_c3:
        ops.nlml_grad(X, y, th, JITTER, G)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, info = ops.nlml_grad(X, y, th, JITTER, G)
        torch.cuda.synchronize()
        gpu_s = time.perf_counter() - t0
        out = out.cpu().numpy()
        t0 = time.perf_counter()
        v_ref, g_ref = o.nlml_and_grad(p, Xh, yh, chunk=128, threads=threads)
        cpu_s = time.perf_counter() - t0
        res["config3"] = {"gpu_nlml": float(out[0]), "oracle_nlml": v_ref, "rel_err_nlml": abs(out[0] - v_ref) / abs(v_ref),
                          "rel_err_grad_vs_max_component": rel(out[1:], g_ref), "info": int(info.item()),
                          "grad_max_component": float(np.max(np.abs(g_ref))), "gpu_seconds": gpu_s,
                          "oracle_seconds": cpu_s,
                          "pass": bool(abs(out[0] - v_ref) <= 1e-9 * abs(v_ref) and rel(out[1:], g_ref) < 1e-9)}
        print(json.dumps(res["config3"]), flush=True)
        del out
        ops.release_workspaces()
        torch.cuda.empty_cache()

    # ---- config 5 ----------------------------------------------------------------------------------------
    var_h = np.random.default_rng(7).uniform(0.01, 0.1, N)
    Xs_h = np.stack((np.linspace(0, 13, TS), np.full(TS, -1.0), np.zeros(TS)), axis=1)
    Xs, var = torch.as_tensor(Xs_h).to(dev), torch.as_tensor(var_h).to(dev)
    t0 = time.perf_counter()
    m, v, info = ops.latent_posterior(X, y, var, th, JITTER, Xs, G)
    torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t0
    m, v = m.cpu().numpy(), v.cpu().numpy()
    ops.release_workspaces()
    torch.cuda.empty_cache()
    idx = np.unique(np.linspace(0, TS - 1, args.tstar_sample).round().astype(np.int64))
    t0 = time.perf_counter()
    # o.latent_predict restricted to the sampled test times, with the Gram built in row chunks over the host threads
    # (o.gram builds N x N through broadcasting temporaries that do not fit at N = 32768) -- same arithmetic, same order
    from concurrent.futures import ThreadPoolExecutor
    S = np.empty((N, N))
    tt, gi = Xh[:, 0], Xh[:, 1].astype(np.int64)

    def build(r0):
        r1 = min(N, r0 + 128)
        S[r0:r1] = o.kernel_xx(p, tt[r0:r1, None], gi[r0:r1, None], tt[None, :], gi[None, :])

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(build, range(0, N, 128)))
    S[np.diag_indices(N)] += var_h
    S[np.diag_indices(N)] += p.jitter
    c, pinfo = dpotrf(S, lower=1, overwrite_a=1, clean=0)
    assert pinfo == 0
    z = yh - o.mean_function(p, Xh)
    alpha = cho_solve((c, True), z)
    Kxf = o.cross_covariance(p, Xh, Xs_h[idx])
    m_ref = Kxf.T @ alpha                                      # mean_t = 0 on latent rows (flag 0)
    V = solve_triangular(c, Kxf, lower=True)
    v_ref = 1.0 + p.jitter - np.sum(V * V, axis=0) + p.jitter   # model.py:456-461 (Q4)
    cpu_s = time.perf_counter() - t0
    res["config5"] = {"sampled_test_points": int(idx.size), "rel_err_mean": rel(m[idx], m_ref), "rel_err_var": rel(v[idx], v_ref),
                      "rel_err_std": rel(np.sqrt(v[idx]), np.sqrt(v_ref)), "var_min": float(v.min()), "var_max": float(v.max()),
                      "mean_abs_max": float(np.max(np.abs(m_ref))), "info": int(info.item()), "gpu_seconds_all_points": gpu_s,
                      "oracle_seconds_sample": cpu_s,
                      "pass": bool(rel(m[idx], m_ref) < 1e-9 and rel(v[idx], v_ref) < 1e-9)}
    print(json.dumps(res["config5"]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
