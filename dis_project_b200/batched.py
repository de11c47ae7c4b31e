"""Batched multi-start fits sharded over GPUs (BASELINE config 4).

``B`` independent ``JaxTrainer.fit`` loops (reference ``src/trainer.py:162-228``) that share (X, y)
and differ in their start point.  Restart b lives on rank ``b * world // B`` (contiguous shards); no
data-path collective is needed because the restarts are independent.  One small all-reduce per
optimiser chunk keeps every rank informed of the global best objective (MIN over a packed
(loss, restart-id) key) and of a few SUM statistics; it runs on a side stream and is consumed one
chunk late so that it never sits on the critical path (NCCL over NVLink on the GPU box, gloo in the
CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np


def make_restarts(theta_init: np.ndarray, B: int, scale: float = 0.5, seed: int = 42,
                  unconstrain=None, constrain=None) -> np.ndarray:
    """Start points: restart 0 is `theta_init`, restart b > 0 perturbs the UNCONSTRAINED leaves by
    scale * N(0, 1) drawn from default_rng(seed + b)."""
    from .model import L_HIGH, L_LOW, _softplus, _softplus_inv

    theta_init = np.asarray(theta_init, dtype=np.float64)
    P = theta_init.shape[0]
    G = (P - 2) // 3
    u0 = _softplus_inv(theta_init)
    r = (theta_init[3 * G] - L_LOW) / (L_HIGH - L_LOW)
    u0[3 * G] = np.log(r) - np.log1p(-r)
    U = np.empty((B, P))
    for b in range(B):
        U[b] = u0 if b == 0 else u0 + scale * np.random.default_rng(seed + b).standard_normal(P)
    TH = _softplus(U)
    TH[:, 3 * G] = L_LOW + (L_HIGH - L_LOW) * 0.5 * (1.0 + np.tanh(0.5 * U[:, 3 * G]))
    return TH


LAST_TIMING = None
_SIDE = {}


def _side_stream(device):
    """One side stream per device for the best-objective all-reduces (created once, not per fit)."""
    import torch

    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def shard_bounds(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous restart range [lo, hi) of `rank`; sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_best(loss: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """(loss, id) pairs -> [best_loss, id_of_best] for a MIN all-reduce; NaN losses never win."""
    loss = np.where(np.isfinite(loss), loss, np.inf)
    if loss.size == 0:
        return np.array([np.inf, -1.0])
    k = int(np.argmin(loss))
    return np.array([loss[k], float(ids[k])])


def reduce_best(local: np.ndarray, dist=None):
    """All-reduce of the packed best: MIN on the loss, then MIN on the id among ranks that hold it."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([local[0]], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    best = float(t.item())
    cand = local[1] if local[0] == best else float("inf")
    t2 = torch.tensor([cand], dtype=torch.float64, device=dev)
    dist.all_reduce(t2, op=dist.ReduceOp.MIN)
    return np.array([best, float(t2.item())])


def reduce_best_gathered(allp: np.ndarray) -> np.ndarray:
    """Rows [loss, id, ...] gathered from every rank -> [best_loss, id_of_best]: MIN on the loss, then the
    smallest id among the rows that hold it (the same rule as `reduce_best`); NaN / inf never win."""
    loss = np.where(np.isfinite(allp[:, 0]), allp[:, 0], np.inf)
    best = float(loss.min()) if loss.size else float("inf")
    if not np.isfinite(best):
        return np.array([np.inf, -1.0])
    return np.array([best, float(allp[loss == best, 1].min())])


@dataclass
class MultiStartResult:
    theta: np.ndarray        # (B_local, P) constrained results of this rank's shard
    history: np.ndarray      # (B_local, steps)
    info: np.ndarray         # (B_local,)
    lo: int
    hi: int
    best_loss: float         # global
    best_id: int             # global restart id
    best_theta: Optional[np.ndarray]  # global winner, broadcast to every rank
    best_trace: np.ndarray   # global best objective after every chunk


def multi_start_fit(X, y, theta0_all, jitter: float, *, num_iters: int = 150, lr: float = 0.01,
                    fix_params: bool = True, num_steps_per_epoch: int = 1000, chunk: int = 1,
                    b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> MultiStartResult:
    """Fit all restarts of this rank's shard on the current CUDA device.

    `theta0_all` is the (B, P) array of constrained start points of the WHOLE job (every rank passes
    the same array); the function slices its own shard.  `y` is (N,) -- B restarts of one data set -- or
    (B, N): LFM b fits y[b] on the shared design X (replicas, gene subsets of equal size, candidate transcription
    factors); the "winner" is then the LFM with the smallest final NLML over all data sets.  `chunk` optimiser steps run per kernel
    launch; after every chunk the global best objective is all-reduced asynchronously.
    """
    import torch
    import torch.distributed as dist

    from . import ops

    import os
    import time
    timing = os.environ.get("LFM_MSF_TIMING") == "1"   # debug: synchronising phase timers (tools/batched_probe3.py)
    tmarks = [time.perf_counter()]

    def mark():
        if timing:
            torch.cuda.synchronize()
            tmarks.append(time.perf_counter())

    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    theta0_all = np.asarray(theta0_all, dtype=np.float64)
    B, P = theta0_all.shape
    G = (P - 2) // 3
    lo, hi = shard_bounds(B, rank, world)
    Xd = ops._rows3(X, "x")
    yh = y if isinstance(y, torch.Tensor) else np.asarray(y, dtype=np.float64)
    if yh.ndim == 2 and yh.shape[0] == B and yh.shape[1] == Xd.shape[0] and B != 1:
        yd = ops._dev(yh[lo:hi])      # one row of observations per LFM (replicas / candidate TFs): this rank's rows
    else:
        yd = ops._dev(yh).reshape(-1)  # multi-start: every LFM fits the same observations
    nchunks = max((num_iters + chunk - 1) // chunk, 1)
    # behind the state of the shard, in the same allocation: [loss, id, theta] of the shard's winner (P + 2), one
    # best-objective word per chunk, and the winners of every rank (world x (P + 2)) -- read back in ONE copy
    n_extra = (P + 2) + nchunks + world * (P + 2)
    st = ops.BatchedFitState(theta0_all[lo:hi], G, num_iters, extra_doubles=n_extra) if hi > lo else None
    if st is not None and not isinstance(X, torch.Tensor):
        st.unique_hint = ops.unique_rows(X)  # from the host copy: no device->host round trip
        st.time_grid = ops.distinct_times(X)
    extra = st.extra if st is not None else torch.empty(n_extra, dtype=torch.float64, device=Xd.device)
    packed = extra[:P + 2]
    # per chunk ONE device word: the fit kernel atomic-mins the order-preserving integer image of every loss into
    # it, the side stream MIN-all-reduces it across ranks -- no extra kernels, nothing the fit ever waits for
    keys = extra[P + 2:P + 2 + nchunks].view(torch.int64)
    keys.fill_(torch.iinfo(torch.int64).max)
    allp = extra[P + 2 + nchunks:].view(world, P + 2)
    main = torch.cuda.current_stream()
    multi = distributed and world > 1
    side = _side_stream(Xd.device) if multi else None
    done = 0
    mark()
    for c in range((num_iters + chunk - 1) // chunk):
        steps = min(chunk, num_iters - done)
        if st is not None:
            ops.batched_fit_steps(st, Xd, yd, jitter, steps, lr=lr, b1=b1, b2=b2, eps=eps, fix_params=fix_params,
                                  steps_per_epoch=num_steps_per_epoch, best_key=keys[c:c + 1])
        done += steps
        if multi:
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                dist.all_reduce(keys[c:c + 1], op=dist.ReduceOp.MIN)
    mark()
    if multi:
        main.wait_stream(side)
    # ---- global winner: one launch packs the shard's [loss, id, theta(P)], ONE all-gather, ONE device->host copy ------
    have = st is not None and num_iters > 0
    ops.batched_best(st.hist if have else None, num_iters - 1 if have else 0, st.theta if have else None, float(lo), packed)
    gathered_on_host = None
    if multi:
        if dist.get_backend() == "nccl":
            dist.all_gather_into_tensor(allp.view(-1), packed)
        else:  # gloo (CPU tests): gather on the host
            buf = [torch.empty(P + 2, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(buf, packed.cpu())
            gathered_on_host = torch.stack(buf).numpy()
    else:
        allp[0].copy_(packed)
    if st is not None:
        theta, hist, info, ex = ops.batched_to_host(st)
    else:
        theta, hist, info = np.zeros((0, P)), np.zeros((0, num_iters)), np.zeros(0, dtype=np.int32)
        ex = extra.cpu().numpy()
    allp_h = gathered_on_host if gathered_on_host is not None else ex[P + 2 + nchunks:].reshape(world, P + 2)
    keys_h = ex[P + 2:P + 2 + nchunks].view(np.int64)[:(num_iters + chunk - 1) // chunk]
    best = reduce_best_gathered(allp_h)
    best_id = int(best[1]) if np.isfinite(best[0]) else -1
    best_theta = None
    if best_id >= 0:
        owner = int(np.flatnonzero((allp_h[:, 0] == best[0]) & (allp_h[:, 1] == best[1]))[0])
        best_theta = allp_h[owner, 2:].copy()
    mark()
    if timing:
        global LAST_TIMING
        LAST_TIMING = [round(1e3 * (b - a), 3) for a, b in zip(tmarks[:-1], tmarks[1:])]
    return MultiStartResult(theta, hist, info, lo, hi, float(best[0]), best_id, best_theta,
                            ops.loss_key_to_float(keys_h))
