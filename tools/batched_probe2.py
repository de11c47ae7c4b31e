"""torchrun probe: where does multi_start_fit spend its time at N > 1 GPUs?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
from dis_project_b200 import ops, batched
from dis_project_b200.batched import make_restarts, multi_start_fit
from dis_project_b200.dataset import JaxP53Data, dataset_3d
x, y, _ = dataset_3d(JaxP53Data.synthetic()); y = y.reshape(-1)
TH = make_restarts(np.concatenate([np.full(5, 0.4), np.ones(5), np.full(5, 0.05), [2.5, 1.0]]), 4096)
def run(chunk, B=4096):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = multi_start_fit(x, y, TH[:B], 1e-4, num_iters=150, chunk=chunk)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), r
for chunk in (10, 150, 10, 50, 1):
    run(chunk)
    dt, r = run(chunk)
    if rank == 0:
        print(f"world {world} chunk {chunk}: {dt*1e3:.2f} ms  best {r.best_loss:.6f} id {r.best_id}", flush=True)
# raw all-reduce latency
t = torch.zeros(1, dtype=torch.float64, device="cuda")
for _ in range(5): dist.all_reduce(t, op=dist.ReduceOp.MIN)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(100): dist.all_reduce(t, op=dist.ReduceOp.MIN)
torch.cuda.synchronize()
if rank == 0: print("all_reduce(1 double) us:", (time.perf_counter() - t0) * 1e4, flush=True)
dist.destroy_process_group()
