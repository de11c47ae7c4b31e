"""torchrun probe: phase timing of multi_start_fit (setup / chunk loop / tail) at 512 restarts per rank."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
from dis_project_b200 import ops, batched
from dis_project_b200.batched import make_restarts, multi_start_fit
from dis_project_b200.dataset import JaxP53Data, dataset_3d
x, y, _ = dataset_3d(JaxP53Data.synthetic()); y = y.reshape(-1)
B = 512 * world
TH = make_restarts(np.concatenate([np.full(5, 0.4), np.ones(5), np.full(5, 0.05), [2.5, 1.0]]), B)
os.environ["LFM_MSF_TIMING"] = "1"
for it in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = multi_start_fit(x, y, TH, 1e-4, num_iters=150, chunk=10)
    torch.cuda.synchronize()
    if rank == 0: print(f"world {world} B {B}: total {1e3*(time.perf_counter()-t0):.2f} ms; phases {getattr(batched, 'LAST_TIMING', None)}", flush=True)
dist.destroy_process_group()
