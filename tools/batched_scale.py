"""torchrun probe: BASELINE config 4 (4096 p53 restarts x 150 Adam steps) sharded over WORLD_SIZE GPUs -- wall clock of
multi_start_fit (barrier before, max over ranks), its phases, and the team size the shard runs with."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
from dis_project_b200 import ops, batched, _lib
from dis_project_b200.batched import make_restarts, multi_start_fit
from dis_project_b200.dataset import JaxP53Data, dataset_3d
x, y, _ = dataset_3d(JaxP53Data.synthetic()); y = y.reshape(-1)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
TH = make_restarts(np.concatenate([np.full(5, 0.4), np.ones(5), np.full(5, 0.05), [2.5, 1.0]]), B)
def once(timing):
    os.environ["LFM_MSF_TIMING"] = "1" if timing else "0"
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = multi_start_fit(x, y, TH, 1e-4, num_iters=150, chunk=10)
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item(), r
for _ in range(2): once(False)
walls = []
for _ in range(7):
    w, r = once(False); walls.append(w)
_, r = once(True)
if rank == 0:
    team = int(_lib.lib().lfm_batched_team_size(r.hi - r.lo, x.shape[0], 5, ops.unique_rows(x), ops.distinct_times(x)))
    print(json.dumps({"world": world, "restarts": B, "restarts_per_gpu": r.hi - r.lo, "warps_per_lfm": team,
                      "wall_ms_min": round(1e3 * min(walls), 3), "wall_ms_median": round(1e3 * float(np.median(walls)), 3),
                      "restarts_per_s": round(B / float(np.median(walls)), 1), "best_nlml": r.best_loss, "best_id": r.best_id,
                      "phases_ms_setup_loop_tail(sync timers)": batched.LAST_TIMING}), flush=True)
dist.destroy_process_group()
