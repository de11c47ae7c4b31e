"""Three plain C = A B^T products through the library's DMMA kernel at n = 8192 (16-warp 128 x 128 tiles,
lfm_dgemm_kernel<0,1,4,4,4>: the variant that carries 94 % of a config-3 evaluation): the `ncu --set full` target."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dis_project_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
A = torch.randn(n, n, dtype=torch.float64, device="cuda"); B = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    C = ops.debug_dgemm_nt(A, B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); C = ops.debug_dgemm_nt(A, B); e1.record(); torch.cuda.synchronize()
print("TF/s", 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
