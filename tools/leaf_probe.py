import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dis_project_b200 import _lib
l=_lib.lib(); dev=torch.device("cuda:0")
A0=torch.randn(128,128,dtype=torch.float64,device=dev); A0=A0@A0.T+128*torch.eye(128,dtype=torch.float64,device=dev)
W=torch.zeros(128,128,dtype=torch.float64,device=dev); info=torch.zeros(1,dtype=torch.int32,device=dev)
st=torch.zeros(32,dtype=torch.int64,device=dev)
for it in range(3):
    A=A0.clone()
    st.zero_()
    l.lfm_debug_leaf_profile(torch.cuda.current_stream().cuda_stream, A.data_ptr(), W.data_ptr(), info.data_ptr(), st.data_ptr())
    torch.cuda.synchronize()
    s=st.cpu().numpy()[:11]
    print("cycles:", np.diff(s), "total", s[10]-s[0], "| D0: load, k0-7, k8-15, k16-23, k24-31, store:", np.diff(st.cpu().numpy()[16:23]))
names=["load(0,0)","win0(D0)","PU0","win1(D1)","PU1","win2(D2)","PU2","win3(D3)","tailF","storeW3"]
print(names)
