import sys, torch, numpy as np
sys.path.insert(0,'/root/repo')
from airiceraytracing_b200 import AirIceSolver
from oracle.ref import ATMOSPHERE
S=AirIceSolver(ATMOSPHERE)
dev=torch.device('cuda')
nr=2048
tr=torch.full((nr,),170.0,dtype=torch.float64,device=dev)-40.0*torch.rand(nr,device=dev,dtype=torch.float64)
hr=torch.full((nr,),20000.0,dtype=torch.float64,device=dev)
px,pz,pc=S.ray_path(tr,hr,-200.0,3000.0); mp=px.shape[1]
torch.cuda.synchronize()
for _ in range(3):
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record(); S.lib.airice_ray_path_device(S.handle,nr,tr.data_ptr(),hr.data_ptr(),-200.0,3000.0,mp,px.data_ptr(),pz.data_ptr(),pc.data_ptr(),torch.cuda.current_stream().cuda_stream); b.record(); torch.cuda.synchronize(); print(a.elapsed_time(b))
