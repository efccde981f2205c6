import sys, os; sys.path.insert(0,'.')
import torch, time
import bench
from torch.profiler import profile, ProfilerActivity
dev=torch.device('cuda:0')
from nsvf_b200 import synthetic
rs, rd = synthetic.camera_rays(800,800,1,radius=4.5,seed=7,device=dev)
rs, rd = rs[None,:,None,0,:].contiguous(), rd[None].contiguous()
pipe, scene = bench.build_model(dev,"C3",train=False,field="trivial",tolerance=0.01,chunk=512,sigma_bias=2.0)
with torch.no_grad():
    pipe(rs,rd); torch.cuda.synchronize()
    t0=time.perf_counter(); out=pipe(rs,rd); torch.cuda.synchronize(); print('wall ms', (time.perf_counter()-t0)*1e3, 'ae', out['ae'])
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        pipe(rs,rd); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
