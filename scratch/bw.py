import torch, time
dev='cuda:0'
def t(fn,n=10):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
N=2**31  # 8 GiB of float32
a=torch.empty(N,dtype=torch.float32,device=dev); b=torch.empty(N,dtype=torch.float32,device=dev)
ms=t(lambda: a.zero_()); print('write-only fill 8GiB: %.3f ms -> %.0f GB/s'%(ms, N*4/ms/1e6))
ms=t(lambda: b.copy_(a)); print('copy 8GiB: %.3f ms -> %.0f GB/s (r+w)'%(ms, 2*N*4/ms/1e6))
ms=t(lambda: a.sum()); print('read-only sum 8GiB: %.3f ms -> %.0f GB/s'%(ms, N*4/ms/1e6))
