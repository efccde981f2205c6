import sys; sys.path.insert(0, '.')
from nsvf_b200 import blas
blas.use_system_cublas()
import torch
dev = torch.device("cuda:0")
M, K, N, S = 65536, 416, 256, 16
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev); dh = torch.randn(M, N, device=dev)
ops = {
  "addmm": lambda: torch.addmm(b, x, w.t()),
  "mm_nn(dx)": lambda: dh @ w,
  "mm_tn(dW plain)": lambda: dh.t() @ x,
  "bmm(dW split)": lambda: torch.bmm(dh.view(S, M // S, N).transpose(1, 2), x.view(S, M // S, K)),
  "bmm+sum": lambda: torch.bmm(dh.view(S, M // S, N).transpose(1, 2), x.view(S, M // S, K)).sum(0),
  "bmm contiguous A": lambda: torch.bmm(dh.view(S, M // S, N).transpose(1, 2).contiguous(), x.view(S, M // S, K)),
}
for name, fn in ops.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(2): fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            y = fn()
        g.replay(); torch.cuda.synchronize()
        print(name, "capture OK")
    except Exception as e:
        print(name, "capture FAILED:", str(e).splitlines()[0][:150])
        torch.cuda.synchronize()
