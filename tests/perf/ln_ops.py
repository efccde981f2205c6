"""Device timings of the fused LayerNorm+ReLU kernels at the field MLP's shape ([65536, 256] fp32), not a pytest file.
   python tests/perf/ln_ops.py [M] [N]   (also the target of the ncu --set full capture in profiles/)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nsvf_b200 import _lib
L = _lib.load(); p = _lib.ptr
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
h = torch.randn(M, N, device=dev); dy = torch.randn(M, N, device=dev)
g = 1 + 0.1 * torch.randn(N, device=dev); bt = 0.1 * torch.randn(N, device=dev)
y = torch.empty_like(h); dh = torch.empty_like(h)
mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
sums = torch.empty(3, N, device=dev)
wsb = L.nsvf_ln_relu_bwd_workspace_bytes(M, N); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream


def t(fn, n=10):
    ms = []
    for i in range(n + 2):
        flush.zero_()                                   # cold L2, as inside a training step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2:
            ms.append(e0.elapsed_time(e1))
    return sum(ms) / len(ms)


f = lambda: L.nsvf_ln_relu_fwd(st, M, N, p(h), p(g), p(bt), 1e-5, p(y), p(mean), p(rstd))
b = lambda: L.nsvf_ln_relu_bwd(st, M, N, p(h), p(dy), p(g), p(bt), p(mean), p(rstd), p(dh), p(sums[0]), p(sums[1]),
                               p(sums[2]), p(ws), wsb)
ms = t(f); print("ln_relu_fwd [%d,%d]: %.4f ms  %.0f GB/s (8 B/element)" % (M, N, ms, (8 * M * N + 8 * M) / ms / 1e6))
ms = t(b); print("ln_relu_bwd [%d,%d] (+partial sum): %.4f ms  %.0f GB/s (12 B/element)" % (M, N, ms, (12 * M * N + 8 * M + wsb) / ms / 1e6))
