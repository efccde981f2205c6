import ctypes, os, sys
ctypes.CDLL('/usr/local/cuda/lib64/libcublasLt.so.12', mode=ctypes.RTLD_GLOBAL)
ctypes.CDLL('/usr/local/cuda/lib64/libcublas.so.12', mode=ctypes.RTLD_GLOBAL)
import torch
dev = torch.device("cuda:0")
torch.manual_seed(0)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
M = 65536
for (K, N) in [(416, 256), (256, 256), (256, 128)]:
    x = torch.randn(M, K, device=dev); dh = torch.randn(M, N, device=dev)
    refw = dh.double().t() @ x.double()
    fl = 2 * M * K * N / 1e9
    ms = t(lambda: dh.t() @ x)
    print("K=%d N=%d plain dW %.3f ms (%.0f TF/s)" % (K, N, ms, fl / ms))
    for S in (8, 16, 32, 64, 128, 256):
        f = lambda: torch.bmm(dh.view(S, M // S, N).transpose(1, 2), x.view(S, M // S, K)).sum(0)
        dw = f(); err = float((dw - refw).abs().max() / refw.abs().max())
        ms = t(f)
        f2 = lambda: torch.bmm(dh.view(S, M // S, N).transpose(1, 2), x.view(S, M // S, K))
        ms2 = t(f2)
        print("   S=%3d bmm+sum %.3f ms (%.0f TF/s) [bmm alone %.3f] err %.2e" % (S, ms, fl / ms, ms2, err))
