import ctypes, os, sys
if os.environ.get("PRELOAD", "1") == "1":
    ctypes.CDLL('/usr/local/cuda/lib64/libcublasLt.so.12', mode=ctypes.RTLD_GLOBAL)
    ctypes.CDLL('/usr/local/cuda/lib64/libcublas.so.12', mode=ctypes.RTLD_GLOBAL)
import torch
dev = torch.device("cuda:0")
print("env", {k: v for k, v in os.environ.items() if k.startswith("CUBLAS") or k == "PRELOAD"}, "cublas", torch.backends.cuda.preferred_blas_library())
torch.manual_seed(0)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
M = 65536
for (K, N) in [(416, 256), (256, 256), (280, 256), (256, 128)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * (2.0 / K) ** 0.5; b = torch.randn(N, device=dev)
    dh = torch.randn(M, N, device=dev)
    y = torch.addmm(b, x, w.t()); ref = torch.addmm(b.double(), x.double(), w.double().t())
    err_f = float((y - ref).abs().max() / ref.abs().max())
    dw = dh.t() @ x; refw = dh.double().t() @ x.double()
    err_w = float((dw - refw).abs().max() / refw.abs().max())
    dx = dh @ w; refx = dh.double() @ w.double()
    err_x = float((dx - refx).abs().max() / refx.abs().max())
    ms_f = t(lambda: torch.addmm(b, x, w.t())); ms_w = t(lambda: dh.t() @ x); ms_x = t(lambda: dh @ w)
    fl = 2 * M * K * N / 1e9
    print("K=%d N=%d  fwd %.3f ms (%.0f TF/s) err %.2e | dW %.3f ms (%.0f TF/s) err %.2e | dx %.3f ms (%.0f TF/s) err %.2e" %
          (K, N, ms_f, fl / ms_f, err_f, ms_w, fl / ms_w, err_w, ms_x, fl / ms_x, err_x))
