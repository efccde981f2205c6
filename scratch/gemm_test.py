import sys; sys.path.insert(0, '.')
from nsvf_b200 import blas
blas.use_system_cublas()
import torch
from nsvf_b200 import _lib
L = _lib.load(); p = _lib.ptr
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
def rel(a, b): return float((a.double() - b).abs().max() / b.abs().max())
print("available", L.nsvf_linear_available())
for (M, K, N) in [(65536, 416, 256), (65536, 256, 256), (65536, 280, 256), (65536, 256, 128), (40003, 416, 256), (100, 256, 256)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * (2.0 / K) ** 0.5; dh = torch.randn(M, N, device=dev)
    S = 16
    wsb = L.nsvf_linear_workspace_bytes(M, N, K, S); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    h = torch.empty(M, N, device=dev); dx = torch.empty(M, K, device=dev); dw = torch.empty(N, K, device=dev)
    f = lambda: _lib.check(L.nsvf_linear_fwd(st, M, N, K, p(x), p(w), p(h), p(ws), wsb))
    bi = lambda: _lib.check(L.nsvf_linear_bwd_input(st, M, N, K, p(dh), p(w), p(dx), p(ws), wsb))
    bw = lambda: _lib.check(L.nsvf_linear_bwd_weight(st, M, N, K, p(dh), p(x), p(dw), S, p(ws), wsb))
    try:
        f(); bi(); bw(); torch.cuda.synchronize()
    except Exception as e:
        print(M, K, N, "FAILED", e); continue
    e1, e2, e3 = rel(h, x.double() @ w.double().t()), rel(dx, dh.double() @ w.double()), rel(dw, dh.double().t() @ x.double())
    fl = 2 * M * K * N / 1e9
    m1, m2, m3 = t(f), t(bi), t(bw)
    c1, c2 = t(lambda: x @ w.t()), t(lambda: dh @ w)
    print("M=%d K=%d N=%d  fwd %.3f ms (%.0f TF/s, cublas %.3f) err %.1e | dx %.3f ms (cublas %.3f) err %.1e | dW %.3f ms (%.0f TF/s) err %.1e" %
          (M, K, N, m1, fl / m1, c1, e1, m2, c2, e2, m3, fl / m3, e3))
