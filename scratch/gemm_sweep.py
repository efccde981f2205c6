import sys, os; sys.path.insert(0, '.')
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
for (M, K, N) in [(65536, 416, 256), (65536, 256, 256), (65536, 280, 256), (65536, 256, 128)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * (2.0 / K) ** 0.5; dh = torch.randn(M, N, device=dev)
    h = torch.empty(M, N, device=dev); dx = torch.empty(M, K, device=dev); dw = torch.empty(N, K, device=dev)
    c1, c2 = t(lambda: x @ w.t()), t(lambda: dh @ w)
    c3 = t(lambda: torch.bmm(dh.view(16, M // 16, N).transpose(1, 2), x.view(16, M // 16, K)).sum(0))
    print("M=%d K=%d N=%d  cuBLAS(+scans): fwd %.3f dx %.3f dW(bmm16+sum) %.3f" % (M, K, N, c1, c2, c3))
    r1, r2 = x.double() @ w.double().t(), dh.double() @ w.double()
    r3 = dh.double().t() @ x.double()
    for v in range(5):
        os.environ["NSVF_GEMM_TN"] = os.environ["NSVF_GEMM_NN"] = os.environ["NSVF_GEMM_NT"] = str(v)
        out = "   v%d:" % v
        for S in (16, 8, 32):
            wsb = L.nsvf_linear_workspace_bytes(M, N, K, S); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            f = lambda: _lib.check(L.nsvf_linear_fwd(st, M, N, K, p(x), p(w), p(h), p(ws), wsb))
            bi = lambda: _lib.check(L.nsvf_linear_bwd_input(st, M, N, K, p(dh), p(w), p(dx), p(ws), wsb))
            bw = lambda: _lib.check(L.nsvf_linear_bwd_weight(st, M, N, K, p(dh), p(x), p(dw), S, p(ws), wsb))
            try:
                if S == 16:
                    f(); bi(); torch.cuda.synchronize()
                    out += " fwd %.3f (%.0e) dx %.3f (%.0e)" % (t(f), rel(h, r1), t(bi), rel(dx, r2))
                bw(); torch.cuda.synchronize()
                out += " dW[S=%d] %.3f (%.0e)" % (S, t(bw), rel(dw, r3))
            except Exception as e:
                out += " FAILED(%s)" % str(e)[-40:]
        print(out)
