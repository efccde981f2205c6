"""Accuracy and speed of the fp32 GEMMs of the field MLP under cuBLAS 12.9's BF16x9 emulation (nsvf_b200/blas.py),
measured against float64 products.  Stand-alone (the cuBLAS choice precedes `import torch`); prints one JSON line.

  python tests/perf/cublas_emulation_check.py          # emulation on (default)
  python tests/perf/cublas_emulation_check.py simt     # torch's bundled cuBLAS, plain SGEMM
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nsvf_b200 import blas
if "simt" not in sys.argv[1:]:
    blas.use_system_cublas()
import torch
import torch.nn.functional as F
from nsvf_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())


out = {"mode": blas.mode(), "emulated": blas.emulated(), "gemm": []}
M = 65536
for (K, N) in [(416, 256), (256, 256), (280, 256), (256, 128)]:
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * (2.0 / K) ** 0.5
    b = torch.randn(N, device=dev)
    dh = torch.randn(M, N, device=dev)
    fl = 2 * M * K * N / 1e9
    r = {"K": K, "N": N}
    r["fwd_err"] = rel(torch.addmm(b, x, w.t()), torch.addmm(b.double(), x.double(), w.double().t()))
    r["dx_err"] = rel(dh @ w, dh.double() @ w.double())
    r["dw_err"] = rel(ops._weight_grad(dh, x), dh.double().t() @ x.double())
    for name, fn in (("fwd", lambda: torch.addmm(b, x, w.t())), ("dx", lambda: dh @ w),
                     ("dw", lambda: ops._weight_grad(dh, x))):
        ms = t(fn)
        r[name + "_ms"] = round(ms, 4)
        r[name + "_tflops"] = round(fl / ms, 1)
    out["gemm"].append(r)

# the fused FCLayer (cuBLAS + csrc/field_norm.cu) against float64 autograd, as tests/test_field_gpu.py does
M, I, N = 70001, 416, 256
x = torch.randn(M, I, device=dev)
w = torch.randn(N, I, device=dev) * (2.0 / I) ** 0.5
b = torch.randn(N, device=dev) * 0.1
g = 1 + 0.2 * torch.randn(N, device=dev)
bt = 0.1 * torch.randn(N, device=dev)
dy = torch.randn(M, N, device=dev)
ins = [v.clone().requires_grad_(True) for v in (x, w, b, g, bt)]
y = ops.linear_layernorm_relu(*ins, 1e-5)
grads = torch.autograd.grad(y, ins, dy)
ins64 = [v.double().requires_grad_(True) for v in (x, w, b, g, bt)]
pre64 = F.layer_norm(F.linear(ins64[0], ins64[1], ins64[2]), (N,), ins64[3], ins64[4], 1e-5)
ref = torch.autograd.grad(pre64 * (y > 0), ins64, dy.double())
out["fc_layer"] = {"y_err": rel(y, torch.relu(pre64))}
for name, a, r_ in zip(("dx", "dW", "db", "dgamma", "dbeta"), grads, ref):
    out["fc_layer"][name + "_err"] = rel(a, r_)
print(json.dumps(out))
