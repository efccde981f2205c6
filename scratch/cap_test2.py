import sys, os; sys.path.insert(0, '.')
from nsvf_b200 import blas
if os.environ.get("EMU", "1") == "1":
    blas.use_system_cublas()
import torch, traceback
import torch.nn.functional as F
from nsvf_b200.field import RadianceField, GraphedField, _FieldCore
dev = torch.device("cuda:0")
rows = int(os.environ.get("ROWS", "65536"))
what = os.environ.get("WHAT", "field")
torch.manual_seed(0)
print("EMU", blas.emulated(), "rows", rows, "what", what)
if what == "field":
    gf = GraphedField(RadianceField().to(dev), rows=rows, slots=2).train()
    try:
        for step in range(2):
            gf.begin_step()
            emb = (torch.randn(rows, 32, device=dev) * 0.2).requires_grad_(True)
            ray = F.normalize(torch.randn(rows, 3, device=dev), dim=-1)
            o = gf({"emb": emb, "ray": ray})
            (o["sigma"].sum() + o["texture"].sum()).backward()
        torch.cuda.synchronize(); print("OK replays", gf.graph_replays)
    except Exception as e:
        print("FAILED", str(e).splitlines()[0][:200])
else:
    from nsvf_b200 import ops
    lin = torch.nn.Linear(416, 256).to(dev); ln = torch.nn.LayerNorm(256).to(dev)
    x = torch.randn(rows, 416, device=dev, requires_grad=True)
    def fn_fc(x): return ops.linear_layernorm_relu(x, lin.weight, lin.bias, ln.weight, ln.bias, 1e-5)
    freq = torch.exp(torch.arange(6, device=dev).float())
    e = torch.randn(rows, 32, device=dev, requires_grad=True)
    def fn_pe(e): return ops.posenc(e, freq, False, True)
    head = torch.nn.Linear(256, 3).to(dev)
    xh = torch.randn(rows, 256, device=dev, requires_grad=True)
    def fn_head(xh): return ops.narrow_linear(xh, head.weight, head.bias)
    class Wrap(torch.nn.Module):
        def __init__(s, f, mods): super().__init__(); s.f = f; s.mods = torch.nn.ModuleList(mods)
        def forward(s, a): return s.f(a)
    cases = []
    for (i, o) in ((416, 256), (256, 256), (280, 256), (256, 128)):
        l2 = torch.nn.Linear(i, o).to(dev); n2 = torch.nn.LayerNorm(o).to(dev)
        cases.append(("fc %d->%d" % (i, o), (lambda l2, n2: lambda a: ops.linear_layernorm_relu(a, l2.weight, l2.bias, n2.weight, n2.bias, 1e-5))(l2, n2),
                      [l2, n2], torch.randn(rows, i, device=dev, requires_grad=True)))
    head1 = torch.nn.Linear(128, 1).to(dev)
    cases.append(("head 128->1", lambda a: ops.narrow_linear(a, head1.weight, head1.bias), [head1], torch.randn(rows, 128, device=dev, requires_grad=True)))
    cases.append(("cat", lambda a: torch.cat([a, a[:, :24]], -1) * 2, [], torch.randn(rows, 256, device=dev, requires_grad=True)))
    for name, f, mods, arg in cases:
        try:
            g = torch.cuda.make_graphed_callables(Wrap(f, mods), (arg,))
            y = g(arg); y.sum().backward(); torch.cuda.synchronize(); print(name, "OK")
        except Exception as ex:
            print(name, "FAILED", str(ex).splitlines()[0][:200]); torch.cuda.synchronize()
