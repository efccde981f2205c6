import torch, numpy as np, oracle
from oracle import wrappers
from nsvf_b200 import ops
import sys; sys.path.insert(0,'.')
from tests.test_ops_gpu import _composite_inputs
cuda=torch.device('cuda:0')
for B,K in [(4096,127),(37,300)]:
    fe,tex,depth=_composite_inputs(B,K,0,cuda)
    mine=ops.composite(fe,tex,depth)
    ref=wrappers.composite_torch(fe,tex,depth)
    orc=oracle.composite_fwd(fe.cpu().numpy(),tex.cpu().numpy(),depth.cpu().numpy())
    # double reference
    fe64=fe.double(); sh=torch.cat([fe64.new_zeros(B,1),fe64[:,:-1]],-1)
    p64=(1-torch.exp(-fe64))*torch.exp(-torch.cumsum(sh,-1))
    d64=(depth.double()*p64).sum(-1)
    for a,b,o,nm in zip(mine,ref,orc,("probs","depth","missed","colors")):
        a=a.double().cpu(); b=b.double().cpu(); o=torch.from_numpy(o).double()
        sc=o.abs().max()
        print(nm,'scale',float(sc),'mine-torch',float((a-b).abs().max()/sc),'mine-oracle',float((a-o).abs().max()/sc),'torch-oracle',float((b-o).abs().max()/sc))
    print('probs vs f64: mine',float((mine[0].double()-p64).abs().max()),'torch',float((ref[0].double()-p64).abs().max()),'oracle',float((torch.from_numpy(orc[0]).double().cuda()-p64).abs().max()))
    print('depth vs f64: mine',float(((mine[1].double()-d64).abs()/d64.abs().clamp(min=1e-9)).max()),'torch',float(((ref[1].double()-d64).abs()/d64.abs().clamp(min=1e-9)).max()), 'max d64', float(d64.max()))
