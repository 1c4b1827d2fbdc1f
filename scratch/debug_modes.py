import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from tests.helpers import small_batch
import mopa_b200.scn as scn
from mopa_b200.unet_scn import UNetSCN
coords, feats = small_batch(2, 200, 9)
net = UNetSCN(1).cuda()
res = {}
for mode in ("0", "1", "0", "1"):
    os.environ["MOPA_SCN_EAGER"] = mode
    net.zero_grad(set_to_none=True)
    f = torch.from_numpy(feats).cuda().requires_grad_(True)
    out = net([torch.from_numpy(coords), f])
    out.square().sum().backward()
    cur = (out.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters()}, f.grad.clone())
    if mode in res:
        prev = res[mode]
        print('mode', mode, 'repeatable:', torch.equal(prev[0], cur[0]), torch.equal(prev[2], cur[2]), all(torch.equal(prev[1][k], cur[1][k]) for k in cur[1]))
    res[mode] = cur
a, b = res["0"], res["1"]
print('out equal', torch.equal(a[0], b[0]), 'fgrad maxdiff', float((a[2]-b[2]).abs().max()), float(a[2].abs().max()))
for k in a[1]:
    d = float((a[1][k]-b[1][k]).abs().max()); m = float(b[1][k].abs().max())
    if d > 0: print('%-60s diff %.3e max %.3e' % (k, d, m))
