import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from oracle import scn_oracle as so
from tests.helpers import small_batch, rel_err
import mopa_b200.scn as scn
from mopa_b200.unet_scn import UNetSCN
def l2(a, b):
    a = a.detach().double().cpu().flatten(); b = b.detach().double().cpu().flatten()
    return float((a-b).norm()/b.norm()), float(torch.dot(a,b)/(a.norm()*b.norm()))
for (nsc, naz, seed) in [(2, 250, 2), (4, 600, 3)]:
    coords, feats = small_batch(nsc, naz, seed)
    state = so.make_unet_state(seed=5)
    o64 = so.OracleUNetSCN(state, dtype=torch.float64); r64 = o64.forward(coords, feats)
    g = torch.randn(r64.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    r64.backward(g)
    o32 = so.OracleUNetSCN(state, dtype=torch.float32); r32 = o32.forward(coords, feats); r32.backward(g.float())
    e = [l2(o32.params[k].grad, o64.params[k].grad) for k in o64.params if 'running' not in k]
    print('N', coords.shape[0], 'oracle f32: worst relL2 %.2e min cos %.6f' % (max(x[0] for x in e), min(x[1] for x in e)))
    for prec in ['fp32', 'tf32']:
        scn.set_precision(prec)
        net = UNetSCN(1).cuda(); net.load_state_dict(state)
        out = net([torch.from_numpy(coords), torch.from_numpy(feats).cuda()])
        out.backward(g.float().cuda())
        e = [(l2(p.grad, o64.params[k].grad), rel_err(p.grad, o64.params[k].grad), k) for k, p in net.named_parameters()]
        print(' ', prec, 'fwd max-rel %.2e relL2 %.2e | grads: worst max-rel %.2e, worst relL2 %.2e, median relL2 %.2e, min cos %.5f' % (
            rel_err(out.detach(), r64.detach()), l2(out, r64)[0], max(x[1] for x in e), max(x[0][0] for x in e),
            float(np.median([x[0][0] for x in e])), min(x[0][1] for x in e)))
        allg = torch.cat([p.grad.flatten().double().cpu() for k, p in net.named_parameters()])
        allr = torch.cat([o64.params[k].grad.flatten() for k, p in net.named_parameters()])
        print('     whole-gradient relL2 %.3e cos %.6f' % l2(allg, allr))
