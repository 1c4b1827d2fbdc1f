"""A/B of the tcgen05 d_weight kernel against the mma.sync one on one submanifold layer: python scratch/dw_debug.py CIN COUT"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mopa_b200.scn as scn
from tests.helpers import random_cloud
torch.set_printoptions(linewidth=200, precision=4, sci_mode=False)
scn.set_precision("tf32")
for cin, cout in [(int(a), int(b)) for a, b in zip(sys.argv[1::2], sys.argv[2::2])]:
    coords = random_cloud(3000, 14, cin * 131 + cout, n_batch=2, dup_frac=0.1)
    for pattern in ("ramp", "randn"):
        n = coords.shape[0]
        if pattern == "ramp":
            feats = torch.arange(cin, dtype=torch.float32).repeat(n, 1) + 1.0      # in[i][c] = c + 1
        else:
            feats = torch.randn(n, cin, generator=torch.Generator().manual_seed(1))
        res = {}
        for mode in ("1", "0"):
            os.environ["MOPA_SCN_NO_DWTC"] = mode
            x = scn.InputLayer(3, 16, mode=4)([torch.from_numpy(coords), feats.cuda().requires_grad_(True)])
            conv = scn.SubmanifoldConvolution(3, cin, cout, 3, False).cuda()
            torch.manual_seed(0)
            with torch.no_grad():
                conv.weight.normal_()
            y = conv(x)
            V = y.features.shape[0]
            if pattern == "ramp":
                g = (torch.arange(cout, dtype=torch.float32).repeat(V, 1) * 100 + 100).cuda()   # dout[o][c] = 100 (c + 1)
            else:
                g = torch.randn(V, cout, generator=torch.Generator().manual_seed(2)).cuda()
            y.features.backward(g)
            torch.cuda.synchronize()
            res[mode] = conv.weight.grad.detach().cpu().reshape(27, cin, cout)
        ref, tc = res["1"], res["0"]
        err = float((tc - ref).abs().max() / ref.abs().max())
        print("== %d->%d %s: V=%d  max|ref| %.4g  max|tc| %.4g  rel err %.3g  nan %d zero-frac %.3f" % (
            cin, cout, pattern, V, ref.abs().max(), tc.abs().max(), err, int(torch.isnan(tc).sum()), float((tc == 0).float().mean())))
        if err > 5e-3:
            for k in (13, 0):
                print("k=%d ref[:6,:6]\n" % k, ref[k, :6, :6], "\n tc[:6,:6]\n", tc[k, :6, :6])
                r = tc[k] / ref[k].clamp_min(1e-20)
                print(" ratio tc/ref [:6,:6]\n", r[:6, :6])
