import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mopa_b200 import synth, _lib
from mopa_b200.unet_scn import UNetSCN
net = UNetSCN(1).cuda()
for bs, naz in [(8, None), (1, None), (8, 300)]:
    c, f = synth.make_batch(bs, 'nuscenes', 0, n_azimuth=naz)
    c = torch.from_numpy(c).cuda(); f = torch.from_numpy(f).cuda()
    for _ in range(5):
        net([c, f]).sum().backward()
    torch.cuda.synchronize()
    K = 20
    t0 = time.time()
    tf = 0
    for _ in range(K):
        a = time.time(); out = net([c, f]); tf += time.time() - a
        out.sum().backward()
    t1 = time.time(); torch.cuda.synchronize(); t2 = time.time()
    print('batch', bs, 'pts', c.shape[0], 'submit ms/step %.2f (fwd submit %.2f) total ms/step %.2f' % ((t1-t0)/K*1e3, tf/K*1e3, (t2-t0)/K*1e3))
