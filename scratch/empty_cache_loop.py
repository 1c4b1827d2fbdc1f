"""MoPA's loop shape: blocking step + torch.cuda.empty_cache() every iteration (train_xmuda_mopa.py:593)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mopa_b200.scn as scn
from mopa_b200 import synth
from mopa_b200.unet_scn import UNetSCN
scn.set_precision("tf32")
torch.manual_seed(0)
net = UNetSCN(1).cuda()
host = [synth.make_batch(8, "nuscenes", s) for s in range(4)]
pinned = [(torch.from_numpy(c).pin_memory(), torch.from_numpy(f).pin_memory()) for c, f in host]
ts = []
for i in range(40):
    t0 = time.perf_counter()
    c, f = pinned[i % 4]
    net.zero_grad(set_to_none=True)
    out = net([c, f.cuda(non_blocking=True)])
    loss = out.sum()
    loss.backward()
    v = float(loss.detach())
    torch.cuda.empty_cache()
    ts.append(1e3 * (time.perf_counter() - t0))
print("arena pool %s: blocking step + empty_cache, median %.3f ms, p90 %.3f ms (40 steps, first 10 dropped)" % (
    os.environ.get("MOPA_SCN_ARENA_POOL", "1"), np.median(ts[10:]), np.percentile(ts[10:], 90)))
