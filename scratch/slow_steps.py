"""Which steps of a fresh process are slow (device time between step-end events, host time per iteration)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mopa_b200.scn as scn
from mopa_b200 import synth, data, parallel
from mopa_b200.unet_scn import UNetSCN
mode = sys.argv[1] if len(sys.argv) > 1 else "default"
import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
scn.set_precision("tf32")
torch.manual_seed(0)
net = UNetSCN(1).cuda()
bucket = parallel.FlatGradBucket(net.parameters()).attach()
host = [synth.make_batch(8, "nuscenes", 4 * rank + s) for s in range(4)]
dev = [(data.mark_ready(torch.from_numpy(c).cuda()), torch.from_numpy(f).cuda()) for c, f in host]
if mode == "reserve":
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    assert rt.cudaMallocAsync(ctypes.byref(p), ctypes.c_size_t(6 << 30), ctypes.c_void_p(0)) == 0
    assert rt.cudaFreeAsync(p, ctypes.c_void_p(0)) == 0
    x = torch.empty(8 << 30, dtype=torch.uint8, device="cuda"); del x
torch.cuda.synchronize()
N = 90
ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
ht = []
ev[0].record()
for i in range(N):
    t0 = time.perf_counter()
    c, f = dev[i % 4]
    bucket.zero()
    out = net([c, f])
    out.sum().backward()
    if world > 1 and mode != "noar":
        bucket.all_reduce()
    ev[i + 1].record()
    ht.append(1e3 * (time.perf_counter() - t0))
torch.cuda.synchronize()
per = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(N)])
med = np.median(per[30:])
if rank == 0:
  print(mode, "median %.3f" % med, "total first 25 after 5: %.3f ms/step" % per[5:25].mean())
  print(" slow device steps:", [(i, round(float(per[i]), 2)) for i in range(N) if per[i] > 1.3 * med])
  print(" slow host iterations:", [(i, round(ht[i], 2)) for i in range(N) if ht[i] > 2.5 * np.median(ht)], "median host %.2f" % np.median(ht))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
