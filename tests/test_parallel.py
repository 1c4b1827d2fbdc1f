"""Host logic of the scan-sharded data-parallel path (mopa_b200/parallel.py) on CPU: world_size-2 gloo processes.
The GPU data path has no collective besides this one gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mopa_b200 import parallel


def test_shard_scans_partitions_every_scan_once():
    for n in (0, 1, 7, 8, 16, 19):
        for world in (1, 2, 4, 8):
            owned = [parallel.shard_scans(n, r, world) for r in range(world)]
            flat = [i for o in owned for i in o]
            assert flat == list(range(n))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)  # ranks start from different weights
        net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.ReLU(), torch.nn.Linear(8, 3))
        parallel.broadcast_parameters(net)
        w0 = torch.cat([p.detach().flatten() for p in net.parameters()])
        bucket = parallel.FlatGradBucket(net.parameters())
        # two backward() calls per step (train_xmuda_mopa.py:417-418,578-579) accumulate locally, one pack + one reduce
        x = torch.full((5, 4), float(rank + 1))
        bucket.zero()
        net(x).sum().backward()
        net(2 * x).sum().backward()
        bucket.pack()
        local = bucket.flat.clone()
        bucket.all_reduce()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        ws = [torch.zeros_like(w0) for _ in range(world)]
        dist.all_gather(ws, w0)
        ok_params = all(torch.equal(w, ws[0]) for w in ws)
        ok_mean = torch.allclose(bucket.flat, sum(gathered) / world, rtol=1e-6, atol=1e-7)
        ok_alias = all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in net.parameters())
        out.put((rank, ok_params, ok_mean, ok_alias, float(local.abs().sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_flat_bucket_all_reduce_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    for _, ok_params, ok_mean, ok_alias, mag in res:
        assert ok_params and ok_mean and ok_alias and mag > 0
    assert res[0][4] != res[1][4]  # ranks really had different scans
