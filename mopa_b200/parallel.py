"""Scan-sharded data parallelism for the 3D branch (new: the reference is single-process).

One process per GPU; every rank holds a full UNetSCN replica and its own scans (a scan is never split); BatchNorm
statistics stay per rank, as in the reference's single-GPU batch of 8. The only exchange is one gradient all-reduce
(mean) per optimizer step over NCCL/NVLink: all gradients are packed into ONE flat fp32 bucket (~10.8 MB for UNetSCN)
so the collective is one call and MoPA's two backward() calls per step (train_xmuda_mopa.py:417-418,578-579) reduce once.
With bucket.attach() the compiled UNetSCN backward writes each gradient straight into its slice of the bucket (the first
backward of a step; later ones are added to it by autograd), so nothing is packed; parameters of other modules are
packed by one multi-tensor copy. (Pointing .grad at bucket slices BEFORE backward made autograd run one small add kernel
per parameter, 78 launches per step; with .grad = None autograd just keeps the tensor the backward returns.)
"""
import torch
import torch.distributed as dist


def shard_scans(n_scans, rank, world_size):
    """Indices of the scans rank `rank` owns: contiguous, sizes differ by at most one, every scan owned exactly once."""
    base, extra = divmod(n_scans, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


class FlatGradBucket:
    """One flat buffer with a slice per parameter. zero() drops the gradients (no kernel); pack() copies the gradients
    autograd produced into their slices with one multi-tensor copy and points .grad at the slices; all_reduce() packs
    and averages the buffer across ranks in place."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dtype = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.zero()

    def attach(self):
        """Let mopa_b200.scn's compiled backward write gradients straight into this bucket's slices (no pack copy): see
        scn.compiler.register_grad_views. Parameters of other modules (e.g. nn.Linear heads) still go through pack()."""
        from .scn import compiler
        compiler.register_grad_views({p: v for p, v in zip(self.params, self.views)})
        return self

    def detach(self):
        from .scn import compiler
        compiler.unregister_grad_views(self.params)

    def zero(self):
        for p in self.params:
            p.grad = None

    def pack(self):
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None:
                v.zero_()  # parameter not reached by this step's backward passes
            elif g.data_ptr() != v.data_ptr():
                src.append(g.detach())
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)
        for p, v in zip(self.params, self.views):
            p.grad = v

    def all_reduce(self, async_op=False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None  # single process: the gradients stay where autograd put them
        self.pack()
        if dist.get_backend() == "nccl":
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, async_op=async_op)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)  # gloo has no AVG
        if async_op:
            raise NotImplementedError("async all-reduce is only wired for nccl")
        self.flat.div_(dist.get_world_size())
        return work


def broadcast_parameters(module, src=0):
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)
