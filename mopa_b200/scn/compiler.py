"""Compiles an scn.Sequential(InputLayer, ..., OutputLayer) module tree into the op list the native executor runs
(csrc/program.cu, `mopa_scn_Program_*` in include/mopa_scn.h), and wraps it in ONE autograd Function.

The module tree stays the source of truth (parameters, buffers, train/eval flags are read from it on every call); only
the Python-level walk over ~90 modules per forward (the reference's scn.Sequential.forward, scn_unet.py:32-34) is
replaced. Trees containing a module the executor does not know (NetworkInNetwork, AddTable, ...) are not compiled and
run module by module instead. `MOPA_SCN_EAGER=1` forces the module-by-module path.
"""
import ctypes
import os
import struct
import weakref

import numpy as np
import torch
from torch.autograd import Function

from .. import _lib
from . import functional as F
from . import modules as M

OP_SUBM, OP_CONV, OP_DECONV, OP_BN = 1, 2, 3, 4



class _ArenaPool:
    """The executor's arenas (activations, gradients, scratch: 0.1-10 GB per pass) kept across steps instead of going back
    to torch's caching allocator. Two reasons: MoPA's loop calls torch.cuda.empty_cache() every iteration
    (train_xmuda_mopa.py:593), which hands every cached block back to the driver and makes the next forward cudaMalloc its
    arenas again; and the caching allocator splits a cached multi-GB block to serve a smaller request, so that the next
    request of the original size allocates afresh (seen as 50 ms - 1 s stalls with two steps in flight). A lease is keyed
    by (device, stream, size class): re-use on the same stream is ordered by the stream itself, exactly like the caching
    allocator's own re-use. At most `keep` idle tensors per key and `max_idle_bytes` in all are kept; `release()` (also
    exported as scn.release_arenas()) drops them. MOPA_SCN_ARENA_POOL=0 turns the pool off."""

    def __init__(self, keep=3, max_idle_bytes=64 << 30):
        self.idle = {}
        self.keep, self.max_idle_bytes, self.idle_bytes = keep, max_idle_bytes, 0
        self.enabled = os.environ.get("MOPA_SCN_ARENA_POOL", "1") != "0"

    def get(self, nbytes, dev):
        nbytes = _arena_bytes(nbytes)
        if not self.enabled:
            return torch.empty(nbytes, dtype=torch.uint8, device=dev)
        key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, nbytes)
        lst = self.idle.get(key)
        if lst:
            self.idle_bytes -= nbytes
            return lst.pop()
        t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        t._mopa_arena_key = key
        return t

    def put(self, t):
        key = getattr(t, "_mopa_arena_key", None)
        if key is None or not self.enabled:
            return
        lst = self.idle.setdefault(key, [])
        if len(lst) < self.keep and self.idle_bytes + key[2] <= self.max_idle_bytes:
            lst.append(t)
            self.idle_bytes += key[2]

    def release(self):
        self.idle.clear()
        self.idle_bytes = 0


_arenas = _ArenaPool()


def release_arenas():
    """Drop the idle arenas of the compiled executor (they survive torch.cuda.empty_cache() by design)."""
    _arenas.release()


class _Lease:
    """Returns its tensors to the arena pool when the autograd context that holds it dies."""

    def __init__(self, *tensors):
        self.tensors = tensors

    def __del__(self):
        try:
            for t in self.tensors:
                _arenas.put(t)
        except Exception:  # interpreter shutdown
            pass


def _arena_bytes(n):
    """Arena sizes rounded up to 1/16 of their power of two: consecutive batches differ by a few per cent in size, and
    torch's caching allocator only reuses a cached block for a request of (nearly) the same size; exact sizes made it
    cudaMalloc a fresh multi-GB segment per distinct batch size and per step in flight (visible as 50 ms - 1 s stalls)."""
    n = int(n)
    if n < (1 << 20):
        return max(n, 1)
    q = 1 << (n.bit_length() - 5)
    return (n + q - 1) // q * q


class _Unsupported(Exception):
    pass


def _fbits(x):
    return struct.unpack("i", struct.pack("f", float(x)))[0]


class CompiledProgram:
    def __init__(self, root):
        mods = list(root._modules.values())
        if len(mods) < 2 or not isinstance(mods[0], M.InputLayer) or not isinstance(mods[-1], M.OutputLayer):
            raise _Unsupported("not an InputLayer ... OutputLayer chain")
        inp = mods[0]
        if inp.mode != 4 or inp.dimension != 3:
            raise _Unsupported("InputLayer mode/dimension")
        self.spatial = M._cube(inp.spatial_size)
        self.ops, self.bufs, self.slots = [], [], []  # slots: (module, attribute name) per param pointer slot
        self.max_level = 0
        self.in_planes = None
        b = self._new_buf(0, None)  # channels filled in by the first consumer
        for mod in mods[1:-1]:
            b = self._walk(mod, b)
        if isinstance(b, list):
            raise _Unsupported("chain ends in a table")
        if self.bufs[b][1] is None:
            raise _Unsupported("no layer between InputLayer and OutputLayer")
        self.out_buf = b
        self.in_planes = self.bufs[0][1]
        self.n_levels = self.max_level + 1
        self.bn_modules = [m for m, a in self.slots if a == "running_mean"]
        self.handle = None
        self._lib = _lib.load()

    # -- tree walk --------------------------------------------------------------------------------------------------
    def _new_buf(self, level, channels):
        self.bufs.append([level, channels, -1, 0])
        self.max_level = max(self.max_level, level)
        return len(self.bufs) - 1

    def _channels(self, b, expected):
        if self.bufs[b][1] is None:
            self.bufs[b][1] = expected
        if self.bufs[b][1] != expected:
            raise _Unsupported("plane count mismatch")

    def _walk(self, mod, b):
        if isinstance(mod, M.ConcatTable):
            return [self._walk(c, b) for c in mod._modules.values()]
        if isinstance(mod, (M.Sequential, torch.nn.Sequential)):
            for c in mod._modules.values():
                b = self._walk(c, b)
            return b
        if isinstance(mod, M.Identity):
            return b
        if isinstance(mod, M.JoinTable):
            if not isinstance(b, list) or any(isinstance(x, list) for x in b):
                raise _Unsupported("JoinTable input")
            level = self.bufs[b[0]][0]
            total = 0
            for x in b:
                if self.bufs[x][2] != -1 or x == 0 or self.bufs[x][0] != level or self.bufs[x][1] is None or x in b[:b.index(x)]:
                    raise _Unsupported("JoinTable of a shared / already joined buffer")
                total += self.bufs[x][1]
            j = self._new_buf(level, total)
            off = 0
            for x in b:
                self.bufs[x][2], self.bufs[x][3] = j, off
                off += self.bufs[x][1]
            return j
        if isinstance(b, list):
            raise _Unsupported("table fed to a non-table module")
        level = self.bufs[b][0]
        if isinstance(mod, M.BatchNormalization):
            if mod.nPlanes > 256:
                raise _Unsupported("BatchNormalization over more than 256 planes")
            self._channels(b, mod.nPlanes)
            out = self._new_buf(level, mod.nPlanes)
            self._op(OP_BN, b, out, level, level, mod.nPlanes, mod.nPlanes, mod.leakiness, mod.eps, mod.momentum)
            self.slots += [(mod, "weight"), (mod, "bias"), (mod, "running_mean"), (mod, "running_var")]
            return out
        if isinstance(mod, M.SubmanifoldConvolution):
            self._channels(b, mod.nIn)
            out = self._new_buf(level, mod.nOut)
            self._op(OP_SUBM, b, out, level, level, mod.nIn, mod.nOut)
            self.slots.append((mod, "weight"))
            return out
        if isinstance(mod, M.Convolution):
            self._channels(b, mod.nIn)
            out = self._new_buf(level + 1, mod.nOut)
            self._op(OP_CONV, b, out, level, level + 1, mod.nIn, mod.nOut)
            self.slots.append((mod, "weight"))
            return out
        if isinstance(mod, M.Deconvolution):
            if level == 0:
                raise _Unsupported("Deconvolution above the input level")
            self._channels(b, mod.nIn)
            out = self._new_buf(level - 1, mod.nOut)
            self._op(OP_DECONV, b, out, level, level - 1, mod.nIn, mod.nOut)
            self.slots.append((mod, "weight"))
            return out
        raise _Unsupported(type(mod).__name__)

    def _op(self, kind, b_in, b_out, l_in, l_out, n_in, n_out, leak=0.0, eps=0.0, momentum=0.0):
        self.ops.append([kind, b_in, b_out, 0, l_in, l_out, n_in, n_out, len(self.slots), _fbits(leak), _fbits(eps),
                         _fbits(momentum)])

    # -- runtime ----------------------------------------------------------------------------------------------------
    def ensure_handle(self, device_index):
        if self.handle is None:
            if self.spatial % (1 << self.max_level):
                raise _lib.ScnError("spatial size %d is not divisible by 2^%d" % (self.spatial, self.max_level))
            ops = np.ascontiguousarray(np.array(self.ops, dtype=np.int32))
            bufs = np.ascontiguousarray(np.array(self.bufs, dtype=np.int32))
            h = self._lib.mopa_scn_Program_new(ops.ctypes.data, len(self.ops), bufs.ctypes.data, len(self.bufs),
                                               self.in_planes, 0, self.out_buf, self.spatial, self.n_levels, device_index)
            if not h:
                raise _lib.ScnError(self._lib.mopa_scn_last_error().decode())
            self.handle, self.device_index = h, device_index
        elif self.device_index != device_index:
            raise _Unsupported("module moved to another device")
        return self.handle

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self._lib.mopa_scn_Program_delete(h)

    def tensors(self):
        return [getattr(m, a) for m, a in self.slots]


_programs = weakref.WeakKeyDictionary()

# parameter -> caller-owned gradient view (mopa_b200.parallel.FlatGradBucket.attach): when a parameter has no .grad yet, the
# compiled backward writes its gradient straight into that view, so a data-parallel step needs no gather / pack of the
# 78 gradient tensors before its single all-reduce
_grad_views = {}  # id(parameter) -> (weakref to the parameter, view); keyed by identity (tensors do not compare by ==)


def register_grad_views(mapping):
    """mapping: parameter -> float32 tensor view of the same shape (e.g. slices of one flat all-reduce buffer)."""
    for p, v in mapping.items():
        if v.shape != p.shape or v.dtype != torch.float32 or not v.is_contiguous():
            raise _lib.ScnError("gradient views must be contiguous float32 tensors of the parameter's shape")
        key = id(p)
        _grad_views[key] = (weakref.ref(p, lambda _, k=key: _grad_views.pop(k, None)), v)


def unregister_grad_views(params):
    for p in params:
        _grad_views.pop(id(p), None)


def _grad_view_of(t):
    entry = _grad_views.get(id(t))
    return entry[1] if entry is not None and entry[0]() is t else None


def compiled_for(root):
    """CompiledProgram for this module tree, or None if it cannot / should not be compiled."""
    if os.environ.get("MOPA_SCN_EAGER") == "1":
        return None
    sig = tuple(id(m) for m in root.modules())
    entry = _programs.get(root)
    if entry is None or entry[0] != sig:
        try:
            entry = (sig, CompiledProgram(root))
        except _Unsupported:
            entry = (sig, None)
        _programs[root] = entry
    prog = entry[1]
    if prog is None:
        return None
    if any(m.training != root.training for m in prog.bn_modules):
        return None  # mixed train/eval BatchNorms: run module by module
    return prog


class _ProgramFunction(Function):
    @staticmethod
    def forward(ctx, prog, coords, train, feats, *trainable):
        F._require_cuda(feats, "InputLayer features")
        L = prog._lib
        dev = feats.device
        handle = prog.ensure_handle(dev.index if dev.index is not None else torch.cuda.current_device())
        meta = F.Metadata(3, dev)
        where = F._coords_where(coords)
        if coords.dtype != torch.int64:
            coords = coords.long()
        coords = coords.contiguous()
        n, ncols = coords.shape
        feats, ld = F._rows(feats)
        if feats.shape[1] != prog.in_planes:
            raise _lib.ScnError("expected %d input planes, got %d" % (prog.in_planes, feats.shape[1]))
        if feats.shape[0] < n:
            raise _lib.ScnError("fewer feature rows than coordinates")
        prec = F._cfg["precision"]
        stream = F._stream(dev)
        n_active = (ctypes.c_int64 * prog.n_levels)()
        sizes = (ctypes.c_uint64 * 3)()
        with torch.cuda.device(dev):
            _lib.check(L.mopa_scn_Program_prepare(handle, meta._h, coords.data_ptr(), n, ncols, where, prec, stream, n_active,
                                                  sizes))
            meta.n_points = n
            act = _arenas.get(sizes[0], dev)
            scratch = _arenas.get(sizes[2], dev)
            out = torch.empty(n, prog.bufs[prog.out_buf][1], dtype=torch.float32, device=dev)
            tensors = prog.tensors()
            params = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
            _lib.check(L.mopa_scn_Program_forward(handle, meta._h, feats.data_ptr(), ld, params, 1 if train else 0, prec,
                                                  act.data_ptr(), scratch.data_ptr(), out.data_ptr(), out.shape[1], stream))
        _arenas.put(scratch)  # (packed weights + statistics blocks: only the forward kernels just queued read them)
        ctx.prog, ctx.meta, ctx.act, ctx.train, ctx.prec = prog, meta, act, train, prec
        ctx.lease = _Lease(act)
        ctx.sizes = (int(sizes[1]), int(sizes[2]))
        ctx.n_rows = feats.shape[0]
        # The tensors the backward pass reads (conv weights for d_input, BatchNorm weight / bias) go through
        # save_for_backward: an in-place update between this forward and its backward (optimizer.step(), .add_()) raises
        # autograd's version-counter error instead of silently differentiating against the new values, exactly as for
        # torch.nn layers. (torch_ema's .data.copy_ swaps bypass version counters for every layer, torch's included.)
        ctx.read_idx = [i for i, (m, a) in enumerate(prog.slots) if not a.startswith("running_")]
        ctx.save_for_backward(*[tensors[i] for i in ctx.read_idx])
        ctx.n_slots = len(tensors)
        ctx.last_metadata = meta
        return out

    @staticmethod
    def backward(ctx, d_out):
        prog, meta, L = ctx.prog, ctx.meta, ctx.prog._lib
        d_out, ld_dout = F._rows(d_out)
        dev = d_out.device
        saved = ctx.saved_tensors  # raises if one of them was modified in place since the forward
        tensors = [None] * ctx.n_slots  # running statistics are not read by the backward pass
        for i, t in zip(ctx.read_idx, saved):
            tensors[i] = t
        need = ctx.needs_input_grad  # (prog, coords, train, feats, *trainable)
        trainable_idx = ctx.read_idx
        sizes = [tensors[i].numel() for i in trainable_idx]
        with torch.cuda.device(dev):
            # gradients land either in the caller's registered views (first backward of a step: .grad is None, autograd
            # keeps the tensor we return) or in one fresh flat buffer (accumulating backward: autograd adds it to .grad)
            direct = [_grad_view_of(tensors[i]) if tensors[i].grad is None else None for i in trainable_idx]
            direct = [v if (v is not None and v.device == dev) else None for v in direct]
            flat = torch.empty(sum(sz for sz, v in zip(sizes, direct) if v is None), dtype=torch.float32, device=dev)
            grads, ptrs, off = [], [None] * len(tensors), 0
            for j, (i, sz) in enumerate(zip(trainable_idx, sizes)):
                if not need[4 + j]:
                    grads.append(None)
                    continue
                if direct[j] is not None:
                    g = direct[j].view_as(direct[j])  # a fresh tensor object: autograd only keeps (instead of cloning) a
                else:                                 # gradient nobody else holds a reference to
                    g = flat[off:off + sz].view_as(tensors[i])
                    off += sz
                ptrs[i] = g.data_ptr()
                grads.append(g)
            params = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() if t is not None else None for t in tensors])
            pgrads = (ctypes.c_void_p * len(tensors))(*ptrs)
            grad_arena = _arenas.get(ctx.sizes[0], dev)
            scratch = _arenas.get(ctx.sizes[1], dev)
            d_feats = torch.zeros(ctx.n_rows, prog.in_planes, dtype=torch.float32, device=dev) if need[3] else None
            _lib.check(L.mopa_scn_Program_backward(
                prog.handle, meta._h, params, pgrads, 1 if ctx.train else 0, ctx.prec, ctx.act.data_ptr(),
                grad_arena.data_ptr(), scratch.data_ptr(), d_out.data_ptr(), ld_dout,
                d_feats.data_ptr() if d_feats is not None else None, prog.in_planes, F._stream(dev)))
            _arenas.put(grad_arena)  # (the d_weight stream has been joined: everything that touches them is queued on this stream)
            _arenas.put(scratch)
        return (None, None, None, d_feats) + tuple(grads)


def run(prog, root, input):
    coords, feats = input[0], input[1]
    if coords.dim() != 2 or coords.shape[1] not in (3, 4):
        raise _lib.ScnError("InputLayer: coords must be (N, 3) or (N, 4)")
    tensors = prog.tensors()
    trainable = [t for t, (m, a) in zip(tensors, prog.slots) if not a.startswith("running_")]
    return _ProgramFunction.apply(prog, coords, root.training, feats, *trainable)
