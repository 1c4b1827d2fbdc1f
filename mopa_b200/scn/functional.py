"""Metadata handle + autograd Functions over the C ABI (include/mopa_scn.h).

Host-side mirror of [UPSTREAM] sparseconvnet/{metadata,ioLayers,submanifoldConvolution,convolution,deconvolution,
batchNormalization}.py: each Function's forward/backward is one or two calls into libmopa_scn.so on the current CUDA
stream. No CPU fallback: a non-CUDA feature tensor raises.
"""
import ctypes
import os

import torch
from torch.autograd import Function

from .. import _lib

_PRECISIONS = {"tf32": _lib.PREC_TF32, "fp32": _lib.PREC_FP32}


def _default_precision():
    """fp32 products by default, as upstream SparseConvNet's SIMT kernels compute (a drop-in must not change the training
    numerics silently). TF32 tensor-core products (the fast path: tcgen05 kernels, ~2^-11 relative error per product) are
    opt-in: scn.set_precision('tf32') or MOPA_SCN_PRECISION=tf32, like torch.backends.cuda.matmul.allow_tf32."""
    name = os.environ.get("MOPA_SCN_PRECISION", "fp32").lower()
    if name not in _PRECISIONS:
        raise _lib.ScnError("MOPA_SCN_PRECISION must be 'fp32' or 'tf32' (got %r)" % name)
    return _PRECISIONS[name]


_cfg = {"precision": _default_precision()}


def set_precision(name):
    """'fp32' (default: 3xTF32 split operands, fp32-equivalent products) or 'tf32' (one TF32 MMA per product on the
    tcgen05 kernels: the fast path, opt-in)."""
    _cfg["precision"] = _PRECISIONS[name]


def get_precision():
    return "tf32" if _cfg["precision"] == _lib.PREC_TF32 else "fp32"


def _stream(device=None):
    """Raw handle of torch's current stream ON THE TENSOR'S DEVICE (not on the current device)."""
    return torch.cuda.current_stream(device).cuda_stream


def _coords_where(coords):
    """coords_on_device argument of the C ABI: 0 host tensor, 1 device tensor (the geometry stream waits for the caller's
    stream), 2 device tensor known to be complete: it carries the event of the copy that produced it
    (mopa_b200.data.DevicePrefetcher / mark_ready) and is passed on as it is, so the geometry of this forward may overlap
    whatever the caller's stream is still running."""
    if not coords.is_cuda:
        return 0
    ready = getattr(coords, "_mopa_ready", None)
    if ready is None or coords.dtype != torch.int64 or not coords.is_contiguous():
        return 1
    ready.synchronize()  # recorded a step ago: returns at once
    return 2


def _on_device(fn):
    """Runs an autograd Function's forward / backward with the device of its first CUDA tensor argument current, so
    that allocations, the stream handle (_stream()) and the library's cudaSetDevice all agree, and the caller's current
    device is restored afterwards (the C entry points call cudaSetDevice(handle device) and do not restore it)."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args):
        for a in args:
            if torch.is_tensor(a) and a.is_cuda:
                with torch.cuda.device(a.device):
                    return fn(*args)
        return fn(*args)

    return wrapped


def _require_cuda(t, what):
    if not t.is_cuda:
        raise _lib.ScnError("%s must be a CUDA tensor: mopa_b200.scn has no CPU path (got device %s)" % (what, t.device))
    if t.dtype != torch.float32:
        raise _lib.ScnError("%s must be float32 (got %s)" % (what, t.dtype))


def _rows(t):
    """Feature matrix with unit plane stride; returns (tensor, row stride in floats)."""
    if t.dim() != 2:
        raise _lib.ScnError("feature tensors are 2-D (rows, planes)")
    if t.shape[0] > 1 and (t.stride(1) != 1 or t.stride(0) < t.shape[1]):
        t = t.contiguous()
    elif t.shape[0] <= 1 and t.stride(1) != 1:
        t = t.contiguous()
    return t, (t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1))


class Metadata:
    """Owns the GPU grids / rulebooks of one forward (replaces [UPSTREAM] sparseconvnet.SCN.Metadata_3)."""

    def __init__(self, dimension=3, device=None):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.ScnError("no CUDA device: mopa_b200.scn runs on the GPU only (no CPU fallback)")
        dev = torch.cuda.current_device() if device is None else torch.device(device).index
        if dev is None:
            dev = torch.cuda.current_device()
        self.device_index = dev
        self.dimension = dimension
        self._h = self._lib.mopa_scn_Metadata_new(dimension, dev)
        if not self._h:
            raise _lib.ScnError(self._lib.mopa_scn_last_error().decode())
        self.n_points = 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.mopa_scn_Metadata_delete(h)

    # -- geometry -----------------------------------------------------------------------------------------------
    def set_locations(self, coords, spatial_size, mode=4):
        if coords.dtype != torch.int64:
            coords = coords.long()
        coords = coords.contiguous()
        n, ncols = (coords.shape[0], coords.shape[1]) if coords.dim() == 2 else (0, 4)
        out = ctypes.c_int64(0)
        with torch.cuda.device(self.device_index):
            _lib.check(self._lib.mopa_scn_InputLayer_setLocations(
                self._h, int(spatial_size), coords.data_ptr(), n, ncols, 1 if coords.is_cuda else 0, mode, _stream(),
                ctypes.byref(out)))
        self.n_points = n
        return out.value

    def prepare_submanifold(self, spatial_size, filter_size=3):
        out = ctypes.c_int64(0)
        with torch.cuda.device(self.device_index):
            _lib.check(self._lib.mopa_scn_Metadata_prepareSubmanifold(self._h, int(spatial_size), filter_size, _stream(),
                                                                      ctypes.byref(out)))
        return out.value

    def prepare_convolution(self, in_size, out_size, filter_size=2, stride=2):
        out = ctypes.c_int64(0)
        with torch.cuda.device(self.device_index):
            _lib.check(self._lib.mopa_scn_Metadata_prepareConvolution(self._h, int(in_size), int(out_size), filter_size,
                                                                      stride, _stream(), ctypes.byref(out)))
        return out.value

    def n_active(self, spatial_size):
        return int(self._lib.mopa_scn_Metadata_getNActive(self._h, int(spatial_size)))

    # -- inspection (parity tests) ------------------------------------------------------------------------------
    def spatial_locations(self, spatial_size):
        v = self.n_active(spatial_size)
        out = torch.empty(v, 4, dtype=torch.int64)
        _lib.check(self._lib.mopa_scn_Metadata_getSpatialLocations(self._h, int(spatial_size), out.data_ptr()))
        return out

    def point_to_voxel(self):
        out = torch.empty(self.n_points, dtype=torch.int32)
        _lib.check(self._lib.mopa_scn_Metadata_getPointToVoxel(self._h, out.data_ptr()))
        return out

    def input_rules(self, spatial_size):
        v = self.n_active(spatial_size)
        off = torch.empty(v + 1, dtype=torch.int32)
        rows = torch.empty(self.n_points, dtype=torch.int32)
        _lib.check(self._lib.mopa_scn_Metadata_getInputRules(self._h, off.data_ptr(), rows.data_ptr()))
        return off, rows

    def _rulebook(self, fn, spatial_size, volume):
        counts = torch.zeros(volume, dtype=torch.int64)
        _lib.check(fn(self._h, int(spatial_size), counts.data_ptr(), None))
        pairs = torch.empty(int(counts.sum()), 2, dtype=torch.int32)
        _lib.check(fn(self._h, int(spatial_size), counts.data_ptr(), pairs.data_ptr()))
        return list(torch.split(pairs, counts.tolist()))

    def submanifold_rule_counts(self, spatial_size):
        """rules per offset (27 ints) of the 3x3x3 rulebook at this level"""
        counts = torch.zeros(27, dtype=torch.int64)
        _lib.check(self._lib.mopa_scn_Metadata_getSubmanifoldRuleBook(self._h, int(spatial_size), counts.data_ptr(), None))
        return counts.tolist()

    def submanifold_rulebook(self, spatial_size):
        """list over the 27 offsets of (R_k, 2) int32 [in, out], ascending out row"""
        return self._rulebook(self._lib.mopa_scn_Metadata_getSubmanifoldRuleBook, spatial_size, 27)

    def convolution_rulebook(self, in_spatial_size):
        """list over the 8 offsets of (R_k, 2) int32 [fine, coarse], ascending coarse row"""
        return self._rulebook(self._lib.mopa_scn_Metadata_getConvolutionRuleBook, in_spatial_size, 8)


    def tile_rulebook(self, spatial_size, kind):
        """(lists int32 (tiles, K, 128), masks int32 (tiles, K, 4)) of the tile rulebook the tcgen05 conv kernels read;
        kind: 'subm' (K = 27), 'child' (Convolution spatial_size -> /2 by coarse rows), 'select' (by fine rows)."""
        code = {"subm": 0, "child": 1, "select": 2}[kind]
        k = 27 if code == 0 else 8
        tiles = ctypes.c_int64(0)
        with torch.cuda.device(self.device_index):
            _lib.check(self._lib.mopa_scn_Metadata_getTileRuleBook(self._h, int(spatial_size), code, ctypes.byref(tiles),
                                                                   None, None))
            lists = torch.empty(tiles.value, k, 128, dtype=torch.int32)
            masks = torch.empty(tiles.value, k, 4, dtype=torch.int32)
            _lib.check(self._lib.mopa_scn_Metadata_getTileRuleBook(self._h, int(spatial_size), code, ctypes.byref(tiles),
                                                                   lists.data_ptr(), masks.data_ptr()))
        return lists, masks


# ---------------------------------------------------------------------------------------------------------------
_bn_ws = {}


def _bn_workspace(device):
    """Persistent zero-initialised scratch per device (the library leaves it zeroed after every call)."""
    ws = _bn_ws.get(device)
    if ws is None:
        nbytes = _lib.load().mopa_scn_bnWorkspaceBytes(256)
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _bn_ws[device] = ws
    return ws


def _mma_ok(c_in, c_out):
    np_ = c_out // 16
    return c_in % 16 == 0 and c_out % 16 == 0 and c_in >= 16 and c_out >= 16 and (np_ <= 8 or np_ in (10, 12))


def _pack(weight, volume, n_in, n_out, transpose, flip):
    """Fragment-order copy of the weights for the MMA kernels (None when the shape takes the generic path).
    Re-packed on every call: parameters can be swapped in place without a version bump (EMA, train_xmuda_mopa.py:266)."""
    c_in, c_out = (n_out, n_in) if transpose else (n_in, n_out)
    if not _mma_ok(c_in, c_out):
        return None
    L = _lib.load()
    prec = _cfg["precision"]
    packed = torch.empty(L.mopa_scn_packedWeightFloats(volume, n_in, n_out, prec), dtype=torch.float32, device=weight.device)
    _lib.check(L.mopa_scn_packWeights(weight.data_ptr(), volume, n_in, n_out, transpose, flip, prec, packed.data_ptr(),
                                      _stream()))
    return packed


def _ptr(t):
    return t.data_ptr() if t is not None else None


class InputLayerFunction(Function):
    @staticmethod
    @_on_device
    def forward(ctx, feats, metadata, n_active):
        _require_cuda(feats, "InputLayer features")
        feats, ld = _rows(feats)
        out = torch.empty(n_active, feats.shape[1], dtype=torch.float32, device=feats.device)
        _lib.check(metadata._lib.mopa_scn_InputLayer_updateOutput(metadata._h, feats.data_ptr(), ld, feats.shape[1],
                                                                  out.data_ptr(), max(feats.shape[1], 1), _stream()))
        ctx.meta = metadata
        ctx.n_rows = feats.shape[0]
        return out

    @staticmethod
    @_on_device
    def backward(ctx, d_out):
        m = ctx.meta
        d_out, ld = _rows(d_out)
        planes = d_out.shape[1]
        d_in = torch.zeros(ctx.n_rows, planes, dtype=torch.float32, device=d_out.device)
        _lib.check(m._lib.mopa_scn_InputLayer_updateGradInput(m._h, d_in.data_ptr(), planes, d_out.data_ptr(), ld, planes,
                                                              _stream()))
        return d_in, None, None


class OutputLayerFunction(Function):
    @staticmethod
    @_on_device
    def forward(ctx, feats, metadata):
        _require_cuda(feats, "OutputLayer features")
        feats, ld = _rows(feats)
        planes = feats.shape[1]
        out = torch.empty(metadata.n_points, planes, dtype=torch.float32, device=feats.device)
        _lib.check(metadata._lib.mopa_scn_OutputLayer_updateOutput(metadata._h, feats.data_ptr(), ld, planes, out.data_ptr(),
                                                                   planes, _stream()))
        ctx.meta = metadata
        ctx.n_active = feats.shape[0]
        return out

    @staticmethod
    @_on_device
    def backward(ctx, d_out):
        m = ctx.meta
        d_out, ld = _rows(d_out)
        planes = d_out.shape[1]
        d_in = torch.empty(ctx.n_active, planes, dtype=torch.float32, device=d_out.device)
        _lib.check(m._lib.mopa_scn_OutputLayer_updateGradInput(m._h, d_in.data_ptr(), planes, d_out.data_ptr(), ld, planes,
                                                               _stream()))
        return d_in, None


class _ConvFunction(Function):
    """kind: 'subm' (sizes = (spatial,)), 'conv' (in, out), 'deconv' (in = coarse, out = fine)."""

    @staticmethod
    @_on_device
    def forward(ctx, feats, weight, metadata, kind, sizes, n_out_rows, filter_size, stride):
        _require_cuda(feats, "convolution features")
        feats, ld_in = _rows(feats)
        volume, n_in, n_out = weight.shape[0], weight.shape[-2], weight.shape[-1]
        if feats.shape[1] != n_in:
            raise _lib.ScnError("convolution expects %d input planes, got %d" % (n_in, feats.shape[1]))
        w = weight.contiguous()
        L = metadata._lib
        out = torch.empty(n_out_rows, n_out, dtype=torch.float32, device=feats.device)
        packed = _pack(w, volume, n_in, n_out, 0, 0)
        prec = _cfg["precision"]
        s = _stream()
        if kind == "subm":
            st = L.mopa_scn_SubmanifoldConvolution_updateOutput(metadata._h, sizes[0], filter_size, feats.data_ptr(), ld_in,
                                                                out.data_ptr(), n_out, w.data_ptr(), _ptr(packed), n_in,
                                                                n_out, prec, s)
        elif kind == "conv":
            st = L.mopa_scn_Convolution_updateOutput(metadata._h, sizes[0], sizes[1], filter_size, stride, feats.data_ptr(),
                                                     ld_in, out.data_ptr(), n_out, w.data_ptr(), _ptr(packed), n_in, n_out,
                                                     prec, s)
        else:
            st = L.mopa_scn_Deconvolution_updateOutput(metadata._h, sizes[0], sizes[1], filter_size, stride,
                                                       feats.data_ptr(), ld_in, out.data_ptr(), n_out, w.data_ptr(),
                                                       _ptr(packed), n_in, n_out, prec, s)
        _lib.check(st)
        ctx.save_for_backward(feats, w)
        ctx.meta, ctx.kind, ctx.sizes, ctx.fs = metadata, kind, sizes, (filter_size, stride)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, d_out):
        feats, w = ctx.saved_tensors
        m, kind, sizes = ctx.meta, ctx.kind, ctx.sizes
        filter_size, stride = ctx.fs
        L = m._lib
        d_out, ld_dout = _rows(d_out)
        feats, ld_in = _rows(feats)
        volume, n_in, n_out = w.shape[0], w.shape[-2], w.shape[-1]
        need_in, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_in = torch.empty_like(feats, memory_format=torch.contiguous_format) if need_in else None
        if need_in and d_in.shape[0] > 0 and kind == "deconv":
            pass  # every coarse row is written (a coarse site always has >= 1 child)
        d_w = torch.empty_like(w) if need_w else None
        packed_t = _pack(w, volume, n_in, n_out, 1, 1 if kind == "subm" else 0) if need_in else None
        n_rows = d_out.shape[0] if kind != "conv" else d_out.shape[0]
        ws_bytes = L.mopa_scn_backwardWorkspaceBytes(volume, n_in, n_out, n_rows) if need_w else 16
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=feats.device)
        prec = _cfg["precision"]
        s = _stream()
        if kind == "subm":
            st = L.mopa_scn_SubmanifoldConvolution_backward(m._h, sizes[0], filter_size, feats.data_ptr(), ld_in, _ptr(d_in),
                                                            n_in, d_out.data_ptr(), ld_dout, w.data_ptr(), _ptr(packed_t),
                                                            _ptr(d_w), n_in, n_out, prec, ws.data_ptr(), ws_bytes, s)
        elif kind == "conv":
            st = L.mopa_scn_Convolution_backward(m._h, sizes[0], sizes[1], filter_size, stride, feats.data_ptr(), ld_in,
                                                 _ptr(d_in), n_in, d_out.data_ptr(), ld_dout, w.data_ptr(), _ptr(packed_t),
                                                 _ptr(d_w), n_in, n_out, prec, ws.data_ptr(), ws_bytes, s)
        else:
            st = L.mopa_scn_Deconvolution_backward(m._h, sizes[0], sizes[1], filter_size, stride, feats.data_ptr(), ld_in,
                                                   _ptr(d_in), n_in, d_out.data_ptr(), ld_dout, w.data_ptr(), _ptr(packed_t),
                                                   _ptr(d_w), n_in, n_out, prec, ws.data_ptr(), ws_bytes, s)
        _lib.check(st)
        return d_in, d_w, None, None, None, None, None, None


def sparse_conv(feats, weight, metadata, kind, sizes, n_out_rows, filter_size, stride):
    return _ConvFunction.apply(feats, weight, metadata, kind, tuple(int(x) for x in sizes), int(n_out_rows), filter_size,
                               stride)


class BatchNormFunction(Function):
    @staticmethod
    @_on_device
    def forward(ctx, feats, weight, bias, running_mean, running_var, eps, momentum, train, leakiness):
        _require_cuda(feats, "BatchNormalization features")
        feats, ld_in = _rows(feats)
        n, planes = feats.shape
        L = _lib.load()
        out = torch.empty(n, planes, dtype=torch.float32, device=feats.device)
        save_mean = torch.empty(planes, dtype=torch.float32, device=feats.device)
        save_invstd = torch.empty(planes, dtype=torch.float32, device=feats.device)
        ws = _bn_workspace(feats.device)
        _lib.check(L.mopa_scn_BatchNormalization_updateOutput(
            feats.data_ptr(), ld_in, out.data_ptr(), planes, save_mean.data_ptr(), save_invstd.data_ptr(),
            running_mean.data_ptr(), running_var.data_ptr(), weight.data_ptr(), bias.data_ptr(), eps, momentum,
            1 if train else 0, leakiness, n, planes, ws.data_ptr(), ws.numel(), _stream()))
        ctx.save_for_backward(feats, weight, bias, save_mean, save_invstd)
        ctx.cfg = (train, leakiness)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, d_out):
        feats, weight, bias, save_mean, save_invstd = ctx.saved_tensors
        train, leakiness = ctx.cfg
        feats, ld_in = _rows(feats)
        d_out, ld_dout = _rows(d_out)
        n, planes = feats.shape
        L = _lib.load()
        d_in = torch.empty(n, planes, dtype=torch.float32, device=feats.device) if ctx.needs_input_grad[0] else None
        d_w = torch.empty_like(weight)
        d_b = torch.empty_like(bias)
        ws = _bn_workspace(feats.device)
        _lib.check(L.mopa_scn_BatchNormalization_backward(
            feats.data_ptr(), ld_in, _ptr(d_in), planes, d_out.data_ptr(), ld_dout, save_mean.data_ptr(),
            save_invstd.data_ptr(), weight.data_ptr(), bias.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), leakiness,
            1 if train else 0, n, planes, ws.data_ptr(), ws.numel(), _stream()))
        return d_in, d_w, d_b, None, None, None, None, None, None
