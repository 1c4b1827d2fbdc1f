"""N2: point-to-pixel feature gather + segmentation heads + cross-modal KL (mopa/models/xmuda_arch.py:62-77,
mopa/train/train_xmuda_mopa.py:389-398, 440-445) as autograd Functions over mopa_xm_PixelGatherHeads_* / mopa_xm_KLDivLoss_*."""
import ctypes

import torch
from torch.autograd import Function

from .. import _lib


def _cuda_f32(t, what):
    if not t.is_cuda:
        raise _lib.ScnError("%s must be a CUDA tensor: mopa_b200.xm has no CPU path (got device %s)" % (what, t.device))
    if t.dtype != torch.float32:
        raise _lib.ScnError("%s must be float32 (got %s)" % (what, t.dtype))
    return t.contiguous()


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t):
    return t.data_ptr() if t is not None else None


class _LiftHeads(Function):
    @staticmethod
    def forward(ctx, x, idx, offsets, w1, b1, w2, b2):
        L = _lib.load()
        x = _cuda_f32(x, "feature map")
        dev = x.device
        b, c, h, w = x.shape
        n, k = idx.shape[0], w1.shape[0]
        with torch.cuda.device(dev):
            feats = torch.empty(n, c, dtype=torch.float32, device=dev)
            logit = torch.empty(n, k, dtype=torch.float32, device=dev)
            logit2 = torch.empty(n, k, dtype=torch.float32, device=dev) if w2 is not None else None
            off = (ctypes.c_int64 * (b + 1))(*offsets)
            _lib.check(L.mopa_xm_PixelGatherHeads_updateOutput(
                x.data_ptr(), b, c, h, w, idx.data_ptr(), off, n, w1.data_ptr(), _ptr(b1), _ptr(w2), _ptr(b2), k,
                feats.data_ptr(), logit.data_ptr(), _ptr(logit2), _stream(dev)))
        ctx.save_for_backward(feats, idx, w1, w2 if w2 is not None else w1)
        ctx.cfg = (tuple(x.shape), tuple(offsets), w2 is not None, b1 is not None, b2 is not None)
        if logit2 is None:
            return feats, logit
        return feats, logit, logit2

    @staticmethod
    def backward(ctx, d_feats, d_logit, d_logit2=None):
        feats, idx, w1, w2 = ctx.saved_tensors
        shape, offsets, has2, has_b1, has_b2 = ctx.cfg
        L = _lib.load()
        dev = feats.device
        b, c, h, w = shape
        n, k = feats.shape[0], w1.shape[0]

        def prep(g):
            return g.contiguous() if g is not None else None

        d_feats, d_logit, d_logit2 = prep(d_feats), prep(d_logit), prep(d_logit2) if has2 else None
        with torch.cuda.device(dev):
            d_x = torch.zeros(shape, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
            d_w1, d_b1 = torch.empty_like(w1), torch.empty(k, dtype=torch.float32, device=dev)
            d_w2 = torch.empty_like(w2) if has2 else None
            d_b2 = torch.empty(k, dtype=torch.float32, device=dev) if has2 else None
            ws_bytes = L.mopa_xm_pixelGatherWorkspaceBytes(c, k)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            off = (ctypes.c_int64 * (b + 1))(*offsets)
            _lib.check(L.mopa_xm_PixelGatherHeads_backward(
                feats.data_ptr(), idx.data_ptr(), off, b, c, h, w, n, w1.data_ptr(), _ptr(w2) if has2 else None, k,
                _ptr(d_feats), _ptr(d_logit), _ptr(d_logit2), _ptr(d_x), d_w1.data_ptr(), d_b1.data_ptr(), _ptr(d_w2),
                _ptr(d_b2), ws.data_ptr(), ws_bytes, _stream(dev)))
        return d_x, None, None, d_w1, d_b1 if has_b1 else None, d_w2, d_b2 if has_b2 else None


def lift_and_classify(x, img_indices, linear, linear2=None):
    """Net2DSeg.forward after the 2D network (xmuda_arch.py:57-77): `x` (B, C, H, W) feature map, `img_indices` list of B
    (N_i, 2) integer [row, col] tensors / arrays, `linear` / `linear2` the nn.Linear heads. Returns the same dict:
    {'feats': (N, C), 'seg_logit': (N, classes)[, 'seg_logit2']} with N = sum N_i, samples concatenated in order."""
    if len(img_indices) != x.shape[0]:
        raise _lib.ScnError("img_indices must hold one index tensor per image (%d != %d)" % (len(img_indices), x.shape[0]))
    idx = [torch.as_tensor(i).to(device=x.device, dtype=torch.int64).reshape(-1, 2) for i in img_indices]
    offsets = [0]
    for i in idx:
        offsets.append(offsets[-1] + i.shape[0])
    idx = torch.cat(idx, 0).contiguous() if idx else torch.zeros(0, 2, dtype=torch.int64, device=x.device)
    w1, b1 = _cuda_f32(linear.weight, "linear.weight"), linear.bias
    w2 = _cuda_f32(linear2.weight, "linear2.weight") if linear2 is not None else None
    b2 = linear2.bias if linear2 is not None else None
    out = _LiftHeads.apply(x, idx, offsets, w1, b1, w2, b2)
    preds = {"feats": out[0], "seg_logit": out[1]}
    if linear2 is not None:
        preds["seg_logit2"] = out[2]
    return preds


class _KLDiv(Function):
    @staticmethod
    def forward(ctx, student, teacher):
        L = _lib.load()
        student, teacher = _cuda_f32(student, "student logits"), _cuda_f32(teacher, "teacher logits")
        if student.shape != teacher.shape or student.dim() != 2:
            raise _lib.ScnError("xm_kl_div expects two (N, classes) logit matrices of the same shape")
        dev = student.device
        n, k = student.shape
        with torch.cuda.device(dev):
            loss = torch.empty((), dtype=torch.float32, device=dev)
            grad = torch.empty_like(student) if ctx.needs_input_grad[0] else None
            _lib.check(L.mopa_xm_KLDivLoss_updateOutput(student.data_ptr(), teacher.data_ptr(), n, k, loss.data_ptr(),
                                                        _ptr(grad), 1.0, _stream(dev)))
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g if grad is not None else None), None


def xm_kl_div(student_logit, teacher_logit):
    """F.kl_div(F.log_softmax(student, 1), F.softmax(teacher.detach(), 1), reduction='none').sum(1).mean()
    (train_xmuda_mopa.py:389-398): one kernel computes the loss and its gradient w.r.t. the student logits."""
    return _KLDiv.apply(student_logit, teacher_logit.detach())
