"""N4: SAM mask-consistency loss (mopa/common/utils/loss.py:241-283) over mopa_xm_MaskConsLoss_*."""
import math

import torch
from torch.autograd import Function

from .. import _lib

MAX_MASK_IDS = 256  # SAM masks are stored as uint8 (nuscenes_dataloader.py:325); -100 marks invalid pixels


class _MaskCons(Function):
    @staticmethod
    def forward(ctx, probs, masks, min_entropy, entropy_norm, max_ids):
        L = _lib.load()
        dev = probs.device
        b, h, w, c = probs.shape
        with torch.cuda.device(dev):
            stats = torch.empty(L.mopa_xm_maskConsStatsBytes(b, max_ids, c), dtype=torch.uint8, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            _lib.check(L.mopa_xm_MaskConsLoss_updateOutput(probs.data_ptr(), masks.data_ptr(), b, h * w, c, max_ids,
                                                           1 if min_entropy else 0, entropy_norm, stats.data_ptr(),
                                                           loss.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(probs, masks, stats)
        ctx.cfg = (min_entropy, entropy_norm, max_ids)
        return loss

    @staticmethod
    def backward(ctx, g):
        probs, masks, stats = ctx.saved_tensors
        min_entropy, entropy_norm, max_ids = ctx.cfg
        L = _lib.load()
        dev = probs.device
        b, h, w, c = probs.shape
        with torch.cuda.device(dev):
            d = torch.empty_like(probs)
            _lib.check(L.mopa_xm_MaskConsLoss_backward(probs.data_ptr(), masks.data_ptr(), b, h * w, c, max_ids,
                                                       1 if min_entropy else 0, entropy_norm, stats.data_ptr(), 1.0,
                                                       d.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        return d * g, None, None, None, None


def mask_cons_loss(all_logits, sam_mask_ls, min_entropy=False, max_ids=MAX_MASK_IDS):
    """Same call and result as the reference's mask_cons_loss (loss.py:241-283): `all_logits` (B, H, W, C) softmaxed
    logits as the train loop passes them (train_xmuda_mopa.py:473-478), `sam_mask_ls` list of B (H, W) integer masks
    (ids < 0 ignored). For every mask id: MSE of the pixels' vectors to their mean [+ entropy of the mean / log2(K)];
    averaged over the ids of an image, then over the images. Reference quirk kept: K is read from all_logits.shape[1]
    (the image height in that layout), loss.py:262. Returns 0 for an empty list, like the reference."""
    if len(sam_mask_ls) == 0:
        return 0
    if not all_logits.is_cuda:
        raise _lib.ScnError("mask_cons_loss: logits must be a CUDA tensor (mopa_b200.xm has no CPU path)")
    if all_logits.dim() != 4 or len(sam_mask_ls) > all_logits.shape[0]:
        raise _lib.ScnError("mask_cons_loss expects (B, H, W, C) logits and at most B masks")
    b = len(sam_mask_ls)
    probs = all_logits[:b].float().contiguous()
    masks = torch.stack([torch.as_tensor(m).to(device=probs.device) for m in sam_mask_ls]).to(torch.int32).contiguous()
    if tuple(masks.shape) != tuple(probs.shape[:3]):
        raise _lib.ScnError("mask_cons_loss: masks must be (H, W) = %s" % (tuple(probs.shape[1:3]),))
    entropy_norm = math.log2(all_logits.shape[1])
    return _MaskCons.apply(probs, masks, bool(min_entropy), float(entropy_norm), int(max_ids))
