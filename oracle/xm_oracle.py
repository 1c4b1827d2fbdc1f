"""CPU ORACLE for the cross-modal operators next to the UNetSCN path -- TEST INFRASTRUCTURE, NOT
PRODUCT CODE (same rules as scn_oracle.py: only tests/, smoke() and bench.py's cpu_baseline legs may import it).

PARITY PINNED for N3 and N4: tests/test_oracle_xm.py checks these restatements against golden vectors produced by the
reference's own functions (tests/golden/make_xm_golden.py runs /root/reference/mopa/common/utils/loss.py::mask_cons_loss
and /root/reference/mopa/data/mixmatch_ss.py::post_process unmodified). N2 is three lines of torch in the reference
(xmuda_arch.py:62-77, train_xmuda_mopa.py:389-398) and is restated verbatim.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------------- N2
def lift_and_classify(x, img_indices, w1, b1, w2=None, b2=None):
    """Net2DSeg.forward after the 2D network, /root/reference/mopa/models/xmuda_arch.py:62-77."""
    img_feats = []
    for i in range(x.shape[0]):
        img_feats.append(x.permute(0, 2, 3, 1)[i][img_indices[i][:, 0], img_indices[i][:, 1]])
    img_feats = torch.cat(img_feats, 0)
    preds = {"feats": img_feats, "seg_logit": F.linear(img_feats, w1, b1)}
    if w2 is not None:
        preds["seg_logit2"] = F.linear(img_feats, w2, b2)
    return preds


def xm_kl_div(student, teacher):
    """/root/reference/mopa/train/train_xmuda_mopa.py:389-398."""
    return F.kl_div(F.log_softmax(student, dim=1), F.softmax(teacher.detach(), dim=1), reduction="none").sum(1).mean()


# ---------------------------------------------------------------------------------------------------------------- N4
def mask_cons_loss(all_logits, sam_mask_ls, min_entropy=False):
    """/root/reference/mopa/common/utils/loss.py:241-283 as segment sums (differentiable, any float dtype).
    all_logits (B, H, W, C); the entropy term is divided by log2(all_logits.shape[1]) as the reference does (:262)."""
    if len(sam_mask_ls) == 0:
        return 0
    norm = math.log2(all_logits.shape[1])
    total = 0
    for b, masks in enumerate(sam_mask_ls):
        x = all_logits[b].reshape(-1, all_logits.shape[-1])
        m = torch.as_tensor(masks).reshape(-1)
        ids = [int(i) for i in torch.unique(m) if int(i) >= 0]
        img = 0
        for i in ids:
            seg = x[m == i]
            mean = seg.mean(0, keepdim=True)
            cur = ((seg - mean) ** 2).mean()
            if min_entropy:
                cur = cur - torch.sum(mean[0] * torch.log2(mean[0] + 1e-30)) / norm
            img = img + cur
        total = total + (img / len(ids) if ids else 0)
    return total / len(sam_mask_ls)


# ---------------------------------------------------------------------------------------------------------------- N3
def vgi_keep_mask(points, obj_mask, fov_up=0.05235, fov_down=-0.43633, proj_W=1024, proj_H=64):
    """range_projection(..., obj_mask)['pres_idx'], /root/reference/mopa/data/utils/augmentation_3d.py:161-280, restated per
    pixel: a range-image pixel that holds an inserted-object point keeps only its nearest point (ties: lowest index, the
    stable lexsort of occulusion_detector :81-111); every other pixel keeps all its points."""
    p = np.asarray(points, np.float64)[:, :3]
    depth = np.linalg.norm(p, 2, axis=1)
    yaw = -np.arctan2(p[:, 1], p[:, 0])
    pitch = np.arcsin(p[:, 2] / depth)
    fov = abs(fov_down) + abs(fov_up)
    px = 0.5 * (yaw / np.pi + 1.0)
    py = 1.0 - (pitch + abs(fov_down)) / fov
    px *= proj_W
    py *= proj_H
    px = np.maximum(0, np.minimum(proj_W - 1, np.floor(px))).astype(np.int64)
    py = np.maximum(0, np.minimum(proj_H - 1, np.floor(py))).astype(np.int64)
    pix = py * proj_W + px
    obj = np.asarray(obj_mask).astype(bool)
    flagged = np.zeros(proj_W * proj_H, bool)
    flagged[pix[obj]] = True
    keep = np.ones(p.shape[0], bool)
    cand = np.nonzero(flagged[pix])[0]
    if cand.size == 0:
        return keep
    order = cand[np.lexsort((cand, depth[cand], pix[cand]))]  # by pixel, then depth, then index
    first = np.concatenate([[True], pix[order][1:] != pix[order][:-1]])
    keep[order[~first]] = False
    return keep


def vgi_post_process_scan(pc, obj_mask, scale, full_scale, rot, rand3, use_proj=True, **proj):
    """One scan of post_process (/root/reference/mopa/data/mixmatch_ss.py:517-538) given the host-drawn rotation matrix
    (or None) and translation factors (or None): returns (coords int64 (M, 3), rows (M,) original indices, aug points)."""
    pc = np.asarray(pc, np.float64)
    obj = np.asarray(obj_mask).astype(bool)
    keep = vgi_keep_mask(pc, obj, **proj) if (use_proj and obj.any()) else np.ones(pc.shape[0], bool)
    rows = np.nonzero(keep)[0]
    pts = pc[rows, :3]
    pts = pts.dot(rot) if rot is not None else pts
    coords = np.round(pts * scale)
    coords -= coords.min(0)
    if rand3 is not None:
        coords += np.clip(full_scale - coords.max(0) - 0.001, a_min=0, a_max=None) * rand3
    idxs = (coords.min(1) >= 0) * (coords.max(1) < full_scale)
    return coords.astype(np.int64)[idxs], rows[idxs], pts[idxs]
