#!/usr/bin/env python
"""Golden vectors for the N3 / N4 operators, produced by the REFERENCE'S OWN functions (read from /root/reference in the
build container; the GPU box has no reference checkout, so the vectors are committed):

  mask_cons_loss   /root/reference/mopa/common/utils/loss.py:241-283          -> xm_mask_cons.npz
  post_process     /root/reference/mopa/data/mixmatch_ss.py:458-559           -> xm_vgi.npz
                   (range_projection / occulusion_detector / augment_and_scale_3d, mopa/data/utils/augmentation_3d.py)

    python tests/golden/make_xm_golden.py

The reference runs unmodified; see _reference_import.py for the import shims (absent third-party modules stubbed,
`.cuda()` a no-op on this CPU box, np.bool8 restored for numpy 2).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import _reference_import as ri  # noqa: E402
from mopa_b200 import synth  # noqa: E402


def make_mask_cons():
    loss_mod = ri.load("mopa.common.utils.loss")
    g = torch.Generator().manual_seed(7)
    out = {}
    for tag, (b, h, w, c) in {"a": (2, 14, 22, 5), "b": (3, 9, 16, 10)}.items():
        logits = torch.randn(b, h, w, c, generator=g) * 2.0
        masks = torch.randint(0, 7, (b, h, w), generator=g)
        masks[torch.rand(b, h, w, generator=g) < 0.15] = -100  # invalid pixels (refine_sam_mask)
        if b == 3:
            masks[2] = -100  # an image without any valid mask contributes 0 but still counts in the mean
        out[tag + "_logits"] = logits.numpy()
        out[tag + "_masks"] = masks.numpy()
        for me in (False, True):
            x = logits.clone().requires_grad_(True)
            probs = torch.softmax(x, dim=3)
            probs.retain_grad()
            loss = loss_mod.mask_cons_loss(probs, [m for m in masks], me)
            loss.backward()
            out["%s_loss_%d" % (tag, me)] = np.float64(loss.item())
            out["%s_dprobs_%d" % (tag, me)] = probs.grad.numpy()
            out["%s_dlogits_%d" % (tag, me)] = x.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "xm_mask_cons.npz"), **out)
    print("mask_cons:", {k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items() if "loss" in k})


def _scan_with_object(seed, n_scene=2500, n_obj=260):
    """A small scene (subsampled synthetic sweep) + an inserted box-shaped object placed on a line of sight, so that it
    occludes scene points behind it, is partly occluded by nearer ones, and occludes itself."""
    rng = np.random.default_rng(seed)
    pts = synth.lidar_points("kitti", seed, n_azimuth=300)
    pts = pts[rng.choice(pts.shape[0], n_scene, replace=False)]
    anchor = pts[rng.integers(0, n_scene)]
    centre = anchor * rng.uniform(0.5, 0.9)  # between the sensor and a scene point
    obj = centre + rng.uniform(-0.8, 0.8, size=(n_obj, 3)) * np.array([1.0, 1.0, 0.6])
    feats = rng.uniform(0, 1, size=(n_scene + n_obj, 1))
    cat = np.concatenate([np.concatenate([pts, obj], 0), feats], 1)
    mask = np.zeros(n_scene + n_obj, dtype=bool)
    mask[n_scene:] = True
    order = rng.permutation(n_scene + n_obj)  # point_mixmatch concatenates; interleave to exercise index tie-breaks too
    label = rng.integers(-1, 10, size=n_scene + n_obj)
    label[label < 0] = -100
    return cat[order], label[order], mask[order]


def make_vgi():
    mm = ri.load("mopa.data.mixmatch_ss")
    scans = [_scan_with_object(s) for s in (11, 12, 13)]
    scans[2] = (scans[2][0], scans[2][1], np.zeros_like(scans[2][2]))  # a scan where nothing was inserted: no occlusion test
    augment = {"noisy_rot": 0.1, "flip_y": 0.5, "rot_z": 6.2831, "transl": True}  # configs/*/xmuda_pl_mopa.yaml
    out = {"n_scans": np.int64(len(scans)), "scale": np.int64(20), "full_scale": np.int64(4096), "seed": np.int64(1234)}
    for i, (pc, lab, msk) in enumerate(scans):
        out["pc%d" % i], out["label%d" % i], out["mask%d" % i] = pc, lab, msk
    for tag, kwargs in {"full": dict(use_proj=True), "noproj": dict(use_proj=False)}.items():
        np.random.seed(1234)
        cat_input, ps_label, obj_mask, aug_pts = mm.post_process(
            [s[0] for s in scans], [s[1] for s in scans], [s[2] for s in scans], 20, 4096, augment,
            scan_pth_ls=["scan%d" % i for i in range(len(scans))], **kwargs)
        out[tag + "_locs"] = cat_input["x"][0].numpy()
        out[tag + "_feats"] = cat_input["x"][1].numpy()
        out[tag + "_label"] = ps_label.numpy()
        out[tag + "_mask"] = obj_mask.numpy()
        for i, a in enumerate(aug_pts):
            out["%s_aug%d" % (tag, i)] = a
        print("vgi", tag, "rows", out[tag + "_locs"].shape, "of", sum(s[0].shape[0] for s in scans))
    # a plain (no-rotation, no-translation) configuration as well
    np.random.seed(99)
    cat_input, ps_label, obj_mask, aug_pts = mm.post_process(
        [scans[0][0]], [scans[0][1]], [scans[0][2]], 20, 4096, {"noisy_rot": 0.0, "rot_z": 0.0, "transl": False},
        scan_pth_ls=["scan0"], use_proj=True)
    out["plain_locs"], out["plain_label"], out["plain_mask"] = cat_input["x"][0].numpy(), ps_label.numpy(), obj_mask.numpy()
    np.savez_compressed(os.path.join(HERE, "xm_vgi.npz"), **out)


if __name__ == "__main__":
    make_mask_cons()
    make_vgi()
