"""N3: VGI post-processing on the GPU (mopa/data/mixmatch_ss.py:458-559) over mopa_xm_VgiPostProcess.

`post_process` keeps the reference's signature and return value. What moves to the GPU, per scan: the range-image
occlusion test (augmentation_3d.py:161-280), rotation / scaling / rounding / random translation (augmentation_3d.py:6-60),
the receptive-field filter, and the gathers of pseudo labels and object masks. The random draws stay on the host and are
taken from numpy's global RNG in EXACTLY the reference's order (randn(3,3), randint, randint, rand, rand(3)), so a seeded
run produces the same voxel coordinates.
"""
import ctypes

import numpy as np
import torch

from .. import _lib


def _rotation_and_translation(augment_3d):
    """The host-side random part of augment_and_scale_3d (augmentation_3d.py:26-57): same calls, same order, same dtypes."""
    noisy_rot = augment_3d["noisy_rot"]
    flip_y = augment_3d["flip_y"] if "flip_y" in augment_3d.keys() else 0.0
    flip_x = augment_3d["flip_x"] if "flip_x" in augment_3d.keys() else 0.0
    rot_z = augment_3d["rot_z"]
    rot = None
    if noisy_rot > 0 or flip_x > 0 or flip_y > 0 or rot_z > 0:
        rot = np.eye(3, dtype=np.float32)
        if noisy_rot > 0:
            rot += np.random.randn(3, 3) * noisy_rot
        if flip_x > 0:
            rot[0][0] *= np.random.randint(0, 2) * 2 - 1
        if flip_y > 0:
            rot[1][1] *= np.random.randint(0, 2) * 2 - 1
        if rot_z > 0:
            theta = np.random.rand() * rot_z
            z_rot = np.array([[np.cos(theta), -np.sin(theta), 0], [np.sin(theta), np.cos(theta), 0], [0, 0, 1]], dtype=np.float32)
            rot = rot.dot(z_rot)
    return rot


def post_process(cat_pc_ls, cat_pslabel_ls, obj_mask_ls, scale, full_scale, augment_3d, proj_W=1024, proj_H=64,
                 fov_up=0.05235, fov_down=-0.43633, scan_pth_ls=None, use_proj=True, backbone="SCN", device=None):
    """Same arguments and return value as the reference's post_process: [cat_input, cat_ps_label, obj_mask, aug_points_ls]
    with cat_input = {'x': [locs (N, 4) int64, feats (N, 1) float32 ones]}. The tensors are returned ON THE GPU (the
    reference returns host tensors and the train loop moves them, train_xmuda_mopa.py:546-551; `.cuda()` on them is a
    no-op and mopa_b200.scn's InputLayer takes device coordinates directly); aug_points_ls holds device tensors (N_i, 3)."""
    if "SCN" not in backbone:
        raise IndexError("The specified backbone is not supported: {}".format(backbone))
    if not torch.cuda.is_available():
        raise _lib.ScnError("post_process: no CUDA device (mopa_b200.xm has no CPU path)")
    L = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    locs, feats, pseudo_label, mask_ls, aug_points_ls = [], [], [], [], []
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        for i in range(len(cat_pc_ls)):
            pc = np.asarray(cat_pc_ls[i])
            if np.any(np.isnan(pc[:, :3])):
                raise AssertionError("Found Nan object points: {}".format(scan_pth_ls[i] if scan_pth_ls else i))
            obj = np.ascontiguousarray(np.asarray(obj_mask_ls[i]).astype(np.uint8))
            n = pc.shape[0]
            pts = torch.from_numpy(np.ascontiguousarray(pc[:, :3], dtype=np.float64)).to(dev)
            obj_d = torch.from_numpy(obj).to(dev)
            proj = bool(use_proj and np.any(obj))
            rot = _rotation_and_translation(augment_3d)
            rand3 = np.random.rand(3) if augment_3d["transl"] else None
            rot_c = np.ascontiguousarray(rot, dtype=np.float64) if rot is not None else None
            keep = torch.empty(n, dtype=torch.uint8, device=dev)
            coords = torch.empty(n, 4, dtype=torch.int64, device=dev)
            sel = torch.empty(n, dtype=torch.int64, device=dev)
            aug = torch.empty(n, 3, dtype=torch.float64, device=dev)
            ws_bytes = L.mopa_xm_vgiWorkspaceBytes(n, proj_H, proj_W)
            ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
            ws_ptr = (ws.data_ptr() + 255) & ~255
            n_out = ctypes.c_int64(0)
            _lib.check(L.mopa_xm_VgiPostProcess(
                pts.data_ptr(), obj_d.data_ptr(), n, 1 if proj else 0, float(fov_up), float(fov_down), int(proj_W), int(proj_H),
                rot_c.ctypes.data if rot_c is not None else None, rand3.ctypes.data if rand3 is not None else None,
                float(scale), int(full_scale), i, keep.data_ptr(), coords.data_ptr(), sel.data_ptr(), aug.data_ptr(),
                ctypes.byref(n_out), ws_ptr, ws_bytes, stream))
            m = n_out.value
            sel_m = sel[:m]
            locs.append(coords[:m])
            feats.append(torch.ones(m, 1, dtype=torch.float32, device=dev))
            pseudo_label.append(torch.from_numpy(np.asarray(cat_pslabel_ls[i])).to(dev)[sel_m])
            mask_ls.append(torch.from_numpy(np.asarray(obj_mask_ls[i])).to(dev)[sel_m])
            aug_points_ls.append(aug[:m])
        cat_input = {"x": [torch.cat(locs, 0), torch.cat(feats, 0)]}
        cat_ps_label = torch.cat(pseudo_label, 0)
        obj_mask = torch.cat(mask_ls, 0)
    return [cat_input, cat_ps_label, obj_mask, aug_points_ls]
