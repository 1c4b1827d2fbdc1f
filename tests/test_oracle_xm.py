"""Pins oracle/xm_oracle.py (N3 / N4 restatements) against golden vectors produced by the REFERENCE'S OWN functions
(tests/golden/make_xm_golden.py; mask_cons_loss and post_process run unmodified from /root/reference), and checks the
host-side RNG replay of mopa_b200/xm/vgi.py (the random draws of augment_and_scale_3d, same calls in the same order)."""
import os

import numpy as np
import torch

from oracle import xm_oracle as xo

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_mask_cons_oracle_matches_reference_golden():
    z = np.load(os.path.join(GOLD, "xm_mask_cons.npz"))
    for tag in ("a", "b"):
        masks = torch.from_numpy(z[tag + "_masks"])
        for me in (0, 1):
            x = torch.from_numpy(z[tag + "_logits"]).double().requires_grad_(True)
            probs = torch.softmax(x, dim=3)
            probs.retain_grad()
            loss = xo.mask_cons_loss(probs, [m for m in masks], bool(me))
            loss.backward()
            assert abs(float(loss) - float(z["%s_loss_%d" % (tag, me)])) < 2e-6 * max(1.0, abs(float(loss)))
            assert np.allclose(probs.grad.numpy(), z["%s_dprobs_%d" % (tag, me)], rtol=2e-4, atol=1e-7)
            assert np.allclose(x.grad.numpy(), z["%s_dlogits_%d" % (tag, me)], rtol=2e-4, atol=1e-7)


def _replay(z, tag, use_proj, seed, augment, scans):
    from mopa_b200.xm import vgi
    np.random.seed(seed)
    locs, labels, masks, augs = [], [], [], []
    for i in scans:
        pc, lab, msk = z["pc%d" % i], z["label%d" % i], z["mask%d" % i]
        rot = vgi._rotation_and_translation(augment)
        rand3 = np.random.rand(3) if augment["transl"] else None
        c, rows, pts = xo.vgi_post_process_scan(pc, msk, 20, 4096, rot, rand3, use_proj=use_proj)
        locs.append(np.concatenate([c, np.full((c.shape[0], 1), len(locs), np.int64)], 1))
        labels.append(lab[rows])
        masks.append(msk[rows])
        augs.append(pts)
    return np.concatenate(locs), np.concatenate(labels), np.concatenate(masks), augs


def test_vgi_oracle_and_rng_replay_match_reference_golden():
    z = np.load(os.path.join(GOLD, "xm_vgi.npz"))
    augment = {"noisy_rot": 0.1, "flip_y": 0.5, "rot_z": 6.2831, "transl": True}
    for tag, use_proj in (("full", True), ("noproj", False)):
        locs, labels, masks, augs = _replay(z, tag, use_proj, int(z["seed"]), augment, range(int(z["n_scans"])))
        assert np.array_equal(locs, z[tag + "_locs"])
        assert np.array_equal(labels, z[tag + "_label"]) and np.array_equal(masks, z[tag + "_mask"])
        for i, a in enumerate(augs):
            assert np.allclose(a, z["%s_aug%d" % (tag, i)], rtol=0, atol=1e-12)
    assert z["full_locs"].shape[0] < z["noproj_locs"].shape[0]  # the occlusion test really removed points
    locs, labels, masks, _ = _replay(z, "plain", True, 99, {"noisy_rot": 0.0, "rot_z": 0.0, "transl": False}, [0])
    assert np.array_equal(locs, z["plain_locs"]) and np.array_equal(labels, z["plain_label"])


def test_vgi_keep_mask_properties():
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(4000, 3)) * np.array([20, 20, 2.0])
    obj = np.zeros(4000, bool)
    assert xo.vgi_keep_mask(pts, obj).all()  # no inserted points: nothing is removed
    obj[:300] = True
    keep = xo.vgi_keep_mask(pts, obj)
    assert 0 < (~keep).sum() < 4000
    # idempotent: applying the test to the survivors removes nothing more
    assert xo.vgi_keep_mask(pts[keep], obj[keep]).all()
