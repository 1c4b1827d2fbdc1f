"""GPU parity of the cross-modal operators next to the UNetSCN path, through the C ABI of
include/mopa_xm.h (mopa_b200.xm):

  N2  lift_and_classify / xm_kl_div   against plain fp32 PyTorch ops on the same GPU (the reference IS three lines of torch:
      mopa/models/xmuda_arch.py:62-77, mopa/train/train_xmuda_mopa.py:389-398): gathers bit-exact, heads 1e-5, grads 1e-4
  N4  mask_cons_loss                  against golden vectors produced by the reference's own function
      (tests/golden/xm_mask_cons.npz) and against the float64 oracle on a full-size image batch: 1e-5 relative
  N3  post_process (VGI)              against golden vectors produced by the reference's own post_process
      (tests/golden/xm_vgi.npz): voxel coordinates, labels, masks BIT-EXACT, augmented points 1e-12; and against the
      oracle at SemanticKITTI scan size
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import xm_oracle as xo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def xm(cuda):
    import mopa_b200.xm as xm
    return xm


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ---------------------------------------------------------------------------------------------------------------- N2
@pytest.mark.parametrize("classes,dual", [(5, True), (10, True), (11, False)])
def test_lift_and_classify_matches_torch(xm, classes, dual):
    torch.manual_seed(classes)
    b, c, h, w = 4, 64, 57, 100  # UNetResNet34 emits 64 channels (xmuda_arch.py:36); img (8, 3, 225, 400) at full size
    x = torch.randn(b, c, h, w, device="cuda", requires_grad=True)
    counts = [900, 0, 1300, 700]  # an image without points in the middle
    idx = [torch.stack([torch.randint(0, h, (n,)), torch.randint(0, w, (n,))], 1).cuda() for n in counts]
    idx[0][:50] = idx[0][50:100]  # points sharing a pixel: their gradients add up
    lin, lin2 = torch.nn.Linear(c, classes).cuda(), (torch.nn.Linear(c, classes).cuda() if dual else None)
    got = xm.lift_and_classify(x, idx, lin, lin2)
    xr = x.detach().clone().requires_grad_(True)
    w1, b1 = lin.weight.detach().clone().requires_grad_(True), lin.bias.detach().clone().requires_grad_(True)
    w2 = lin2.weight.detach().clone().requires_grad_(True) if dual else None
    b2 = lin2.bias.detach().clone().requires_grad_(True) if dual else None
    ref = xo.lift_and_classify(xr, idx, w1, b1, w2, b2)
    assert set(got) == set(ref)
    assert torch.equal(got["feats"], ref["feats"])  # a gather: bit-exact
    assert _rel(got["seg_logit"], ref["seg_logit"]) < 1e-5
    g = torch.Generator(device="cuda").manual_seed(1)

    def loss(p):
        out = (p["feats"] * torch.randn(p["feats"].shape, device="cuda", generator=g)).sum()
        out = out + (p["seg_logit"] * torch.randn(p["seg_logit"].shape, device="cuda", generator=g)).sum()
        if dual:
            out = out + (p["seg_logit2"] * torch.randn(p["seg_logit2"].shape, device="cuda", generator=g)).sum()
        return out

    loss(got).backward()
    g.manual_seed(1)
    loss(ref).backward()
    assert _rel(x.grad, xr.grad) < 1e-4
    assert _rel(lin.weight.grad, w1.grad) < 1e-4 and _rel(lin.bias.grad, b1.grad) < 1e-4
    if dual:
        assert _rel(got["seg_logit2"], ref["seg_logit2"]) < 1e-5
        assert _rel(lin2.weight.grad, w2.grad) < 1e-4 and _rel(lin2.bias.grad, b2.grad) < 1e-4


def test_lift_negative_indices_wrap_and_out_of_range_is_reported(xm):
    from mopa_b200 import _lib
    x = torch.randn(1, 64, 8, 9, device="cuda")
    lin = torch.nn.Linear(64, 5).cuda()
    idx = [torch.tensor([[-1, -2], [3, 4]]).cuda()]
    got = xm.lift_and_classify(x, idx, lin)
    assert torch.equal(got["feats"], xo.lift_and_classify(x, idx, lin.weight, lin.bias)["feats"])
    xm.lift_and_classify(x, [torch.tensor([[8, 0]]).cuda()], lin)  # row 8 of an 8-row map
    with pytest.raises(_lib.ScnError):
        _lib.check(_lib.load().mopa_xm_checkAsyncError(torch.cuda.current_stream().cuda_stream))


@pytest.mark.parametrize("n,classes", [(1, 5), (3001, 5), (260000, 10)])
def test_xm_kl_div_matches_torch(xm, n, classes):
    torch.manual_seed(n)
    s = (torch.randn(n, classes, device="cuda") * 3).requires_grad_(True)
    t = torch.randn(n, classes, device="cuda") * 3
    got = xm.xm_kl_div(s, t)
    sr = s.detach().clone().requires_grad_(True)
    ref = xo.xm_kl_div(sr, t)
    assert abs(float(got) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    (got * 0.7).backward()
    (ref * 0.7).backward()
    assert _rel(s.grad, sr.grad) < 1e-4


# ---------------------------------------------------------------------------------------------------------------- N4
def test_mask_cons_loss_matches_reference_golden(xm):
    z = np.load(os.path.join(GOLD, "xm_mask_cons.npz"))
    for tag in ("a", "b"):
        masks = [torch.from_numpy(m).cuda() for m in z[tag + "_masks"]]
        for me in (0, 1):
            x = torch.from_numpy(z[tag + "_logits"]).cuda().requires_grad_(True)
            probs = torch.softmax(x, dim=3)
            probs.retain_grad()
            loss = xm.mask_cons_loss(probs, masks, bool(me))
            loss.backward()
            want = float(z["%s_loss_%d" % (tag, me)])
            assert abs(float(loss) - want) < 1e-5 * max(1.0, abs(want)), (tag, me, float(loss), want)
            assert np.allclose(probs.grad.cpu().numpy(), z["%s_dprobs_%d" % (tag, me)], rtol=1e-3, atol=1e-7)
            assert np.allclose(x.grad.cpu().numpy(), z["%s_dlogits_%d" % (tag, me)], rtol=1e-3, atol=1e-7)


def test_mask_cons_loss_full_size_vs_oracle_and_edge_cases(xm):
    from mopa_b200 import _lib
    torch.manual_seed(3)
    b, h, w, c = 8, 225, 400, 5  # the nuScenes image batch of a training step (xmuda.py:98)
    logits = torch.randn(b, h, w, c, device="cuda")
    # SAM-like masks: blocks of pixels share an id (uint8 range), -100 = invalid
    ids = torch.randint(0, 200, (b, h // 15 + 1, w // 20 + 1), device="cuda")
    masks = ids.repeat_interleave(15, 1).repeat_interleave(20, 2)[:, :h, :w].clone()
    masks[torch.rand(b, h, w, device="cuda") < 0.1] = -100
    for me in (False, True):
        x = logits.clone().requires_grad_(True)
        loss = xm.mask_cons_loss(torch.softmax(x, dim=3), [m for m in masks], me)
        loss.backward()
        xr = logits.double().cpu().requires_grad_(True)
        ref = xo.mask_cons_loss(torch.softmax(xr, dim=3), [m.cpu() for m in masks], me)
        ref.backward()
        assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
        assert _rel(x.grad, xr.grad) < 1e-3
    assert xm.mask_cons_loss(torch.softmax(logits, 3), [], True) == 0  # the reference returns the int 0 here
    only_invalid = xm.mask_cons_loss(torch.softmax(logits[:1], 3), [torch.full((h, w), -100, device="cuda")], True)
    assert float(only_invalid) == 0.0
    xm.mask_cons_loss(torch.softmax(logits[:1], 3), [torch.full((h, w), 300, device="cuda")], True)  # id >= 256
    with pytest.raises(_lib.ScnError):
        _lib.check(_lib.load().mopa_xm_checkAsyncError(torch.cuda.current_stream().cuda_stream))


# ---------------------------------------------------------------------------------------------------------------- N3
def test_vgi_post_process_matches_reference_golden(xm):
    z = np.load(os.path.join(GOLD, "xm_vgi.npz"))
    n = int(z["n_scans"])
    scans = [(z["pc%d" % i], z["label%d" % i], z["mask%d" % i]) for i in range(n)]
    augment = {"noisy_rot": 0.1, "flip_y": 0.5, "rot_z": 6.2831, "transl": True}
    for tag, use_proj in (("full", True), ("noproj", False)):
        np.random.seed(int(z["seed"]))
        cat_input, label, mask, aug = xm.post_process([s[0] for s in scans], [s[1] for s in scans], [s[2] for s in scans], 20,
                                                      4096, augment, use_proj=use_proj)
        assert np.array_equal(cat_input["x"][0].cpu().numpy(), z[tag + "_locs"])  # voxel coordinates + batch index: exact
        assert np.array_equal(cat_input["x"][1].cpu().numpy(), z[tag + "_feats"])
        assert np.array_equal(label.cpu().numpy(), z[tag + "_label"]) and np.array_equal(mask.cpu().numpy(), z[tag + "_mask"])
        for i, a in enumerate(aug):
            assert np.allclose(a.cpu().numpy(), z["%s_aug%d" % (tag, i)], rtol=0, atol=1e-12)
    np.random.seed(99)
    cat_input, label, mask, _ = xm.post_process([scans[0][0]], [scans[0][1]], [scans[0][2]], 20, 4096,
                                                {"noisy_rot": 0.0, "rot_z": 0.0, "transl": False}, use_proj=True)
    assert np.array_equal(cat_input["x"][0].cpu().numpy(), z["plain_locs"])
    assert np.array_equal(label.cpu().numpy(), z["plain_label"]) and np.array_equal(mask.cpu().numpy(), z["plain_mask"])


def test_vgi_post_process_full_scan_vs_oracle_and_feeds_unet(xm):
    """SemanticKITTI-size scan (config 4) + an inserted object: same rows as the oracle, and the result goes straight into
    UNetSCN (device coordinates) as the third forward of a MoPA step does (train_xmuda_mopa.py:556-558)."""
    from mopa_b200 import synth
    from mopa_b200.xm import vgi
    from mopa_b200.unet_scn import UNetSCN
    rng = np.random.default_rng(5)
    pts = synth.lidar_points("kitti", 21)
    anchor = pts[rng.integers(0, pts.shape[0])]
    obj = anchor * 0.7 + rng.uniform(-1.0, 1.0, size=(1500, 3)) * np.array([1.0, 1.0, 0.7])
    pc = np.concatenate([np.concatenate([pts, obj]), rng.uniform(size=(pts.shape[0] + 1500, 1))], 1)
    mask = np.zeros(pc.shape[0], bool)
    mask[pts.shape[0]:] = True
    label = rng.integers(0, 10, pc.shape[0])
    augment = {"noisy_rot": 0.1, "flip_y": 0.5, "rot_z": 6.2831, "transl": True}
    np.random.seed(7)
    cat_input, lab, msk, aug = xm.post_process([pc], [label], [mask], 20, 4096, augment)
    np.random.seed(7)
    rot = vgi._rotation_and_translation(augment)
    rand3 = np.random.rand(3)
    c, rows, p = xo.vgi_post_process_scan(pc, mask, 20, 4096, rot, rand3)
    assert c.shape[0] < pc.shape[0]  # occluded points were removed
    assert np.array_equal(cat_input["x"][0][:, :3].cpu().numpy(), c)
    assert np.array_equal(lab.cpu().numpy(), label[rows]) and np.array_equal(msk.cpu().numpy(), mask[rows])
    assert np.allclose(aug[0].cpu().numpy(), p, rtol=0, atol=1e-12)
    import mopa_b200.scn as scn
    scn.set_precision("tf32")
    out = UNetSCN(1).cuda()(cat_input["x"])
    assert out.shape == (c.shape[0], 16) and torch.isfinite(out).all()
