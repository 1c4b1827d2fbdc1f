"""The 3D branch of a MoPA / xMUDA training step on the GPU, as the reference drives it (BASELINE.json configs 2 and 3;
mopa/train/train_xmuda_mopa.py:260-343, 417-418, 420-427, 556-558, 578-593; mopa/train/train_xmuda.py:226-333):

    zero_grad
    [EMA teacher] no-grad, eval-mode forward on the un-augmented target scans with the EMA weights swapped in (.data)
    forward(source) -> loss -> backward                                    (first backward() of the step)
    forward(target); forward(target + inserted objects) -> one backward    (second backward(): gradients ACCUMULATE)
    optimizer step, torch.cuda.empty_cache()

Several forwards are alive before a backward, an eval forward with swapped parameter storage sits in between, the BatchNorm
running statistics move three times, and two backward() calls accumulate into the same .grad tensors. The whole sequence
is replayed on the float64 oracle (Net3DSeg = UNetSCN + two Linear heads, mopa/models/xmuda_arch.py:82-126) and
every accumulated gradient, running statistic and logit is compared.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mopa_b200 import synth
from oracle import scn_oracle as so
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

NUM_CLASSES = 5  # nuScenes xM configs (configs/nuscenes/*/xmuda_pl_mopa.yaml)


@pytest.fixture(scope="module")
def scn(cuda):
    import mopa_b200.scn as scn
    keep = scn.get_precision()
    yield scn
    scn.set_precision(keep)


def _rel_l2_cos(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm()), float(torch.dot(a, b) / (a.norm() * b.norm()))


class OracleNet3DSeg:
    """float64 replay of Net3DSeg.forward (xmuda_arch.py:114-126) on top of the oracle UNetSCN."""

    def __init__(self, state):
        unet = {k[len("net_3d."):]: v for k, v in state.items() if k.startswith("net_3d.")}
        self.unet = so.OracleUNetSCN(unet, dtype=torch.float64)
        self.heads = {k: v.detach().clone().double().requires_grad_(True) for k, v in state.items() if k.startswith("linear")}

    def forward(self, coords, feats, train=True):
        f = self.unet.forward(coords, feats, train=train)
        return {"feats": f,
                "seg_logit": f @ self.heads["linear.weight"].t() + self.heads["linear.bias"],
                "seg_logit2": f @ self.heads["linear2.weight"].t() + self.heads["linear2.bias"]}

    def params(self):
        out = {"net_3d." + k: v for k, v in self.unet.params.items()}
        out.update(self.heads)
        return out


def _losses(preds, labels, other_logit):
    """CE on the main head + KL(second head || other modality's prediction), the two 3D-side losses of a step
    (train_xmuda_mopa.py:365-398): both heads and the shared UNetSCN receive gradients."""
    ce = F.cross_entropy(preds["seg_logit"], labels)
    kl = F.kl_div(F.log_softmax(preds["seg_logit2"], dim=1), F.softmax(other_logit, dim=1), reduction="none").sum(1).mean()
    return ce + 0.1 * kl


def _make_inputs(seed, n_scans=2, n_az=160, extra=0):
    coords, feats = synth.make_batch(n_scans, "nuscenes", seed, n_azimuth=n_az)
    if extra:  # "cat" batch: target scans + an inserted object of `extra` points per scan (VGI, train_xmuda_mopa.py:483-558)
        rng = np.random.default_rng(seed)
        parts_c, parts_f = [], []
        for b in range(n_scans):
            sel = coords[:, 3] == b
            c = coords[sel]
            centre = c[rng.integers(0, c.shape[0]), :3]
            obj = np.clip(centre + rng.integers(-12, 13, size=(extra, 3)), 0, 4095)
            obj = np.concatenate([obj, np.full((extra, 1), b, np.int64)], 1)
            parts_c += [c, obj]
            parts_f += [feats[sel], np.ones((extra, 1), np.float32)]
        coords, feats = np.concatenate(parts_c, 0), np.concatenate(parts_f, 0)
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, NUM_CLASSES, (coords.shape[0],), generator=g)
    other = torch.randn(coords.shape[0], NUM_CLASSES, generator=g, dtype=torch.float64)
    return coords, feats, labels, other


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_mopa_step_pattern_accumulated_gradients_match_oracle(scn, precision):
    from mopa_b200.unet_scn import Net3DSeg
    scn.set_precision(precision)
    torch.manual_seed(0)
    model = Net3DSeg(NUM_CLASSES, dual_head=True, backbone_3d="SCN", backbone_3d_kwargs={"in_channels": 1}).cuda()
    unet_state = so.make_unet_state(seed=21)
    model.net_3d.load_state_dict(unet_state)
    state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    oracle = OracleNet3DSeg(state)
    # EMA teacher weights: a perturbed copy of the parameters (torch_ema keeps shadow params, swaps .data in place)
    g = torch.Generator().manual_seed(5)
    shadow = {n: (p.detach().cpu() * (1 + 0.05 * torch.randn(p.shape, generator=g))) for n, p in model.named_parameters()}

    src, trg, cat, ori = _make_inputs(1), _make_inputs(2), _make_inputs(2, extra=150), _make_inputs(3)

    def gpu_forward(inp):
        coords, feats = inp[0], inp[1]
        return model({"x": [torch.from_numpy(coords), torch.from_numpy(feats).cuda()]})

    # ---- GPU, in the reference's order ------------------------------------------------------------------------
    model.train()
    model.zero_grad(set_to_none=True)
    with torch.no_grad():  # ema_model_3d.average_parameters(): store, copy_to, ..., restore (all through .data)
        stored = {n: p.data.clone() for n, p in model.named_parameters()}
        for n, p in model.named_parameters():
            p.data.copy_(shadow[n].to(p.device))
        model.eval()
        ema_logit = gpu_forward(ori)["seg_logit"]
        for n, p in model.named_parameters():
            p.data.copy_(stored[n])
        model.train()
    p_src = gpu_forward(src)
    _losses(p_src, src[2].cuda(), src[3].float().cuda()).backward()
    p_trg = gpu_forward(trg)
    p_cat = gpu_forward(cat)  # two forwards alive
    loss_trg = _losses(p_trg, trg[2].cuda(), trg[3].float().cuda()) + 0.1 * F.cross_entropy(p_cat["seg_logit"], cat[2].cuda())
    loss_trg.backward()
    torch.cuda.empty_cache()  # train_xmuda_mopa.py:593

    # ---- oracle, same sequence ---------------------------------------------------------------------------------
    teacher = OracleNet3DSeg({**state, **{k: v for k, v in shadow.items()}})
    with torch.no_grad():
        ref_ema = teacher.forward(ori[0], ori[1], train=False)["seg_logit"]
    r_src = oracle.forward(src[0], src[1])
    _losses(r_src, src[2], src[3]).backward()
    r_trg = oracle.forward(trg[0], trg[1])
    r_cat = oracle.forward(cat[0], cat[1])
    (_losses(r_trg, trg[2], trg[3]) + 0.1 * F.cross_entropy(r_cat["seg_logit"], cat[2])).backward()

    # bars: <= 2x the worst value measured for THIS test (fp32 mode: rel-L2 3.4e-2 / cos 0.99945 on a level-4 BatchNorm weight;
    # the scans here are small -- ~3k points each, levels 5-6 hold a few dozen voxels -- so the backward pass is worse
    # conditioned than at full size, where the worst tensor sits at 9.9e-3: profiles/r02_grad_errors.txt)
    tol_f, tol_l2, tol_cos = {"fp32": (1e-4, 6e-2, 0.999), "tf32": (1e-2, 0.2, 0.98)}[precision]
    assert rel_err(ema_logit, ref_ema) < tol_f  # eval-mode forward with swapped weights, running stats untouched by it
    assert rel_err(p_src["seg_logit"], r_src["seg_logit"]) < tol_f
    assert rel_err(p_cat["seg_logit2"], r_cat["seg_logit2"]) < tol_f
    ref_params = oracle.params()
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        l2, cos = _rel_l2_cos(p.grad, ref_params[name].grad)
        assert l2 < tol_l2 and cos > tol_cos, (name, l2, cos)
    for name, buf in model.named_buffers():  # three train-mode forwards moved every running statistic three times
        assert rel_err(buf, ref_params[name]) < tol_f, name
    # parameters are back to the student's values after the EMA swap
    for n, p in model.named_parameters():
        assert torch.equal(p.detach().cpu(), state[n]), n


def test_in_place_weight_update_between_forward_and_backward_raises(scn):
    """The compiled executor saves the tensors its backward reads through save_for_backward: optimizer-style in-place
    updates between a forward and its backward raise autograd's version error instead of silently using new weights."""
    from mopa_b200.unet_scn import UNetSCN
    scn.set_precision("tf32")
    coords, feats = synth.make_batch(1, "nuscenes", 0, n_azimuth=100)
    net = UNetSCN(1).cuda()
    out = net([torch.from_numpy(coords), torch.from_numpy(feats).cuda()])
    with torch.no_grad():
        net.sparseModel[1].weight.add_(1.0)  # what optimizer.step() does
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        out.sum().backward()


def test_net3dseg_dual_head_forward_backward(scn):
    """Net3DSeg (xmuda_arch.py:82-126) with dual_head=True, C = 10 (A2D2/SemanticKITTI configs): dict keys, shapes, and
    both heads' gradients reach the shared backbone; no-batch-column coords as in xmuda_arch.py:171."""
    from mopa_b200.unet_scn import Net3DSeg
    scn.set_precision("tf32")
    torch.manual_seed(1)
    model = Net3DSeg(10, dual_head=True, backbone_3d="SCN", backbone_3d_kwargs={"in_channels": 1}).cuda()
    coords, feats = synth.make_scan("nuscenes", seed=4, n_azimuth=200)  # (N, 3): single sample, no batch column
    preds = model({"x": [torch.from_numpy(coords), torch.from_numpy(feats).cuda()]})
    n = coords.shape[0]
    assert set(preds) == {"feats", "seg_logit", "seg_logit2"}
    assert preds["feats"].shape == (n, 16) and preds["seg_logit"].shape == (n, 10) and preds["seg_logit2"].shape == (n, 10)
    state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    oracle = OracleNet3DSeg(state)
    ref = oracle.forward(coords, feats)
    assert rel_err(preds["seg_logit"], ref["seg_logit"]) < 5e-2 and rel_err(preds["seg_logit2"], ref["seg_logit2"]) < 5e-2
    only2 = torch.autograd.grad(preds["seg_logit2"].sum(), model.net_3d.sparseModel[1].weight, retain_graph=True)[0]
    assert torch.isfinite(only2).all() and float(only2.abs().max()) > 0
    (preds["seg_logit"].square().mean() + preds["seg_logit2"].square().mean()).backward()
    (ref["seg_logit"].square().mean() + ref["seg_logit2"].square().mean()).backward()
    for name in ("linear.weight", "linear2.weight", "linear.bias", "linear2.bias"):
        l2, cos = _rel_l2_cos(dict(model.named_parameters())[name].grad, oracle.heads[name].grad)
        assert l2 < 0.1 and cos > 0.99, (name, l2, cos)
    l2, cos = _rel_l2_cos(model.net_3d.sparseModel[3].weight.grad, oracle.unet.params["sparseModel.3.weight"].grad)
    assert l2 < 0.1 and cos > 0.99
