"""mopa_b200.data: the one-step-deep host <-> device pipeline of a training loop, and the optional fusion of the BatchNorm
backward sums into the d_input convolution's epilogue (MOPA_SCN_BNBWD_FUSION=1, off by default)."""
import os

import numpy as np
import pytest
import torch

from mopa_b200 import synth
from oracle import scn_oracle as so

pytestmark = pytest.mark.gpu


def test_prefetcher_hands_out_the_batches_in_order_and_unchanged(cuda):
    from mopa_b200.data import DevicePrefetcher
    host = []
    for seed in range(5):
        c, f = synth.make_batch(1, "nuscenes", seed, n_azimuth=64 + 16 * seed)
        host.append([torch.from_numpy(c), torch.from_numpy(f).pin_memory()])  # coords unpinned on purpose: pinned on the way
    got = list(DevicePrefetcher(iter(host), depth=2))
    assert len(got) == len(host)
    for (c, f), (hc, hf) in zip(got, host):
        assert c.is_cuda and f.is_cuda and c.dtype == torch.int64 and f.dtype == torch.float32
        assert torch.equal(c.cpu(), hc) and torch.equal(f.cpu(), hf)


def test_prefetched_batches_feed_the_network(cuda):
    import mopa_b200.scn as scn
    from mopa_b200.data import DevicePrefetcher, LaggedScalar
    from mopa_b200.unet_scn import UNetSCN
    keep = scn.get_precision()
    scn.set_precision("tf32")
    try:
        torch.manual_seed(0)
        net = UNetSCN(1).cuda()
        host = [[torch.from_numpy(c), torch.from_numpy(f)] for c, f in
                (synth.make_batch(2, "nuscenes", s, n_azimuth=120) for s in range(3))]
        ref = []
        for c, f in host:  # the blocking loop
            net.zero_grad(set_to_none=True)
            loss = net([c, f.cuda()]).square().mean()
            loss.backward()
            ref.append(float(loss))
        lag, seen = LaggedScalar(), []
        for c, f in DevicePrefetcher(iter(host)):
            net.zero_grad(set_to_none=True)
            loss = net([c, f]).square().mean()
            loss.backward()
            seen.append(lag.push(loss))
        seen.append(lag.last())
        assert seen[0] is None
        np.testing.assert_allclose(seen[1:], ref, rtol=1e-5)
    finally:
        scn.set_precision(keep)


def test_bn_backward_sums_from_the_conv_epilogue_match_the_two_kernel_path(cuda):
    """Same network, same inputs, backward with and without MOPA_SCN_BNBWD_FUSION (read per call): all 78 parameter
    gradients agree to accumulation-order noise, and both agree with the float64 oracle within the tf32 bars."""
    import mopa_b200.scn as scn
    from mopa_b200.unet_scn import UNetSCN
    keep = scn.get_precision()
    scn.set_precision("tf32")
    old = os.environ.get("MOPA_SCN_BNBWD_FUSION")
    try:
        coords, feats = synth.make_batch(2, "nuscenes", 7, n_azimuth=500)
        state = so.make_unet_state(seed=3)
        net = UNetSCN(1).cuda()
        net.load_state_dict(state)
        g = torch.randn(coords.shape[0], 16, generator=torch.Generator().manual_seed(2)).cuda()
        grads = {}
        for mode in ("0", "1"):
            os.environ["MOPA_SCN_BNBWD_FUSION"] = mode
            net.zero_grad(set_to_none=True)
            net.train()
            net.load_state_dict(state)  # running statistics back to the start
            out = net([torch.from_numpy(coords), torch.from_numpy(feats).cuda()])
            out.backward(g)
            grads[mode] = {n: p.grad.detach().double().cpu().clone() for n, p in net.named_parameters()}
        for n in grads["0"]:
            a, b = grads["0"][n].flatten(), grads["1"][n].flatten()
            rel = float((a - b).norm() / a.norm().clamp_min(1e-30))
            assert rel < 2e-3, (n, rel)  # measured: <= 3e-4 (fp32 partial sums in a different order)
        oracle = so.OracleUNetSCN(state, dtype=torch.float64)
        ref = oracle.forward(coords, feats)
        ref.backward(g.double().cpu())
        for n, p in net.named_parameters():
            a, b = grads["1"][n].flatten(), oracle.params[n].grad.flatten()
            l2 = float((a - b).norm() / b.norm())
            cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
            assert l2 < 0.27 and cos > 0.98, (n, l2, cos)
    finally:
        scn.set_precision(keep)
        if old is None:
            os.environ.pop("MOPA_SCN_BNBWD_FUSION", None)
        else:
            os.environ["MOPA_SCN_BNBWD_FUSION"] = old


def test_prefetcher_moves_a_nested_data_batch(cuda):
    """A MoPA-style data_batch (dict with 'x': [coords, feats], labels, non-tensor fields): tensors arrive on the device
    carrying the copy's event, everything else is passed through; the coordinates are accepted as complete."""
    import numpy as np
    from mopa_b200.data import DevicePrefetcher
    from mopa_b200.scn.functional import _coords_where
    c, f = synth.make_batch(1, "nuscenes", 2, n_azimuth=80)
    batch = {"x": [torch.from_numpy(c), torch.from_numpy(f)], "seg_label": torch.zeros(c.shape[0], dtype=torch.int64),
             "img_indices": [np.zeros((c.shape[0], 2), np.int64)], "name": "scan-0"}
    (out,) = list(DevicePrefetcher(iter([batch])))
    assert out["x"][0].is_cuda and out["x"][1].is_cuda and out["seg_label"].is_cuda
    assert out["name"] == "scan-0" and out["img_indices"][0] is batch["img_indices"][0]
    assert torch.equal(out["x"][0].cpu(), batch["x"][0])
    assert _coords_where(out["x"][0]) == 2 and _coords_where(out["x"][0].clone()) == 1
