"""GPU numerics of the reference's model classes on the `sparseconvnet` surface: the unrolled UNetSCN_ED
(mopa/models/scn_unet.py:38-134, module-by-module path, 3 input channels as in its smoke test :222-239) against the
float64 oracle; and, when the reference checkout is present next to a GPU, the reference's own files imported unchanged.
(tests/test_surface.py proves on CPU that the mirrors used here have the reference's exact module trees.)"""
import os
import sys

import numpy as np
import pytest
import torch

from mopa_b200 import synth
from oracle import scn_oracle as so
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scn(cuda):
    import mopa_b200.scn as scn
    keep = scn.get_precision()
    yield scn
    scn.set_precision(keep)


def _rel_l2_cos(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm()), float(torch.dot(a, b) / (a.norm() * b.norm()))


def _check_ed(net_cls, precision, scn):
    scn.set_precision(precision)
    coords, _ = synth.make_batch(2, "nuscenes", 6, n_azimuth=220)
    feats = np.random.default_rng(0).uniform(size=(coords.shape[0], 3)).astype(np.float32)  # torch.rand(b * n, 3), :232
    state = so.make_ed_state(3, seed=4)
    net = net_cls(3).cuda()
    net.load_state_dict(state)
    out = net([torch.from_numpy(coords), torch.from_numpy(feats).cuda()])
    oracle = so.OracleUNetSCN_ED(state)
    ref = oracle.forward(coords, feats)
    tol_f, tol_l2, tol_cos = {"fp32": (1e-4, 3e-2, 0.9995), "tf32": (1e-2, 0.2, 0.98)}[precision]
    assert out.shape == ref.shape == (coords.shape[0], 16)
    assert rel_err(out, ref) < tol_f
    g = torch.randn(ref.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    out.backward(g.float().cuda())
    ref.backward(g)
    for name, p in net.named_parameters():
        l2, cos = _rel_l2_cos(p.grad, oracle.params[name].grad)
        assert l2 < tol_l2 and cos > tol_cos, (name, l2, cos)
    for name, buf in net.named_buffers():
        assert rel_err(buf, oracle.params[name]) < tol_f, name


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_unet_scn_ed_forward_backward_matches_oracle(scn, precision):
    from tests.ed_mirror import UNetSCN_ED
    _check_ed(UNetSCN_ED, precision, scn)


def test_reference_files_unchanged_through_the_shim(scn):
    """Only where /root/reference and a GPU coexist (not the round-end box, which has no reference checkout)."""
    ref_root = "/root/reference"
    if not os.path.isdir(os.path.join(ref_root, "mopa", "models")):
        pytest.skip("the reference checkout is not present on this box")
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import importlib
    ref = importlib.import_module("mopa.models.scn_unet")
    _check_ed(ref.UNetSCN_ED, "tf32", scn)
    scn.set_precision("tf32")
    coords, feats = synth.make_batch(2, "nuscenes", 2, n_azimuth=250)
    state = so.make_unet_state(seed=5)
    net = ref.UNetSCN(1).cuda()
    net.load_state_dict(state)
    out = net([torch.from_numpy(coords), torch.from_numpy(feats).cuda()])
    assert rel_err(out, so.OracleUNetSCN(state, dtype=torch.float64).forward(coords, feats)) < 5e-2
