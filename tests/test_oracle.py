"""CPU tests: the oracle against its independent arbiters (dense conv3d equivalence, the C hash-map restatement,
float64) -- SURVEY.md 8(c). No GPU, no product code."""
import numpy as np
import pytest
import torch

from oracle import c_rules, dense_equiv as de, scn_oracle as so
from tests.helpers import random_cloud, small_batch


def test_input_rules_first_occurrence():
    coords = np.array([[5, 5, 5, 0], [1, 2, 3, 0], [5, 5, 5, 0], [1, 2, 3, 1], [1, 2, 3, 0]], np.int64)
    vc, p2v, off, rows = so.input_layer_rules(coords)
    assert p2v.tolist() == [0, 1, 0, 2, 1]
    assert vc.tolist() == [[5, 5, 5, 0], [1, 2, 3, 0], [1, 2, 3, 1]]
    assert off.tolist() == [0, 2, 4, 5] and rows.tolist() == [0, 2, 1, 4, 3]


def test_three_column_coords_get_batch_zero():
    coords = np.array([[1, 1, 1], [2, 2, 2], [1, 1, 1]], np.int64)
    vc, p2v, _, _ = so.input_layer_rules(coords)
    assert vc[:, 3].tolist() == [0, 0] and p2v.tolist() == [0, 1, 0]


@pytest.mark.parametrize("seed", [0, 1])
def test_c_and_numpy_rule_builders_agree(seed):
    coords, _ = small_batch(2, 150, seed)
    g, c = so.Geometry(coords), c_rules.CGeometry(coords)
    assert (g.p2v == c.p2v).all() and (g.level_coords[0] == c.level_coords[0]).all()
    for level in range(4):
        assert (g.subm_table(level) == c.subm_table(level)).all()
        (p, k), (pc, kc) = g.down_rules(level), c.down_rules(level)
        assert (p == pc).all() and (k == kc).all()
        assert (g.level_coords[level + 1] == c.level_coords[level + 1]).all()


def test_submanifold_symmetry_and_borders():
    coords = random_cloud(400, 6, 3, n_batch=2)
    coords[:5, :3] = 0
    vc, _, _, _ = so.input_layer_rules(coords)
    nbr = so.submanifold_rules(vc, 6)
    assert (nbr[13] == np.arange(vc.shape[0])).all()
    for k in range(27):  # rule (i, o, k) <-> rule (o, i, 26 - k): what the GPU d_input pass relies on
        o = np.nonzero(nbr[k] >= 0)[0]
        assert (nbr[26 - k][nbr[k][o]] == o).all()
    # no cross-sample neighbours
    for k in range(27):
        o = np.nonzero(nbr[k] >= 0)[0]
        assert (vc[o, 3] == vc[nbr[k][o], 3]).all()


@pytest.mark.parametrize("cin,cout", [(1, 4), (3, 5)])
def test_subm_conv_equals_dense_conv3d(cin, cout):
    torch.manual_seed(0)
    coords = random_cloud(300, 8, 1, n_batch=2)
    geo = so.Geometry(coords, 8)
    x = torch.randn(geo.n_active(0), cin, dtype=torch.float64)
    w = torch.randn(27, 1, cin, cout, dtype=torch.float64)
    sparse = so.submanifold_conv(geo, 0, x, w)
    dense = de.subm_conv_dense(geo.level_coords[0], x, w, 8, 2)
    assert torch.allclose(sparse, dense, atol=1e-12)


def test_strided_conv_and_deconv_equal_dense():
    torch.manual_seed(1)
    coords = random_cloud(200, 8, 2, n_batch=2)
    geo = so.Geometry(coords, 8)
    x = torch.randn(geo.n_active(0), 3, dtype=torch.float64)
    w = torch.randn(8, 1, 3, 5, dtype=torch.float64)
    y = so.strided_conv(geo, 0, x, w)
    dense = de.strided_conv_dense(geo.level_coords[0], x, w, geo.level_coords[1], 8, 2)
    assert torch.allclose(y, dense, atol=1e-12)
    wd = torch.randn(8, 1, 5, 3, dtype=torch.float64)
    z = so.strided_deconv(geo, 0, y, wd)
    dense = de.strided_deconv_dense(geo.level_coords[1], y, wd, geo.level_coords[0], 4, 2)
    assert torch.allclose(z, dense, atol=1e-12)


@pytest.mark.parametrize("train", [True, False])
def test_batchnorm_equals_torch(train):
    torch.manual_seed(2)
    x = torch.randn(500, 6, dtype=torch.float64) * 3 + 1
    w, b = torch.randn(6, dtype=torch.float64), torch.randn(6, dtype=torch.float64)
    rm, rv = torch.randn(6, dtype=torch.float64), torch.rand(6, dtype=torch.float64) + 0.5
    rm2, rv2 = rm.clone(), rv.clone()
    y = so.batchnorm_leakyrelu(x, w, b, rm, rv, train, leakiness=0.0)
    y2 = de.bn_relu_dense(x, w, b, rm2, rv2, train, 0.0)
    assert torch.allclose(y, y2, atol=1e-10)
    assert torch.allclose(rm, rm2, atol=1e-12) and torch.allclose(rv, rv2, atol=1e-12)


def test_input_output_layers():
    coords = np.array([[1, 1, 1, 0], [2, 2, 2, 0], [1, 1, 1, 0], [1, 1, 1, 0]], np.int64)
    geo = so.Geometry(coords)
    f = torch.tensor([[1.0], [10.0], [2.0], [6.0]], dtype=torch.float64)
    v = so.input_layer_forward(geo, f)
    assert torch.allclose(v, torch.tensor([[3.0], [10.0]], dtype=torch.float64))
    out = so.output_layer_forward(geo, v)
    assert out[:, 0].tolist() == [3.0, 10.0, 3.0, 3.0]


def test_unet_float32_tracks_float64_and_grads_flow():
    coords, feats = small_batch(2, 60, 0)
    st = so.make_unet_state(seed=3)
    n32, n64 = so.OracleUNetSCN(st), so.OracleUNetSCN(st, dtype=torch.float64)
    o32, o64 = n32.forward(coords, feats), n64.forward(coords, feats)
    assert o32.shape == (coords.shape[0], 16)
    assert float((o32.double() - o64).abs().max()) < 1e-3 * float(o64.abs().max())
    o64.square().sum().backward()
    for k, v in n64.params.items():
        if "running" not in k:
            assert v.grad is not None and torch.isfinite(v.grad).all(), k
    # running stats moved
    assert float(n64.params["sparseModel.3.running_mean"].abs().max()) > 0


def test_unet_is_translation_invariant_for_multiples_of_64():
    coords, feats = small_batch(1, 60, 1)
    st = so.make_unet_state(seed=4)
    a = so.OracleUNetSCN(st).forward(coords, feats)
    shifted = coords.copy()
    shifted[:, :3] += np.array([64, 128, 192]) - (coords[:, :3].min(0) // 64) * 64
    b = so.OracleUNetSCN(st).forward(shifted, feats)
    assert torch.equal(a, b)


def test_golden_fixture_reproduces_on_cpu():
    """The committed fixture is what the oracle computes today (guards against silent oracle drift)."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "unet_small.npz"))
    st = so.make_unet_state(seed=int(z["state_seed"]))
    net = so.OracleUNetSCN(st, dtype=torch.float64)
    out = net.forward(z["coords"], z["feats"])
    assert np.array_equal(net.geo.p2v, z["p2v"])
    assert [net.geo.n_active(l) for l in range(7)] == z["n_active"].tolist()
    assert np.allclose(out.detach().numpy(), z["out"], rtol=0, atol=1e-9)
    out.backward(torch.from_numpy(z["grad_out"]).double())
    assert np.allclose(net.params["sparseModel.1.weight"].grad.numpy(), z["grad_w1"], rtol=1e-9, atol=1e-9)
    # float32 oracle within the fp32 bar of the GPU tests
    o32 = so.OracleUNetSCN(st).forward(z["coords"], z["feats"])
    assert float((o32.double() - out).abs().max() / out.abs().max()) < 5e-4


# ------------------------------------------------------------------------------------------------ properties (SURVEY 8(c)(3,4))
def test_layer_gradients_pass_gradcheck_in_float64():
    """torch.autograd.gradcheck of the oracle's conv / strided conv / deconv / BatchNorm on a tiny grid (float64)."""
    coords = random_cloud(40, 6, 5, n_batch=1, dup_frac=0.0)
    geo = so.Geometry(coords, 8)
    v0 = geo.n_active(0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(v0, 2, dtype=torch.float64, generator=g, requires_grad=True)
    w = torch.randn(27, 1, 2, 3, dtype=torch.float64, generator=g, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b: so.submanifold_conv(geo, 0, a, b), (x, w), eps=1e-6, atol=1e-6)
    wd = torch.randn(8, 1, 2, 3, dtype=torch.float64, generator=g, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b: so.strided_conv(geo, 0, a, b), (x, wd), eps=1e-6, atol=1e-6)
    xc = torch.randn(geo.n_active(1), 3, dtype=torch.float64, generator=g, requires_grad=True)
    wu = torch.randn(8, 1, 3, 2, dtype=torch.float64, generator=g, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b: so.strided_deconv(geo, 0, a, b), (xc, wu), eps=1e-6, atol=1e-6)
    gam = torch.rand(2, dtype=torch.float64, generator=g).add(0.5).requires_grad_(True)
    bet = torch.randn(2, dtype=torch.float64, generator=g, requires_grad=True)

    def bn(a, ga, be):
        return so.batchnorm_leakyrelu(a, ga, be, torch.zeros(2, dtype=torch.float64), torch.ones(2, dtype=torch.float64), True,
                                      leakiness=0.3)
    assert torch.autograd.gradcheck(bn, (x, gam, bet), eps=1e-6, atol=1e-5)


def test_permuting_the_points_permutes_the_outputs():
    """Row i of the output belongs to input point i whatever the order of the points (voxel ids change with the order, the
    per-point result must not). float64: the summation order inside a voxel / over rules changes with the permutation."""
    coords, feats = small_batch(2, 50, 2)
    feats = np.random.default_rng(0).normal(size=feats.shape).astype(np.float32)
    st = so.make_unet_state(seed=5)
    a = so.OracleUNetSCN(st, dtype=torch.float64).forward(coords, feats)
    perm = np.random.default_rng(1).permutation(coords.shape[0])
    b = so.OracleUNetSCN(st, dtype=torch.float64).forward(coords[perm], feats[perm])
    assert float((a[perm] - b).abs().max()) < 1e-9 * float(a.abs().max())


def test_duplicate_points_get_identical_outputs_and_samples_do_not_interact():
    coords, feats = small_batch(2, 50, 3)
    st = so.make_unet_state(seed=6)
    # duplicates: append copies of the first 20 points
    c2 = np.concatenate([coords, coords[:20]], 0)
    f2 = np.concatenate([feats, feats[:20]], 0)
    out = so.OracleUNetSCN(st, dtype=torch.float64).forward(c2, f2, train=False)
    assert torch.equal(out[:20], out[-20:])
    # batch separation (eval mode: BatchNorm uses running statistics, the only cross-sample coupling is gone): a sample's
    # rows do not change when the other sample is replaced
    keep = coords[:, 3] == 0
    other, _ = small_batch(1, 70, 9)
    other = other.copy()
    other[:, 3] = 1
    c3 = np.concatenate([coords[keep], other], 0)
    f3 = np.concatenate([feats[keep], np.ones((other.shape[0], 1), np.float32)], 0)
    a = so.OracleUNetSCN(st, dtype=torch.float64).forward(coords, feats, train=False)[torch.from_numpy(keep)]
    b = so.OracleUNetSCN(st, dtype=torch.float64).forward(c3, f3, train=False)[: int(keep.sum())]
    assert torch.equal(a, b)
