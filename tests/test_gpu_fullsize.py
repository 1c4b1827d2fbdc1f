"""GPU parity at BASELINE.json sizes: full nuScenes-shaped (N32, ~32k points) and SemanticKITTI-shaped
(K64, ~126k points) scans through the whole UNetSCN against the float64 oracle, every parameter gradient included, and
every rulebook of a batch-8 step bit-exact (all 7 submanifold levels, all 6 strided links) -- not checksums.

Gradient bars come from the measured table profiles/r02_grad_errors.txt (tools/grad_error_table.py): per parameter tensor,
fp32 oracle / GPU fp32 mode / GPU tf32 mode against the float64 oracle; the bars here are <= 3x the worst measured value.
"""
import numpy as np
import pytest
import torch

from mopa_b200 import synth
from oracle import scn_oracle as so
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

# (forward max-abs / max-magnitude, gradient rel-L2, gradient cosine) per precision mode; see module docstring
# measured worst case over the 78 tensors (profiles/r02_grad_errors*.txt): fp32 mode forward 1.3e-5, rel-L2 9.9e-3, cos 0.99995;
# tf32 mode forward 1.2e-3, rel-L2 0.132, cos 0.9915 (the float32 ORACLE itself: 6.3e-4)
BARS = {"fp32": (5e-5, 3e-2, 0.9998), "tf32": (5e-3, 0.27, 0.98)}


@pytest.fixture(scope="module")
def scn(cuda):
    import mopa_b200.scn as scn
    keep = scn.get_precision()
    yield scn
    scn.set_precision(keep)


def _rel_l2_cos(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm()), float(torch.dot(a, b) / (a.norm() * b.norm()))


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("sensor", ["nuscenes", "kitti"])
def test_full_scan_whole_net_forward_and_all_gradients_vs_fp64_oracle(scn, sensor, precision):
    """BASELINE configs 1 and 4 at their real size: ONE full scan (BatchNorm over one scan, as the CPU reference arm),
    forward features, running statistics and all 78 parameter gradients against the float64 oracle."""
    from mopa_b200.unet_scn import UNetSCN
    scn.set_precision(precision)
    coords, feats = synth.make_scan(sensor, seed=3)
    assert coords.shape[0] > (100000 if sensor == "kitti" else 30000)
    state = so.make_unet_state(seed=11)
    net = UNetSCN(1).cuda()
    net.load_state_dict(state)
    out = net([torch.from_numpy(coords), torch.from_numpy(feats).cuda()])
    oracle = so.OracleUNetSCN(state, dtype=torch.float64)
    ref = oracle.forward(coords, feats)
    tol_f, tol_l2, tol_cos = BARS[precision]
    assert out.shape == ref.shape == (coords.shape[0], 16)
    assert rel_err(out, ref) < tol_f
    g = torch.randn(ref.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    out.backward(g.float().cuda())
    ref.backward(g)
    worst = (0.0, 1.0, "")
    for name, p in net.named_parameters():
        l2, cos = _rel_l2_cos(p.grad, oracle.params[name].grad)
        if l2 > worst[0]:
            worst = (l2, cos, name)
        assert l2 < tol_l2 and cos > tol_cos, (name, l2, cos)
    for name, buf in net.named_buffers():
        assert rel_err(buf, oracle.params[name]) < tol_f, name
    print("%s %s: forward %.2e, worst gradient %s rel-L2 %.2e cos %.6f" % (sensor, precision, rel_err(out, ref), worst[2],
                                                                         worst[0], worst[1]))


@pytest.mark.parametrize("sensor,batch", [("nuscenes", 8), ("kitti", 2)])
def test_every_rulebook_of_a_full_batch_is_bit_exact(scn, sensor, batch):
    """All 7 submanifold rulebooks (27 offset lists each) and all 6 strided rulebooks + coarse voxel coordinates of a
    BASELINE-size batch, element for element against the oracle."""
    coords, feats = synth.make_batch(batch, sensor, 0)
    layer = scn.InputLayer(3, 4096, mode=4)
    x = layer([torch.from_numpy(coords), torch.from_numpy(feats).cuda()])
    m = x.metadata
    geo = so.Geometry(coords)
    assert np.array_equal(m.point_to_voxel().numpy(), geo.p2v)
    assert np.array_equal(m.spatial_locations(4096).numpy(), geo.level_coords[0])
    size = 4096
    for level in range(7):
        assert m.prepare_submanifold(size, 3) == geo.n_active(level)
        ref = so.table_to_rulebook(geo.subm_table(level))
        got = m.submanifold_rulebook(size)
        for k in range(27):
            assert np.array_equal(got[k].numpy(), ref[k]), (level, k)
        geo.subm.pop(level)  # free the oracle's dense table (27 x V int32)
        if level == 6:
            break
        parent, kidx = geo.down_rules(level)
        assert m.prepare_convolution(size, size // 2, 2, 2) == geo.n_active(level + 1)
        assert np.array_equal(m.spatial_locations(size // 2).numpy(), geo.level_coords[level + 1])
        ref = so.strided_rulebook(parent, kidx)
        got = m.convolution_rulebook(size)
        for k in range(8):
            assert np.array_equal(got[k].numpy(), ref[k]), (level, k)
        size //= 2


def test_tile_rulebooks_agree_with_the_dense_tables(scn):
    """The per-(128-row tile, offset) compact lists + row masks the tcgen05 conv kernel consumes (geometry.cu::k_tile_lists_batch)
    against the dense neighbour / child / parent tables they are built from, through the inspection entry point."""
    coords, feats = synth.make_batch(2, "nuscenes", 5, n_azimuth=400)
    x = scn.InputLayer(3, 4096, mode=4)([torch.from_numpy(coords), torch.from_numpy(feats).cuda()])
    m = x.metadata
    geo = so.Geometry(coords)
    m.prepare_submanifold(4096, 3)
    m.prepare_convolution(4096, 2048, 2, 2)
    parent, kidx = geo.down_rules(0)
    tables = {
        "subm": geo.subm_table(0),
        "child": so.child_table(parent, kidx, geo.n_active(1)),
        "select": np.stack([np.where(kidx == k, parent, -1) for k in range(8)]).astype(np.int32),
    }
    for kind, table in tables.items():
        lists, masks = m.tile_rulebook(4096, kind)
        K, V = table.shape
        tiles = (V + 127) // 128
        assert lists.shape == (tiles, K, 128) and masks.shape == (tiles, K, 4)
        pad = np.full((K, tiles * 128), -1, np.int32)
        pad[:, :V] = table
        live = pad.reshape(K, tiles, 128).transpose(1, 0, 2) >= 0  # (tiles, K, 128)
        bits = (masks.numpy().astype(np.uint32)[..., None] >> np.arange(32, dtype=np.uint32)) & 1  # (tiles, K, 4, 32)
        assert np.array_equal(bits.reshape(tiles, K, 128).astype(bool), live), kind
        src = pad.reshape(K, tiles, 128).transpose(1, 0, 2)
        ln = lists.numpy()
        for t in range(0, tiles, max(1, tiles // 40)):  # a spread of tiles, every offset
            for k in range(K):
                rows = np.nonzero(live[t, k])[0]
                want = (src[t, k, rows].astype(np.int64) | (rows.astype(np.int64) << 25)).astype(np.int32)
                assert np.array_equal(ln[t, k, :rows.size], want), (kind, t, k)
