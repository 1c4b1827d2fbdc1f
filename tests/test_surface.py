"""CPU checks of the `sparseconvnet` module surface MoPA imports (mopa/models/scn_unet.py:4,25-30):
constructor signatures, module nesting and therefore state_dict keys / shapes of [UPSTREAM] scn.UNet, so that the
published xMUDA / MoPA 3D checkpoints (`net_3d.sparseModel.*`, checkpoint.py:39-87) load; error behaviour of what the path
does not implement. Modules are plain torch.nn.Modules: building them needs no GPU."""
import copy

import pytest
import torch


def _expected_unet_keys(planes, reps=1):
    """Keys of [UPSTREAM] networkArchitectures.UNet(dimension, reps, nPlanes, residual_blocks=False) restated from its
    construction order: per level `reps` VGG blocks Sequential(BatchNormLeakyReLU, SubmanifoldConvolution); then
    ConcatTable(Identity, Sequential(BNLeakyReLU, Convolution, U(rest), BNLeakyReLU, Deconvolution)), JoinTable, and `reps`
    blocks on the joined planes."""
    bn = ("weight", "bias", "running_mean", "running_var")

    def block(prefix, idx, a, b, out):
        out += [("%s%d.0.%s" % (prefix, idx, k), (a,)) for k in bn]
        out.append(("%s%d.1.weight" % (prefix, idx), (27, 1, a, b)))

    def u(prefix, pl, out):
        idx = 0
        for _ in range(reps):
            block(prefix, idx, pl[0], pl[0], out)
            idx += 1
        if len(pl) > 1:
            inner = "%s%d.1." % (prefix, idx)  # ConcatTable child 1 (child 0 is Identity: no parameters)
            out += [(inner + "0." + k, (pl[0],)) for k in bn]
            out.append((inner + "1.weight", (8, 1, pl[0], pl[1])))
            u(inner + "2.", pl[1:], out)
            out += [(inner + "3." + k, (pl[1],)) for k in bn]
            out.append((inner + "4.weight", (8, 1, pl[1], pl[0])))
            idx += 2  # ConcatTable, JoinTable
            for i in range(reps):
                block(prefix, idx, pl[0] * (2 if i == 0 else 1), pl[0], out)
                idx += 1
        return out

    return u("", planes, [])


def test_unet_scn_state_dict_matches_upstream_naming():
    from mopa_b200.unet_scn import UNetSCN
    net = UNetSCN(1)  # in_channels=1, m=16, 7 levels, reps=1 (xmuda.py:217-224)
    sd = net.state_dict()
    planes = [16 * (i + 1) for i in range(7)]
    want = [("sparseModel.1.weight", (27, 1, 1, 16))]
    want += [("sparseModel.2." + k, shape) for k, shape in _expected_unet_keys(planes)]
    want += [("sparseModel.3." + k, (16,)) for k in ("weight", "bias", "running_mean", "running_var")]
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == want
    assert sum(p.numel() for p in net.parameters()) == 2688656  # ~2.69 M
    # a checkpoint saved from an upstream-shaped model loads strictly, also behind the `net_3d.` prefix MoPA uses
    other = UNetSCN(1)
    other.load_state_dict({k: torch.full_like(v, 0.5) for k, v in sd.items()}, strict=True)
    assert all(bool((v == 0.5).all()) for v in other.state_dict().values())


def test_module_surface_signatures_and_containers():
    import sparseconvnet as scn  # the shim MoPA's `import sparseconvnet as scn` resolves to
    seq = scn.Sequential()
    assert seq.add(scn.InputLayer(3, 4096, mode=4)) is seq  # .add returns self (scn_unet.py:25-30 chains it)
    seq.add(scn.SubmanifoldConvolution(3, 1, 16, 3, False)).add(scn.UNet(3, 1, [16, 32], False))
    seq.add(scn.BatchNormReLU(16)).add(scn.OutputLayer(3))
    assert len(list(seq.children())) == 5
    for cls in ("Convolution", "Deconvolution", "BatchNormLeakyReLU", "JoinTable", "ConcatTable", "AddTable", "Identity",
                "NetworkInNetwork", "SparseConvNetTensor"):
        assert hasattr(scn, cls), cls
    bn = scn.BatchNormLeakyReLU(32, leakiness=0.2)
    assert bn.eps == 1e-4 and bn.momentum == 0.9 and bn.leakiness == 0.2  # upstream defaults (appendix A.3)
    conv = scn.Convolution(3, 16, 32, 2, 2, False)
    assert tuple(conv.weight.shape) == (8, 1, 16, 32)
    assert tuple(scn.Deconvolution(3, 32, 16, 2, 2, False).weight.shape) == (8, 1, 32, 16)
    # train / eval, deepcopy (train_xmuda_mopa.py:78) and in-place .data swaps (torch_ema) work on the module tree
    twin = copy.deepcopy(seq).eval()
    assert not twin.training and seq.training
    with torch.no_grad():
        for p in twin.parameters():
            p.data = p.data * 0
    assert all(float(p.abs().sum()) == 0 for p in twin.parameters())
    assert any(float(p.abs().sum()) > 0 for p in seq.parameters())


def test_unimplemented_features_raise():
    import sparseconvnet as scn
    with pytest.raises(NotImplementedError):
        scn.SubmanifoldConvolution(3, 16, 16, 3, True)  # bias
    with pytest.raises(NotImplementedError):
        scn.Convolution(3, 16, 32, 3, 2, False)  # only size-2 stride-2
    with pytest.raises(NotImplementedError):
        scn.SubmanifoldConvolution(3, 16, 16, 5, False)  # only 3x3x3


def test_cpu_tensors_fail_loudly():
    """No CPU fallback: features on the host raise instead of running somewhere else."""
    import sparseconvnet as scn
    from mopa_b200._lib import ScnError
    with pytest.raises(ScnError):
        scn.InputLayer(3, 4096, mode=4)([torch.zeros(4, 4, dtype=torch.long), torch.ones(4, 1)])


# ---------------------------------------------------------------------------------------------------------------
# The reference's OWN model files through the shim (the drop-in boundary claim). /root/reference exists in
# the build container only (never on the GPU box), so these run on CPU: they prove that mopa/models/scn_unet.py imports and
# constructs unchanged on `import sparseconvnet as scn`, and that the module trees the GPU tests exercise
# (mopa_b200/unet_scn.py, tests/ed_mirror.py) are identical to the reference's: same state_dict keys, shapes, module types
# in traversal order, and the same compiled op list. The numerics of those trees are then checked on the GPU
# (tests/test_gpu_parity.py, tests/test_gpu_reference_models.py).
# ---------------------------------------------------------------------------------------------------------------
REFERENCE = "/root/reference"


def _reference_scn_unet():
    import importlib
    import os
    import sys
    if not os.path.isdir(os.path.join(REFERENCE, "mopa", "models")):
        pytest.skip("the reference checkout is not present on this box")
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    return importlib.import_module("mopa.models.scn_unet")  # its `import sparseconvnet as scn` resolves to the shim


def _tree(net):
    return ([(k, tuple(v.shape)) for k, v in net.state_dict().items()], [type(m).__name__ for m in net.modules()])


def test_reference_unet_scn_builds_through_the_shim_and_equals_the_mirror():
    ref = _reference_scn_unet()
    from mopa_b200.unet_scn import UNetSCN
    from mopa_b200.scn import compiler
    theirs, ours = ref.UNetSCN(1), UNetSCN(1)  # xmuda.py:217-224 defaults
    assert _tree(theirs) == _tree(ours)
    pa, pb = compiler.CompiledProgram(theirs.sparseModel), compiler.CompiledProgram(ours.sparseModel)
    assert pa.ops == pb.ops and pa.bufs == pb.bufs and pa.n_levels == pb.n_levels == 7
    assert len(pa.ops) == 52  # 26 convolutions + 26 BatchNorms in one native program
    ours.load_state_dict(theirs.state_dict(), strict=True)


def test_ed_mirror_matches_the_reference_module_tree():
    ref = _reference_scn_unet()
    from tests.ed_mirror import UNetSCN_ED
    theirs, ours = ref.UNetSCN_ED(3), UNetSCN_ED(3)  # in_channels = 3 as in the reference's own smoke test (:232)
    ka, ma = _tree(theirs)
    kb, mb = _tree(ours)
    assert sorted(ka) == sorted(kb) and ma == mb
    ours.load_state_dict(theirs.state_dict(), strict=True)


def test_oracle_ed_state_uses_the_reference_key_layout():
    from oracle import scn_oracle as so
    from tests.ed_mirror import UNetSCN_ED
    st = so.make_ed_state(3)
    net = UNetSCN_ED(3)
    assert {k: tuple(v.shape) for k, v in st.items()} == {k: tuple(v.shape) for k, v in net.state_dict().items()}


def test_compiled_for_caches_and_respects_eager_switch(monkeypatch):
    """Host logic of the executor selection (no GPU needed): one CompiledProgram per module tree, rebuilt when the tree
    changes, none under MOPA_SCN_EAGER=1 or with mixed train/eval BatchNorms; gradient views are keyed by parameter identity."""
    from mopa_b200.unet_scn import UNetSCN
    from mopa_b200.scn import compiler
    net = UNetSCN(1)
    a = compiler.compiled_for(net.sparseModel)
    assert a is not None and compiler.compiled_for(net.sparseModel) is a
    net.sparseModel[3].eval()
    assert compiler.compiled_for(net.sparseModel) is None  # mixed modes -> module-by-module path
    net.train()
    monkeypatch.setenv("MOPA_SCN_EAGER", "1")
    assert compiler.compiled_for(net.sparseModel) is None
    monkeypatch.delenv("MOPA_SCN_EAGER")
    params = list(net.parameters())
    views = {p: torch.zeros_like(p) for p in params[:3]}
    compiler.register_grad_views(views)
    assert compiler._grad_view_of(params[0]) is views[params[0]] and compiler._grad_view_of(params[5]) is None
    compiler.unregister_grad_views(params[:3])
    assert compiler._grad_view_of(params[0]) is None


def test_arena_sizes_fall_into_coarse_classes():
    """Arena sizes are rounded up to 1/16 of their power of two (compiler._arena_bytes): batches a few per cent apart share a
    size class, the request is never below the real size and never more than 1/16 above it."""
    from mopa_b200.scn.compiler import _arena_bytes
    for n in (1, 1000, (1 << 20) - 1, (1 << 20) + 1, 123456789, 1000000000, 1010000000, 6100000000):
        r = _arena_bytes(n)
        assert r >= n and r <= n + max(1, n // 16) + 1
    assert _arena_bytes(1000000000) == _arena_bytes(1005000000)


def test_host_coordinates_are_passed_as_host_pointers():
    """coords_on_device of the C ABI: 0 for host tensors whatever attributes they carry (the ready-event shortcut is for
    device tensors only)."""
    import torch
    from mopa_b200.scn.functional import _coords_where
    c = torch.zeros(4, 4, dtype=torch.int64)
    c._mopa_ready = object()
    assert _coords_where(c) == 0


def test_map_tensors_walks_a_mopa_data_batch():
    """mopa_b200.data.map_tensors: every tensor of a nested batch (dict / list / tuple, the shape of MoPA's collate output)
    is transformed, everything else is passed through, the structure is preserved."""
    import numpy as np
    import torch
    from mopa_b200.data import map_tensors
    batch = {"x": [torch.zeros(3, 4, dtype=torch.int64), torch.ones(3, 1)], "seg_label": torch.arange(3),
             "img_indices": [np.zeros((3, 2), np.int64)], "id": ("a", 7), "nested": {"t": (torch.ones(2),)}}
    out = map_tensors(batch, lambda t: t + 1)
    assert set(out) == set(batch) and isinstance(out["x"], list) and isinstance(out["nested"]["t"], tuple)
    assert torch.equal(out["x"][0], batch["x"][0] + 1) and torch.equal(out["nested"]["t"][0], torch.full((2,), 2.0))
    assert out["img_indices"][0] is batch["img_indices"][0] and out["id"] == ("a", 7)


def test_arena_pool_keeps_and_bounds_idle_tensors():
    """compiler._ArenaPool (host logic, exercised with CPU tensors standing in for device arenas): a returned tensor is
    handed out again for the same key, at most `keep` idle tensors per key and `max_idle_bytes` in all are kept, tensors
    that did not come from the pool are ignored, release() drops everything."""
    import torch
    from mopa_b200.scn.compiler import _ArenaPool
    pool = _ArenaPool(keep=2, max_idle_bytes=3 << 20)
    pool.enabled = True

    def lease(nbytes):
        t = torch.empty(nbytes, dtype=torch.uint8)
        t._mopa_arena_key = (0, 0, nbytes)
        return t
    a, b, c = lease(1 << 20), lease(1 << 20), lease(1 << 20)
    for t in (a, b, c):
        pool.put(t)
    assert len(pool.idle[(0, 0, 1 << 20)]) == 2 and pool.idle_bytes == 2 << 20  # keep = 2
    big = lease(2 << 20)
    pool.put(big)  # would exceed max_idle_bytes
    assert (0, 0, 2 << 20) not in pool.idle or not pool.idle[(0, 0, 2 << 20)]
    pool.put(torch.empty(8, dtype=torch.uint8))  # foreign tensor: ignored
    assert pool.idle_bytes == 2 << 20
    got = pool.idle[(0, 0, 1 << 20)].pop()
    assert got is b or got is a
    pool.release()
    assert not pool.idle and pool.idle_bytes == 0
