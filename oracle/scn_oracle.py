"""CPU ORACLE for the UNetSCN / SparseConvNet hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this package. The product path (`mopa_b200/`) never routes through it and has no CPU fallback.

**PARITY UNPINNED.** The arithmetic of this path lives in `facebookresearch/SparseConvNet`, installed
by the reference un-pinned from git HEAD (`/root/reference/install.sh:1`); its source is not under
`/root/reference`, it is not installable offline, and neither it nor MoPA ships a single golden vector or
test for this path. This file therefore restates SparseConvNet's *published
algorithm* (labelled [UPSTREAM] below, file names from that project's `sparseconvnet/SCN/` tree), anchored on
the reference's only call sites, `mopa/models/scn_unet.py:25-30`, and on the input layout of
`mopa/data/collate.py:182-186`. What pins it instead is `oracle/dense_equiv.py`: an independent
restatement through dense `F.conv3d` / `F.conv_transpose3d` / `F.batch_norm`, checked in tests/test_oracle.py.

Two layers:
  * integer work (voxel ids, rulebooks) in numpy -- exact;
  * floating point (mean pool, per-offset gather -> matmul -> scatter-add, BatchNorm, unpool) as torch-CPU
    ops in float32 or float64, differentiable, so backward comes from autograd over the restated forward
    ([UPSTREAM] CPU/Convolution.cpp computes exactly those products: dW[k] = rows^T dOut, dIn += dOut W[k]^T).
"""
import numpy as np
import torch

# ----------------------------------------------------------------------------------------------------------
# integer part: grids and rulebooks
# ----------------------------------------------------------------------------------------------------------


def _keys(coords4):
    """(V, 4) int64 [x, y, z, b] -> one int64 per site (16 bits per axis; oracle-private packing)."""
    c = np.asarray(coords4, np.int64)
    return (c[:, 3] << 48) | (c[:, 0] << 32) | (c[:, 1] << 16) | c[:, 2]


def with_batch_column(coords):
    """(N, 3) -> (N, 4) with batch 0; (N, 4) unchanged. scn_unet.py callers pass either (xmuda_arch.py:171)."""
    coords = np.asarray(coords, np.int64)
    if coords.shape[1] == 3:
        coords = np.concatenate([coords, np.zeros((coords.shape[0], 1), np.int64)], 1)
    return coords


def input_layer_rules(coords):
    """[UPSTREAM] Metadata/IOLayersRules.h::inputLayerRules, mode 4 (scn_unet.py:26).

    Scan rows in order; a site's first occurrence gets the next voxel id. Returns
      voxel_coords (V, 4) int64 in id order, p2v (N,) int32, csr_off (V+1,) int32,
      csr_rows (N,) int32 (rows of each voxel ascending).
    """
    coords = with_batch_column(coords)
    n = coords.shape[0]
    if n == 0:
        return coords.reshape(0, 4), np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32)
    keys = _keys(coords)
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")  # unique sites ordered by first occurrence
    rank = np.empty(order.size, np.int64)
    rank[order] = np.arange(order.size)
    p2v = rank[inv.reshape(-1)].astype(np.int32)
    voxel_coords = coords[first[order]]
    csr_rows = np.argsort(p2v, kind="stable").astype(np.int32)
    counts = np.bincount(p2v, minlength=order.size)
    csr_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return voxel_coords, p2v, csr_off, csr_rows


def _lookup(sorted_keys, sorted_ids, query):
    pos = np.searchsorted(sorted_keys, query)
    pos = np.minimum(pos, sorted_keys.size - 1)
    hit = sorted_keys[pos] == query
    return np.where(hit, sorted_ids[pos], -1)


def submanifold_rules(voxel_coords, spatial_size, filter_size=3):
    """[UPSTREAM] Metadata/SubmanifoldConvolutionRules.h. Returns the dense neighbour table
    nbr (K, V) int32 (-1 = inactive) with nbr[k, o] = id of the site at coord(o) + delta_k,
    k = (dx+1)*9 + (dy+1)*3 + (dz+1) for filter 3 (last axis fastest)."""
    vc = np.asarray(voxel_coords, np.int64)
    v = vc.shape[0]
    f = filter_size
    nbr = np.full((f ** 3, v), -1, np.int32)
    if v == 0:
        return nbr
    keys = _keys(vc)
    order = np.argsort(keys)
    sk, sid = keys[order], order.astype(np.int64)
    h = f // 2
    k = 0
    for dx in range(-h, f - h):
        for dy in range(-h, f - h):
            for dz in range(-h, f - h):
                q = vc.copy()
                q[:, 0] += dx
                q[:, 1] += dy
                q[:, 2] += dz
                ok = ((q[:, :3] >= 0) & (q[:, :3] < spatial_size)).all(1)
                ids = _lookup(sk, sid, _keys(q))
                nbr[k] = np.where(ok, ids, -1)
                k += 1
    return nbr


def strided_rules(fine_coords, filter_size=2, stride=2):
    """[UPSTREAM] Metadata/ConvolutionRules.h for size == stride == 2 (the only strided shape scn.UNet builds).

    parent = coord >> 1 per axis, k = (x&1)*4 + (y&1)*2 + (z&1). Coarse ids are canonical: first occurrence
    when scanning fine ids ascending (upstream's numbering follows dense_hash_map iteration order and carries no
    semantics -- SURVEY appendix A.2). Returns coarse_coords (Vc, 4), parent (Vf,) int32, kidx (Vf,) int32.
    """
    assert filter_size == 2 and stride == 2, "only the k2/s2 shape of scn.UNet is restated"
    fc = np.asarray(fine_coords, np.int64)
    pc = fc.copy()
    pc[:, :3] >>= 1
    kidx = ((fc[:, 0] & 1) * 4 + (fc[:, 1] & 1) * 2 + (fc[:, 2] & 1)).astype(np.int32)
    coarse_coords, parent, _, _ = input_layer_rules(pc)
    return coarse_coords, parent, kidx


def child_table(parent, kidx, n_coarse, filter_volume=8):
    """Dense (K, Vc) table: child[k, p] = fine id whose parent is p at filter position k, or -1."""
    tab = np.full((filter_volume, n_coarse), -1, np.int32)
    tab[kidx, parent] = np.arange(parent.shape[0], dtype=np.int32)
    return tab


def table_to_rulebook(table):
    """Dense (K, Vout) table -> SparseConvNet-style rulebook: list over k of (R_k, 2) int32 [in, out],
    canonical order = ascending out row."""
    rules = []
    for k in range(table.shape[0]):
        out = np.nonzero(table[k] >= 0)[0].astype(np.int32)
        rules.append(np.stack([table[k, out], out], 1))
    return rules


def strided_rulebook(parent, kidx, filter_volume=8):
    """Rulebook of a Convolution(k2, s2): rules[k] = [(fine, coarse)], ascending coarse row."""
    rules = []
    fine = np.arange(parent.shape[0], dtype=np.int32)
    for k in range(filter_volume):
        sel = fine[kidx == k]
        o = np.argsort(parent[sel], kind="stable")
        rules.append(np.stack([sel[o], parent[sel][o]], 1).astype(np.int32))
    return rules


class Geometry:
    """What [UPSTREAM] Metadata<3> holds for one forward: per-level grids and rulebooks."""

    def __init__(self, coords, spatial_size=4096):
        self.spatial_size = int(spatial_size)
        self.n_points = int(np.asarray(coords).shape[0])
        vc, self.p2v, self.csr_off, self.csr_rows = input_layer_rules(coords)
        self.level_coords = [vc]
        self.subm = {}
        self.down = {}

    def n_active(self, level):
        return self.level_coords[level].shape[0]

    def subm_table(self, level):
        if level not in self.subm:
            self.subm[level] = submanifold_rules(self.level_coords[level], self.spatial_size >> level)
        return self.subm[level]

    def down_rules(self, level):
        """level -> level+1; returns (parent, kidx)."""
        if level not in self.down:
            cc, parent, kidx = strided_rules(self.level_coords[level])
            assert len(self.level_coords) == level + 1
            self.level_coords.append(cc)
            self.down[level] = (parent, kidx)
        return self.down[level]


# ----------------------------------------------------------------------------------------------------------
# floating-point part (torch-CPU, differentiable)
# ----------------------------------------------------------------------------------------------------------


def input_layer_forward(geo, feats):
    """[UPSTREAM] CPU/IOLayers.cpp InputLayer mode 4: out[v] = sum_i (1/n_v) * in[i] (multiply, then add, rows in order)."""
    n = geo.n_points
    feats = feats[:n]  # rows beyond coords.shape[0] are tolerated (nuscenes_dataloader.py:426 quirk)
    v = geo.n_active(0)
    counts = torch.from_numpy(np.diff(geo.csr_off).astype(np.int64))
    inv = (1.0 / counts.to(feats.dtype))
    p2v = torch.from_numpy(geo.p2v.astype(np.int64))
    out = torch.zeros(v, feats.shape[1], dtype=feats.dtype)
    # index_add_ on CPU accumulates in ascending row order, i.e. the upstream order.
    return out.index_add(0, p2v, feats * inv[p2v].unsqueeze(1))


def output_layer_forward(geo, vox_feats):
    """[UPSTREAM] CPU/IOLayers.cpp OutputLayer: out[i] = in[voxel(i)] (copy; backward sums)."""
    return vox_feats.index_select(0, torch.from_numpy(geo.p2v.astype(np.int64)))


def conv_from_table(x, table, weight, n_out_rows):
    """[UPSTREAM] CPU/Convolution.cpp: out = 0; for k ascending with non-empty rules:
    out[rules_k.out] += x[rules_k.in] @ W[k]. `table` is the dense (K, Vout) in-row table; `weight` is
    (K, 1, nIn, nOut) or (K, nIn, nOut)."""
    w = weight.reshape(weight.shape[0], weight.shape[-2], weight.shape[-1])
    out = torch.zeros(n_out_rows, w.shape[2], dtype=x.dtype)
    for k in range(table.shape[0]):
        o = np.nonzero(table[k] >= 0)[0]
        if o.size == 0:
            continue
        i = torch.from_numpy(table[k, o].astype(np.int64))
        out = out.index_add(0, torch.from_numpy(o.astype(np.int64)), x.index_select(0, i) @ w[k])
    return out


def submanifold_conv(geo, level, x, weight):
    return conv_from_table(x, geo.subm_table(level), weight, geo.n_active(level))


def strided_conv(geo, level, x, weight):
    """Convolution(k2, s2) level -> level+1: out[p] = sum_{children c} x[c] @ W[k(c)]."""
    parent, kidx = geo.down_rules(level)
    return conv_from_table(x, child_table(parent, kidx, geo.n_active(level + 1)), weight, geo.n_active(level + 1))


def strided_deconv(geo, level, x_coarse, weight):
    """Deconvolution(k2, s2) level+1 -> level on the existing fine grid: out[c] = x[parent(c)] @ Wd[k(c)]."""
    parent, kidx = geo.down_rules(level)
    w = weight.reshape(weight.shape[0], weight.shape[-2], weight.shape[-1])
    n_fine = parent.shape[0]
    out = torch.zeros(n_fine, w.shape[2], dtype=x_coarse.dtype)
    for k in range(w.shape[0]):
        c = np.nonzero(kidx == k)[0]
        if c.size == 0:
            continue
        p = torch.from_numpy(parent[c].astype(np.int64))
        out = out.index_add(0, torch.from_numpy(c.astype(np.int64)), x_coarse.index_select(0, p) @ w[k])
    return out


def batchnorm_leakyrelu(x, weight, bias, running_mean, running_var, train, eps=1e-4, momentum=0.9, leakiness=0.0):
    """[UPSTREAM] CPU/BatchNormalization.cpp. `momentum` is the KEEP fraction.
    Updates running stats in place when train. Returns y."""
    n = x.shape[0]
    if train:
        mean = x.sum(0) / n
        d = x - mean
        sq = (d * d).sum(0)
        invstd = (sq / n + eps) ** -0.5
        with torch.no_grad():
            running_mean.mul_(momentum).add_((1 - momentum) * mean.detach())
            running_var.mul_(momentum).add_((1 - momentum) * sq.detach() / max(n - 1, 1))
    else:
        mean = running_mean
        invstd = (running_var + eps) ** -0.5
    w = invstd * weight
    b = bias - mean * w
    y = x * w + b
    return torch.where(y > 0, y, y * leakiness)


# ----------------------------------------------------------------------------------------------------------
# UNetSCN as the reference builds it (mopa/models/scn_unet.py:23-30 + [UPSTREAM] networkArchitectures.py::UNet)
# ----------------------------------------------------------------------------------------------------------


class OracleUNetSCN:
    """Functional restatement driven by a state dict with SparseConvNet's key layout
    (`sparseModel.1.weight`, `sparseModel.2.0.0.running_mean`, ...).

    reps=1, residual_blocks=False (the MoPA config, mopa/config/xmuda.py:217-224)."""

    def __init__(self, state, m=16, num_planes=7, full_scale=4096, dtype=torch.float32, prefix="sparseModel."):
        self.planes = [(i + 1) * m for i in range(num_planes)]
        self.full_scale = full_scale
        self.dtype = dtype
        self.prefix = prefix
        self.params = {k: v.detach().clone().to(dtype) for k, v in state.items()}
        for k, v in self.params.items():
            if not ("running_" in k):
                v.requires_grad_(True)

    def p(self, name):
        return self.params[self.prefix + name]

    def _bn(self, x, path, train):
        return batchnorm_leakyrelu(x, self.p(path + ".weight"), self.p(path + ".bias"),
                                   self.p(path + ".running_mean"), self.p(path + ".running_var"), train)

    def _u(self, geo, level, x, path, train, taps):
        # block: Sequential(BN, Subm) at <path>.0
        x = self._bn(x, path + ".0.0", train)
        x = submanifold_conv(geo, level, x, self.p(path + ".0.1.weight"))
        taps.append((path + ".0.1", x))
        if level + 1 < len(self.planes):
            q = path + ".1.1"  # ConcatTable[1] = Sequential(BN, Conv, U, BN, Deconv)
            y = self._bn(x, q + ".0", train)
            y = strided_conv(geo, level, y, self.p(q + ".1.weight"))
            taps.append((q + ".1", y))
            y = self._u(geo, level + 1, y, q + ".2", train, taps)
            y = self._bn(y, q + ".3", train)
            y = strided_deconv(geo, level, y, self.p(q + ".4.weight"))
            taps.append((q + ".4", y))
            x = torch.cat([x, y], 1)  # JoinTable at <path>.2
            x = self._bn(x, path + ".3.0", train)
            x = submanifold_conv(geo, level, x, self.p(path + ".3.1.weight"))
            taps.append((path + ".3.1", x))
        return x

    def forward(self, coords, feats, train=True, geo=None):
        geo = geo or Geometry(coords, self.full_scale)
        self.geo = geo
        taps = []
        x = input_layer_forward(geo, torch.as_tensor(feats).to(self.dtype))
        taps.append(("0", x))
        x = submanifold_conv(geo, 0, x, self.p("1.weight"))
        taps.append(("1", x))
        x = self._u(geo, 0, x, "2", train, taps)
        x = self._bn(x, "3", train)
        taps.append(("3", x))
        out = output_layer_forward(geo, x)
        self.taps = taps
        return out


def make_unet_state(in_channels=1, m=16, num_planes=7, seed=0, dtype=torch.float32, prefix="sparseModel."):
    """Random-init parameters with SparseConvNet's shapes and init:
    conv weight (volume, 1, nIn, nOut) ~ N(0, sqrt(2/(nIn*volume))); BN weight=1, bias=0, mean=0, var=1.
    BN affine terms are perturbed so gradients w.r.t. them are exercised."""
    g = torch.Generator().manual_seed(seed)
    planes = [(i + 1) * m for i in range(num_planes)]
    st = {}

    def conv(name, vol, nin, nout):
        st[prefix + name + ".weight"] = (torch.randn(vol, 1, nin, nout, generator=g, dtype=torch.float64)
                                         * (2.0 / nin / vol) ** 0.5).to(dtype)

    def bn(name, c):
        st[prefix + name + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g, dtype=torch.float64)).to(dtype)
        st[prefix + name + ".bias"] = (0.1 * torch.randn(c, generator=g, dtype=torch.float64)).to(dtype)
        st[prefix + name + ".running_mean"] = torch.zeros(c, dtype=dtype)
        st[prefix + name + ".running_var"] = torch.ones(c, dtype=dtype)

    def u(level, path):
        a = planes[level]
        bn(path + ".0.0", a)
        conv(path + ".0.1", 27, a, a)
        if level + 1 < num_planes:
            b = planes[level + 1]
            q = path + ".1.1"
            bn(q + ".0", a)
            conv(q + ".1", 8, a, b)
            u(level + 1, q + ".2")
            bn(q + ".3", b)
            conv(q + ".4", 8, b, a)
            bn(path + ".3.0", 2 * a)
            conv(path + ".3.1", 27, 2 * a, a)

    conv("1", 27, in_channels, m)
    u(0, "2")
    bn("3", m)
    return st


# ----------------------------------------------------------------------------------------------------------
# UNetSCN_ED: the reference's unrolled encoder / decoder variant (mopa/models/scn_unet.py:38-134, VGG blocks,
# residual_blocks=False, the configuration its own smoke test :222-239 builds with in_channels=3)
# ----------------------------------------------------------------------------------------------------------
_BN_KEYS = ("weight", "bias", "running_mean", "running_var")


def make_ed_state(in_channels=3, m=16, seed=0, dtype=torch.float32):
    """Random-init parameters under UNetSCN_ED's attribute names (`main_block2.1.0.weight`, `deconv6.3.weight`, ...)."""
    g = torch.Generator().manual_seed(seed)
    st = {}

    def conv(name, vol, nin, nout):
        st[name + ".weight"] = (torch.randn(vol, 1, nin, nout, generator=g, dtype=torch.float64) * (2.0 / nin / vol) ** 0.5).to(dtype)

    def bn(name, c):
        st[name + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g, dtype=torch.float64)).to(dtype)
        st[name + ".bias"] = (0.1 * torch.randn(c, generator=g, dtype=torch.float64)).to(dtype)
        st[name + ".running_mean"] = torch.zeros(c, dtype=dtype)
        st[name + ".running_var"] = torch.ones(c, dtype=dtype)

    conv("down_in", 27, in_channels, m)
    bn("main_block1.0", m)
    conv("main_block1.1", 27, m, m)
    for l in range(2, 8):  # scn_unet.py:197-207: BN(nPlanes), Sequential(Convolution k2 s2, BN(n), Submanifold n -> n)
        a, b = (l - 1) * m, l * m
        bn("main_block%d.0" % l, a)
        conv("main_block%d.1.0" % l, 8, a, b)
        bn("main_block%d.1.1" % l, b)
        conv("main_block%d.1.2" % l, 27, b, b)
    bn("deconv7.0", 7 * m)
    conv("deconv7.1", 8, 7 * m, 6 * m)
    for l in range(6, 1, -1):  # decoder(2 l m, (l - 1) m), scn_unet.py:136-144
        a, b = 2 * l * m, (l - 1) * m
        bn("deconv%d.0" % l, a)
        conv("deconv%d.1" % l, 27, a, a // 2)
        bn("deconv%d.2" % l, a // 2)
        conv("deconv%d.3" % l, 8, a // 2, b)
    bn("deconv1.0", 2 * m)
    conv("deconv1.1", 27, 2 * m, m)
    bn("output.0", m)
    return st


class OracleUNetSCN_ED:
    """Functional restatement of UNetSCN_ED.forward (scn_unet.py:97-133) from the oracle's layer primitives."""

    def __init__(self, state, m=16, full_scale=4096, dtype=torch.float64):
        self.m, self.full_scale, self.dtype = m, full_scale, dtype
        self.params = {k: v.detach().clone().to(dtype) for k, v in state.items()}
        for k, v in self.params.items():
            if "running_" not in k:
                v.requires_grad_(True)

    def _bn(self, x, name, train):
        p = self.params
        return batchnorm_leakyrelu(x, p[name + ".weight"], p[name + ".bias"], p[name + ".running_mean"],
                                   p[name + ".running_var"], train)

    def forward(self, coords, feats, train=True):
        p = self.params
        geo = self.geo = Geometry(coords, self.full_scale)
        x = input_layer_forward(geo, torch.as_tensor(feats).to(self.dtype))
        x = submanifold_conv(geo, 0, x, p["down_in.weight"])
        feat = [None] * 8
        feat[1] = submanifold_conv(geo, 0, self._bn(x, "main_block1.0", train), p["main_block1.1.weight"])
        for l in range(2, 8):  # feature_l lives on level l - 1
            y = self._bn(feat[l - 1], "main_block%d.0" % l, train)
            y = strided_conv(geo, l - 2, y, p["main_block%d.1.0.weight" % l])
            y = self._bn(y, "main_block%d.1.1" % l, train)
            feat[l] = submanifold_conv(geo, l - 1, y, p["main_block%d.1.2.weight" % l])
        d = strided_deconv(geo, 5, self._bn(feat[7], "deconv7.0", train), p["deconv7.1.weight"])
        d = torch.cat([feat[6], d], 1)
        for l in range(6, 1, -1):  # d: joined planes on level l - 1
            y = submanifold_conv(geo, l - 1, self._bn(d, "deconv%d.0" % l, train), p["deconv%d.1.weight" % l])
            y = strided_deconv(geo, l - 2, self._bn(y, "deconv%d.2" % l, train), p["deconv%d.3.weight" % l])
            d = torch.cat([feat[l - 1], y], 1)
        d = submanifold_conv(geo, 0, self._bn(d, "deconv1.0", train), p["deconv1.1.weight"])
        d = self._bn(d, "output.0", train)
        return output_layer_forward(geo, d)
