"""The `sparseconvnet` module surface MoPA touches (mopa/models/scn_unet.py:4,25-30,52-216), same constructor
signatures, parameter names/shapes and module nesting as [UPSTREAM] sparseconvnet/*.py so that
`scn_unet.UNetSCN`'s state_dict keys and call convention are unchanged:

    module([coords LongTensor (N, 3|4) on the host, feats float32 (N, C) on the GPU]) -> float32 (N, m) on the GPU
"""
import torch
import torch.nn as nn

from . import functional as F


def _to_size(dimension, x):
    """[UPSTREAM] utils.toLongTensor: int -> LongTensor([x] * dimension)."""
    if isinstance(x, torch.Tensor):
        return x.long().clone()
    if isinstance(x, (list, tuple)):
        return torch.LongTensor(list(x))
    return torch.LongTensor([int(x)] * dimension)


def _cube(size):
    s = [int(v) for v in size.tolist()]
    if any(v != s[0] for v in s):
        raise F._lib.ScnError("only cubic spatial sizes are supported (got %s)" % (s,))
    return s[0]


class SparseConvNetTensor:
    """[UPSTREAM] sparseConvNetTensor.py: (features, metadata, spatial_size)."""

    def __init__(self, features=None, metadata=None, spatial_size=None):
        self.features = features
        self.metadata = metadata
        self.spatial_size = spatial_size

    def get_spatial_locations(self, spatial_size=None):
        return self.metadata.spatial_locations(_cube(self.spatial_size if spatial_size is None else spatial_size))

    def cuda(self):
        self.features = self.features.cuda()
        return self

    def cpu(self):
        self.features = self.features.cpu()
        return self

    def __repr__(self):
        return "SparseConvNetTensor<<features=%s, spatial_size=%s>>" % (
            tuple(self.features.shape) if self.features is not None else None,
            None if self.spatial_size is None else self.spatial_size.tolist())


class Sequential(nn.Sequential):
    """[UPSTREAM] sequential.py: nn.Sequential with .add() chaining (scn_unet.py:25-30)."""

    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def forward(self, input):
        # [coords, feats] entering an InputLayer ... OutputLayer chain: run the compiled program (one native call for
        # the whole forward, one for the backward) when every module is known to the executor; else module by module.
        if isinstance(input, (list, tuple)) and len(self._modules) >= 2 and isinstance(self._modules["0"], InputLayer) \
                and torch.is_tensor(input[1]) and input[1].is_cuda:
            from . import compiler
            prog = compiler.compiled_for(self)
            if prog is not None:
                return compiler.run(prog, self, input)
        for module in self._modules.values():
            input = module(input)
        return input

    def input_spatial_size(self, out_size):
        for m in reversed(self._modules.values()):
            out_size = m.input_spatial_size(out_size)
        return out_size

    def reweight(self, input):
        for module in self._modules.values():
            input = module(input)
        return input


class Identity(nn.Module):
    def forward(self, input):
        return input

    def input_spatial_size(self, out_size):
        return out_size


class ConcatTable(nn.Sequential):
    """[UPSTREAM] tables.py: applies every child to the same input, returns the list."""

    def forward(self, input):
        return [module(input) for module in self._modules.values()]

    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def input_spatial_size(self, out_size):
        return self._modules["0"].input_spatial_size(out_size)


class JoinTable(nn.Module):
    """[UPSTREAM] tables.py: concatenates the feature planes of tensors on the same grid."""

    def forward(self, input):
        out = SparseConvNetTensor(metadata=input[0].metadata, spatial_size=input[0].spatial_size)
        out.features = torch.cat([i.features for i in input], 1) if input[0].features.numel() else input[0].features
        return out

    def input_spatial_size(self, out_size):
        return out_size


class AddTable(nn.Module):
    def forward(self, input):
        out = SparseConvNetTensor(metadata=input[0].metadata, spatial_size=input[0].spatial_size)
        out.features = sum(i.features for i in input)
        return out

    def input_spatial_size(self, out_size):
        return out_size


class InputLayer(nn.Module):
    """[UPSTREAM] ioLayers.py InputLayer(dimension, spatial_size, mode). mode 4 = mean of duplicate points
    (the only mode scn_unet.py:26 uses; others raise). input = [coords, features] or [coords, features, batch_size]."""

    def __init__(self, dimension, spatial_size, mode=3):
        super().__init__()
        self.dimension = dimension
        self.spatial_size = _to_size(dimension, spatial_size)
        self.mode = mode

    def forward(self, input):
        coords, feats = input[0], input[1]
        if not feats.is_cuda:
            raise F._lib.ScnError("InputLayer: features must live on the GPU (mopa_b200.scn has no CPU path)")
        metadata = F.Metadata(self.dimension, feats.device)
        if coords.dim() != 2 or coords.shape[1] not in (self.dimension, self.dimension + 1):
            raise F._lib.ScnError("InputLayer: coords must be (N, %d) or (N, %d)" % (self.dimension, self.dimension + 1))
        n_active = metadata.set_locations(coords, _cube(self.spatial_size), self.mode)
        out = SparseConvNetTensor(metadata=metadata, spatial_size=self.spatial_size)
        # rows of `feats` beyond coords.shape[0] are ignored (nuscenes_dataloader.py:426 emits such tensors)
        out.features = F.InputLayerFunction.apply(feats, metadata, n_active)
        return out

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        return "InputLayer(spatial_size=%s, mode=%d)" % (self.spatial_size.tolist(), self.mode)


class OutputLayer(nn.Module):
    """[UPSTREAM] ioLayers.py OutputLayer(dimension): one output row per input point, in input order."""

    def __init__(self, dimension):
        super().__init__()
        self.dimension = dimension

    def forward(self, input):
        return F.OutputLayerFunction.apply(input.features, input.metadata)

    def input_spatial_size(self, out_size):
        return out_size


class _ConvBase(nn.Module):
    def _init_weight(self, volume, nIn, nOut, bias, groups):
        if groups != 1:
            raise NotImplementedError("groups != 1 is not used by MoPA and not implemented")
        if bias:
            raise NotImplementedError("bias=True is not used by MoPA (scn_unet.py passes bias=False) and not implemented")
        std = (2.0 / nIn / volume) ** 0.5
        self.weight = nn.Parameter(torch.empty(volume, groups, nIn // groups, nOut // groups).normal_(0, std))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # older SparseConvNet checkpoints store (volume, nIn, nOut) without the groups axis
        key = prefix + "weight"
        if key in state_dict and state_dict[key].dim() == 3:
            state_dict[key] = state_dict[key].unsqueeze(1)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class SubmanifoldConvolution(_ConvBase):
    """[UPSTREAM] submanifoldConvolution.py: SubmanifoldConvolution(dimension, nIn, nOut, filter_size, bias, groups=1)."""

    def __init__(self, dimension, nIn, nOut, filter_size, bias, groups=1):
        super().__init__()
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size = _to_size(dimension, filter_size)
        self.filter_volume = int(self.filter_size.prod().item())
        if _cube(self.filter_size) != 3:
            raise NotImplementedError("only 3^d submanifold filters are implemented (the shape scn.UNet builds)")
        self._init_weight(self.filter_volume, nIn, nOut, bias, groups)

    def forward(self, input):
        assert input.features.numel() == 0 or input.features.size(1) == self.nIn, (self.nIn, self.nOut, input)
        size = _cube(input.spatial_size)
        n = input.metadata.prepare_submanifold(size, 3)
        out = SparseConvNetTensor(metadata=input.metadata, spatial_size=input.spatial_size)
        out.features = F.sparse_conv(input.features, self.weight, input.metadata, "subm", (size,), n, 3, 1)
        return out

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        return "SubmanifoldConvolution %d->%d C%d" % (self.nIn, self.nOut, _cube(self.filter_size))


class Convolution(_ConvBase):
    """[UPSTREAM] convolution.py: Convolution(dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1)."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size = _to_size(dimension, filter_size)
        self.filter_stride = _to_size(dimension, filter_stride)
        self.filter_volume = int(self.filter_size.prod().item())
        if _cube(self.filter_size) != 2 or _cube(self.filter_stride) != 2:
            raise NotImplementedError("only size-2 stride-2 convolutions are implemented (the shape scn.UNet builds)")
        self._init_weight(self.filter_volume, nIn, nOut, bias, groups)

    def forward(self, input):
        assert input.features.numel() == 0 or input.features.size(1) == self.nIn
        in_size = _cube(input.spatial_size)
        out_sz = (input.spatial_size - self.filter_size) // self.filter_stride + 1
        assert ((out_sz - 1) * self.filter_stride + self.filter_size == input.spatial_size).all(), \
            "Convolution: (out - 1) * stride + size must equal the input spatial size"
        n = input.metadata.prepare_convolution(in_size, _cube(out_sz), 2, 2)
        out = SparseConvNetTensor(metadata=input.metadata, spatial_size=out_sz)
        out.features = F.sparse_conv(input.features, self.weight, input.metadata, "conv", (in_size, _cube(out_sz)), n, 2, 2)
        return out

    def input_spatial_size(self, out_size):
        return (out_size - 1) * self.filter_stride + self.filter_size

    def __repr__(self):
        return "Convolution %d->%d C%d/%d" % (self.nIn, self.nOut, _cube(self.filter_size), _cube(self.filter_stride))


class Deconvolution(_ConvBase):
    """[UPSTREAM] deconvolution.py: transposed use of the paired Convolution's rulebook, onto the existing finer grid."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size = _to_size(dimension, filter_size)
        self.filter_stride = _to_size(dimension, filter_stride)
        self.filter_volume = int(self.filter_size.prod().item())
        if _cube(self.filter_size) != 2 or _cube(self.filter_stride) != 2:
            raise NotImplementedError("only size-2 stride-2 deconvolutions are implemented (the shape scn.UNet builds)")
        self._init_weight(self.filter_volume, nIn, nOut, bias, groups)

    def forward(self, input):
        assert input.features.numel() == 0 or input.features.size(1) == self.nIn
        in_size = _cube(input.spatial_size)
        out_sz = (input.spatial_size - 1) * self.filter_stride + self.filter_size
        n = input.metadata.n_active(_cube(out_sz))
        if n < 0:
            raise F._lib.ScnError("Deconvolution: no grid at spatial size %d (it must follow the paired Convolution)" % _cube(out_sz))
        out = SparseConvNetTensor(metadata=input.metadata, spatial_size=out_sz)
        out.features = F.sparse_conv(input.features, self.weight, input.metadata, "deconv", (in_size, _cube(out_sz)), n, 2, 2)
        return out

    def input_spatial_size(self, out_size):
        return (out_size - self.filter_size) // self.filter_stride + 1

    def __repr__(self):
        return "Deconvolution %d->%d C%d/%d" % (self.nIn, self.nOut, _cube(self.filter_size), _cube(self.filter_stride))


class NetworkInNetwork(nn.Module):
    """[UPSTREAM] networkInNetwork.py: 1x1 convolution = a dense matmul over the active rows (library GEMM; only the
    unused residual variants of scn_unet.py:150 reach it)."""

    def __init__(self, nIn, nOut, bias):
        super().__init__()
        self.nIn, self.nOut = nIn, nOut
        self.weight = nn.Parameter(torch.empty(nIn, nOut).normal_(0, (2.0 / nIn) ** 0.5))
        self.bias = nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, input):
        out = SparseConvNetTensor(metadata=input.metadata, spatial_size=input.spatial_size)
        out.features = input.features @ self.weight
        if self.bias is not None:
            out.features = out.features + self.bias
        return out

    def input_spatial_size(self, out_size):
        return out_size


class BatchNormalization(nn.Module):
    """[UPSTREAM] batchNormalization.py: eps=1e-4, momentum=0.9 (the KEEP fraction), optional fused leaky ReLU.
    Parameters weight, bias; buffers running_mean, running_var (no num_batches_tracked)."""

    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1):
        super().__init__()
        self.nPlanes, self.eps, self.momentum, self.affine, self.leakiness = nPlanes, eps, momentum, affine, leakiness
        self.register_buffer("running_mean", torch.zeros(nPlanes))
        self.register_buffer("running_var", torch.ones(nPlanes))
        if affine:
            self.weight = nn.Parameter(torch.ones(nPlanes))
            self.bias = nn.Parameter(torch.zeros(nPlanes))
        else:
            raise NotImplementedError("affine=False is not used by MoPA and not implemented")

    def forward(self, input):
        assert input.features.numel() == 0 or input.features.size(1) == self.nPlanes
        out = SparseConvNetTensor(metadata=input.metadata, spatial_size=input.spatial_size)
        out.features = F.BatchNormFunction.apply(input.features, self.weight, self.bias, self.running_mean, self.running_var,
                                                 self.eps, self.momentum, self.training, self.leakiness)
        return out

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        s = "BatchNorm(%d,eps=%g,momentum=%g,affine=%s" % (self.nPlanes, self.eps, self.momentum, self.affine)
        return s + (",leakiness=%g)" % self.leakiness if self.leakiness != 1 else ")")


class BatchNormReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9):
        super().__init__(nPlanes, eps, momentum, True, 0)


class BatchNormLeakyReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, leakiness=0.333):
        super().__init__(nPlanes, eps, momentum, True, leakiness)


def UNet(dimension, reps, nPlanes, residual_blocks=False, downsample=[2, 2], leakiness=0, n_input_planes=-1):
    """[UPSTREAM] networkArchitectures.py::UNet -- the recursive U-Net scn_unet.py:28 instantiates as
    UNet(3, 1, [16, 32, ..., 112], False). Module nesting (and therefore state_dict keys) follows upstream exactly."""

    def block(m, a, b):
        if residual_blocks:  # ResNet style
            m.add(ConcatTable()
                  .add(Identity() if a == b else NetworkInNetwork(a, b, False))
                  .add(Sequential()
                       .add(BatchNormLeakyReLU(a, leakiness=leakiness))
                       .add(SubmanifoldConvolution(dimension, a, b, 3, False))
                       .add(BatchNormLeakyReLU(b, leakiness=leakiness))
                       .add(SubmanifoldConvolution(dimension, b, b, 3, False)))).add(AddTable())
        else:  # VGG style
            m.add(Sequential()
                  .add(BatchNormLeakyReLU(a, leakiness=leakiness))
                  .add(SubmanifoldConvolution(dimension, a, b, 3, False)))

    def U(planes, n_in=-1):
        m = Sequential()
        for _ in range(reps):
            block(m, n_in if n_in != -1 else planes[0], planes[0])
            n_in = -1
        if len(planes) > 1:
            m.add(ConcatTable()
                  .add(Identity())
                  .add(Sequential()
                       .add(BatchNormLeakyReLU(planes[0], leakiness=leakiness))
                       .add(Convolution(dimension, planes[0], planes[1], downsample[0], downsample[1], False))
                       .add(U(planes[1:]))
                       .add(BatchNormLeakyReLU(planes[1], leakiness=leakiness))
                       .add(Deconvolution(dimension, planes[1], planes[0], downsample[0], downsample[1], False))))
            m.add(JoinTable())
            for i in range(reps):
                block(m, planes[0] * (2 if i == 0 else 1), planes[0])
        return m

    return U(list(nPlanes), n_input_planes)
