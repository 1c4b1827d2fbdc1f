"""Drop-in for the `sparseconvnet` surface on MoPA's UNetSCN path (mopa/models/scn_unet.py:4). B200 / CUDA only."""
from .functional import Metadata, get_precision, set_precision
from .modules import (AddTable, BatchNormalization, BatchNormLeakyReLU, BatchNormReLU, ConcatTable, Convolution,
                      Deconvolution, Identity, InputLayer, JoinTable, NetworkInNetwork, OutputLayer, Sequential,
                      SparseConvNetTensor, SubmanifoldConvolution, UNet)



def release_arenas():
    """Drop the idle activation / gradient / scratch arenas the compiled executor keeps across steps (they survive
    torch.cuda.empty_cache() by design: scn/compiler.py::_ArenaPool)."""
    from .compiler import release_arenas as _release
    _release()


__all__ = ["AddTable", "BatchNormalization", "BatchNormLeakyReLU", "BatchNormReLU", "ConcatTable", "Convolution",
           "Deconvolution", "Identity", "InputLayer", "JoinTable", "Metadata", "NetworkInNetwork", "OutputLayer",
           "Sequential", "SparseConvNetTensor", "SubmanifoldConvolution", "UNet", "get_precision", "release_arenas", "set_precision"]
