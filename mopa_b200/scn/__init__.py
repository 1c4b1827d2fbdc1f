"""Drop-in for the `sparseconvnet` surface on MoPA's UNetSCN path (mopa/models/scn_unet.py:4). B200 / CUDA only."""
from .functional import Metadata, get_precision, set_precision
from .modules import (AddTable, BatchNormalization, BatchNormLeakyReLU, BatchNormReLU, ConcatTable, Convolution,
                      Deconvolution, Identity, InputLayer, JoinTable, NetworkInNetwork, OutputLayer, Sequential,
                      SparseConvNetTensor, SubmanifoldConvolution, UNet)

__all__ = ["AddTable", "BatchNormalization", "BatchNormLeakyReLU", "BatchNormReLU", "ConcatTable", "Convolution",
           "Deconvolution", "Identity", "InputLayer", "JoinTable", "Metadata", "NetworkInNetwork", "OutputLayer",
           "Sequential", "SparseConvNetTensor", "SubmanifoldConvolution", "UNet", "get_precision", "set_precision"]
