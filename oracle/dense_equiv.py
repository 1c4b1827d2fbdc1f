"""Dense-equivalence arbiter for the CPU oracle (TEST INFRASTRUCTURE; PARITY UNPINNED -- see scn_oracle.py).

Independent of anyone's memory of SparseConvNet's internals: on a small crop the sparse layers must equal
ordinary dense torch ops read back at the active sites:
  SubmanifoldConvolution(3)   == F.conv3d(padding=1)            at the SAME active set
  Convolution(k2, s2)         == F.conv3d(stride=2)             at the coarse active set
  Deconvolution(k2, s2)       == F.conv_transpose3d(stride=2)   at the fine active set
  BatchNorm(Leaky)ReLU        == F.batch_norm(eps=1e-4, momentum=0.1) + leaky_relu over the active rows
  InputLayer(mode 4)          == per-site mean;  OutputLayer == index_select
"""
import numpy as np
import torch
import torch.nn.functional as F


def densify(voxel_coords, feats, size, n_batch):
    vc = torch.as_tensor(np.asarray(voxel_coords), dtype=torch.long)
    dense = torch.zeros(n_batch, feats.shape[1], size, size, size, dtype=feats.dtype)
    dense[vc[:, 3], :, vc[:, 0], vc[:, 1], vc[:, 2]] = feats
    return dense


def read_sites(dense, voxel_coords):
    vc = torch.as_tensor(np.asarray(voxel_coords), dtype=torch.long)
    return dense[vc[:, 3], :, vc[:, 0], vc[:, 1], vc[:, 2]]


def subm_conv_dense(voxel_coords, feats, weight, size, n_batch):
    w = weight.reshape(3, 3, 3, weight.shape[-2], weight.shape[-1]).permute(4, 3, 0, 1, 2)
    out = F.conv3d(densify(voxel_coords, feats, size, n_batch), w, padding=1)
    return read_sites(out, voxel_coords)


def strided_conv_dense(fine_coords, feats, weight, coarse_coords, size, n_batch):
    w = weight.reshape(2, 2, 2, weight.shape[-2], weight.shape[-1]).permute(4, 3, 0, 1, 2)
    out = F.conv3d(densify(fine_coords, feats, size, n_batch), w, stride=2)
    return read_sites(out, coarse_coords)


def strided_deconv_dense(coarse_coords, feats, weight, fine_coords, size_coarse, n_batch):
    w = weight.reshape(2, 2, 2, weight.shape[-2], weight.shape[-1]).permute(3, 4, 0, 1, 2)
    out = F.conv_transpose3d(densify(coarse_coords, feats, size_coarse, n_batch), w, stride=2)
    return read_sites(out, fine_coords)


def bn_relu_dense(x, weight, bias, running_mean, running_var, train, leakiness=0.0):
    y = F.batch_norm(x, running_mean, running_var, weight, bias, training=train, momentum=0.1, eps=1e-4)
    return F.leaky_relu(y, leakiness)
