"""Generates tests/golden/unet_small.npz from the CPU oracle (float64). PARITY UNPINNED: the reference ships no golden
vectors and SparseConvNet cannot be imported here, so these are outputs of oracle/scn_oracle.py, which is
itself pinned only by the dense-conv3d arbiter in tests/test_oracle.py.   Run: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import scn_oracle as so  # noqa: E402
from tests.helpers import small_batch  # noqa: E402

STATE_SEED = 21


def main():
    coords, feats = small_batch(2, 80, 11)
    state = so.make_unet_state(seed=STATE_SEED)
    net = so.OracleUNetSCN(state, dtype=torch.float64)
    out = net.forward(coords, feats)
    grad_out = torch.randn(out.shape, generator=torch.Generator().manual_seed(5)).float()
    out.backward(grad_out.double())
    geo = net.geo
    np.savez_compressed(
        os.path.join(os.path.dirname(os.path.abspath(__file__)), "unet_small.npz"),
        coords=coords, feats=feats, state_seed=STATE_SEED, p2v=geo.p2v,
        n_active=np.array([geo.n_active(l) for l in range(7)]),
        subm_rule_counts=np.array([[int((geo.subm_table(l)[k] >= 0).sum()) for k in range(27)] for l in range(7)]),
        out=out.detach().numpy(), grad_out=grad_out.numpy(),
        grad_w1=net.params["sparseModel.1.weight"].grad.numpy(),
        grad_bn3_w=net.params["sparseModel.3.weight"].grad.numpy(),
        running_mean3=net.params["sparseModel.3.running_mean"].numpy())


if __name__ == "__main__":
    main()
