"""Host-side mirror of the reference's 3D branch on top of mopa_b200.scn.

`UNetSCN` follows mopa/models/scn_unet.py:9-34 (same constructor arguments, `.sparseModel` child, `.out_channels`),
`Net3DSeg` follows mopa/models/xmuda_arch.py:82-126 (UNetSCN + one or two linear heads; returns the same dict).
With the repo root on PYTHONPATH the reference's own files run unchanged through the `sparseconvnet` shim; these
classes exist so the tests and the bench do not need /root/reference at run time.
"""
import torch.nn as nn

from . import scn


class UNetSCN(nn.Module):
    def __init__(self, in_channels, m=16, block_reps=1, residual_blocks=False, full_scale=4096, num_planes=7,
                 pretrained=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, m
        planes = [m * (level + 1) for level in range(num_planes)]
        net = scn.Sequential()
        net.add(scn.InputLayer(3, full_scale, mode=4))
        net.add(scn.SubmanifoldConvolution(3, in_channels, m, 3, False))
        net.add(scn.UNet(3, block_reps, planes, residual_blocks))
        net.add(scn.BatchNormReLU(m))
        net.add(scn.OutputLayer(3))
        self.sparseModel = net

    def forward(self, x):
        return self.sparseModel(x)


class Net3DSeg(nn.Module):
    def __init__(self, num_classes, dual_head, backbone_3d="SCN", backbone_3d_kwargs=None, da_method=None,
                 pretrained=False):
        super().__init__()
        if backbone_3d != "SCN":
            raise NotImplementedError("3D backbone {} not supported".format(backbone_3d))
        self.backbone_3d = backbone_3d
        self.net_3d = UNetSCN(**(backbone_3d_kwargs or {"in_channels": 1}))
        self.linear = nn.Linear(self.net_3d.out_channels, num_classes)
        self.dual_head = dual_head
        if dual_head:
            self.linear2 = nn.Linear(self.net_3d.out_channels, num_classes)
        self.da_method = da_method
        if da_method == "MCD":
            self.linear3 = nn.Linear(self.net_3d.out_channels, num_classes)

    def forward(self, data_batch):
        feats = self.net_3d(data_batch["x"])
        preds = {"feats": feats, "seg_logit": self.linear(feats)}
        if self.dual_head:
            preds["seg_logit2"] = self.linear2(feats)
        return preds
