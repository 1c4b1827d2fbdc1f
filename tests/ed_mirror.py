"""UNetSCN_ED, the reference's unrolled encoder/decoder variant (mopa/models/scn_unet.py:38-134; VGG blocks only, the
configuration its own smoke test :222-239 builds), written against the `sparseconvnet` surface so that the GPU tests can
run it without /root/reference. Same attribute names and module nesting, hence the same state_dict keys: the CPU test
tests/test_surface.py::test_ed_mirror_matches_the_reference_module_tree checks that against the reference's own file
whenever /root/reference is present. This class takes the module-by-module (eager) path: its forward is plain Python."""
import torch.nn as nn

import sparseconvnet as scn


class UNetSCN_ED(nn.Module):
    def __init__(self, in_channels, m=16, full_scale=4096):
        super().__init__()

        def bn(c):
            return scn.BatchNormLeakyReLU(c, leakiness=0)

        def subm(a, b):
            return scn.SubmanifoldConvolution(3, a, b, 3, False)

        self.input = scn.InputLayer(3, full_scale, mode=4)
        self.down_in = subm(in_channels, m)
        self.main_block1 = scn.Sequential().add(bn(m)).add(subm(m, m))
        for l in range(2, 8):  # BN, then Sequential(strided conv, BN, submanifold conv)
            a, b = (l - 1) * m, l * m
            inner = scn.Sequential().add(scn.Convolution(3, a, b, 2, 2, False)).add(bn(b)).add(subm(b, b))
            setattr(self, "main_block%d" % l, scn.Sequential().add(bn(a)).add(inner))
        self.deconv7 = scn.Sequential().add(bn(7 * m)).add(scn.Deconvolution(3, 7 * m, 6 * m, 2, 2, False))
        self.join7 = scn.JoinTable()
        for l in range(6, 1, -1):
            a, b = 2 * l * m, (l - 1) * m
            dec = scn.Sequential().add(bn(a)).add(subm(a, a // 2)).add(bn(a // 2)).add(scn.Deconvolution(3, a // 2, b, 2, 2, False))
            setattr(self, "deconv%d" % l, dec)
            setattr(self, "join%d" % l, scn.JoinTable())
        self.deconv1 = scn.Sequential().add(bn(2 * m)).add(subm(2 * m, m))
        self.output = scn.Sequential().add(scn.BatchNormReLU(m)).add(scn.OutputLayer(3))

    def forward(self, x):
        x = self.down_in(self.input(x))
        feats = [None, self.main_block1(x)]
        for l in range(2, 8):
            feats.append(getattr(self, "main_block%d" % l)(feats[-1]))
        d = self.join7([feats[6], self.deconv7(feats[7])])
        for l in range(6, 1, -1):
            d = getattr(self, "join%d" % l)([feats[l - 1], getattr(self, "deconv%d" % l)(d)])
        return self.output(self.deconv1(d))
