"""Phases of one mid-grid CTA of the LAST k_conv_tc launch of a UNetSCN forward+backward (= d_input of the level-0
16->16 submanifold convolution, in front of the first BatchNormReLU). Trace build: scratch/build_trace_lib.sh."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mopa_b200.scn as scn
from mopa_b200 import _lib, synth
from mopa_b200.unet_scn import UNetSCN
scn.set_precision("tf32")
lib = _lib.load()
lib.mopa_scn_debug_tc_trace.argtypes = [ctypes.c_void_p]
coords, feats = synth.make_batch(8, "nuscenes", 0)
net = UNetSCN(1).cuda()
c, f = torch.from_numpy(coords), torch.from_numpy(feats).cuda()
buf = np.zeros((8, 512), np.int64)
for it in range(3):
    out = net([c, f]); out.square().mean().backward()
    torch.cuda.synchronize()
    lib.mopa_scn_debug_tc_trace(buf.ctypes.data)
ph = buf[7, :6].astype(np.float64)
print("CTA phases (cycles): setup %d | gather loop (warp 0) %d | wait d_full %d | epilogue %d | final sync %d | total %d" % (
    ph[1] - ph[0], ph[2] - ph[1], ph[3] - ph[2], ph[4] - ph[3], ph[5] - ph[4], ph[5] - ph[0]))
e = buf[7, :12].astype(np.float64)
print("epilogue of warp 0, q = 0 (cycles after d_full seen): TMEM loaded %d | rows stored %d | half 0: summands %d, folded %d | half 1: summands %d, folded %d | end %d" % (
    e[6] - e[3], e[7] - e[3], e[8] - e[3], e[9] - e[3], e[10] - e[3], e[11] - e[3], e[4] - e[3]))
