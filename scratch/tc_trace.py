"""Timeline of CTA 0 of one k_conv_tc launch (MOPA_TC_DBG=32): python scratch/tc_trace.py CIN COUT ROWS"""
import ctypes, os, sys
os.environ["MOPA_TC_DBG"] = str(32 | int(os.environ.get("EXTRA_DBG", "0")))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mopa_b200.scn as scn
from mopa_b200 import _lib
from tests.helpers import small_batch
cin, cout = int(sys.argv[1]), int(sys.argv[2])
naz = int(sys.argv[3]) if len(sys.argv) > 3 else 300
coords, _ = small_batch(4, naz, 0)
feats = torch.randn(coords.shape[0], cin).cuda()
x = scn.InputLayer(3, 4096, mode=4)([torch.from_numpy(coords), feats])
conv = scn.SubmanifoldConvolution(3, cin, cout, 3, False).cuda()
for _ in range(3):
    y = conv(x)
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros((8, 512), np.int64)
lib.mopa_scn_debug_tc_trace.argtypes = [ctypes.c_void_p]
print("rc", lib.mopa_scn_debug_tc_trace(buf.ctypes.data), "rows", x.features.shape[0])
n = 27 * ((cin + 31) // 32)
t0 = buf[buf > 0].min()
b = buf - t0
names = ["g:loop top", "g:landed", "g:a_empty ok", "g:arrived", "i:step top", "i:b_full ok", "i:a_full ok", "i:committed"]
print("step " + " ".join("%12s" % s for s in names))
for i in list(range(0, min(n, 14))) + list(range(max(14, n - 4), n)):
    print("%4d " % i + " ".join("%12d" % b[r, i] for r in range(8)))
d = np.diff(b[3, :n])
print("gather warp0 arrive-to-arrive cycles: mean %.0f median %.0f" % (d.mean(), np.median(d)))
print("mean phase cycles per step: issue->landed %.0f, landed->a_empty %.0f, a_empty->arrived %.0f" % (
    (b[1, :n] - b[0, :n]).mean(), (b[2, :n] - b[1, :n]).mean(), (b[3, :n] - b[2, :n]).mean()))
print("issuer: top->b_full %.0f, b_full->a_full(t0) %.0f, a_full->commit %.0f" % (
    (b[5, :n] - b[4, :n]).mean(), (b[6, :n] - b[5, :n]).mean(), (b[7, :n] - b[6, :n]).mean()))

if int(os.environ.get("EXTRA_DBG", "0")) & 64:
    print("issuer fine (cycles): a_full(t0)->fenced %.0f, fenced->4 MMAs issued %.0f, ->commit(a_empty t0) %.0f, ->tile1 done %.0f, ->commit(b) %.0f" % (
        (b[0, :n] - b[6, :n]).mean(), (b[1, :n] - b[0, :n]).mean(), (b[2, :n] - b[1, :n]).mean(), (b[3, :n] - b[2, :n]).mean(), (b[7, :n] - b[3, :n]).mean()))
