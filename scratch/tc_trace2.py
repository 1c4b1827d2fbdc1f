"""Timeline of one mid-grid CTA of k_conv_tc (trace build: scratch/build_trace_lib.sh, MOPA_SCN_LIB=scratch/bin/libmopa_scn_trace.so).
usage: python scratch/tc_trace2.py LEVEL CIN COUT [LEVEL CIN COUT ...]   (level = how many strided convs below the input)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mopa_b200.scn as scn
from mopa_b200 import _lib, synth
scn.set_precision("tf32")
lib = _lib.load()
lib.mopa_scn_debug_tc_trace.argtypes = [ctypes.c_void_p]
args = [int(a) for a in sys.argv[1:]]
coords, _ = synth.make_batch(8, "nuscenes", 0)
for level, cin, cout in zip(args[0::3], args[1::3], args[2::3]):
    x = scn.InputLayer(3, 4096, mode=4)([torch.from_numpy(coords), torch.ones(coords.shape[0], 16 if level else cin).cuda()])
    with torch.no_grad():
        for l in range(level):
            a = x.features.shape[1]
            x = scn.Convolution(3, a, cin if l == level - 1 else a, 2, 2, False).cuda()(x)
        conv = scn.SubmanifoldConvolution(3, cin, cout, 3, False).cuda()
        x.features = torch.randn_like(x.features)
        for _ in range(3):
            y = conv(x)
        torch.cuda.synchronize()
        buf = np.zeros((8, 512), np.int64)
        lib.mopa_scn_debug_tc_trace(buf.ctypes.data)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); y = conv(x); t1.record()
        torch.cuda.synchronize()
    print("== level %d %d->%d rows %d kernel+pack %.1f us" % (level, cin, cout, x.features.shape[0], 1e3 * t0.elapsed_time(t1)))
    lib.mopa_scn_debug_tc_trace(buf.ctypes.data)
    ph = buf[7, :6].astype(np.float64)
    print("CTA phases (cycles): setup %d | gather loop (warp 0) %d | wait d_full %d | epilogue %d | final sync %d | total %d" % (
        ph[1] - ph[0], ph[2] - ph[1], ph[3] - ph[2], ph[4] - ph[3], ph[5] - ph[4], ph[5] - ph[0]))
    ni = int((buf[4] != 0).sum()); ng = int((buf[0] != 0).sum())
    i_top, i_full, i_com = (buf[r, :ni].astype(np.float64) for r in (4, 5, 6))
    i_fence = buf[3, :ni].astype(np.float64)
    if ni > 1 and i_fence.any():
        print("issuer split: a_full seen -> fences done mean %.0f | fences -> MMAs + commits issued mean %.0f" % (
            (i_fence - i_full).mean(), (i_com - i_fence).mean()))
    g_top, g_emp, g_arr = (buf[r, :ng].astype(np.float64) for r in (0, 1, 2))
    print("issuer steps %d (tile 0), gather-warp-0 steps %d" % (ni, ng))
    if ni > 1:
        print("issuer: step-to-step mean %.0f median %.0f | wait a_full mean %.0f | a_full->committed mean %.0f | first a_full at %d after setup" % (
            np.diff(i_top).mean(), np.median(np.diff(i_top)), (i_full - i_top).mean(), (i_com - i_full).mean(), i_full[0] - ph[1]))
    if ng > 1:
        print("gather warp 0: step-to-step mean %.0f median %.0f | wait a_empty mean %.0f | issue (a_empty->arrive) mean %.0f max %.0f" % (
            np.diff(g_top).mean(), np.median(np.diff(g_top)), (g_emp - g_top).mean(), (g_arr - g_emp).mean(), (g_arr - g_emp).max()))
    print("issuer per-step (a_full wait, issue):", " ".join("%d/%d" % (a, b) for a, b in list(zip(i_full - i_top, i_com - i_full))[:30]))
