"""Timeline of one mid-grid CTA of a k_conv_tc launch. Needs a trace build: MOPA_BUILD_DEFS=MOPA_TC_TRACE python -m mopa_b200._build --force ; python scratch/tc_trace.py CIN COUT [CIN COUT ...]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mopa_b200.scn as scn
from mopa_b200 import _lib, synth
scn.set_precision("tf32")
lib = _lib.load()
lib.mopa_scn_debug_tc_trace.argtypes = [ctypes.c_void_p]
args = [int(a) for a in sys.argv[1:]]
nscan = 8
coords, _ = synth.make_batch(nscan, "nuscenes", 0)
for cin, cout in zip(args[0::2], args[1::2]):
    feats = torch.randn(coords.shape[0], cin).cuda()
    x = scn.InputLayer(3, 4096, mode=4)([torch.from_numpy(coords), feats])
    conv = scn.SubmanifoldConvolution(3, cin, cout, 3, False).cuda()
    with torch.no_grad():
        for _ in range(3):
            y = conv(x)
    torch.cuda.synchronize()
    buf = np.zeros((8, 512), np.int64)
    lib.mopa_scn_debug_tc_trace(buf.ctypes.data)   # clear
    with torch.no_grad():
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); y = conv(x); t1.record()
    torch.cuda.synchronize()
    print("== %d->%d rows %d kernel+pack %.1f us" % (cin, cout, x.features.shape[0], 1e3 * t0.elapsed_time(t1)))
    lib.mopa_scn_debug_tc_trace(buf.ctypes.data)
    n = 27 * ((cin + 31) // 32)
    g_top, g_emp, g_arr, i_top, i_full, i_com = (buf[r, :n].astype(np.float64) for r in (0, 1, 2, 4, 5, 6))
    base = g_top[0]
    print("step   g:top  g:a_empty  g:arrived | i:top  i:a_full  i:committed   (cycles from first step)")
    for i in list(range(0, min(n, 12))) + list(range(max(12, n - 3), n)):
        print("%4d %7d %9d %9d | %7d %8d %8d" % (i, g_top[i] - base, g_emp[i] - base, g_arr[i] - base, i_top[i] - base, i_full[i] - base, i_com[i] - base))
    d = np.diff(g_arr)
    print("gather warp 0: arrive-to-arrive mean %.0f median %.0f cycles; wait for a_empty mean %.0f; issue (a_empty -> arrive) mean %.0f" % (
        d.mean(), np.median(d), (g_emp - g_top).mean(), (g_arr - g_emp).mean()))
    print("issuer: wait for a_full mean %.0f; a_full -> committed mean %.0f; step-to-step mean %.0f" % (
        (i_full - i_top).mean(), (i_com - i_full).mean(), np.diff(i_top).mean()))
    print("gather arrive -> issuer sees a_full: mean %.0f median %.0f (includes the data landing)" % ((i_full - g_arr).mean(), np.median(i_full - g_arr)))
    sa = int(os.environ.get("TRACE_RING", "3"))
    if n > sa:
        lat = g_emp[sa:] - i_com[:-sa]
        print("issuer commit of step s -> gather acquires the stage for step s+%d: mean %.0f median %.0f min %.0f" % (sa, lat.mean(), np.median(lat), lat.min()))
