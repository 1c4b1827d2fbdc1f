#!/usr/bin/env python
"""Per-op device times of one UNetSCN fwd+bwd (CUDA events on the launching stream, library-side: mopa_scn_Profile_*),
against each op's algorithmic bytes / flops (SURVEY.md 8(d) formulas). Run on the GPU box:
    python tools/layer_table.py [--batch 8] [--steps 5] [--precision tf32] [--out gpurun_out/layers.json]
"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CLASS = {1: "conv_fwd", 2: "conv_dinput", 3: "conv_dweight", 4: "bn_fwd", 5: "bn_bwd", 6: "io"}
OP = {0: "", 1: "subm", 2: "conv", 3: "deconv"}


def algorithmic(rec):
    """(bytes, flops) of one op per SURVEY.md 8(d)."""
    cls = rec["tag"] // 10
    r, ci, co, k = rec["rules"], rec["c_in"], rec["c_out"], rec["volume"]
    if cls in (1, 2):  # gather conv: R (4 Cin + 8) + 4 Vout Cout + 4 K Cin Cout   (d_input: same kernel, roles swapped)
        return r * (4 * ci + 8) + 4 * rec["rows_out"] * co + 4 * k * ci * co, 2.0 * r * ci * co
    if cls == 3:  # d_weight: R (4 (Cin + Cout) + 8) + 4 K Cin Cout
        return r * (4 * (ci + co) + 8) + 4 * k * ci * co, 2.0 * r * ci * co
    if cls == 4:  # BN + ReLU forward (train): stats read, normalise read, write
        return 3 * 4 * rec["rows_out"] * ci, 0.0
    if cls == 5:
        return 5 * 4 * rec["rows_out"] * ci, 0.0
    return 0, 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--sensor", default="nuscenes")
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "layers.json"))
    a = ap.parse_args()
    # per-op event times: keep the d_weight kernels on the main stream (in normal runs they overlap the d_input kernels)
    os.environ["MOPA_SCN_NO_DW_OVERLAP"] = "1"
    from mopa_b200 import _lib, synth
    from mopa_b200.unet_scn import UNetSCN
    import mopa_b200.scn as scn
    scn.set_precision(a.precision)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("hbm_gbs", 6650.0)
    torch.manual_seed(0)
    net = UNetSCN(1).cuda()
    batches = [synth.make_batch(a.batch, a.sensor, seed=1000 * i) for i in range(4)]
    dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()) for c, f in batches]
    for i in range(3):
        net(list(dev[i % 4])).sum().backward()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for i in range(a.steps):
        net(list(dev[i % 4])).sum().backward()
    recs = _lib.profile_read()
    _lib.profile_enable(False)
    agg = collections.OrderedDict()
    for r in recs:
        key = (CLASS.get(r["tag"] // 10, "?"), OP.get(r["tag"] % 10, "?"), r["c_in"], r["c_out"], r["volume"])
        b, fl = algorithmic(r)
        e = agg.setdefault(key, {"n": 0, "ms": 0.0, "bytes": 0.0, "flops": 0.0, "rows": 0, "rules": 0})
        e["n"] += 1; e["ms"] += r["ms"]; e["bytes"] += b; e["flops"] += fl; e["rows"] += r["rows_out"]; e["rules"] += r["rules"]
    rows, by_class = [], collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    print("%-12s %-6s %4s %4s %3s %9s %9s %8s %8s %6s %7s" % ("class", "op", "cin", "cout", "K", "rows", "rules", "us/call", "GB/s", "frac", "TF/s"))
    for (cls, op, ci, co, k), e in agg.items():
        us = 1e3 * e["ms"] / e["n"]
        gbs = e["bytes"] / (e["ms"] * 1e-3) / 1e9 if e["ms"] > 0 else 0.0
        tf = e["flops"] / (e["ms"] * 1e-3) / 1e12 if e["ms"] > 0 else 0.0
        print("%-12s %-6s %4d %4d %3d %9d %9d %8.1f %8.0f %6.3f %7.1f" % (cls, op, ci, co, k, e["rows"] // e["n"], e["rules"] // e["n"], us, gbs, gbs / peak, tf))
        rows.append({"class": cls, "op": op, "c_in": ci, "c_out": co, "volume": k, "calls": e["n"], "rows": e["rows"] // e["n"],
                     "rules": e["rules"] // e["n"], "us_per_call": us, "gbs": gbs, "frac_of_hbm_peak": gbs / peak, "tflops": tf})
        c = by_class[cls]
        c[0] += e["ms"]; c[1] += e["bytes"]; c[2] += e["flops"]
    print()
    summary = {}
    for cls, (ms, b, fl) in by_class.items():
        gbs = b / (ms * 1e-3) / 1e9 if ms else 0.0
        summary[cls] = {"ms_per_step": ms / a.steps, "gbs": gbs, "frac_of_hbm_peak": gbs / peak, "tflops": fl / (ms * 1e-3) / 1e12 if ms else 0.0}
        print("%-12s %8.3f ms/step  %7.0f GB/s  (%.3f of measured %.0f GB/s)  %6.1f TF/s" % (cls, ms / a.steps, gbs, gbs / peak, peak, summary[cls]["tflops"]))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"batch": a.batch, "sensor": a.sensor, "precision": a.precision, "steps": a.steps, "peak_gbs": peak,
               "ops": rows, "classes": summary}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
