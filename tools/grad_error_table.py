#!/usr/bin/env python
"""Measured gradient / feature error table: ONE full nuScenes-shaped scan through UNetSCN, per
parameter tensor: float32 ORACLE, GPU fp32 mode (3xTF32) and GPU tf32 mode, each against the float64 oracle.
Metrics: relative L2, cosine, and an element-wise metric that small entries cannot hide from:
median and 99th percentile of |a - b| / (|b| + 1e-3 * rms(b)).

    python tools/grad_error_table.py [--sensor nuscenes] [--out profiles/r02_grad_errors]      (run on the GPU box)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def metrics(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    rms = float(b.pow(2).mean().sqrt())
    ew = (a - b).abs() / (b.abs() + 1e-3 * rms + 1e-300)
    return {"rel_l2": float((a - b).norm() / b.norm()), "cos": float(torch.dot(a, b) / (a.norm() * b.norm())),
            "maxabs_over_max": float((a - b).abs().max() / b.abs().max()),
            "ew_median": float(ew.median()), "ew_p99": float(torch.quantile(ew, 0.99)) if ew.numel() < 2 ** 24 else float(np.quantile(ew.numpy(), 0.99))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sensor", default="nuscenes")
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "grad_errors"))
    a = ap.parse_args()
    from mopa_b200 import synth
    from mopa_b200.unet_scn import UNetSCN
    import mopa_b200.scn as scn
    from oracle import scn_oracle as so

    coords, feats = synth.make_scan(a.sensor, seed=a.seed)
    state = so.make_unet_state(seed=11)
    g64 = torch.randn(coords.shape[0], 16, dtype=torch.float64, generator=torch.Generator().manual_seed(1))

    def run_oracle(dtype):
        o = so.OracleUNetSCN(state, dtype=dtype)
        out = o.forward(coords, feats)
        out.backward(g64.to(dtype))
        return out.detach(), {k: v.grad for k, v in o.params.items() if v.grad is not None}

    def run_gpu(precision):
        scn.set_precision(precision)
        net = UNetSCN(1).cuda()
        net.load_state_dict(state)
        out = net([torch.from_numpy(coords), torch.from_numpy(feats).cuda()])
        out.backward(g64.float().cuda())
        return out.detach(), {k: p.grad for k, p in net.named_parameters()}

    ref_out, ref_g = run_oracle(torch.float64)
    arms = {"oracle_fp32": run_oracle(torch.float32), "gpu_fp32": run_gpu("fp32"), "gpu_tf32": run_gpu("tf32")}
    table = {"points": int(coords.shape[0]), "sensor": a.sensor, "forward": {}, "params": {}}
    for arm, (out, _) in arms.items():
        table["forward"][arm] = metrics(out, ref_out)
    for name in ref_g:
        table["params"][name] = {arm: metrics(g[name], ref_g[name]) for arm, (_, g) in arms.items()}
    worst = {arm: {m: (max if m != "cos" else min)(table["params"][n][arm][m] for n in ref_g)
                   for m in ("rel_l2", "cos", "maxabs_over_max", "ew_median", "ew_p99")} for arm in arms}
    table["worst"] = worst
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(table, open(a.out + ".json", "w"), indent=1)
    with open(a.out + ".txt", "w") as f:
        f.write("# tools/grad_error_table.py: one %s-shaped scan (%d points), UNetSCN fwd+bwd, every tensor vs the float64 oracle\n"
                % (a.sensor, coords.shape[0]))
        f.write("# columns per arm: rel-L2 | 1-cos | elementwise |a-b|/(|b|+1e-3 rms) median / p99\n")
        f.write("%-34s %s\n" % ("forward features", "   ".join("%s %.2e" % (arm, table["forward"][arm]["maxabs_over_max"]) for arm in arms)))
        f.write("%-34s " % "parameter gradient" + "".join("| %-38s" % arm for arm in arms) + "\n")
        for n in ref_g:
            row = "%-34s " % n.replace("sparseModel.", "")
            for arm in arms:
                m = table["params"][n][arm]
                row += "| %.1e %.1e  %.1e / %.1e " % (m["rel_l2"], 1 - m["cos"], m["ew_median"], m["ew_p99"])
            f.write(row + "\n")
        f.write("\nworst over the %d tensors:\n" % len(ref_g))
        for arm in arms:
            w = worst[arm]
            f.write("  %-12s rel-L2 %.2e   min cos %.6f   max-abs/max %.2e   elementwise median %.2e  p99 %.2e\n"
                    % (arm, w["rel_l2"], w["cos"], w["maxabs_over_max"], w["ew_median"], w["ew_p99"]))
    print(open(a.out + ".txt").read().split("worst over")[1])


if __name__ == "__main__":
    main()
