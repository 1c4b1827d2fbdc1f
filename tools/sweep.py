#!/usr/bin/env python
"""BASELINE.json configs 4 and 5 in one script: UNetSCN fwd+bwd on SemanticKITTI-shaped scans and a
point-count sweep (10k .. 200k points per scan, batch 8), each through bench.py's own timed region (subprocess, so every
point is a fresh process); writes one JSON with points/s, ms/step (mean, median, p10, p90), voxel counts and the per-scan
imbalance.      python tools/sweep.py [--out profiles/r02_sweep.json] [--steps 40]       (run on the GPU box)"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra, steps):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(steps), "--warmup", "10", "--no-cpu-baseline",
           "--no-roofline", "--no-fp32"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    for line in reversed(out.stdout.splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("bench.py printed no JSON line: %s\n%s" % (" ".join(cmd), out.stderr[-2000:]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--steps", type=int, default=40)
    a = ap.parse_args()
    sys.path.insert(0, ROOT)
    from mopa_b200 import synth
    rows = []
    cases = [("nuscenes N32 (config 1/2/3 shape)", ["--sensor", "nuscenes"]), ("kitti K64 full sweep (config 4)", ["--sensor", "kitti"]),
             ("kitti K64 ~100k points (config 4 as quoted)", ["--sensor", "kitti", "--points", "100000"])]
    cases += [("nuscenes %dk points/scan (config 5)" % (p // 1000), ["--sensor", "nuscenes", "--points", str(p)])
              for p in (10000, 20000, 50000, 100000, 200000)]
    for name, extra in cases:
        d = run(extra, a.steps)
        sensor = extra[1]
        pts = int(extra[3]) if len(extra) > 3 else 0
        n_az = synth.azimuth_for_points(pts, sensor) if pts else None
        per_scan = [synth.make_scan(sensor, seed=b, n_azimuth=n_az)[0].shape[0] for b in range(8)]
        rows.append({"case": name, "args": extra, "points_per_step": d["config"]["points_per_step"], "points_per_s": d["value"],
                     "ms_per_step": d["ms_per_step"], "step_ms": d.get("step_ms"), "e2e_ms_per_step": d["e2e"]["ms_per_step"],
                     "e2e_points_per_s": d["e2e"]["value"], "scan_points_min_max": [min(per_scan), max(per_scan)],
                     "clocks": d.get("clocks")})
        print("%-46s %9.0f pts/step  %7.3f ms/step (median %.3f)  %6.1f M points/s   e2e %7.3f ms" % (
            name, rows[-1]["points_per_step"], d["ms_per_step"], (d.get("step_ms") or {}).get("median", float("nan")),
            d["value"] / 1e6, d["e2e"]["ms_per_step"]), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"what": "tools/sweep.py: UNetSCN fwd+bwd, batch 8 per GPU, tf32 mode, 1 GPU; bench.py timed region per case",
               "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
