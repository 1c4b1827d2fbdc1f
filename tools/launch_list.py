#!/usr/bin/env python
"""Step-aligned launch list from an `ncu --metrics gpu__time_duration.sum` CSV of bench.py: takes exactly ONE step (from the
step's first kernel, k_insert<0> of the voxel hashing, to the launch before the next one), skipping `--skip` warm-up steps,
and aggregates per kernel: launches, total / average device time, share of the step. (ncu serialises launches and runs them
cold-cache: the SHARES are meaningful, the absolute times are not.)
    python tools/launch_list.py gpurun_out/r02_launches.csv [--skip 3] > profiles/r02_launches.txt"""
import argparse
import collections
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--skip", type=int, default=3)
    a = ap.parse_args()
    with open(a.csv) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    seq = [(re.sub(r"\(.*", "", row[ki]), float(row[vi].replace(",", ""))) for row in r if row[mi] == "gpu__time_duration.sum"]
    starts = [i for i, (n, _) in enumerate(seq) if n.startswith("void mopa::k_insert<0>") or n.startswith("k_insert<0>")
              or "k_insert<0>" in n]
    if len(starts) < a.skip + 2:
        raise SystemExit("not enough steps in the capture (%d step starts)" % len(starts))
    lo, hi = starts[a.skip], starts[a.skip + 1]
    step = seq[lo:hi]
    agg = collections.OrderedDict()
    for n, t in step:
        n = n.replace("void ", "").replace("mopa::", "")
        e = agg.setdefault(n, [0, 0.0])
        e[0] += 1
        e[1] += t
    tot = sum(v[1] for v in agg.values())
    print("# one step (launches %d..%d of the capture; step %d after %d warm-up steps): %d launches, %.1f us of serialised device time"
          % (lo, hi - 1, a.skip + 1, a.skip, len(step), tot / 1e3))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%9.1f us %5.1f%%  n=%3d  avg %7.1f us  %s" % (v[1] / 1e3, 100 * v[1] / tot, v[0], v[1] / 1e3 / v[0], k[:110]))


if __name__ == "__main__":
    main()
