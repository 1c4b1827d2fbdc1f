"""Turns the raw files of one scratch/profile_round2.sh run (gpurun_out/<tag>_*) into the committed profiles/r02_* set."""
import csv, json, os, re, subprocess, sys
tag = sys.argv[1]
G, P = "gpurun_out", "profiles"
def cp(src, dst):
    open(os.path.join(P, dst), "w").write(open(os.path.join(G, src)).read())
cp(tag + "_bench.json", "r02_bench.json")
cp(tag + "_layers.json", "r02_layers.json")
cp(tag + "_sweep.json", "r02_sweep.json")
open(os.path.join(P, "r02_layers.txt"), "w").write(
    "# tools/layer_table.py on a B200 (batch 8 x ~32.9k points, tf32 mode, d_weight kernels on the main stream): per-op device time,\n"
    "# algorithmic bytes per SURVEY 8(d), achieved GB/s and fraction of the measured copy peak\n\n" + open(os.path.join(G, tag + "_layers.log")).read())
out = subprocess.run([sys.executable, "tools/launch_list.py", os.path.join(G, tag + "_launches.csv"), "--skip", "3"], capture_output=True, text=True).stdout
open(os.path.join(P, "r02_launches.txt"), "w").write(
    "# ncu --metrics gpu__time_duration.sum --clock-control none over `bench.py --steps 1 --warmup 3`; tools/launch_list.py cut out ONE step\n" + out)
# conv_tc traffic: the launches of one step (skip 3 warm-up steps x 50 launches)
rows = [l for l in open(os.path.join(G, tag + "_conv_tc_traffic.csv")) if l.startswith('"')]
r = csv.reader(rows); hdr = next(r)
ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = {}
for row in r:
    per.setdefault(int(row[idi]), {})[row[mi]] = float(row[vi].replace(",", ""))
ids = sorted(per)
step = ids[150:200] if len(ids) >= 200 else ids[-50:]
def mean(m): return sum(per[i].get(m, 0.0) for i in step) / len(step)
unit_fix = 1.0
d = {"launches": len(step), "what": "mean over the 50 k_conv_tc launches of one step (4th step of the capture), ncu --clock-control none, d_weight kernels on the main stream",
     "dram_bytes_per_launch": mean("dram__bytes_read.sum") + mean("dram__bytes_write.sum"),
     "lts_bytes_per_launch": mean("lts__t_bytes.sum"), "time_us_per_launch_under_ncu": mean("gpu__time_duration.sum") / 1e3}
# ncu prints bytes in scaled units per row in some versions: detect via the unit column
ui = hdr.index("Metric Unit")
units = {}
for row in csv.reader(rows[1:]):
    units[row[mi]] = row[ui]
d["units_reported_by_ncu"] = units
json.dump(d, open(os.path.join(P, "r02_conv_tc_traffic.json"), "w"), indent=1)
print(json.dumps(d)[:600])
for name, short in (("conv_tc", "conv_tc"), ("dw_tc", "dw_tc"), ("bn_stats", "bn")):
    raw, src = os.path.join(G, "%s_%s_raw.csv" % (tag, name)), os.path.join(G, "%s_%s_src1.csv" % (tag, name))
    if os.path.getsize(raw) < 1000:
        print("no ncu report for", name); continue
    a = subprocess.run([sys.executable, "scratch/ncu_raw_summary.py", raw], capture_output=True, text=True).stdout
    b = subprocess.run([sys.executable, "scratch/ncu_hot_spots.py", src, "0.015"], capture_output=True, text=True).stdout
    open(os.path.join(P, "r02_ncu_%s.txt" % short), "w").write(
        "# ncu --set full --clock-control none --import-source on, raw-page excerpt (scratch/ncu_raw_summary.py) of the captured launches\n" + a +
        "\n# source page of the first captured launch: stall reasons summed over the kernel, then the SASS lines with > 1.5 % of the samples (scratch/ncu_hot_spots.py)\n" + b)
