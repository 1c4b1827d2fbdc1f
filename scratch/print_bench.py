import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    e = d["e2e"]
    print("%s: N=%d %.3f ms/step median %.3f max %.2f = %.1f M pts/s | e2e %.3f (median %.3f) = %.1f M | sync loop %.3f | all-reduce %.0f us | slow %s" % (
        f.split("/")[-1], d["n_gpus"], d["ms_per_step"], d["step_ms"]["median"], d["step_ms"]["max"], d["value"] / 1e6, e["ms_per_step"], e["median_ms"],
        e["value"] / 1e6, e.get("sync_loop", {}).get("ms_per_step", float("nan")), d.get("all_reduce_us_median") or 0, d["step_ms"].get("slow_steps_rank0")))
