#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/c17_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c17_smoke.log
if ! grep -q "xm operators ok" $O/c17_smoke.log; then echo "SMOKE FAILED - stopping"; cat $O/c17_smoke.log | tail -20; exit 1; fi
$T 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/c17_tests.log; grep -E "passed|failed|FAILED" $O/c17_tests.log
for sensor in nuscenes kitti; do
  f="$O/c17_bench_${sensor}.json"
  $T 200 python bench.py --sensor $sensor --no-cpu-baseline --no-fp32 --steps 50 --warmup 10 > "$f" 2>$O/c17_err.txt; echo "$sensor: $(python -c "import json,sys; d=json.load(open('$f')); print('%.3f ms/step (median %.3f) e2e %.3f (median %.3f) geometry %.3f ms launches/step %.0f' % (d['ms_per_step'], d['step_ms']['median'], d['e2e']['ms_per_step'], d['e2e']['median_ms'], d['geometry']['ms_per_forward'], d['gpu_launches']/d['steps']))" 2>&1 | tail -1)"
done
