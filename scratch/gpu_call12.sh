#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/c12_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c12_smoke.log
if ! grep -q "xm operators ok" $O/c12_smoke.log; then echo "SMOKE FAILED - stopping"; cat $O/c12_smoke.log | tail -20; exit 1; fi
$T 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for sensor in nuscenes kitti; do
  f="$O/c12_bench_${sensor}.json"
  $T 200 python bench.py --sensor $sensor --no-cpu-baseline --no-fp32 --steps 50 --warmup 10 > "$f" 2>$O/c12_err.txt; echo "$sensor: $(python -c "import json,sys; d=json.load(open('$f')); print('%.3f ms/step (median %.3f) e2e %.3f (median %.3f) geometry %.3f ms' % (d['ms_per_step'], d['step_ms']['median'], d['e2e']['ms_per_step'], d['e2e']['median_ms'], d['geometry']['ms_per_forward']))" 2>&1 | tail -1)"
done
