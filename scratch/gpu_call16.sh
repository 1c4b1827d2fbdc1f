#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/c27_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c27_smoke.log
if ! grep -q "^smoke:" $O/c27_smoke.log; then echo "SMOKE FAILED - stopping"; tail -30 $O/c27_smoke.log; exit 1; fi
for cfg in "X=0" "MOPA_TC_SPLIT=0" "MOPA_SCN_BNBWD_FUSION=1"; do
  tag="${cfg// /_}"
  for rep in 1 2; do
  f="$O/c27_bench_${tag}_$rep.json"
  env $cfg $T 150 python bench.py --no-cpu-baseline --no-roofline --no-fp32 --steps 60 > "$f" 2>$O/c27_err.txt; echo "$cfg: $(python -c "import json,sys; d=json.load(open('$f')); print('%.3f ms/step median %.3f e2e %.3f geometry %.3f' % (d['ms_per_step'], d['step_ms']['median'], d['e2e']['median_ms'], d['geometry']['ms_per_forward']))" 2>&1 | tail -1)"
  done
done
$T 900 python -m pytest tests -x -q -m gpu > $O/c27_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/c27_tests.log
