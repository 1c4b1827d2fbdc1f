#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/c29_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c29_smoke.log
if ! grep -q "^smoke:" $O/c29_smoke.log; then echo "SMOKE FAILED - stopping"; tail -30 $O/c29_smoke.log; exit 1; fi
for cfg in "MOPA_SCN_PDL=1" "MOPA_SCN_PDL=0" "MOPA_SCN_PDL=1"; do
  f="$O/c29_bench_$cfg.json"
  env $cfg $T 200 python bench.py --no-cpu-baseline --no-roofline --no-fp32 --steps 60 > "$f" 2>$O/c29_err.txt; echo "$cfg: $(python -c "import json,sys; d=json.load(open('$f')); print('%.3f ms/step median %.3f | e2e %.3f median %.3f | sync loop %.3f' % (d['ms_per_step'], d['step_ms']['median'], d['e2e']['ms_per_step'], d['e2e']['median_ms'], d['e2e']['sync_loop']['ms_per_step']))" 2>&1 | tail -1)"
done
tail -3 $O/c29_err.txt
$T 900 python -m pytest tests -x -q -m gpu > $O/c29_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/c29_tests.log
