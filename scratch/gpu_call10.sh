#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/c22_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c22_smoke.log
if ! grep -q "^smoke:" $O/c22_smoke.log; then echo "SMOKE FAILED - stopping"; tail -30 $O/c22_smoke.log; exit 1; fi
for cfg in "X=0" "MOPA_TC_BN_RING=0"; do
  tag="${cfg// /_}"
  env $cfg $T 200 python tools/layer_table.py --out "$O/c22_layers_$tag.json" > "$O/c22_layers_$tag.log" 2>&1; echo "== $cfg"; tail -6 "$O/c22_layers_$tag.log"
done
for cfg in "X=0" "MOPA_SCN_NO_BNSTATS_FUSION=1"; do
  f="$O/c22_bench_${cfg// /_}.json"
  env $cfg $T 150 python bench.py --no-cpu-baseline --no-roofline --no-fp32 --steps 50 > "$f" 2>$O/c22_err.txt; echo "$cfg: $(python -c "import json,sys; d=json.load(open('$f')); print('%.3f ms/step  e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']))" 2>&1 | tail -1)"
done
$T 900 python -m pytest tests -x -q -m gpu > $O/c22_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/c22_tests.log
