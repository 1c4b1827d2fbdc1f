#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/c7_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c7_smoke.log
if ! grep -q "^smoke:" $O/c7_smoke.log; then echo "SMOKE FAILED - stopping"; exit 1; fi
for cfg in "MOPA_TC_GW=2" "MOPA_TC_GW=1" "MOPA_TC_GW=2 MOPA_TC_ACC=1"; do
  f="$O/c7_bench_${cfg// /_}.json"
  env $cfg $T 120 python bench.py --no-cpu-baseline --no-roofline --steps 30 > "$f" 2>/dev/null; echo "$cfg: $(python -c "import json,sys; d=json.load(open('$f')); print('%.3f ms/step  e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']))" 2>&1 | tail -1)"
done
$T 200 python tools/layer_table.py --out $O/c7_layers.json > $O/c7_layers.log 2>&1; tail -7 $O/c7_layers.log
export MOPA_SCN_LIB=$PWD/scratch/bin/libmopa_scn_trace.so
$T 200 python scratch/tc_trace2.py 0 16 16 2 96 48 > $O/c7_trace.txt 2>&1
grep -v "per-step" $O/c7_trace.txt
