#!/bin/bash
# round-2 call 6: gather groups (GW) x accumulators sweep, d_weight prefetch; every command under a hard kill
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/c6_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c6_smoke.log
if ! grep -q "^smoke:" $O/c6_smoke.log; then echo "SMOKE FAILED - stopping"; exit 1; fi
$T 400 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5 > $O/c6_tests.log; tail -3 $O/c6_tests.log
for cfg in "MOPA_TC_GW=1" "MOPA_TC_GW=2" "MOPA_TC_GW=1 MOPA_TC_CTAS=1 MOPA_TC_ACC=2" "MOPA_TC_GW=1 MOPA_TC_CTAS=1 MOPA_TC_ACC=4" "MOPA_TC_GW=2 MOPA_TC_CTAS=1 MOPA_TC_ACC=4" "MOPA_TC_GW=1 MOPA_TC_CTAS=1 MOPA_TC_ACC=3 MOPA_TC_SB=3"; do
  f="$O/c6_bench_${cfg// /_}.json"
  env $cfg $T 120 python bench.py --no-cpu-baseline --no-roofline --steps 30 > "$f" 2>/dev/null; echo "$cfg: $(python -c "import json,sys; d=json.load(open('$f')); print('%.3f ms/step  e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']))" 2>&1 | tail -1)"
done
$T 200 python tools/layer_table.py --out $O/c6_layers.json > $O/c6_layers.log 2>&1; tail -7 $O/c6_layers.log
MOPA_TC_CTAS=1 MOPA_TC_ACC=4 $T 200 python tools/layer_table.py --out $O/c6_layers_wide.json > $O/c6_layers_wide.log 2>&1; tail -7 $O/c6_layers_wide.log
export MOPA_SCN_LIB=$PWD/scratch/bin/libmopa_scn_trace.so
$T 200 python scratch/tc_trace2.py 0 16 16 1 32 32 2 96 48 4 160 80 > $O/c6_trace.txt 2>&1
grep -v "per-step" $O/c6_trace.txt
