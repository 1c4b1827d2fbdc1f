#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/c10_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c10_smoke.log
if ! grep -q "xm operators ok" $O/c10_smoke.log; then echo "SMOKE FAILED - stopping"; cat $O/c10_smoke.log | tail -20; exit 1; fi
for cfg in "X=0" "MOPA_TC_CTAS=3" "MOPA_SCN_BN_FUSED_MIN=0" "MOPA_SCN_BN_FUSED_MIN=16777216" "MOPA_TC_CTAS=3 MOPA_SCN_BN_FUSED_MIN=16777216"; do
  f="$O/c10_bench_${cfg// /_}.json"
  env $cfg $T 120 python bench.py --no-cpu-baseline --no-roofline --no-fp32 --steps 40 --warmup 10 > "$f" 2>$O/c10_err.txt; echo "$cfg: $(python -c "import json,sys; d=json.load(open('$f')); print('%.3f ms/step (median %.3f) e2e %.3f' % (d['ms_per_step'], d['step_ms']['median'], d['e2e']['ms_per_step']))" 2>&1 | tail -1)"
done
MOPA_TC_CTAS=3 $T 200 python tools/layer_table.py --out $O/c10_layers3.json > $O/c10_layers3.log 2>&1; tail -7 $O/c10_layers3.log
$T 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
MOPA_TC_CTAS=3 $T 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_level or unet_scn_forward or full_size" 2>&1 | tail -4
