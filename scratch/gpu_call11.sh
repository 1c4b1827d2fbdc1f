#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
for sensor in nuscenes kitti; do
for cfg in "MOPA_SCN_BN_FUSED_MIN=16777216" "MOPA_SCN_BN_FUSED_MIN=1000000000000" "MOPA_SCN_BN_FUSED_MIN=1000000000000 MOPA_TC_CTAS=3" "MOPA_SCN_BN_FUSED_MIN=1000000000000 MOPA_SCN_NO_DW_OVERLAP=1"; do
  f="$O/c11_bench_${sensor}_${cfg// /_}.json"
  env $cfg $T 120 python bench.py --sensor $sensor --no-cpu-baseline --no-roofline --no-fp32 --steps 40 --warmup 10 > "$f" 2>$O/c11_err.txt; echo "$sensor $cfg: $(python -c "import json,sys; d=json.load(open('$f')); print('%.3f ms/step (median %.3f) e2e %.3f (median %.3f)' % (d['ms_per_step'], d['step_ms']['median'], d['e2e']['ms_per_step'], d['e2e']['median_ms']))" 2>&1 | tail -1)"
done
done
MOPA_SCN_BN_FUSED_MIN=1000000000000 $T 200 python tools/layer_table.py --out $O/c11_layers.json > $O/c11_layers.log 2>&1; tail -7 $O/c11_layers.log
