#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 200 python bench.py --no-cpu-baseline --no-fp32 --no-roofline --steps 50 --warmup 10 > $O/c18_bench.json 2>$O/c18_err.txt; python -c "import json,sys; d=json.load(open('$O/c18_bench.json')); print('%.3f ms/step (median %.3f max %.3f) e2e %.3f (median %.3f p90 %.3f max %.3f)' % (d['ms_per_step'], d['step_ms']['median'], d['step_ms']['max'], d['e2e']['ms_per_step'], d['e2e']['median_ms'], d['e2e']['p90_ms'], d['e2e']['max_ms']))"
$T 200 python tools/layer_table.py --out $O/c18_layers.json > $O/c18_layers.log 2>&1; tail -7 $O/c18_layers.log
$T 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dweight or strided or submanifold_conv or large_level" 2>&1 | tail -3
